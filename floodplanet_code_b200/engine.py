"""Execution plan of the UNet hot path on one B200: which kernel runs on which buffer.

Mirrors the wiring of the reference network (st_water_seg/models/unet.py:80-111): inc, four
Down stages (MaxPool2d(2) + DoubleConv), four Up stages (bilinear x2 + pad + cat([skip, up])
+ DoubleConv with mid = in/2) and the 1x1 OutConv -- but as a flat schedule of C-ABI kernel
launches over NHWC bf16 buffers:

  * skip tensors are written straight into the first half of their concat buffer and the
    upsampled tensor into the second half ("virtual concat": torch.cat never runs);
  * BatchNorm batch statistics come out of the conv epilogue; normalise+ReLU is one pass,
    fused with the 2x2 max-pool where a Down stage follows;
  * in eval mode BatchNorm (+conv bias) folds into the conv epilogue (scale/shift/ReLU);
  * backward walks the same schedule in reverse: BN/ReLU backward (reduce + apply), wgrad,
    dgrad, and the structural gradients (concat split, upsample gather, pool scatter + skip
    add).  Parameter gradients land in ONE flat fp32 slab laid out in reverse-forward order
    so data-parallel buckets can be all-reduced while earlier layers are still running.

Only torch allocation / stream plumbing happens here; every FLOP is in csrc/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def pad_channels(c: int) -> int:
    """Channel padding of the network input: 16 / 32 / multiple of 64 (TMA swizzle widths)."""
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    return ((c + 63) // 64) * 64


@dataclass
class ConvSpec:
    conv: str      # state_dict prefix of the conv   (e.g. 'inc.double_conv.0')
    bn: str        # state_dict prefix of its BatchNorm
    cin: int
    cout: int
    level: int     # resolution level 0..4


def unet_conv_specs(n_channels: int) -> List[ConvSpec]:
    """The 18 conv3x3+BN+ReLU layers in forward order (unet.py:88-97, bilinear=True)."""
    specs: List[ConvSpec] = []

    def dc(prefix: str, cin: int, cout: int, level: int, mid: Optional[int] = None):
        mid = mid or cout
        specs.append(ConvSpec(f"{prefix}.0", f"{prefix}.1", cin, mid, level))
        specs.append(ConvSpec(f"{prefix}.3", f"{prefix}.4", mid, cout, level))

    dc("inc.double_conv", n_channels, 64, 0)
    dc("down1.maxpool_conv.1.double_conv", 64, 128, 1)
    dc("down2.maxpool_conv.1.double_conv", 128, 256, 2)
    dc("down3.maxpool_conv.1.double_conv", 256, 512, 3)
    dc("down4.maxpool_conv.1.double_conv", 512, 512, 4)
    dc("up1.conv.double_conv", 1024, 256, 3, 512)
    dc("up2.conv.double_conv", 512, 128, 2, 256)
    dc("up3.conv.double_conv", 256, 64, 1, 128)
    dc("up4.conv.double_conv", 128, 64, 0, 64)
    return specs


def param_names(n_channels: int) -> List[str]:
    """Trainable parameter names in forward order (conv w, conv b, bn w, bn b per layer, head)."""
    names: List[str] = []
    for s in unet_conv_specs(n_channels):
        names += [f"{s.conv}.weight", f"{s.conv}.bias", f"{s.bn}.weight", f"{s.bn}.bias"]
    names += ["outc.conv.weight", "outc.conv.bias"]
    return names


class PackedWeights:
    """bf16 GEMM-operand copies of the fp32 OIHW master weights, refreshed when a parameter's
    version counter moves (optimizer step, load_state_dict)."""

    def __init__(self):
        self._fprop: Dict[str, Tuple[int, torch.Tensor]] = {}
        self._dgrad: Dict[str, Tuple[int, torch.Tensor]] = {}
        self.generation = 0
        # when True every lookup re-packs into the SAME buffer (CUDA-graph capture: the replayed
        # graph must contain the repack kernels because the optimiser changes the masters)
        self.always_repack = False

    def invalidate(self) -> None:
        """For updates torch cannot see (raw-pointer kernels such as the fused Adam)."""
        self.generation += 1

    def _key(self, w: torch.Tensor):
        return (w._version, w.data_ptr(), w.device, self.generation)

    def fprop(self, name: str, w: torch.Tensor, cin_pad: int) -> torch.Tensor:
        key = self._key(w)
        hit = self._fprop.get(name)
        if hit is None or hit[0] != key or self.always_repack:
            buf = hit[1] if hit is not None and hit[1].device == w.device else None
            self._fprop[name] = (key, ops.repack_fprop(w, cin_pad, buf))
        return self._fprop[name][1]

    def dgrad(self, name: str, w: torch.Tensor) -> torch.Tensor:
        key = self._key(w)
        hit = self._dgrad.get(name)
        if hit is None or hit[0] != key or self.always_repack:
            buf = hit[1] if hit is not None and hit[1].device == w.device else None
            self._dgrad[name] = (key, ops.repack_dgrad(w, buf))
        return self._dgrad[name][1]


@dataclass
class LayerSaved:
    x: torch.Tensor                 # conv input view (NHWC bf16)
    y: torch.Tensor                 # raw conv output (NHWC bf16, pre-BN)
    scale: torch.Tensor
    shift: torch.Tensor
    mean: torch.Tensor
    invstd: torch.Tensor


@dataclass
class ForwardState:
    """Everything backward needs (owned by the autograd node)."""
    n: int = 0
    sizes: List[Tuple[int, int]] = field(default_factory=list)
    layers: List[LayerSaved] = field(default_factory=list)
    pool_idx: List[torch.Tensor] = field(default_factory=list)
    head_in: Optional[torch.Tensor] = None


class UNetEngine:
    """Stateless w.r.t. parameters: they are passed in as a name -> tensor dict each call."""

    def __init__(self, n_channels: int, n_classes: int):
        if n_classes < 1 or n_classes > 8:
            raise RuntimeError(f"floodplanet_b200: n_classes={n_classes} unsupported (1..8)")
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.cin_pad = pad_channels(n_channels)
        if self.cin_pad > 64:
            raise RuntimeError(f"floodplanet_b200: {n_channels} input channels unsupported (max 64)")
        self.specs = unet_conv_specs(n_channels)
        self.names = param_names(n_channels)
        self.packed = PackedWeights()
        # optional hook(flat_grad_slab, start, end): called in backward as soon as the grads
        # in slab[start:end] are final (used for bucketed data-parallel all-reduce)
        self.grad_ready_hook: Optional[Callable[[torch.Tensor, int, int], None]] = None
        # optional hook(flat_grad_slab, total): called once at the end of backward
        self.grad_done_hook: Optional[Callable[[torch.Tensor, int], None]] = None
        # optional profiler: when a list, (tag, flops, start_event, end_event) per conv launch
        self.conv_events: Optional[list] = None
        # backward: optionally run every wgrad on a second stream.  wgrad is off the critical path
        # (only the optimiser needs dW) and could overlap the HBM-bound BatchNorm-backward passes
        # of the next layer.  Measured on B200 at batch 64: <= 1 % gain (both kernels are persistent
        # and the step is power-capped, so overlap buys no energy) and more run-to-run variance
        # from cross-stream allocator bookkeeping -- off by default.
        self.overlap_wgrad = False
        self._side_stream: Optional[torch.cuda.Stream] = None
        self.launches = 0  # kernels launched by the last forward/backward (for bench accounting)

    # -------------------------------------------------------------------------------------
    @staticmethod
    def _level_sizes(h: int, w: int) -> List[Tuple[int, int]]:
        sizes = [(h, w)]
        for _ in range(4):
            h, w = h // 2, w // 2
            sizes.append((h, w))
        if sizes[-1][0] < 1 or sizes[-1][1] < 1:
            raise RuntimeError("floodplanet_b200: input must be at least 16x16")
        return sizes

    def _cin_pad_of(self, i: int) -> int:
        return self.cin_pad if i == 0 else self.specs[i].cin

    def _timed(self, tag: str, i: int, n: int, hw: int, fn) -> None:
        """Run fn(); when profiling, bracket it with CUDA events on the launching stream and
        record the layer's algorithmic FLOPs (2 * pixels * Cout * 9 * Cin, real channels)."""
        if self.conv_events is None:
            fn()
            return
        s = self.specs[i]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.conv_events.append((tag, i, 2.0 * n * hw * s.cout * 9 * s.cin, e0, e1))

    # -------------------------------------------------------------------------------------
    def forward(self, images: Optional[Sequence[torch.Tensor]], params: Dict[str, torch.Tensor],
                buffers: Dict[str, torch.Tensor], training: bool, save: bool,
                ingested: Optional[torch.Tensor] = None):
        """images: list of NCHW fp32 tensors (concatenated along C in order), or `ingested`: an
        already channel-padded NHWC bf16 batch (tile-sharded inference ingests straight from the
        resident scene).  Returns (logits fp32 NCHW, ForwardState or None)."""
        if ingested is not None:
            dev = ingested.device
            n, h, w, cp = ingested.shape
            if cp != self.cin_pad:
                raise RuntimeError(f"floodplanet_b200: ingested batch has {cp} channels, expected {self.cin_pad}")
        else:
            dev = images[0].device
            n, _, h, w = images[0].shape
            if sum(int(t.shape[1]) for t in images) != self.n_channels:
                raise RuntimeError(
                    f"floodplanet_b200: expected {self.n_channels} input channels, got "
                    f"{[int(t.shape[1]) for t in images]}")
        sizes = self._level_sizes(h, w)
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        st = ForwardState(n=n, sizes=sizes) if save else None
        launches = 0

        x = ingested if ingested is not None else ops.ingest(images, self.cin_pad)
        launches += 1

        # concat buffers for the four Up stages: [skip | upsampled], at levels 3,2,1,0
        cat: Dict[int, torch.Tensor] = {}
        for lvl, c in ((3, 512), (2, 256), (1, 128), (0, 64)):
            hh, ww = sizes[lvl]
            cat[lvl] = torch.empty((n, hh, ww, 2 * c), **bf)

        stat_rows = ops.stat_rows()

        def conv_bn_relu(i: int, xin: torch.Tensor, out_view: Optional[torch.Tensor],
                         pool_to: Optional[torch.Tensor], defer_apply: bool = False):
            """Layer i on xin.  Activation goes to out_view (or a fresh tensor); if pool_to is
            given the 2x2 max-pool of the activation is written there too.  With defer_apply
            (training only) the normalise+ReLU pass is left to the consumer kernel and the raw
            conv output plus its (scale, shift) are returned instead."""
            nonlocal launches
            s = self.specs[i]
            hh, ww = sizes[s.level]
            wt = params[f"{s.conv}.weight"]
            wp = self.packed.fprop(s.conv, wt, self._cin_pad_of(i))
            gamma, beta = params[f"{s.bn}.weight"], params[f"{s.bn}.bias"]
            bias = params[f"{s.conv}.bias"]
            scale = torch.empty(s.cout, **f32)
            shift = torch.empty(s.cout, **f32)
            a = out_view if out_view is not None else torch.empty((n, hh, ww, s.cout), **bf)
            if training:
                y = torch.empty((n, hh, ww, s.cout), **bf)
                parts = torch.empty((stat_rows, 2, s.cout), **f32)
                self._timed("fprop", i, n, hh * ww,
                            lambda: ops.conv3x3_fprop(xin, wp, y, stat_partials=parts))
                mean = torch.empty(s.cout, **f32)
                invstd = torch.empty(s.cout, **f32)
                ops.bn_stats_finalize(parts, n * hh * ww, gamma, beta, bias, BN_EPS, BN_MOMENTUM,
                                      buffers[f"{s.bn}.running_mean"], buffers[f"{s.bn}.running_var"],
                                      scale, shift, mean, invstd)
                buffers[f"{s.bn}.num_batches_tracked"].add_(1)
                if defer_apply:
                    a = None
                elif pool_to is not None:
                    idx = torch.empty(pool_to.shape, dtype=torch.uint8, device=dev)
                    ops.bn_apply_relu_maxpool2(y, a, pool_to, idx, scale, shift)
                    if save:
                        st.pool_idx.append(idx)
                else:
                    ops.bn_apply_relu(y, a, scale, shift)
                launches += 4  # memset+conv counted as conv(2), finalize, apply
                if save:
                    st.layers.append(LayerSaved(xin, y, scale, shift, mean, invstd))
                if defer_apply:
                    return y, scale, shift
            else:
                ops.bn_fold_eval(gamma, beta, bias, buffers[f"{s.bn}.running_mean"],
                                 buffers[f"{s.bn}.running_var"], BN_EPS, scale, shift)
                self._timed("fprop", i, n, hh * ww,
                            lambda: ops.conv3x3_fprop(xin, wp, a, scale=scale, shift=shift, relu=True))
                launches += 2
                if pool_to is not None:
                    idx = torch.empty(pool_to.shape, dtype=torch.uint8, device=dev)
                    ops.bn_apply_relu_maxpool2(a, None, pool_to, idx, None, None)
                    launches += 1
            return a

        # ---------------- encoder ----------------
        cur = x
        li = 0
        enc_ch = (64, 128, 256, 512)
        for lvl in range(4):
            c = enc_ch[lvl]
            cur = conv_bn_relu(li, cur, None, None); li += 1
            hp, wp_ = sizes[lvl + 1]
            pooled = torch.empty((n, hp, wp_, c), **bf)
            conv_bn_relu(li, cur, cat[lvl][..., :c], pooled); li += 1
            cur = pooled
        cur = conv_bn_relu(li, cur, None, None); li += 1
        cur = conv_bn_relu(li, cur, None, None); li += 1  # x5 (bottleneck)

        # ---------------- decoder ----------------
        for lvl in (3, 2, 1, 0):
            c = enc_ch[lvl]
            ops.upsample2x_pad_concat_fwd(cur, cat[lvl][..., c:])
            launches += 1
            cur = conv_bn_relu(li, cat[lvl], None, None); li += 1
            if lvl == 0 and training:
                # last layer: its normalise+ReLU is fused into the head kernel (forward) and its
                # BatchNorm-backward reduction into the head backward -- no activation is stored
                cur, head_scale, head_shift = conv_bn_relu(li, cur, None, None, defer_apply=True)
            else:
                cur, head_scale, head_shift = conv_bn_relu(li, cur, None, None), None, None
            li += 1

        # ---------------- head ----------------
        logits = torch.empty((n, self.n_classes, h, w), **f32)
        wh = params["outc.conv.weight"].detach().reshape(self.n_classes, 64)
        ops.head1x1_fwd(cur, wh, params["outc.conv.bias"].detach(), logits, head_scale, head_shift)
        launches += 1
        if save:
            st.head_in = cur   # raw output y of the last conv (training)
        self.launches = launches
        return logits, st

    # -------------------------------------------------------------------------------------
    def grad_layout(self, params: Dict[str, torch.Tensor]) -> Tuple[Dict[str, Tuple[int, int]], int]:
        """Offsets of every parameter's gradient inside the flat slab, in REVERSE forward order
        (the order backward produces them), each 16-byte aligned."""
        off = 0
        layout: Dict[str, Tuple[int, int]] = {}
        for name in reversed(self.names):
            nel = params[name].numel()
            layout[name] = (off, nel)
            off += (nel + 3) // 4 * 4
        return layout, off

    def backward(self, st: ForwardState, dlogits: torch.Tensor, params: Dict[str, torch.Tensor]):
        """Returns {param name: fp32 gradient view} (views of one flat slab)."""
        dev = dlogits.device
        n = st.n
        sizes = st.sizes
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        layout, total = self.grad_layout(params)
        slab = torch.zeros(total, **f32)  # conv-bias grads stay exactly 0 (cancelled by BN)
        grads = {k: slab[o:o + nel].view(params[k].shape) for k, (o, nel) in layout.items()}
        launches = 1
        ready_upto = 0
        main = torch.cuda.current_stream(dev)
        side = None
        if self.overlap_wgrad and self.conv_events is None:
            if self._side_stream is None or self._side_stream.device != dev:
                self._side_stream = torch.cuda.Stream(device=dev)
            side = self._side_stream
        # one split-K workspace for all layers (the wgrads are serialised on one stream)
        ws_need = max(ops.wgrad_workspace_bytes(n, sizes[sp.level][0], sizes[sp.level][1],
                                                self._cin_pad_of(j), sp.cout)
                      for j, sp in enumerate(self.specs)) // 4
        ws = torch.empty(ws_need, **f32)

        def mark_ready(name_last: str):
            """All grads from slab[ready_upto] through `name_last` are final (enqueued)."""
            nonlocal ready_upto
            o, nel = layout[name_last]
            end = (o + nel + 3) // 4 * 4
            if self.grad_ready_hook is not None and end > ready_upto:
                if side is not None:
                    main.wait_stream(side)   # weight grads of this bucket come from the side stream
                self.grad_ready_hook(slab, ready_upto, end)
            ready_upto = end

        # ---------------- head ----------------
        h0, w0 = sizes[0]
        d_cur = torch.empty((n, h0, w0, 64), **bf)
        parts = torch.empty((ops.head_bwd_rows(), self.n_classes * 65), **f32)
        wh = params["outc.conv.weight"].detach().reshape(self.n_classes, 64)
        last = st.layers[-1]
        head_bn_parts = torch.empty((ops.head_bwd_rows(), 2, 64), **f32)
        ops.head1x1_bwd(dlogits.contiguous(), st.head_in, wh, d_cur,
                        grads["outc.conv.weight"].view(self.n_classes, 64), grads["outc.conv.bias"],
                        parts, bn=(last.scale, last.shift, last.mean, last.invstd),
                        bn_partials=head_bn_parts)
        launches += 2
        mark_ready("outc.conv.weight")

        bn_rows = ops.bn_bwd_rows()
        stat_rows_bwd = ops.stat_rows()
        fused_parts: Dict[int, torch.Tensor] = {}

        def layer_backward(i: int, da: torch.Tensor, need_dx: bool,
                           dx_out: Optional[torch.Tensor] = None,
                           bn_parts: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
            """bn_parts: BatchNorm-backward partial sums already produced by the kernel that
            wrote `da` (then the separate reduction pass is skipped)."""
            nonlocal launches
            s = self.specs[i]
            sv = st.layers[i]
            hh, ww = sizes[s.level]
            if bn_parts is None:
                bn_parts = fused_parts.pop(i, None)
            if bn_parts is not None:
                parts = bn_parts
            else:
                parts = torch.empty((bn_rows, 2, s.cout), **f32)
                ops.bn_relu_bwd_reduce(da, sv.y, sv.scale, sv.shift, sv.mean, sv.invstd, parts)
            coef = torch.empty((2, s.cout), **f32)
            ops.bn_bwd_finalize(parts, n * hh * ww, sv.scale, sv.mean, sv.invstd,
                                grads[f"{s.bn}.weight"], grads[f"{s.bn}.bias"], coef)
            dy = torch.empty((n, hh, ww, s.cout), **bf)
            ops.bn_relu_bwd_apply(da, sv.y, dy, sv.scale, sv.shift, coef)
            if side is not None:
                side.wait_stream(main)                      # dy is ready
                with torch.cuda.stream(side):
                    ops.conv3x3_wgrad(sv.x, dy, grads[f"{s.conv}.weight"], ws, s.cin)
                dy.record_stream(side)                      # keep dy alive until the side stream is done
            else:
                self._timed("wgrad", i, n, hh * ww,
                            lambda: ops.conv3x3_wgrad(sv.x, dy, grads[f"{s.conv}.weight"], ws, s.cin))
            launches += 5
            dx = None
            if need_dx:
                dx = dx_out if dx_out is not None else torch.empty((n, hh, ww, s.cin), **bf)
                wd = self.packed.dgrad(s.conv, params[f"{s.conv}.weight"])
                if i % 2 == 1 and dx_out is None and (s.cin == 64 or s.cout >= 256):
                    # second conv of a DoubleConv: dx IS the activation gradient of layer i-1, so
                    # that layer's BatchNorm-backward reduction rides in this dgrad's epilogue.
                    # (Measured: pays off when the epilogue has slack -- 8 epilogue warps for
                    # 64-channel outputs, or >= 2304-deep GEMMs; for the 128-channel layers with
                    # short K the epilogue becomes critical and the separate pass is cheaper.)
                    pv = st.layers[i - 1]
                    fparts = torch.empty((stat_rows_bwd, 2, s.cin), **f32)
                    self._timed("dgrad", i, n, hh * ww,
                                lambda: ops.conv3x3_dgrad(dy, wd, dx, bn_y=pv.y,
                                                          bn=(pv.scale, pv.shift, pv.mean, pv.invstd),
                                                          bn_partials=fparts))
                    fused_parts[i - 1] = fparts
                else:
                    self._timed("dgrad", i, n, hh * ww, lambda: ops.conv3x3_dgrad(dy, wd, dx))
                launches += 1
            mark_ready(f"{s.conv}.weight")
            return dx

        enc_ch = (64, 128, 256, 512)
        li = len(self.specs) - 1
        dcat: Dict[int, torch.Tensor] = {}
        # ---------------- decoder (reverse) ----------------
        for lvl in (0, 1, 2, 3):
            c = enc_ch[lvl]
            hh, ww = sizes[lvl]
            d_mid = layer_backward(li, d_cur, True, bn_parts=head_bn_parts if lvl == 0 else None); li -= 1
            dcat[lvl] = torch.empty((n, hh, ww, 2 * c), **bf)
            layer_backward(li, d_mid, True, dcat[lvl]); li -= 1
            hl, wl = sizes[lvl + 1]
            d_cur = torch.empty((n, hl, wl, c), **bf)
            ops.upsample2x_pad_concat_bwd(dcat[lvl][..., c:], d_cur)
            launches += 1
        # ---------------- bottleneck + encoder (reverse) ----------------
        d_mid = layer_backward(li, d_cur, True); li -= 1          # down4 second conv
        d_pool = layer_backward(li, d_mid, True); li -= 1         # down4 first conv -> d(pooled x4)
        for lvl in (3, 2, 1, 0):
            c = enc_ch[lvl]
            hh, ww = sizes[lvl]
            d_skip = torch.empty((n, hh, ww, c), **bf)
            # the skip layer's BatchNorm-backward reduction rides in the pool-backward pass
            sk = st.layers[li]
            pool_parts = torch.empty((2 * bn_rows, 2, c), **f32)
            ops.maxpool2_bwd(d_pool, st.pool_idx[lvl], dcat[lvl][..., :c], d_skip, bn_y=sk.y,
                             bn=(sk.scale, sk.shift, sk.mean, sk.invstd), bn_partials=pool_parts)
            launches += 1
            del dcat[lvl]
            d_mid = layer_backward(li, d_skip, True, bn_parts=pool_parts); li -= 1
            d_pool = layer_backward(li, d_mid, lvl > 0); li -= 1
        assert li == -1
        if side is not None:
            main.wait_stream(side)
            ws.record_stream(side)
        mark_ready(self.names[0])
        if self.grad_done_hook is not None:
            self.grad_done_hook(slab, total)
        self.launches = launches
        return grads, slab
