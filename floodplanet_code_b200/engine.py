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
MAX_INPUT_CHANNELS = 256     # after early fusion, padded (csrc/elementwise.cu: kIngestMaxC)


def pad_channels(c: int) -> int:
    """Channel padding of the network input: 16 / 32 / multiple of 64 (TMA swizzle widths)."""
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    return ((c + 63) // 64) * 64


@dataclass
class ConvSpec:
    conv: str      # state_dict name of the conv   (e.g. 'inc.double_conv.0')
    bn: str        # state_dict name of its BatchNorm
    cin: int
    cout: int
    level: int     # resolution level 0..4
    second: bool = False   # second conv of its DoubleConv
    idx: int = 0           # position in the owning engine's layer list (profiling tag)


ENC_CH = (64, 128, 256, 512)           # skip-feature channels at levels 0..3
FEAT_CH = (64, 128, 256, 512, 512)     # the five encoder features x1..x5 (unet.py:113-120)


def _double_conv_specs(specs: List[ConvSpec], prefix: str, cin: int, cout: int, level: int,
                       mid: Optional[int] = None) -> None:
    mid = mid or cout
    specs.append(ConvSpec(f"{prefix}.0", f"{prefix}.1", cin, mid, level, False))
    specs.append(ConvSpec(f"{prefix}.3", f"{prefix}.4", mid, cout, level, True))


def encoder_conv_specs(n_channels: int, prefix: str = "") -> List[ConvSpec]:
    """The 10 encoder layers (unet.py:88-93 / UNetEncoder :143-148), names prefixed."""
    specs: List[ConvSpec] = []
    _double_conv_specs(specs, f"{prefix}inc.double_conv", n_channels, 64, 0)
    _double_conv_specs(specs, f"{prefix}down1.maxpool_conv.1.double_conv", 64, 128, 1)
    _double_conv_specs(specs, f"{prefix}down2.maxpool_conv.1.double_conv", 128, 256, 2)
    _double_conv_specs(specs, f"{prefix}down3.maxpool_conv.1.double_conv", 256, 512, 3)
    _double_conv_specs(specs, f"{prefix}down4.maxpool_conv.1.double_conv", 512, 512, 4)
    return specs


def decoder_conv_specs(prefix: str = "") -> List[ConvSpec]:
    """The 8 decoder layers (unet.py:94-97 / UNetDecoder :176-182, bilinear, channel_factor 1)."""
    specs: List[ConvSpec] = []
    _double_conv_specs(specs, f"{prefix}up1.conv.double_conv", 1024, 256, 3, 512)
    _double_conv_specs(specs, f"{prefix}up2.conv.double_conv", 512, 128, 2, 256)
    _double_conv_specs(specs, f"{prefix}up3.conv.double_conv", 256, 64, 1, 128)
    _double_conv_specs(specs, f"{prefix}up4.conv.double_conv", 128, 64, 0, 64)
    return specs


def unet_conv_specs(n_channels: int) -> List[ConvSpec]:
    """The 18 conv3x3+BN+ReLU layers in forward order (unet.py:88-97, bilinear=True)."""
    return encoder_conv_specs(n_channels) + decoder_conv_specs()


def conv_param_names(specs: Sequence[ConvSpec]) -> List[str]:
    names: List[str] = []
    for s in specs:
        names += [f"{s.conv}.weight", f"{s.conv}.bias", f"{s.bn}.weight", f"{s.bn}.bias"]
    return names


def param_names(n_channels: int) -> List[str]:
    """Trainable parameter names in forward order (conv w, conv b, bn w, bn b per layer, head)."""
    return conv_param_names(unet_conv_specs(n_channels)) + ["outc.conv.weight", "outc.conv.bias"]


class _RawWriteEpochs:
    """Process-wide counters for writes torch's version counters cannot see: kernels that update
    parameters through raw pointers (fused Adam, CUDA-graph replays of it), `.data` writes
    (`dist.broadcast(p.data)`) and training-mode BatchNorm kernels rewriting the running
    statistics.  Every engine of every module keys its caches (packed bf16 weight copies, folded
    eval-mode BatchNorm coefficients) on them, so one bump invalidates the `_engine`,
    `_enc_engine` and `_dec_engine` views of the same parameters alike."""
    param = 0
    bn = 0


def note_raw_parameter_write() -> None:
    """Call after parameters / buffers were modified behind torch's back (see _RawWriteEpochs)."""
    _RawWriteEpochs.param += 1
    _RawWriteEpochs.bn += 1


class PackedWeights:
    """bf16 GEMM-operand copies of the fp32 OIHW master weights, refreshed when a parameter's
    version counter moves (optimizer step, load_state_dict) or a raw write was announced."""

    def __init__(self):
        self._fprop: Dict[str, Tuple[int, torch.Tensor]] = {}
        self._dgrad: Dict[str, Tuple[int, torch.Tensor]] = {}
        # when True every lookup re-packs into the SAME buffer (CUDA-graph capture: the replayed
        # graph must contain the repack kernels because the optimiser changes the masters)
        self.always_repack = False
        # one-launch refresh of every 3x3 copy: name -> (w, cin_pad) per table, device record array
        self._meta: Dict[Tuple[int, str], Tuple[torch.Tensor, int]] = {}
        self._batch = None          # (signature, table tensor, n_entries, total_blocks, [(table, name)])

    def refresh_all(self) -> int:
        """Re-pack every known 3x3 operand copy in ONE kernel launch and mark the copies current (the
        fused Adam calls this right after its update: all masters changed).  Returns launches made."""
        if self.always_repack or not self._meta:
            return 0
        items = sorted(self._meta.items())
        tables = (self._fprop, self._dgrad)
        sig = tuple((k, w.data_ptr(), tables[k[0]][k[1]][1].data_ptr()) for k, (w, _) in items)
        if self._batch is None or self._batch[0] != sig:
            entries = [(w, tables[k[0]][k[1]][1], cin_pad, k[0]) for k, (w, cin_pad) in items]
            table, total = ops.repack_batch_table(entries, items[0][1][0].device)
            self._batch = (sig, table, len(entries), total)
        _, table, n, total = self._batch
        ops.repack_batch(table, n, total)
        for k, (w, _) in items:
            tbl = tables[k[0]]
            tbl[k[1]] = (self._key(w), tbl[k[1]][1])
        return 1

    def invalidate(self) -> None:
        """For updates torch cannot see (raw-pointer kernels such as the fused Adam)."""
        note_raw_parameter_write()

    @property
    def generation(self) -> int:
        return _RawWriteEpochs.param

    def _key(self, w: torch.Tensor):
        return (w._version, w.data_ptr(), w.device, _RawWriteEpochs.param)

    def _lookup(self, table, name: str, w: torch.Tensor, pack):
        key = self._key(w)
        hit = table.get(name)
        if hit is None or hit[0] != key or self.always_repack:
            buf = hit[1] if hit is not None and hit[1].device == w.device else None
            table[name] = (key, pack(buf))
        return table[name][1]

    def fprop(self, name: str, w: torch.Tensor, cin_pad: int) -> torch.Tensor:
        self._meta[(0, name)] = (w, cin_pad)
        return self._lookup(self._fprop, name, w, lambda buf: ops.repack_fprop(w, cin_pad, buf))

    def dgrad(self, name: str, w: torch.Tensor) -> torch.Tensor:
        self._meta[(1, name)] = (w, w.shape[1])
        return self._lookup(self._dgrad, name, w, lambda buf: ops.repack_dgrad(w, buf))

    def fprop_1x1(self, name: str, w: torch.Tensor) -> torch.Tensor:
        return self._lookup(self._fprop, name, w, lambda buf: ops.repack_1x1(w, False, buf))

    def dgrad_1x1(self, name: str, w: torch.Tensor) -> torch.Tensor:
        return self._lookup(self._dgrad, name, w, lambda buf: ops.repack_1x1(w, True, buf))


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
    frozen: bool = False            # forward ran with eval-mode (running-statistics) BatchNorm
    layers: List[LayerSaved] = field(default_factory=list)          # UNet / decoder / single encoder
    pool_idx: List[torch.Tensor] = field(default_factory=list)
    head_in: Optional[torch.Tensor] = None
    # late fusion: one (layers, pool_idx) pair per modality encoder and the concatenated features
    enc_layers: List[List[LayerSaved]] = field(default_factory=list)
    enc_pool_idx: List[List[torch.Tensor]] = field(default_factory=list)
    fused_in: List[torch.Tensor] = field(default_factory=list)
    order: List[str] = field(default_factory=list)                  # modality order of this batch


class _Fwd:
    """Per-call forward context."""

    def __init__(self, n, sizes, dev, params, buffers, training, save):
        self.n, self.sizes, self.dev = n, sizes, dev
        self.params, self.buffers = params, buffers
        self.training, self.save = training, save
        # eval-mode forward that must be differentiable ("frozen BatchNorm"): the two-pass training schedule
        # (raw conv output kept, normalise + ReLU as its own pass) with the running statistics as constants
        self.frozen = save and not training
        self.bf = dict(dtype=torch.bfloat16, device=dev)
        self.f32 = dict(dtype=torch.float32, device=dev)
        self.launches = 0
        self.stat_rows = ops.stat_rows()
        self.nbt: List[torch.Tensor] = []      # num_batches_tracked counters bumped by this call

    def flush_nbt(self) -> None:
        """`num_batches_tracked += 1` of every BatchNorm run so far, as one multi-tensor op instead of
        one tiny launch per layer."""
        if self.nbt:
            torch._foreach_add_(self.nbt, 1)
            self.nbt = []


class _Bwd:
    """Per-call backward context: the flat gradient slab and its bookkeeping."""

    def __init__(self, n, sizes, dev, params, layout, total, frozen=False):
        self.n, self.sizes, self.dev = n, sizes, dev
        self.params = params
        self.frozen = frozen
        self.bf = dict(dtype=torch.bfloat16, device=dev)
        self.f32 = dict(dtype=torch.float32, device=dev)
        self.layout, self.total = layout, total
        self.slab = torch.zeros(total, **self.f32)  # conv-bias grads stay exactly 0 (cancelled by BN)
        self.grads = {k: self.slab[o:o + nel].view(params[k].shape) for k, (o, nel) in layout.items()}
        self.launches = 1
        self.ready_upto = 0
        self.main = torch.cuda.current_stream(dev)
        self.side: Optional[torch.cuda.Stream] = None
        self.ws: Optional[torch.Tensor] = None
        self.bn_rows = ops.bn_bwd_rows()
        self.stat_rows = ops.stat_rows()
        self.fused_parts: Dict[str, torch.Tensor] = {}


class _Schedule:
    """Kernel schedules shared by the engines: one conv3x3+BN+ReLU layer forward/backward, the
    encoder and the decoder halves of the network.  Stateless w.r.t. parameters: they are passed
    in as a name -> tensor dict each call."""

    def __init__(self):
        self.packed = PackedWeights()
        # optional hook(flat_grad_slab, start, end, side_stream): called in backward as soon as the grads
        # in slab[start:end] are final = ENQUEUED on the current stream and, for weight gradients when
        # wgrads overlap, on side_stream (None otherwise); used for the bucketed data-parallel all-reduce
        self.grad_ready_hook: Optional[Callable] = None
        # optional hook(flat_grad_slab, total): called once at the end of backward
        self.grad_done_hook: Optional[Callable[[torch.Tensor, int], None]] = None
        # optional profiler: when a list, (tag, flops, start_event, end_event) per conv launch
        self.conv_events: Optional[list] = None
        # backward: every wgrad runs on a second stream.  wgrad is off the critical path (only the
        # optimiser needs dW) and overlaps the HBM-bound BatchNorm-backward passes of the next layer.
        # The step is power-capped, so the overlap raises power and lowers the SM clock (1608 -> 1567 MHz);
        # net gain measured in round 2 on one box, alternating A/B, 3 x 30 steps each: 900.8 -> 912.3
        # chips/s (+1.3 %), end to end 892 -> 908 (+1.8 %), every overlapped run above every serial one
        # (profiles/r02_wgrad_overlap_ab.md).  Set False for a strictly single-stream backward.
        self.overlap_wgrad = True
        self.wgrad_after_dgrad = True      # see _layer_backward: launch order of the two tensor kernels of a layer
        self._side_stream: Optional[torch.cuda.Stream] = None
        self._fold_cache: Dict[str, Tuple[tuple, torch.Tensor, torch.Tensor]] = {}
        self.launches = 0  # kernels launched by the last forward/backward (for bench accounting)
        self.names: List[str] = []
        # optional debug trace: when a list, every kernel step appends a record {"op": ..., tensors}
        # holding REFERENCES to the buffers it read and wrote (nothing is copied or recomputed, the
        # schedule is unchanged).  The teacher-forced parity tests walk it: each step's own inputs are
        # fed to the fp32 oracle op and compared with the step's own outputs.
        self.trace: Optional[list] = None

    def _tr(self, op: str, **rec) -> None:
        if self.trace is not None:
            rec["op"] = op
            self.trace.append(rec)

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

    def _timed(self, tag: str, s: ConvSpec, n: int, hw: int, fn) -> None:
        """Run fn(); when profiling, bracket it with CUDA events on the launching stream and
        record the layer's algorithmic FLOPs (2 * pixels * Cout * 9 * Cin, real channels)."""
        if self.conv_events is None:
            fn()
            return
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.conv_events.append((tag, s.idx, 2.0 * n * hw * s.cout * 9 * s.cin, e0, e1))

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

    # ------------------------------------------------------------------------------ forward
    def _conv_bn_relu(self, fw: _Fwd, s: ConvSpec, cin_pad: int, xin: torch.Tensor,
                      out_view: Optional[torch.Tensor], pool_to: Optional[torch.Tensor],
                      layers: Optional[List[LayerSaved]], pool_idx: Optional[List[torch.Tensor]],
                      defer_apply: bool = False):
        """One layer on xin.  Activation goes to out_view (or a fresh tensor); if pool_to is
        given the 2x2 max-pool of the activation is written there too.  With defer_apply
        (training only) the normalise+ReLU pass is left to the consumer kernel and the raw
        conv output plus its (scale, shift) are returned instead."""
        n, dev = fw.n, fw.dev
        params, buffers = fw.params, fw.buffers
        hh, ww = fw.sizes[s.level]
        wt = params[f"{s.conv}.weight"]
        wp = self.packed.fprop(s.conv, wt, cin_pad)
        gamma, beta = params[f"{s.bn}.weight"], params[f"{s.bn}.bias"]
        bias = params[f"{s.conv}.bias"]
        scale = torch.empty(s.cout, **fw.f32)
        shift = torch.empty(s.cout, **fw.f32)
        a = out_view if out_view is not None else torch.empty((n, hh, ww, s.cout), **fw.bf)
        if fw.training or fw.frozen:
            y = torch.empty((n, hh, ww, s.cout), **fw.bf)
            mean = torch.empty(s.cout, **fw.f32)
            invstd = torch.empty(s.cout, **fw.f32)
            if fw.training:
                _RawWriteEpochs.bn += 1    # running statistics are about to be rewritten by a kernel
                parts = torch.empty((fw.stat_rows, 2, s.cout), **fw.f32)
                self._timed("fprop", s, n, hh * ww,
                            lambda: ops.conv3x3_fprop(xin, wp, y, stat_partials=parts))
                ops.bn_stats_finalize(parts, n * hh * ww, gamma, beta, bias, BN_EPS, BN_MOMENTUM,
                                      buffers[f"{s.bn}.running_mean"], buffers[f"{s.bn}.running_var"],
                                      scale, shift, mean, invstd)
                fw.nbt.append(buffers[f"{s.bn}.num_batches_tracked"])
            else:
                # frozen: raw conv output, coefficients folded from the running statistics (conv bias included in
                # the shift); nothing is written to the BatchNorm buffers
                rm, rv = buffers[f"{s.bn}.running_mean"], buffers[f"{s.bn}.running_var"]
                self._timed("fprop", s, n, hh * ww, lambda: ops.conv3x3_fprop(xin, wp, y))
                ops.bn_fold_eval(gamma, beta, bias, rm, rv, BN_EPS, scale, shift)
                ops.bn_eval_stats(bias, rm, rv, BN_EPS, mean, invstd)
            if defer_apply:
                a = None
            elif pool_to is not None:
                idx = torch.empty(pool_to.shape, dtype=torch.uint8, device=dev)
                ops.bn_apply_relu_maxpool2(y, a, pool_to, idx, scale, shift)
                if fw.save:
                    pool_idx.append(idx)
            else:
                ops.bn_apply_relu(y, a, scale, shift)
            fw.launches += 4  # memset+conv counted as conv(2), finalize, apply
            if fw.save:
                layers.append(LayerSaved(xin, y, scale, shift, mean, invstd))
            self._tr("conv_bn_relu", spec=s, x=xin, y=y, scale=scale, shift=shift, mean=mean, invstd=invstd,
                     a=a, pooled=pool_to, pool_idx=(idx if pool_to is not None and not defer_apply else None))
            if defer_apply:
                return y, scale, shift
        else:
            # folded eval-mode coefficients are cached per layer until a parameter / running statistic
            # can have changed: tensor version counters, plus the two process-wide epochs that cover
            # raw-pointer writers (fused Adam / graph replay / broadcast -> param, training-mode
            # BatchNorm of ANY engine -> bn)
            rm, rv = buffers[f"{s.bn}.running_mean"], buffers[f"{s.bn}.running_var"]
            key = tuple((t._version, t.data_ptr()) for t in (gamma, beta, bias, rm, rv)) + (
                _RawWriteEpochs.param, _RawWriteEpochs.bn)
            hit = self._fold_cache.get(s.bn)
            if hit is not None and hit[0] == key:
                scale, shift = hit[1], hit[2]
            else:
                ops.bn_fold_eval(gamma, beta, bias, rm, rv, BN_EPS, scale, shift)
                self._fold_cache[s.bn] = (key, scale, shift)
                fw.launches += 1
            self._timed("fprop", s, n, hh * ww,
                        lambda: ops.conv3x3_fprop(xin, wp, a, scale=scale, shift=shift, relu=True))
            fw.launches += 1
            if pool_to is not None:
                idx = torch.empty(pool_to.shape, dtype=torch.uint8, device=dev)
                ops.bn_apply_relu_maxpool2(a, None, pool_to, idx, None, None)
                fw.launches += 1
        return a

    def _run_encoder(self, fw: _Fwd, specs: Sequence[ConvSpec], cin_pad: int, x: torch.Tensor,
                     skip_views: Dict[int, torch.Tensor], x5_view: Optional[torch.Tensor],
                     layers: Optional[List[LayerSaved]], pool_idx: Optional[List[torch.Tensor]]
                     ) -> torch.Tensor:
        """inc + down1..down4 (unet.py:101-105).  The skip activation of level l lands in
        skip_views[l] (a channel slice of a concat / fusion buffer), the bottleneck x5 in x5_view
        (or a fresh tensor).  Returns x5."""
        n = fw.n
        cur = x
        li = 0
        for lvl in range(4):
            c = ENC_CH[lvl]
            cur = self._conv_bn_relu(fw, specs[li], cin_pad if li == 0 else specs[li].cin, cur, None,
                                     None, layers, pool_idx); li += 1
            hp, wp_ = fw.sizes[lvl + 1]
            pooled = torch.empty((n, hp, wp_, c), **fw.bf)
            self._conv_bn_relu(fw, specs[li], specs[li].cin, cur, skip_views[lvl], pooled, layers,
                               pool_idx); li += 1
            cur = pooled
        cur = self._conv_bn_relu(fw, specs[li], specs[li].cin, cur, None, None, layers, pool_idx); li += 1
        cur = self._conv_bn_relu(fw, specs[li], specs[li].cin, cur, x5_view, None, layers, pool_idx)
        fw.flush_nbt()
        return cur

    def _alloc_cat(self, fw: _Fwd) -> Dict[int, torch.Tensor]:
        """Concat buffers of the four Up stages: [skip | upsampled], at levels 3,2,1,0."""
        cat: Dict[int, torch.Tensor] = {}
        for lvl, c in ((3, 512), (2, 256), (1, 128), (0, 64)):
            hh, ww = fw.sizes[lvl]
            cat[lvl] = torch.empty((fw.n, hh, ww, 2 * c), **fw.bf)
        return cat

    def _run_decoder(self, fw: _Fwd, specs: Sequence[ConvSpec], cat: Dict[int, torch.Tensor],
                     x5: torch.Tensor, layers: Optional[List[LayerSaved]], head_prefix: str,
                     n_classes: int, head: bool = True):
        """up1..up4 + outc (unet.py:106-111).  cat[l][..., :c] already holds the skip features.
        Returns (logits fp32 NCHW, head input) -- or, with head=False, the final 64-channel
        activation (UNetDecoder.get_output_feats, unet.py:185-191)."""
        cur = x5
        li = 0
        head_scale = head_shift = None
        for lvl in (3, 2, 1, 0):
            c = ENC_CH[lvl]
            ops.upsample2x_pad_concat_fwd(cur, cat[lvl][..., c:])
            fw.launches += 1
            self._tr("upsample_concat", level=lvl, x=cur, cat=cat[lvl], c=c)
            cur = self._conv_bn_relu(fw, specs[li], specs[li].cin, cat[lvl], None, None, layers, None); li += 1
            if lvl == 0 and (fw.training or fw.frozen) and head:
                # last layer: its normalise+ReLU is fused into the head kernel (forward) and its
                # BatchNorm-backward reduction into the head backward -- no activation is stored
                cur, head_scale, head_shift = self._conv_bn_relu(fw, specs[li], specs[li].cin, cur, None,
                                                                 None, layers, None, defer_apply=True)
            else:
                cur = self._conv_bn_relu(fw, specs[li], specs[li].cin, cur, None, None, layers, None)
            li += 1
        fw.flush_nbt()
        if not head:
            return cur, None
        h, w = fw.sizes[0]
        logits = torch.empty((fw.n, n_classes, h, w), **fw.f32)
        wh = fw.params[f"{head_prefix}outc.conv.weight"].detach().reshape(n_classes, 64)
        ops.head1x1_fwd(cur, wh, fw.params[f"{head_prefix}outc.conv.bias"].detach(), logits,
                        head_scale, head_shift)
        fw.launches += 1
        self._tr("head", x=cur, scale=head_scale, shift=head_shift, logits=logits, prefix=head_prefix)
        return logits, cur

    # ----------------------------------------------------------------------------- backward
    def _begin_backward(self, n, sizes, dev, params, wgrad_shapes, frozen: bool = False) -> _Bwd:
        layout, total = self.grad_layout(params)
        bw = _Bwd(n, sizes, dev, params, layout, total, frozen)
        if self.overlap_wgrad and self.conv_events is None:
            if self._side_stream is None or self._side_stream.device != dev:
                self._side_stream = torch.cuda.Stream(device=dev)
            bw.side = self._side_stream
        # one split-K workspace for all layers (the wgrads are serialised on one stream)
        bw.ws = torch.empty(max(wgrad_shapes) // 4, **bw.f32)
        return bw

    def _wgrad_ws_bytes(self, n, sizes, specs: Sequence[ConvSpec], cin_pad0: int) -> List[int]:
        return [ops.wgrad_workspace_bytes(n, sizes[sp.level][0], sizes[sp.level][1],
                                          cin_pad0 if j == 0 and cin_pad0 else sp.cin, sp.cout)
                for j, sp in enumerate(specs)]

    def _mark_ready(self, bw: _Bwd, name_last: str) -> None:
        """All grads from slab[ready_upto] through `name_last` are final (enqueued)."""
        o, nel = bw.layout[name_last]
        end = (o + nel + 3) // 4 * 4
        if self.grad_ready_hook is not None and end > bw.ready_upto:
            # weight grads come from the side stream when wgrads overlap: the hook gets that stream and orders its
            # collective after BOTH producers (the main stream is not made to wait for wgrad here)
            self.grad_ready_hook(bw.slab, bw.ready_upto, end, bw.side)
        bw.ready_upto = max(bw.ready_upto, end)

    def _finish_backward(self, bw: _Bwd) -> None:
        if bw.side is not None:
            bw.main.wait_stream(bw.side)
            bw.ws.record_stream(bw.side)
        self._mark_ready(bw, self.names[0])
        if self.grad_done_hook is not None:
            self.grad_done_hook(bw.slab, bw.total)
        self.launches = bw.launches

    def _layer_backward(self, bw: _Bwd, s: ConvSpec, sv: LayerSaved, cin_pad: int, da: torch.Tensor,
                        need_dx: bool, dx_out: Optional[torch.Tensor] = None,
                        bn_parts: Optional[torch.Tensor] = None,
                        prev: Optional[Tuple[ConvSpec, LayerSaved]] = None) -> Optional[torch.Tensor]:
        """bn_parts: BatchNorm-backward partial sums already produced by the kernel that
        wrote `da` (then the separate reduction pass is skipped).  prev: the layer feeding this
        one inside the same DoubleConv (its BatchNorm-backward reduction may ride in this
        layer's dgrad epilogue)."""
        n = bw.n
        grads, params = bw.grads, bw.params
        hh, ww = bw.sizes[s.level]
        if bn_parts is None:
            bn_parts = bw.fused_parts.pop(s.conv, None)
        if bn_parts is not None:
            parts = bn_parts
        else:
            parts = torch.empty((bw.bn_rows, 2, s.cout), **bw.f32)
            ops.bn_relu_bwd_reduce(da, sv.y, sv.scale, sv.shift, sv.mean, sv.invstd, parts)
        coef = torch.empty((2, s.cout), **bw.f32)
        if bw.frozen:
            # eval-mode BatchNorm: dy = scale * g, and the conv bias gets a real gradient (sum dy)
            ops.bn_bwd_finalize_frozen(parts, sv.scale, grads[f"{s.bn}.weight"], grads[f"{s.bn}.bias"],
                                       grads[f"{s.conv}.bias"], coef)
        else:
            ops.bn_bwd_finalize(parts, n * hh * ww, sv.scale, sv.mean, sv.invstd,
                                grads[f"{s.bn}.weight"], grads[f"{s.bn}.bias"], coef)
        dy = torch.empty((n, hh, ww, s.cout), **bw.bf)
        ops.bn_relu_bwd_apply(da, sv.y, dy, sv.scale, sv.shift, coef)
        dy_ready = None
        if bw.side is not None:
            dy_ready = torch.cuda.Event()
            dy_ready.record(bw.main)

        def launch_wgrad():
            if bw.side is not None:
                bw.side.wait_event(dy_ready)                  # dy is ready
                with torch.cuda.stream(bw.side):
                    ops.conv3x3_wgrad(sv.x, dy, grads[f"{s.conv}.weight"], bw.ws, s.cin)
                dy.record_stream(bw.side)                  # keep dy alive until the side stream is done
            else:
                self._timed("wgrad", s, n, hh * ww,
                            lambda: ops.conv3x3_wgrad(sv.x, dy, grads[f"{s.conv}.weight"], bw.ws, s.cin))

        # launch ORDER matters when wgrads overlap: two tensor kernels cannot co-reside (one CTA per SM), the
        # block scheduler serves the kernel that was launched first.  dgrad is on the critical path, so it goes
        # first; the wgrad's CTAs then run when the main stream is in its next HBM-bound phase (BatchNorm-backward,
        # pool / upsample backward), whose blocks DO fit beside a wgrad CTA.  (Measured neutral within noise in the
        # power-capped step, profiles/r02_wgrad_overlap_ab.md: 881.1 vs 882.4 chips/s; kept for the rationale.)
        wgrad_first = bw.side is None or not self.wgrad_after_dgrad
        if wgrad_first:
            launch_wgrad()
        bw.launches += 5
        dx = None
        if need_dx:
            dx = dx_out if dx_out is not None else torch.empty((n, hh, ww, s.cin), **bw.bf)
            wd = self.packed.dgrad(s.conv, params[f"{s.conv}.weight"])
            if s.second and prev is not None and dx_out is None and (s.cin == 64 or s.cout >= 256):
                # second conv of a DoubleConv: dx IS the activation gradient of the first conv, so
                # that layer's BatchNorm-backward reduction rides in this dgrad's epilogue.
                # (Measured: pays off when the epilogue has slack -- 8 epilogue warps for
                # 64-channel outputs, or >= 2304-deep GEMMs; for the 128-channel layers with
                # short K the epilogue becomes critical and the separate pass is cheaper.  Re-measured
                # with the 8-epilogue-warp short-K variant: in isolation fusing wins for every second
                # conv (128->64 @256^2: 0.69 vs 0.45 + 0.42 ms), inside the power-capped step it does
                # not (dgrad family +1.0 ms for 1.06 ms of removed passes, 855 vs 857 chips/s at a
                # higher clock), so the rule stays.)
                ps, pv = prev
                fparts = torch.empty((bw.stat_rows, 2, s.cin), **bw.f32)
                self._timed("dgrad", s, n, hh * ww,
                            lambda: ops.conv3x3_dgrad(dy, wd, dx, bn_y=pv.y,
                                                      bn=(pv.scale, pv.shift, pv.mean, pv.invstd),
                                                      bn_partials=fparts))
                bw.fused_parts[ps.conv] = fparts
            else:
                self._timed("dgrad", s, n, hh * ww, lambda: ops.conv3x3_dgrad(dy, wd, dx))
            bw.launches += 1
        if not wgrad_first:
            launch_wgrad()
        self._tr("layer_bwd", spec=s, da=da, y=sv.y, x=sv.x, dy=dy, dx=dx, coef=coef,
                 dw=grads[f"{s.conv}.weight"], dgamma=grads[f"{s.bn}.weight"], dbeta=grads[f"{s.bn}.bias"],
                 fused_reduce=bn_parts is not None)
        self._mark_ready(bw, f"{s.conv}.weight")
        return dx

    def _decoder_backward(self, bw: _Bwd, specs: Sequence[ConvSpec], layers: Sequence[LayerSaved],
                          head_in: torch.Tensor, dlogits: torch.Tensor, head_prefix: str,
                          n_classes: int) -> Tuple[Dict[int, torch.Tensor], torch.Tensor]:
        """Head + up4..up1 in reverse.  Returns (dcat: level -> [N,H,W,2C] gradient of the concat
        buffer, d_x5: gradient w.r.t. the bottleneck feature)."""
        n, sizes = bw.n, bw.sizes
        grads, params = bw.grads, bw.params
        h0, w0 = sizes[0]
        d_cur = torch.empty((n, h0, w0, 64), **bw.bf)
        parts = torch.empty((ops.head_bwd_rows(), n_classes * 65), **bw.f32)
        wh = params[f"{head_prefix}outc.conv.weight"].detach().reshape(n_classes, 64)
        last = layers[-1]
        head_bn_parts = torch.empty((ops.head_bwd_rows(), 2, 64), **bw.f32)
        ops.head1x1_bwd(dlogits.contiguous(), head_in, wh, d_cur,
                        grads[f"{head_prefix}outc.conv.weight"].view(n_classes, 64),
                        grads[f"{head_prefix}outc.conv.bias"],
                        parts, bn=(last.scale, last.shift, last.mean, last.invstd),
                        bn_partials=head_bn_parts)
        bw.launches += 2
        self._tr("head_bwd", dlogits=dlogits, x=head_in, bn=last, d_act=d_cur,
                 dw=grads[f"{head_prefix}outc.conv.weight"], db=grads[f"{head_prefix}outc.conv.bias"],
                 prefix=head_prefix)
        self._mark_ready(bw, f"{head_prefix}outc.conv.weight")

        li = len(specs) - 1
        dcat: Dict[int, torch.Tensor] = {}
        for lvl in (0, 1, 2, 3):
            c = ENC_CH[lvl]
            hh, ww = sizes[lvl]
            d_mid = self._layer_backward(bw, specs[li], layers[li], specs[li].cin, d_cur, True,
                                         bn_parts=head_bn_parts if lvl == 0 else None,
                                         prev=(specs[li - 1], layers[li - 1])); li -= 1
            dcat[lvl] = torch.empty((n, hh, ww, 2 * c), **bw.bf)
            self._layer_backward(bw, specs[li], layers[li], specs[li].cin, d_mid, True, dcat[lvl]); li -= 1
            hl, wl = sizes[lvl + 1]
            d_cur = torch.empty((n, hl, wl, c), **bw.bf)
            ops.upsample2x_pad_concat_bwd(dcat[lvl][..., c:], d_cur)
            bw.launches += 1
            self._tr("upsample_concat_bwd", level=lvl, dcat=dcat[lvl], c=c, dx=d_cur)
        assert li == -1
        return dcat, d_cur

    def _encoder_backward(self, bw: _Bwd, specs: Sequence[ConvSpec], layers: Sequence[LayerSaved],
                          pool_idx: Sequence[torch.Tensor], cin_pad: int, d_x5: torch.Tensor,
                          d_skip_of: Callable[[int], torch.Tensor], release=None) -> None:
        """down4..inc in reverse.  d_skip_of(l): gradient arriving at the level-l skip feature
        from outside the encoder (the skip half of dcat for the UNet, a slice of the fusion
        gradient for late fusion); release(l) is called once it has been consumed."""
        n, sizes = bw.n, bw.sizes
        li = len(specs) - 1
        d_mid = self._layer_backward(bw, specs[li], layers[li], specs[li].cin, d_x5, True,
                                     prev=(specs[li - 1], layers[li - 1])); li -= 1   # down4 second conv
        d_pool = self._layer_backward(bw, specs[li], layers[li], specs[li].cin, d_mid, True); li -= 1
        for lvl in (3, 2, 1, 0):
            c = ENC_CH[lvl]
            hh, ww = sizes[lvl]
            d_skip = torch.empty((n, hh, ww, c), **bw.bf)
            # the skip layer's BatchNorm-backward reduction rides in the pool-backward pass
            sk = layers[li]
            pool_parts = torch.empty((2 * bw.bn_rows, 2, c), **bw.f32)
            d_from_skip = d_skip_of(lvl)
            ops.maxpool2_bwd(d_pool, pool_idx[lvl], d_from_skip, d_skip, bn_y=sk.y,
                             bn=(sk.scale, sk.shift, sk.mean, sk.invstd), bn_partials=pool_parts)
            bw.launches += 1
            self._tr("maxpool_bwd", level=lvl, d_pooled=d_pool, pool_idx=pool_idx[lvl], d_skip_in=d_from_skip,
                     d_act=d_skip)
            if release is not None:
                release(lvl)
            d_mid = self._layer_backward(bw, specs[li], layers[li], specs[li].cin, d_skip, True,
                                         bn_parts=pool_parts,
                                         prev=(specs[li - 1], layers[li - 1])); li -= 1
            d_pool = self._layer_backward(bw, specs[li], layers[li], cin_pad if li == 0 else specs[li].cin,
                                          d_mid, lvl > 0); li -= 1
        assert li == -1


def _number(specs: Sequence[ConvSpec], start: int = 0) -> None:
    for i, s in enumerate(specs):
        s.idx = start + i


class UNetEngine(_Schedule):
    """The whole UNet (unet.py:80-111) as one schedule."""

    def __init__(self, n_channels: int, n_classes: int):
        super().__init__()
        if n_classes < 1 or n_classes > 8:
            raise RuntimeError(f"floodplanet_b200: n_classes={n_classes} unsupported (1..8)")
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.cin_pad = pad_channels(n_channels)
        if self.cin_pad > MAX_INPUT_CHANNELS:
            raise RuntimeError(f"floodplanet_b200: {n_channels} input channels unsupported (max {MAX_INPUT_CHANNELS})")
        self.enc_specs = encoder_conv_specs(n_channels)
        self.dec_specs = decoder_conv_specs()
        self.specs = self.enc_specs + self.dec_specs
        _number(self.specs)
        self.names = param_names(n_channels)

    def _cin_pad_of(self, i: int) -> int:
        return self.cin_pad if i == 0 else self.specs[i].cin

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
        fw = _Fwd(n, sizes, dev, params, buffers, training, save)
        st = ForwardState(n=n, sizes=sizes, frozen=fw.frozen) if save else None
        layers = st.layers if save else None
        pool_idx = st.pool_idx if save else None

        x = ingested if ingested is not None else ops.ingest(images, self.cin_pad)
        fw.launches += 1
        cat = self._alloc_cat(fw)
        x5 = self._run_encoder(fw, self.enc_specs, self.cin_pad, x,
                               {lvl: cat[lvl][..., :ENC_CH[lvl]] for lvl in range(4)}, None,
                               layers, pool_idx)
        logits, head_in = self._run_decoder(fw, self.dec_specs, cat, x5, layers, "", self.n_classes)
        if save:
            st.head_in = head_in   # raw output y of the last conv (training)
        self.launches = fw.launches
        return logits, st

    def backward(self, st: ForwardState, dlogits: torch.Tensor, params: Dict[str, torch.Tensor]):
        """Returns {param name: fp32 gradient view} (views of one flat slab)."""
        n, sizes = st.n, st.sizes
        bw = self._begin_backward(n, sizes, dlogits.device, params,
                                  self._wgrad_ws_bytes(n, sizes, self.specs, self.cin_pad), st.frozen)
        ne = len(self.enc_specs)
        dcat, d_x5 = self._decoder_backward(bw, self.dec_specs, st.layers[ne:], st.head_in, dlogits, "",
                                            self.n_classes)
        self._encoder_backward(bw, self.enc_specs, st.layers[:ne], st.pool_idx, self.cin_pad, d_x5,
                               lambda lvl: dcat[lvl][..., :ENC_CH[lvl]],
                               release=lambda lvl: dcat.pop(lvl, None))
        self._finish_backward(bw)
        return bw.grads, bw.slab


def _check_images(images: Sequence[torch.Tensor], n_channels: int) -> Tuple[torch.device, int, int, int]:
    dev = images[0].device
    n, _, h, w = images[0].shape
    if sum(int(t.shape[1]) for t in images) != n_channels:
        raise RuntimeError(f"floodplanet_b200: expected {n_channels} input channels, got "
                           f"{[int(t.shape[1]) for t in images]}")
    return dev, n, h, w


class EncoderEngine(_Schedule):
    """inc + down1..4 alone (UNet.encode, unet.py:113-120; UNetEncoder.forward :150-159): the
    five features leave as NHWC bf16 tensors; the module layer converts them to the fp32 NCHW
    list the reference API returns."""

    def __init__(self, n_channels: int, prefix: str = ""):
        super().__init__()
        self.n_channels = n_channels
        self.cin_pad = pad_channels(n_channels)
        if self.cin_pad > MAX_INPUT_CHANNELS:
            raise RuntimeError(f"floodplanet_b200: {n_channels} input channels unsupported (max {MAX_INPUT_CHANNELS})")
        self.specs = encoder_conv_specs(n_channels, prefix)
        _number(self.specs)
        self.names = conv_param_names(self.specs)

    def forward(self, images, params, buffers, training: bool, save: bool):
        dev, n, h, w = _check_images(images, self.n_channels)
        sizes = self._level_sizes(h, w)
        fw = _Fwd(n, sizes, dev, params, buffers, training, save)
        st = ForwardState(n=n, sizes=sizes, frozen=fw.frozen) if save else None
        x = ops.ingest(images, self.cin_pad)
        fw.launches += 1
        feats = [torch.empty((n, sizes[l][0], sizes[l][1], FEAT_CH[l]), **fw.bf) for l in range(5)]
        self._run_encoder(fw, self.specs, self.cin_pad, x, {l: feats[l] for l in range(4)}, feats[4],
                          st.layers if save else None, st.pool_idx if save else None)
        self.launches = fw.launches
        return feats, st

    def backward(self, st: ForwardState, d_feats: Sequence[torch.Tensor], params):
        """d_feats: NHWC bf16 gradients of the five features."""
        n, sizes = st.n, st.sizes
        bw = self._begin_backward(n, sizes, d_feats[0].device, params,
                                  self._wgrad_ws_bytes(n, sizes, self.specs, self.cin_pad), st.frozen)
        self._encoder_backward(bw, self.specs, st.layers, st.pool_idx, self.cin_pad, d_feats[4],
                               lambda lvl: d_feats[lvl])
        self._finish_backward(bw)
        return bw.grads, bw.slab


class DecoderEngine(_Schedule):
    """up1..4 + outc alone (UNet.decode, unet.py:122-131; UNetDecoder.forward :176-183)."""

    def __init__(self, n_classes: int, prefix: str = ""):
        super().__init__()
        if n_classes < 1 or n_classes > 8:
            raise RuntimeError(f"floodplanet_b200: n_classes={n_classes} unsupported (1..8)")
        self.n_classes = n_classes
        self.prefix = prefix
        self.specs = decoder_conv_specs(prefix)
        _number(self.specs)
        self.names = conv_param_names(self.specs) + [f"{prefix}outc.conv.weight", f"{prefix}outc.conv.bias"]

    def forward(self, feats_nchw: Sequence[torch.Tensor], params, buffers, training: bool, save: bool,
                head: bool = True):
        """feats_nchw: the five fp32 NCHW features [x1..x5]; the skips are converted straight
        into the skip halves of the concat buffers."""
        if len(feats_nchw) != 5:
            raise RuntimeError(f"floodplanet_b200: decode expects 5 feature maps, got {len(feats_nchw)}")
        dev = feats_nchw[0].device
        n, _, h, w = feats_nchw[0].shape
        sizes = self._level_sizes(h, w)
        for l, f in enumerate(feats_nchw):
            if tuple(f.shape) != (n, FEAT_CH[l], sizes[l][0], sizes[l][1]):
                raise RuntimeError(f"floodplanet_b200: feature {l} has shape {tuple(f.shape)}, expected "
                                   f"{(n, FEAT_CH[l], sizes[l][0], sizes[l][1])}")
        fw = _Fwd(n, sizes, dev, params, buffers, training, save)
        st = ForwardState(n=n, sizes=sizes, frozen=fw.frozen) if save else None
        cat = self._alloc_cat(fw)
        for l in range(4):
            ops.nchw_f32_to_nhwc_bf16(feats_nchw[l], cat[l][..., :ENC_CH[l]])
        x5 = torch.empty((n, sizes[4][0], sizes[4][1], FEAT_CH[4]), **fw.bf)
        ops.nchw_f32_to_nhwc_bf16(feats_nchw[4], x5)
        fw.launches += 5
        out, head_in = self._run_decoder(fw, self.specs, cat, x5, st.layers if save else None,
                                         self.prefix, self.n_classes, head=head)
        if save:
            st.head_in = head_in
        self.launches = fw.launches
        return out, st

    def backward(self, st: ForwardState, dlogits: torch.Tensor, params):
        """Returns (grads, slab, [d_x1..d_x5] as NHWC bf16 views)."""
        n, sizes = st.n, st.sizes
        bw = self._begin_backward(n, sizes, dlogits.device, params,
                                  self._wgrad_ws_bytes(n, sizes, self.specs, 0), st.frozen)
        dcat, d_x5 = self._decoder_backward(bw, self.specs, st.layers, st.head_in, dlogits, self.prefix,
                                            self.n_classes)
        self._finish_backward(bw)
        return bw.grads, bw.slab, [dcat[l][..., :ENC_CH[l]] for l in range(4)] + [d_x5]


class LateFusionEngine(_Schedule):
    """Late fusion (lf_model.py:29-92): one encoder per modality, the five feature levels of all
    modalities concatenated along C, a 1x1 `concat_conv` per level back to the UNet widths, then
    the ordinary decoder.  The concatenation is virtual: encoder m writes its level-l feature
    into channels [m*fs, (m+1)*fs) of ONE fusion buffer per level, the pointwise convolution
    (tensor cores) reads that buffer and writes the skip half of the decoder's concat buffer."""

    def __init__(self, in_channels: "Dict[str, int]", n_classes: int):
        super().__init__()
        if n_classes < 1 or n_classes > 8:
            raise RuntimeError(f"floodplanet_b200: n_classes={n_classes} unsupported (1..8)")
        self.n_classes = n_classes
        self.modalities = list(in_channels.keys())          # nn.ModuleDict order
        self.in_channels = dict(in_channels)
        self.enc_specs: Dict[str, List[ConvSpec]] = {}
        self.cin_pad: Dict[str, int] = {}
        self.names = []
        k = 0
        for name, c in in_channels.items():
            self.cin_pad[name] = pad_channels(c)
            if self.cin_pad[name] > MAX_INPUT_CHANNELS:
                raise RuntimeError(f"floodplanet_b200: {c} input channels unsupported (max {MAX_INPUT_CHANNELS})")
            self.enc_specs[name] = encoder_conv_specs(c, f"encoders.{name}.")
            _number(self.enc_specs[name], k)
            k += len(self.enc_specs[name])
            self.names += conv_param_names(self.enc_specs[name])
        for l in range(5):
            self.names += [f"concat_convs.{l}.weight", f"concat_convs.{l}.bias"]
        self.dec_specs = decoder_conv_specs("decoder.")
        _number(self.dec_specs, k)
        self.names += conv_param_names(self.dec_specs) + ["decoder.outc.conv.weight", "decoder.outc.conv.bias"]
        self._ones: Dict[Tuple[int, torch.device], torch.Tensor] = {}

    def _one(self, c: int, dev) -> torch.Tensor:
        key = (c, dev)
        if key not in self._ones:
            self._ones[key] = torch.ones(c, dtype=torch.float32, device=dev)
        return self._ones[key]

    def forward(self, images: "Dict[str, torch.Tensor]", params, buffers, training: bool, save: bool):
        """images: modality name -> NCHW fp32 tensor, in concatenation order (lf_model.py:60-81:
        ms_image first, then dem, slope, preflood, pre_post_difference, hand)."""
        order = list(images.keys())
        k = len(order)
        first = images[order[0]]
        dev = first.device
        n, _, h, w = first.shape
        for name in order:
            if name not in self.enc_specs:
                raise KeyError(name)
            _check_images([images[name]], self.in_channels[name])
        for l in range(5):
            wshape = params[f"concat_convs.{l}.weight"].shape
            if wshape[1] != FEAT_CH[l] * k:
                raise RuntimeError(
                    f"floodplanet_b200: concat_convs.{l} expects {wshape[1]} input channels but the "
                    f"batch provides {k} modalities x {FEAT_CH[l]} (lf_model.py:44-45 sizes the "
                    "fusion convs by len(in_channels))")
        sizes = self._level_sizes(h, w)
        fw = _Fwd(n, sizes, dev, params, buffers, training, save)
        st = ForwardState(n=n, sizes=sizes, frozen=fw.frozen) if save else None
        fused_in = [torch.empty((n, sizes[l][0], sizes[l][1], FEAT_CH[l] * k), **fw.bf) for l in range(5)]
        for m, name in enumerate(order):
            x = ops.ingest([images[name]], self.cin_pad[name])
            fw.launches += 1
            layers: Optional[List[LayerSaved]] = [] if save else None
            pidx: Optional[List[torch.Tensor]] = [] if save else None
            sl = lambda l: fused_in[l][..., m * FEAT_CH[l]:(m + 1) * FEAT_CH[l]]
            self._run_encoder(fw, self.enc_specs[name], self.cin_pad[name], x,
                              {l: sl(l) for l in range(4)}, sl(4), layers, pidx)
            if save:
                st.enc_layers.append(layers)
                st.enc_pool_idx.append(pidx)
        cat = self._alloc_cat(fw)
        x5 = torch.empty((n, sizes[4][0], sizes[4][1], FEAT_CH[4]), **fw.bf)
        for l in range(5):
            wname = f"concat_convs.{l}"
            wp = self.packed.fprop_1x1(wname, params[f"{wname}.weight"])
            out = x5 if l == 4 else cat[l][..., :ENC_CH[l]]
            ops.conv1x1(fused_in[l], wp, out, self._one(FEAT_CH[l], dev), params[f"{wname}.bias"].detach())
            fw.launches += 1
        logits, head_in = self._run_decoder(fw, self.dec_specs, cat, x5, st.layers if save else None,
                                            "decoder.", self.n_classes)
        if save:
            st.head_in = head_in
            st.fused_in = fused_in
            st.order = order
        self.launches = fw.launches
        return logits, st

    def backward(self, st: ForwardState, dlogits: torch.Tensor, params):
        n, sizes = st.n, st.sizes
        order = st.order
        k = len(order)
        ws_bytes = self._wgrad_ws_bytes(n, sizes, self.dec_specs, 0)
        for name in order:
            ws_bytes += self._wgrad_ws_bytes(n, sizes, self.enc_specs[name], self.cin_pad[name])
        ws_bytes += [ops.conv1x1_wgrad_workspace_bytes(n, sizes[l][0], sizes[l][1], FEAT_CH[l] * k, FEAT_CH[l])
                     for l in range(5)]
        bw = self._begin_backward(n, sizes, dlogits.device, params, ws_bytes, st.frozen)
        dcat, d_x5 = self._decoder_backward(bw, self.dec_specs, st.layers, st.head_in, dlogits, "decoder.",
                                            self.n_classes)
        # fusion convs, levels 4..0 (reverse forward order = slab order)
        d_fused: Dict[int, torch.Tensor] = {}
        for l in (4, 3, 2, 1, 0):
            wname = f"concat_convs.{l}"
            dfeat = d_x5 if l == 4 else dcat[l][..., :ENC_CH[l]]
            ops.conv1x1_wgrad(st.fused_in[l], dfeat, bw.grads[f"{wname}.weight"], bw.ws)
            ops.channel_sum(dfeat, bw.grads[f"{wname}.bias"])
            d_fused[l] = torch.empty((n, sizes[l][0], sizes[l][1], FEAT_CH[l] * k), **bw.bf)
            ops.conv1x1(dfeat, self.packed.dgrad_1x1(wname, params[f"{wname}.weight"]), d_fused[l])
            bw.launches += 5
            dcat.pop(l, None)
            self._mark_ready(bw, f"{wname}.weight")
        del dfeat
        # encoders in reverse parameter order (the forward check guarantees every encoder ran)
        for name in reversed(self.modalities):
            m = order.index(name)
            sl = lambda l: d_fused[l][..., m * FEAT_CH[l]:(m + 1) * FEAT_CH[l]]
            self._encoder_backward(bw, self.enc_specs[name], st.enc_layers[m], st.enc_pool_idx[m],
                                   self.cin_pad[name], sl(4), sl)
        self._finish_backward(bw)
        return bw.grads, bw.slab
