"""Tensor-level wrappers over the C ABI (``capi``): they only translate torch tensors into
(pointer, pitch, dims, stream) and raise on a non-zero status.  All arithmetic happens in
the CUDA kernels of ``csrc/``; nothing here computes on the CPU or through torch ops.

NHWC bf16 activation *views* are torch tensors of shape [N, H, W, C] whose last dimension is
contiguous and whose pixel pitch ``stride(2)`` may exceed C (a channel slice of a concat
buffer).
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Optional, Sequence, Tuple

import torch

from . import capi


def _lib():
    return capi.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "floodplanet_b200: tensors must live on a CUDA device (sm_100a); there is no "
                "CPU fallback for this path")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def nhwc_view(t: torch.Tensor) -> Tuple[int, int]:
    """(device pointer, pixel pitch) of an NHWC bf16 view; validates the layout."""
    if t.dtype != torch.bfloat16 or t.dim() != 4:
        raise RuntimeError(f"expected a bf16 [N,H,W,C] view, got {t.dtype} {tuple(t.shape)}")
    n, h, w, c = t.shape
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
    ok = t.stride(3) == 1 and (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == ld * w) \
        and (n == 1 or t.stride(0) == ld * w * h)
    if not ok or ld % 8 != 0 or t.data_ptr() % 16 != 0:
        raise RuntimeError(f"not a dense-pitch NHWC view: shape {tuple(t.shape)} strides {t.stride()}")
    return t.data_ptr(), ld


def stat_rows() -> int:
    return _lib().fpb200_conv_stat_rows()


def bn_bwd_rows() -> int:
    return _lib().fpb200_bn_bwd_rows()


def head_bwd_rows() -> int:
    return _lib().fpb200_head_bwd_rows()


def ce_rows() -> int:
    return _lib().fpb200_ce_rows()


# ------------------------------------------------------------------------------------------
def ingest(srcs: Sequence[torch.Tensor], c_pad: int) -> torch.Tensor:
    """NCHW fp32 image(s) -> NHWC bf16 [N,H,W,c_pad] (early-fusion concat + cast + pad)."""
    _require_cuda(*srcs)
    n, _, h, w = srcs[0].shape
    srcs = [s.contiguous() if s.dtype == torch.float32 else s.float().contiguous() for s in srcs]
    for s in srcs:
        if s.dim() != 4 or s.shape[0] != n or s.shape[2] != h or s.shape[3] != w:
            raise RuntimeError(f"ingest: inconsistent source shapes {[tuple(x.shape) for x in srcs]}")
    out = torch.empty((n, h, w, c_pad), dtype=torch.bfloat16, device=srcs[0].device)
    k = len(srcs)
    ptrs = (C.c_void_p * k)(*[s.data_ptr() for s in srcs])
    chans = (C.c_int * k)(*[s.shape[1] for s in srcs])
    st = _lib().fpb200_ingest_nchw_f32_to_nhwc_bf16(ptrs, chans, k, out.data_ptr(), c_pad, n, h, w,
                                                    _stream())
    capi.check(st, "ingest_nchw_f32_to_nhwc_bf16", N=n, H=h, W=w, c_pad=c_pad,
               channels=[s.shape[1] for s in srcs])
    return out


def repack_fprop(w: torch.Tensor, cin_pad: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty((cout, 9, cin_pad), dtype=torch.bfloat16, device=w.device)
    st = _lib().fpb200_repack_weights_fprop(w.detach().contiguous().data_ptr(), out.data_ptr(), cout,
                                            cin, cin_pad, _stream())
    capi.check(st, "repack_weights_fprop", Cout=cout, Cin=cin, cin_pad=cin_pad)
    return out


def repack_batch_table(entries, device) -> Tuple[torch.Tensor, int]:
    """entries: [(w fp32 OIHW, packed bf16 buffer, cin_pad, kind)] -> (device table, total blocks) for
    `repack_batch`; the record layout is the one documented in include/floodplanet_b200.h."""
    import struct
    blob, first = bytearray(), 0
    for w, out, cin_pad, kind in entries:
        cout, cin = w.shape[0], w.shape[1]
        blob += struct.pack("<QQiiiiq", w.data_ptr(), out.data_ptr(), cout, cin, cin_pad, kind, first)
        first += cout if kind == 0 else cin      # one block per output (fprop) / input (dgrad) channel
        if cin > 1024:
            raise RuntimeError(f"repack_batch: Cin {cin} exceeds the 1024 channels the batch kernel stages")
    table = torch.frombuffer(blob, dtype=torch.uint8).clone().to(device)
    return table, first


def repack_batch(table: torch.Tensor, n_entries: int, total_blocks: int) -> None:
    st = _lib().fpb200_repack_weights_batch(table.data_ptr(), n_entries, total_blocks, _stream())
    capi.check(st, "repack_weights_batch", n_entries=n_entries, total_blocks=total_blocks)


def repack_dgrad(w: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty((cin, 9, cout), dtype=torch.bfloat16, device=w.device)
    st = _lib().fpb200_repack_weights_dgrad(w.detach().contiguous().data_ptr(), out.data_ptr(), cout,
                                            cin, _stream())
    capi.check(st, "repack_weights_dgrad", Cout=cout, Cin=cin)
    return out


def conv3x3_fprop(x: torch.Tensor, w_packed: torch.Tensor, y: torch.Tensor,
                  scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                  relu: bool = False, stat_partials: Optional[torch.Tensor] = None, _fn=None) -> None:
    """`_fn`: a C function with the same prototype to call instead (the test-only cross-check kernel)."""
    _require_cuda(x, w_packed, y)
    xp, ldx = nhwc_view(x)
    yp, ldy = nhwc_view(y)
    n, h, w, cin = x.shape
    cout = y.shape[3]
    if w_packed.shape != (cout, 9, cin):
        raise RuntimeError(f"conv3x3_fprop: packed weight {tuple(w_packed.shape)} != ({cout}, 9, {cin})")
    fn = _fn if _fn is not None else _lib().fpb200_conv3x3_fprop_bf16_nhwc
    st = fn(xp, ldx, w_packed.data_ptr(), yp, ldy, n, h, w, cin, cout, _ptr(scale), _ptr(shift),
            int(relu), _ptr(stat_partials), _stream())
    capi.check(st, "conv3x3_fprop_bf16_nhwc", N=n, H=h, W=w, Cin=cin, Cout=cout)


def conv3x3_dgrad(dy: torch.Tensor, w_packed_dgrad: torch.Tensor, dx: torch.Tensor,
                  bn_y: Optional[torch.Tensor] = None, bn=None,
                  bn_partials: Optional[torch.Tensor] = None) -> None:
    """bn_y / bn=(scale, shift, mean, invstd) / bn_partials: fuse the BatchNorm-backward
    reduction of the layer whose activation gradient `dx` is (see the header)."""
    _require_cuda(dy, w_packed_dgrad, dx)
    dyp, lddy = nhwc_view(dy)
    dxp, lddx = nhwc_view(dx)
    n, h, w, cout = dy.shape
    cin = dx.shape[3]
    byp, ldby = nhwc_view(bn_y) if bn_y is not None else (None, 0)
    sc, sh, mu, istd = bn if bn is not None else (None, None, None, None)
    st = _lib().fpb200_conv3x3_dgrad_bf16_nhwc(dyp, lddy, w_packed_dgrad.data_ptr(), dxp, lddx, n, h,
                                               w, cout, cin, byp, ldby, _ptr(sc), _ptr(sh), _ptr(mu),
                                               _ptr(istd), _ptr(bn_partials), _stream())
    capi.check(st, "conv3x3_dgrad_bf16_nhwc", N=n, H=h, W=w, Cout=cout, Cin=cin)


def wgrad_workspace_bytes(n: int, h: int, w: int, cin: int, cout: int) -> int:
    b = _lib().fpb200_conv3x3_wgrad_workspace_bytes(n, h, w, cin, cout)
    if b < 0:
        raise RuntimeError(f"conv3x3_wgrad: unsupported shape Cin={cin} Cout={cout}")
    return b


def conv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, workspace: torch.Tensor,
                  cin_real: int) -> None:
    """dw (fp32 [Cout, cin_real, 3, 3], contiguous) is overwritten."""
    _require_cuda(x, dy, dw, workspace)
    xp, ldx = nhwc_view(x)
    dyp, lddy = nhwc_view(dy)
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    need = wgrad_workspace_bytes(n, h, w, cin, cout)
    if workspace.numel() * workspace.element_size() < need or not dw.is_contiguous():
        raise RuntimeError("conv3x3_wgrad: workspace too small or dw not contiguous")
    st = _lib().fpb200_conv3x3_wgrad_bf16_nhwc(xp, ldx, dyp, lddy, dw.data_ptr(), workspace.data_ptr(),
                                               n, h, w, cin, cin_real, cout, _stream())
    capi.check(st, "conv3x3_wgrad_bf16_nhwc", N=n, H=h, W=w, Cin=cin, Cout=cout)


# ------------------------------------------------------------------------------------------
# pointwise (1x1) convolution and the feature-level seams (late fusion, encode/decode API)
def repack_1x1(w: torch.Tensor, transpose: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [Cout, Cin, 1, 1] -> bf16 [Cout, Cin] (forward) or [Cin, Cout] (data gradient)."""
    _require_cuda(w)
    cout, cin = w.shape[0], w.shape[1]
    if out is None:
        out = torch.empty((cin, cout) if transpose else (cout, cin), dtype=torch.bfloat16, device=w.device)
    st = _lib().fpb200_repack_weights_1x1(w.detach().contiguous().data_ptr(), out.data_ptr(), cout, cin,
                                          int(transpose), _stream())
    capi.check(st, "repack_weights_1x1", Cout=cout, Cin=cin, transpose=transpose)
    return out


def conv1x1(x: torch.Tensor, w_packed: torch.Tensor, y: torch.Tensor,
            scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
            relu: bool = False) -> None:
    """y = x . w_packed^T per pixel (w_packed bf16 [Cout_gemm, Cin_gemm]), optional affine."""
    _require_cuda(x, w_packed, y)
    xp, ldx = nhwc_view(x)
    yp, ldy = nhwc_view(y)
    n, h, w, cin = x.shape
    cout = y.shape[3]
    if tuple(w_packed.shape) != (cout, cin):
        raise RuntimeError(f"conv1x1: packed weight {tuple(w_packed.shape)} != ({cout}, {cin})")
    st = _lib().fpb200_conv1x1_bf16_nhwc(xp, ldx, w_packed.data_ptr(), yp, ldy, n, h, w, cin, cout,
                                         _ptr(scale), _ptr(shift), int(relu), _stream())
    capi.check(st, "conv1x1_bf16_nhwc", N=n, H=h, W=w, Cin=cin, Cout=cout)


def conv1x1_wgrad_workspace_bytes(n: int, h: int, w: int, cin: int, cout: int) -> int:
    b = _lib().fpb200_conv1x1_wgrad_workspace_bytes(n, h, w, cin, cout)
    if b < 0:
        raise RuntimeError(f"conv1x1_wgrad: unsupported shape Cin={cin} Cout={cout}")
    return b


def conv1x1_wgrad(x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, workspace: torch.Tensor) -> None:
    """dw (fp32 [Cout, Cin, 1, 1] or [Cout, Cin], contiguous) is overwritten."""
    _require_cuda(x, dy, dw, workspace)
    xp, ldx = nhwc_view(x)
    dyp, lddy = nhwc_view(dy)
    n, h, w, cin = x.shape
    cout = dy.shape[3]
    need = conv1x1_wgrad_workspace_bytes(n, h, w, cin, cout)
    if workspace.numel() * workspace.element_size() < need or not dw.is_contiguous() \
            or dw.numel() != cout * cin:
        raise RuntimeError("conv1x1_wgrad: workspace too small or dw not a contiguous [Cout, Cin]")
    st = _lib().fpb200_conv1x1_wgrad_bf16_nhwc(xp, ldx, dyp, lddy, dw.data_ptr(), workspace.data_ptr(),
                                               n, h, w, cin, cout, _stream())
    capi.check(st, "conv1x1_wgrad_bf16_nhwc", N=n, H=h, W=w, Cin=cin, Cout=cout)


def channel_sum(x: torch.Tensor, out: torch.Tensor) -> None:
    """out[c] (fp32) = sum over pixels of the NHWC bf16 view x."""
    _require_cuda(x, out)
    xp, ld = nhwc_view(x)
    n, h, w, c = x.shape
    rows = _lib().fpb200_channel_sum_rows()
    partials = torch.empty((rows, c), dtype=torch.float32, device=x.device)
    st = _lib().fpb200_channel_sum_bf16_nhwc(xp, ld, partials.data_ptr(), out.data_ptr(), n * h * w, c,
                                             _stream())
    capi.check(st, "channel_sum_bf16_nhwc", N=n, H=h, W=w, C=c)


def nchw_f32_to_nhwc_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    """fp32 NCHW tensor -> bf16 NHWC view `dst` ([N,H,W,C], may be a channel slice)."""
    _require_cuda(src, dst)
    src = src.contiguous() if src.dtype == torch.float32 else src.float().contiguous()
    dp, ld = nhwc_view(dst)
    n, c, h, w = src.shape
    if tuple(dst.shape) != (n, h, w, c):
        raise RuntimeError(f"nchw_f32_to_nhwc_bf16: {tuple(src.shape)} vs view {tuple(dst.shape)}")
    st = _lib().fpb200_nchw_f32_to_nhwc_bf16(src.data_ptr(), dp, ld, n, c, h, w, _stream())
    capi.check(st, "nchw_f32_to_nhwc_bf16", N=n, C=c, H=h, W=w)


def nhwc_bf16_to_nchw_f32(src: torch.Tensor) -> torch.Tensor:
    """bf16 NHWC view -> fresh fp32 NCHW tensor."""
    _require_cuda(src)
    sp, ld = nhwc_view(src)
    n, h, w, c = src.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    st = _lib().fpb200_nhwc_bf16_to_nchw_f32(sp, ld, out.data_ptr(), n, c, h, w, _stream())
    capi.check(st, "nhwc_bf16_to_nchw_f32", N=n, C=c, H=h, W=w)
    return out


# ------------------------------------------------------------------------------------------
def bn_stats_finalize(partials, count, gamma, beta, conv_bias, eps, momentum, running_mean,
                      running_var, scale, shift, save_mean, save_invstd) -> None:
    c = gamma.numel()
    st = _lib().fpb200_bn_stats_finalize(partials.data_ptr(), partials.shape[0], c, float(count),
                                         gamma.data_ptr(), beta.data_ptr(), _ptr(conv_bias), eps,
                                         momentum, _ptr(running_mean), _ptr(running_var),
                                         scale.data_ptr(), shift.data_ptr(), _ptr(save_mean),
                                         _ptr(save_invstd), _stream())
    capi.check(st, "bn_stats_finalize", C=c, count=count)


def bn_fold_eval(gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift) -> None:
    c = gamma.numel()
    st = _lib().fpb200_bn_fold_eval(gamma.data_ptr(), beta.data_ptr(), _ptr(conv_bias),
                                    running_mean.data_ptr(), running_var.data_ptr(), eps, c,
                                    scale.data_ptr(), shift.data_ptr(), _stream())
    capi.check(st, "bn_fold_eval", C=c)


def bn_eval_stats(conv_bias, running_mean, running_var, eps, mean_eff, invstd) -> None:
    c = running_mean.numel()
    st = _lib().fpb200_bn_eval_stats(_ptr(conv_bias), running_mean.data_ptr(), running_var.data_ptr(), eps, c,
                                     mean_eff.data_ptr(), invstd.data_ptr(), _stream())
    capi.check(st, "bn_eval_stats", C=c)


def bn_bwd_finalize_frozen(partials, scale, dgamma, dbeta, dbias, coef) -> None:
    c = scale.numel()
    st = _lib().fpb200_bn_bwd_finalize_frozen(partials.data_ptr(), partials.shape[0], c, scale.data_ptr(),
                                              _ptr(dgamma), _ptr(dbeta), _ptr(dbias), coef.data_ptr(), _stream())
    capi.check(st, "bn_bwd_finalize_frozen", C=c)


def bn_apply_relu(y, a, scale, shift) -> None:
    yp, ldy = nhwc_view(y)
    ap, lda = nhwc_view(a)
    n, h, w, c = y.shape
    st = _lib().fpb200_bn_apply_relu(yp, ldy, ap, lda, scale.data_ptr(), shift.data_ptr(), n * h * w, c,
                                     _stream())
    capi.check(st, "bn_apply_relu", N=n, H=h, W=w, C=c)


def bn_apply_relu_maxpool2(y, a, pooled, pool_idx, scale, shift) -> None:
    """y -> a = relu(y*scale+shift) (a may be None), pooled = maxpool2(a), pool_idx (uint8)."""
    yp, ldy = nhwc_view(y)
    ap, lda = nhwc_view(a) if a is not None else (None, 0)
    pp, ldp = nhwc_view(pooled)
    n, h, w, c = y.shape
    st = _lib().fpb200_bn_apply_relu_maxpool2(yp, ldy, ap, lda, pp, ldp, pool_idx.data_ptr(),
                                              _ptr(scale), _ptr(shift), n, h, w, c, _stream())
    capi.check(st, "bn_apply_relu_maxpool2", N=n, H=h, W=w, C=c)


def maxpool2_bwd(dpooled, pool_idx, dskip, dx, bn_y=None, bn=None, bn_partials=None) -> None:
    """bn_y / bn=(scale, shift, mean, invstd) / bn_partials ([2*bn_bwd_rows()][2][C]): also reduce the
    BatchNorm-backward sums of the layer whose activation gradient dx is."""
    dpp, lddp = nhwc_view(dpooled)
    dsp, ldds = nhwc_view(dskip) if dskip is not None else (None, 0)
    dxp, lddx = nhwc_view(dx)
    n, h, w, c = dx.shape
    byp, ldby = nhwc_view(bn_y) if bn_y is not None else (None, 0)
    sc, sh, mu, istd = bn if bn is not None else (None, None, None, None)
    st = _lib().fpb200_maxpool2_bwd(dpp, lddp, pool_idx.data_ptr(), dsp, ldds, dxp, lddx, n, h, w, c,
                                    byp, ldby, _ptr(sc), _ptr(sh), _ptr(mu), _ptr(istd),
                                    _ptr(bn_partials), _stream())
    capi.check(st, "maxpool2_bwd", N=n, H=h, W=w, C=c)


def bn_relu_bwd_reduce(da, y, scale, shift, save_mean, save_invstd, partials) -> None:
    dap, ldda = nhwc_view(da)
    yp, ldy = nhwc_view(y)
    n, h, w, c = y.shape
    st = _lib().fpb200_bn_relu_bwd_reduce(dap, ldda, yp, ldy, scale.data_ptr(), shift.data_ptr(),
                                          save_mean.data_ptr(), save_invstd.data_ptr(),
                                          partials.data_ptr(), n * h * w, c, _stream())
    capi.check(st, "bn_relu_bwd_reduce", N=n, H=h, W=w, C=c)


def bn_bwd_finalize(partials, count, scale, save_mean, save_invstd, dgamma, dbeta, coef) -> None:
    c = scale.numel()
    st = _lib().fpb200_bn_bwd_finalize(partials.data_ptr(), partials.shape[0], c, float(count),
                                       scale.data_ptr(), save_mean.data_ptr(), save_invstd.data_ptr(),
                                       _ptr(dgamma), _ptr(dbeta), coef.data_ptr(), _stream())
    capi.check(st, "bn_bwd_finalize", C=c)


def bn_relu_bwd_apply(da, y, dy, scale, shift, coef) -> None:
    dap, ldda = nhwc_view(da)
    yp, ldy = nhwc_view(y)
    dyp, lddy = nhwc_view(dy)
    n, h, w, c = y.shape
    st = _lib().fpb200_bn_relu_bwd_apply(dap, ldda, yp, ldy, dyp, lddy, scale.data_ptr(),
                                         shift.data_ptr(), coef.data_ptr(), n * h * w, c, _stream())
    capi.check(st, "bn_relu_bwd_apply", N=n, H=h, W=w, C=c)


# ------------------------------------------------------------------------------------------
def upsample2x_pad_concat_fwd(x, out) -> None:
    """out (view [N,Ho,Wo,C]) <- zero-padded bilinear x2 (align_corners=True) of x [N,h,w,C]."""
    xp, ldx = nhwc_view(x)
    op, ldo = nhwc_view(out)
    n, h, w, c = x.shape
    ho, wo = out.shape[1], out.shape[2]
    st = _lib().fpb200_upsample2x_pad_concat_fwd(xp, ldx, op, ldo, n, h, w, ho, wo, c, _stream())
    capi.check(st, "upsample2x_pad_concat_fwd", N=n, h=h, w=w, Ho=ho, Wo=wo, C=c)


def upsample2x_pad_concat_bwd(dout, dx) -> None:
    dop, lddo = nhwc_view(dout)
    dxp, lddx = nhwc_view(dx)
    n, h, w, c = dx.shape
    ho, wo = dout.shape[1], dout.shape[2]
    st = _lib().fpb200_upsample2x_pad_concat_bwd(dop, lddo, dxp, lddx, n, h, w, ho, wo, c, _stream())
    capi.check(st, "upsample2x_pad_concat_bwd", N=n, h=h, w=w, Ho=ho, Wo=wo, C=c)


# ------------------------------------------------------------------------------------------
def head1x1_fwd(x, w, b, logits, bn_scale=None, bn_shift=None) -> None:
    """With bn_scale/bn_shift, x is the raw conv output and relu(x*scale+shift) is fused in."""
    xp, ldx = nhwc_view(x)
    n, h, wd, c = x.shape
    ncls = w.shape[0]
    st = _lib().fpb200_head1x1_fwd(xp, ldx, w.data_ptr(), b.data_ptr(), logits.data_ptr(), n, h, wd, c,
                                   ncls, _ptr(bn_scale), _ptr(bn_shift), _stream())
    capi.check(st, "head1x1_fwd", N=n, H=h, W=wd, C=c, n_classes=ncls)


def head1x1_bwd(dlogits, x, w, dx, dw, db, partials, bn=None, bn_partials=None) -> None:
    """bn = (scale, shift, mean, invstd) of the layer whose RAW output x is: fuses the activation
    recompute and that layer's BatchNorm-backward reduction (-> bn_partials [rows][2][C])."""
    xp, ldx = nhwc_view(x)
    dxp, lddx = nhwc_view(dx)
    n, h, wd, c = x.shape
    ncls = w.shape[0]
    sc, sh, mu, istd = bn if bn is not None else (None, None, None, None)
    st = _lib().fpb200_head1x1_bwd(dlogits.data_ptr(), xp, ldx, w.data_ptr(), dxp, lddx, dw.data_ptr(),
                                   db.data_ptr(), partials.data_ptr(), n, h, wd, c, ncls, _ptr(sc),
                                   _ptr(sh), _ptr(mu), _ptr(istd), _ptr(bn_partials), _stream())
    capi.check(st, "head1x1_bwd", N=n, H=h, W=wd, C=c, n_classes=ncls)


def softmax_ce_argmax_fwd(logits, target, ignore_index, result, pred, confusion, partials) -> None:
    n, ncls = logits.shape[0], logits.shape[1]
    hw = logits.numel() // (n * ncls)
    st = _lib().fpb200_softmax_ce_argmax_fwd(logits.data_ptr(), target.data_ptr(), int(ignore_index),
                                             result.data_ptr(), _ptr(pred), _ptr(confusion),
                                             partials.data_ptr(), n, ncls, hw, _stream())
    capi.check(st, "softmax_ce_argmax_fwd", N=n, n_classes=ncls, hw=hw)


def softmax_ce_bwd(logits, target, ignore_index, result, grad_out, dlogits) -> None:
    n, ncls = logits.shape[0], logits.shape[1]
    hw = logits.numel() // (n * ncls)
    st = _lib().fpb200_softmax_ce_bwd(logits.data_ptr(), target.data_ptr(), int(ignore_index),
                                      result.data_ptr(), _ptr(grad_out), dlogits.data_ptr(), n, ncls,
                                      hw, _stream())
    capi.check(st, "softmax_ce_bwd", N=n, n_classes=ncls, hw=hw)


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0) -> None:
    st = _lib().fpb200_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr,
                                 beta1, beta2, eps, step, grad_scale, _stream())
    capi.check(st, "adam_step", n=p.numel())


def adam_step_graphable(p, g, m, v, lr, beta1, beta2, eps, step_state, grad_scale=1.0) -> None:
    """Adam with the step counter on the device (int32[2]); CUDA-graph replayable."""
    st = _lib().fpb200_adam_step_graphable(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                           p.numel(), lr, beta1, beta2, eps, step_state.data_ptr(),
                                           grad_scale, _stream())
    capi.check(st, "adam_step_graphable", n=p.numel())


# ------------------------------------------------------------------------------------------
def ingest_scene_tiles(scene: torch.Tensor, tiles: torch.Tensor, th: int, tw: int, c_pad: int) -> torch.Tensor:
    """scene [C,H,W] fp32 (device) + tiles int32 [n,4] (device: h0,w0,valid_h,valid_w) -> NHWC bf16."""
    _require_cuda(scene, tiles)
    c, h, w = scene.shape
    n = tiles.shape[0]
    out = torch.empty((n, th, tw, c_pad), dtype=torch.bfloat16, device=scene.device)
    st = _lib().fpb200_ingest_scene_tiles(scene.data_ptr(), c, h, w, tiles.data_ptr(), n, th, tw,
                                          out.data_ptr(), c_pad, _stream())
    capi.check(st, "ingest_scene_tiles", C=c, H=h, W=w, n_tiles=n, th=th, tw=tw)
    return out


def softmax_stitch_add(logits, canvas, weight, tiles) -> None:
    n, ncls, th, tw = logits.shape
    h, w = weight.shape
    st = _lib().fpb200_softmax_stitch_add(logits.data_ptr(), canvas.data_ptr(), weight.data_ptr(),
                                          tiles.data_ptr(), n, ncls, th, tw, h, w, _stream())
    capi.check(st, "softmax_stitch_add", n_tiles=n, n_classes=ncls, th=th, tw=tw, H=h, W=w)


def canvas_to_mask_u8(canvas, weight, mask) -> None:
    h, w = weight.shape
    st = _lib().fpb200_canvas_to_mask_u8(canvas.data_ptr(), weight.data_ptr(), mask.data_ptr(), h * w,
                                         canvas.shape[2], _stream())
    capi.check(st, "canvas_to_mask_u8", H=h, W=w, n_classes=canvas.shape[2])


# ------------------------------------------------------------------------------------------
def augment(image: torch.Tensor, target: Optional[torch.Tensor], theta: torch.Tensor, flags: torch.Tensor,
            xgrid: torch.Tensor, ygrid: torch.Tensor, mean: Optional[torch.Tensor] = None,
            std: Optional[torch.Tensor] = None, *, want_f32: bool = True, c_pad: int = 0):
    """Fused normalise + hflip / vflip / rotate(nearest) gather over a batch (base_dataset.py:77-113,
    532-555).  image fp32 NCHW, target int64 [N,H,W] or None; theta fp32 [N,6], flags int32 [N],
    mean/std float64 [N,C] or None.  Returns (image fp32 NCHW | None, NHWC bf16 [N,H,W,c_pad] | None,
    target int64 | None)."""
    _require_cuda(image, target, theta, flags, xgrid, ygrid, mean, std)
    if image.dim() != 4 or image.dtype != torch.float32:
        raise RuntimeError(f"augment: image must be fp32 NCHW, got {image.dtype} {tuple(image.shape)}")
    n, c, h, w = image.shape
    image = image.contiguous()
    if target is not None:
        if target.dtype != torch.int64 or tuple(target.shape) != (n, h, w):
            raise RuntimeError(f"augment: target must be int64 [N,H,W], got {target.dtype} {tuple(target.shape)}")
        target = target.contiguous()
    if tuple(theta.shape) != (n, 6) or theta.dtype != torch.float32 or tuple(flags.shape) != (n,) \
            or flags.dtype != torch.int32 or xgrid.numel() != w or ygrid.numel() != h:
        raise RuntimeError("augment: theta [N,6] fp32, flags [N] int32, xgrid [W], ygrid [H] expected")
    if (mean is None) != (std is None):
        raise RuntimeError("augment: mean and std go together")
    if mean is not None and (mean.dtype != torch.float64 or std.dtype != torch.float64
                             or mean.numel() != n * c or std.numel() != n * c):
        raise RuntimeError("augment: mean / std must be float64 [N,C]")
    out_f32 = torch.empty_like(image) if want_f32 else None
    out_bf16 = torch.empty((n, h, w, c_pad), dtype=torch.bfloat16, device=image.device) if c_pad else None
    tgt_out = torch.empty_like(target) if target is not None else None
    st = _lib().fpb200_augment_nchw_f32(
        image.data_ptr(), _ptr(out_f32), _ptr(out_bf16), c_pad, _ptr(target), _ptr(tgt_out),
        theta.contiguous().data_ptr(), flags.contiguous().data_ptr(), xgrid.contiguous().data_ptr(),
        ygrid.contiguous().data_ptr(), _ptr(mean.contiguous() if mean is not None else None),
        _ptr(std.contiguous() if std is not None else None), n, c, h, w, _stream())
    capi.check(st, "augment_nchw_f32", N=n, C=c, H=h, W=w, c_pad=c_pad)
    return out_f32, out_bf16, tgt_out


def plane_mean_std(image: torch.Tensor):
    """Per-(sample, channel) mean and population std of an fp32 NCHW batch -> two float64 [N,C]."""
    _require_cuda(image)
    n, c, h, w = image.shape
    image = image.contiguous()
    mean = torch.empty((n, c), dtype=torch.float64, device=image.device)
    std = torch.empty_like(mean)
    st = _lib().fpb200_plane_mean_std_f32(image.data_ptr(), mean.data_ptr(), std.data_ptr(), n * c, h * w,
                                          _stream())
    capi.check(st, "plane_mean_std_f32", planes=n * c, hw=h * w)
    return mean, std


# ------------------------------------------------------------------------------------------
# Device guard: every launcher runs on the device (and that device's current stream) of its
# tensors, not on whatever device happens to be current -- a model on cuda:1 while cuda:0 is the
# current device must not launch on cuda:0's stream.  Same-device calls pay two attribute reads.
def _first_cuda_device(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (list, tuple)):
            for b in a:
                if isinstance(b, torch.Tensor) and b.is_cuda:
                    return b.device
    return None


def _device_guarded(fn):
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _first_cuda_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


for _name, _fn in list(globals().items()):
    if (callable(_fn) and getattr(_fn, "__module__", None) == __name__ and not _name.startswith("_")
            and _name not in ("nhwc_view", "stat_rows", "bn_bwd_rows", "head_bwd_rows", "ce_rows",
                              "wgrad_workspace_bytes", "conv1x1_wgrad_workspace_bytes")):
        globals()[_name] = _device_guarded(_fn)
del _name, _fn
