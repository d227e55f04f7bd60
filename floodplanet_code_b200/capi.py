"""ctypes binding of the C ABI declared in ``include/floodplanet_b200.h``.

This is the *only* way Python reaches the CUDA kernels: plain pointers and sizes, no torch
types cross the boundary.  Loading fails loudly when the library has not been built; there
is no fallback path of any kind.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

_vp, _i, _l, _f, _d = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double

#: name -> (restype, argtypes).  Mirrors include/floodplanet_b200.h one to one; the
#: `not gpu` test-suite checks that every symbol declared in the header is listed here and
#: exported by the built library.
SIGNATURES = {
    "fpb200_abi_version": (_i, []),
    "fpb200_ingest_nchw_f32_to_nhwc_bf16": (_i, [C.POINTER(_vp), C.POINTER(_i), _i, _vp, _i, _i, _i, _i, _vp]),
    "fpb200_ingest_scene_tiles": (_i, [_vp, _i, _l, _l, _vp, _i, _i, _i, _vp, _i, _vp]),
    "fpb200_repack_weights_fprop": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "fpb200_repack_weights_dgrad": (_i, [_vp, _vp, _i, _i, _vp]),
    "fpb200_repack_weights_batch": (_i, [_vp, _i, _l, _vp]),
    "fpb200_conv_stat_rows": (_i, []),
    "fpb200_conv3x3_fprop_bf16_nhwc": (_i, [_vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "fpb200_conv3x3_dgrad_bf16_nhwc": (_i, [_vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _i, _vp, _l, _vp, _vp, _vp, _vp,
                                            _vp, _vp]),
    "fpb200_conv3x3_wgrad_workspace_bytes": (_l, [_i, _i, _i, _i, _i]),
    "fpb200_conv3x3_wgrad_bf16_nhwc": (_i, [_vp, _l, _vp, _l, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fpb200_bn_stats_finalize": (_i, [_vp, _i, _i, _d, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fpb200_bn_fold_eval": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "fpb200_bn_eval_stats": (_i, [_vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "fpb200_bn_bwd_finalize_frozen": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fpb200_bn_apply_relu": (_i, [_vp, _l, _vp, _l, _vp, _vp, _l, _i, _vp]),
    "fpb200_bn_apply_relu_maxpool2": (_i, [_vp, _l, _vp, _l, _vp, _l, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fpb200_maxpool2_bwd": (_i, [_vp, _l, _vp, _vp, _l, _vp, _l, _i, _i, _i, _i, _vp, _l, _vp, _vp, _vp, _vp, _vp,
                                 _vp]),
    "fpb200_bn_bwd_rows": (_i, []),
    "fpb200_bn_relu_bwd_reduce": (_i, [_vp, _l, _vp, _l, _vp, _vp, _vp, _vp, _vp, _l, _i, _vp]),
    "fpb200_bn_bwd_finalize": (_i, [_vp, _i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fpb200_bn_relu_bwd_apply": (_i, [_vp, _l, _vp, _l, _vp, _l, _vp, _vp, _vp, _l, _i, _vp]),
    "fpb200_upsample2x_pad_concat_fwd": (_i, [_vp, _l, _vp, _l, _i, _i, _i, _i, _i, _i, _vp]),
    "fpb200_upsample2x_pad_concat_bwd": (_i, [_vp, _l, _vp, _l, _i, _i, _i, _i, _i, _i, _vp]),
    "fpb200_head1x1_fwd": (_i, [_vp, _l, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "fpb200_head_bwd_rows": (_i, []),
    "fpb200_head1x1_bwd": (_i, [_vp, _vp, _l, _vp, _vp, _l, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                                _vp, _vp]),
    "fpb200_ce_rows": (_i, []),
    "fpb200_softmax_ce_argmax_fwd": (_i, [_vp, _vp, _l, _vp, _vp, _vp, _vp, _i, _i, _l, _vp]),
    "fpb200_softmax_ce_bwd": (_i, [_vp, _vp, _l, _vp, _vp, _vp, _i, _i, _l, _vp]),
    "fpb200_softmax_stitch_add": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _l, _l, _vp]),
    "fpb200_canvas_to_mask_u8": (_i, [_vp, _vp, _vp, _l, _i, _vp]),
    "fpb200_nchw_f32_to_nhwc_bf16": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _vp]),
    "fpb200_nhwc_bf16_to_nchw_f32": (_i, [_vp, _l, _vp, _i, _i, _i, _i, _vp]),
    "fpb200_repack_weights_1x1": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "fpb200_conv1x1_bf16_nhwc": (_i, [_vp, _l, _vp, _vp, _l, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "fpb200_conv1x1_wgrad_workspace_bytes": (_l, [_i, _i, _i, _i, _i]),
    "fpb200_conv1x1_wgrad_bf16_nhwc": (_i, [_vp, _l, _vp, _l, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fpb200_channel_sum_rows": (_i, []),
    "fpb200_channel_sum_bf16_nhwc": (_i, [_vp, _l, _vp, _vp, _l, _i, _vp]),
    "fpb200_augment_nchw_f32": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fpb200_plane_mean_std_f32": (_i, [_vp, _vp, _vp, _i, _l, _vp]),
    "fpb200_adam_step": (_i, [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _i, _f, _vp]),
    "fpb200_adam_step_graphable": (_i, [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _vp, _f, _vp]),
    "fpb200_nccl_version": (_i, []),
    "fpb200_nccl_unique_id": (_i, [_vp]),
    "fpb200_nccl_comm_create": (_i, [C.POINTER(_vp), _i, _i, _vp, _i]),
    "fpb200_nccl_comm_destroy": (_i, [_vp]),
    "fpb200_allreduce_f32": (_i, [_vp, _vp, _l, _i, _vp]),
}

_ERRORS = {
    -1: "unsupported or inconsistent shape",
    -2: "pointer / pitch alignment violated",
    -3: "CUDA launch or runtime error",
    -4: "CUDA driver entry point unavailable",
    -5: "TMA tensor-map encode rejected the view",
    -6: "NCCL runtime missing or NCCL call failed",
}

_LIB = None


def library_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the shared library and attach prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if build_if_missing:
        path = _build.build_library()
    if not path.exists():
        raise RuntimeError(
            f"floodplanet_b200: CUDA library {path} is missing; run `python -m "
            "floodplanet_code_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError -> symbol missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(status: int, what: str, **shape) -> None:
    """Raise RuntimeError(kernel name + shape) on a non-zero status (header error convention)."""
    if status != 0:
        desc = ", ".join(f"{k}={v}" for k, v in shape.items())
        raise RuntimeError(
            f"floodplanet_b200.{what} failed: {_ERRORS.get(status, 'status ' + str(status))} ({desc})")
