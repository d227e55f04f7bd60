"""Per-kernel parity tests (GPU): every C-ABI kernel against the torch op the reference
dispatches to, on the same seeded inputs.

Tolerances (BASELINE.json north_star): bit-exact for integer outputs (pool indices, argmax,
confusion counts); <= 1e-2 relative for bf16-compute forward values; <= 2e-2 for gradients.
Inputs are rounded to bf16 before the fp32 torch op runs, so the comparison isolates the
kernel's own arithmetic (fp32 accumulate, one bf16 rounding on store).
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from floodplanet_code_b200 import ops as _ops
    return _ops


@pytest.fixture(params=["halo", "per_tap"])
def conv_impl(request):
    """Every fprop case runs through the product's halo kernel AND through an independent
    implementation: the first-generation one-box-per-tap kernel, built as a TEST-ONLY library
    (tests/csrc/conv_pertap_crosscheck.cu -> tests/lib/libfpb200_crosscheck.so); the product library
    does not contain it.  Returns the `_fn` argument of ops.conv3x3_fprop (None = product kernel)."""
    if request.param == "halo":
        return None
    import ctypes as C
    from floodplanet_code_b200 import build, capi
    lib = C.CDLL(str(build.build_test_library()))
    fn = lib.fpb200_test_conv3x3_pertap_bf16_nhwc
    fn.restype, fn.argtypes = capi.SIGNATURES["fpb200_conv3x3_fprop_bf16_nhwc"]
    return fn


def rel(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


def rand_act(n, h, w, c, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(n, h, w, c, generator=g, device="cuda") * scale).to(torch.bfloat16)


def rand_w(cout, cin, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / math.sqrt(9 * cin)
    return w.to(torch.bfloat16).float()  # exactly representable in bf16


CONV_SHAPES = [
    # n, h, w, cin, cout
    (2, 16, 16, 64, 64),
    (1, 32, 32, 64, 128),
    (2, 20, 24, 128, 256),     # ragged tiles
    (1, 37, 37, 64, 64),       # odd size from the 300-px crop path
    (2, 16, 16, 16, 64),       # first layer, KCH=16 (32B swizzle)
    (1, 16, 24, 32, 64),       # early fusion pad 32 (64B swizzle)
    (1, 8, 8, 512, 512),       # two N blocks, deep K
    (1, 128, 128, 64, 64),     # wide rows: 1x128 tiles
    (3, 18, 18, 256, 128),
]


@pytest.mark.parametrize("n,h,w,cin,cout", CONV_SHAPES)
def test_conv3x3_fprop_and_stats(ops, conv_impl, n, h, w, cin, cout):
    x = rand_act(n, h, w, cin, 1)
    wt = rand_w(cout, cin, 2)
    wp = ops.repack_fprop(wt, cin)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    parts = torch.empty(ops.stat_rows(), 2, cout, dtype=torch.float32, device="cuda")
    ops.conv3x3_fprop(x, wp, y, stat_partials=parts, _fn=conv_impl)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x.float()), wt, padding=1)
    err = rel(nchw(y.float()), ref)
    assert err < 1e-2, f"fprop rel err {err}"
    s = parts.double().sum(0)
    ref_sum = ref.double().sum((0, 2, 3))
    ref_sq = (ref.double() ** 2).sum((0, 2, 3))
    assert rel(s[0], ref_sum) < 2e-3 or float((s[0] - ref_sum).abs().max()) < 1e-2 * float(ref_sq.sqrt().max())
    assert rel(s[1], ref_sq) < 2e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (1, 20, 24, 128, 256)])
def test_conv3x3_fprop_affine_relu_into_concat_view(ops, conv_impl, n, h, w, cin, cout):
    x = rand_act(n, h, w, cin, 3)
    wt = rand_w(cout, cin, 4)
    wp = ops.repack_fprop(wt, cin)
    g = torch.Generator(device="cuda").manual_seed(5)
    scale = torch.rand(cout, generator=g, device="cuda") + 0.5
    shift = torch.randn(cout, generator=g, device="cuda") * 0.2
    buf = torch.full((n, h, w, 2 * cout), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3_fprop(x, wp, buf[..., :cout], scale=scale, shift=shift, relu=True, _fn=conv_impl)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(nchw(x.float()), wt, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    assert rel(nchw(buf[..., :cout].float()), ref) < 1e-2
    assert bool((buf[..., cout:] == 7.0).all()), "wrote outside the channel slice"


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (1, 20, 24, 128, 256), (1, 19, 19, 256, 64),
                                            (1, 16, 16, 1024, 512)])
def test_conv3x3_dgrad(ops, n, h, w, cin, cout):
    dy = rand_act(n, h, w, cout, 6)
    wt = rand_w(cout, cin, 7)
    wd = ops.repack_dgrad(wt)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3_dgrad(dy, wd, dx)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((n, cin, h, w), wt, nchw(dy.float()), padding=1)
    err = rel(nchw(dx.float()), ref)
    assert err < 2e-2, f"dgrad rel err {err}"


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (1, 20, 24, 128, 256), (1, 37, 37, 128, 64),
                                            (2, 16, 16, 256, 128)])
def test_conv3x3_dgrad_fused_bn_backward_reduce(ops, n, h, w, cin, cout):
    """dgrad whose epilogue also reduces the BatchNorm-backward sums of the layer that produced
    the conv's input: same dx bit for bit, same sums as the separate reduction pass."""
    dy = rand_act(n, h, w, cout, 50)
    wt = rand_w(cout, cin, 51)
    wd = ops.repack_dgrad(wt)
    y_prev = rand_act(n, h, w, cin, 52)
    g = torch.Generator(device="cuda").manual_seed(53)
    scale = torch.rand(cin, generator=g, device="cuda") + 0.5
    shift = torch.randn(cin, generator=g, device="cuda") * 0.3
    mean = torch.randn(cin, generator=g, device="cuda") * 0.2
    invstd = torch.rand(cin, generator=g, device="cuda") + 0.5
    dx0 = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
    dx1 = torch.empty_like(dx0)
    ops.conv3x3_dgrad(dy, wd, dx0)
    parts = torch.empty(ops.stat_rows(), 2, cin, device="cuda")
    ops.conv3x3_dgrad(dy, wd, dx1, bn_y=y_prev, bn=(scale, shift, mean, invstd), bn_partials=parts)
    torch.cuda.synchronize()
    assert torch.equal(dx0, dx1)
    ref = torch.empty(ops.bn_bwd_rows(), 2, cin, device="cuda")
    ops.bn_relu_bwd_reduce(dx0, y_prev, scale, shift, mean, invstd, ref)
    a, b = parts.double().sum(0), ref.double().sum(0)
    assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()), (a - b).abs().max()


WGRAD_SHAPES = [
    # n, h, w, cin(padded), cin_real, cout
    (2, 16, 16, 128, 128, 128),   # MODE_X_SHIFT, NB=2
    (1, 24, 20, 64, 64, 128),     # MODE_X_SHIFT, NB=1
    (2, 16, 16, 64, 64, 64),      # MODE_DY_SHIFT, 64
    (2, 16, 16, 16, 4, 64),       # first layer (pad 16)
    (1, 16, 16, 32, 21, 64),      # early fusion pad 32
    (1, 37, 29, 16, 6, 64),       # MODE_X_STACK: odd size (ragged tiles, halo beyond the image)
    (2, 20, 24, 48, 40, 64),      # MODE_X_STACK: three 16-channel input blocks
    (1, 16, 32, 16, 16, 128),     # MODE_X_STACK: two 64-channel output blocks
    (1, 37, 37, 128, 128, 64),    # odd size, Cout 64 / Cin 128
    (3, 18, 18, 256, 256, 512),
    (1, 64, 64, 64, 64, 64),
]


@pytest.mark.parametrize("n,h,w,cin,cin_real,cout", WGRAD_SHAPES)
def test_conv3x3_wgrad(ops, n, h, w, cin, cin_real, cout):
    x = rand_act(n, h, w, cin, 8)
    if cin_real < cin:
        x[..., cin_real:] = 0
    dy = rand_act(n, h, w, cout, 9)
    dw = torch.empty(cout, cin_real, 3, 3, dtype=torch.float32, device="cuda")
    ws = torch.empty(ops.wgrad_workspace_bytes(n, h, w, cin, cout) // 4, dtype=torch.float32, device="cuda")
    ops.conv3x3_wgrad(x, dy, dw, ws, cin_real)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(nchw(x.float())[:, :cin_real], (cout, cin_real, 3, 3), nchw(dy.float()),
                                      padding=1)
    err = rel(dw, ref)
    assert err < 2e-3, f"wgrad rel err {err}"


def test_wgrad_through_concat_view(ops):
    n, h, w, c = 2, 16, 16, 64
    buf = rand_act(n, h, w, 2 * c, 10)
    dy = rand_act(n, h, w, 128, 11)
    dw = torch.empty(128, 2 * c, 3, 3, dtype=torch.float32, device="cuda")
    ws = torch.empty(ops.wgrad_workspace_bytes(n, h, w, 2 * c, 128) // 4, dtype=torch.float32, device="cuda")
    ops.conv3x3_wgrad(buf, dy, dw, ws, 2 * c)
    ref = torch.nn.grad.conv2d_weight(nchw(buf.float()), (128, 2 * c, 3, 3), nchw(dy.float()), padding=1)
    assert rel(dw, ref) < 2e-3


# ------------------------------------------------------------------------------------------
def test_ingest_early_fusion(ops):
    g = torch.Generator(device="cuda").manual_seed(0)
    img = torch.rand(2, 4, 20, 24, generator=g, device="cuda")
    dem = torch.rand(2, 1, 20, 24, generator=g, device="cuda")
    s2 = torch.rand(2, 10, 20, 24, generator=g, device="cuda")
    out = ops.ingest([img, dem, s2], 16)
    ref = torch.cat([img, dem, s2], 1).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out[..., :15], ref)
    assert bool((out[..., 15:] == 0).all())


def test_repack(ops):
    wt = rand_w(64, 4, 1)
    wp = ops.repack_fprop(wt, 16)
    ref = torch.zeros(64, 9, 16, device="cuda")
    ref[:, :, :4] = wt.permute(0, 2, 3, 1).reshape(64, 9, 4)
    assert torch.equal(wp.float(), ref)
    wt = rand_w(128, 64, 2)
    wd = ops.repack_dgrad(wt)
    ref = wt.flip(2, 3).permute(1, 2, 3, 0).reshape(64, 9, 128)
    assert torch.equal(wd.float(), ref)


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 37, 37, 128), (2, 9, 75, 64)])
def test_bn_train_forward_chain(ops, n, h, w, c):
    """conv stats -> finalize -> apply+relu(+maxpool) against F.batch_norm/relu/max_pool2d."""
    cin = 64
    x = rand_act(n, h, w, cin, 12)
    wt = rand_w(c, cin, 13)
    g = torch.Generator(device="cuda").manual_seed(14)
    gamma = torch.rand(c, generator=g, device="cuda") + 0.5
    beta = torch.randn(c, generator=g, device="cuda") * 0.1
    bias = torch.randn(c, generator=g, device="cuda") * 0.1
    rm = torch.zeros(c, device="cuda")
    rv = torch.ones(c, device="cuda")
    y = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    parts = torch.empty(ops.stat_rows(), 2, c, device="cuda")
    ops.conv3x3_fprop(x, ops.repack_fprop(wt, cin), y, stat_partials=parts)
    scale, shift, mean, invstd = (torch.empty(c, device="cuda") for _ in range(4))
    ops.bn_stats_finalize(parts, n * h * w, gamma, beta, bias, 1e-5, 0.1, rm, rv, scale, shift, mean, invstd)
    a = torch.empty_like(y)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device="cuda")
    ops.bn_apply_relu_maxpool2(y, a, pooled, idx, scale, shift)
    a2 = torch.empty_like(y)
    ops.bn_apply_relu(y, a2, scale, shift)
    torch.cuda.synchronize()
    # reference in fp32 from the same bf16 inputs
    conv = F.conv2d(nchw(x.float()), wt, bias, padding=1)
    rm_ref, rv_ref = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    ref = F.relu(F.batch_norm(conv, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5))
    assert rel(rm, rm_ref) < 1e-3 and rel(rv, rv_ref) < 1e-3
    assert rel(nchw(a.float()), ref) < 1e-2
    assert torch.equal(a, a2)
    # pooling: bit-exact against torch on the SAME (stored) pre-pool activation
    pref, iref = F.max_pool2d(nchw(a.float()), 2, return_indices=True)
    assert torch.equal(nchw(pooled.float()), pref)
    ii = nchw(idx).long()
    hh = torch.arange(h // 2, device="cuda")[None, None, :, None] * 2 + ii // 2
    ww = torch.arange(w // 2, device="cuda")[None, None, None, :] * 2 + ii % 2
    assert torch.equal(hh * w + ww, iref), "max-pool argmax differs from torch"


def test_maxpool_ties_and_nan(ops):
    a = torch.zeros(1, 4, 4, 8, dtype=torch.bfloat16, device="cuda")
    a[0, 1, 1, 0] = float("nan")
    a[0, 2, 3, 1] = 3.0
    pooled = torch.empty(1, 2, 2, 8, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(1, 2, 2, 8, dtype=torch.uint8, device="cuda")
    ops.bn_apply_relu_maxpool2(a, None, pooled, idx, None, None)
    pref, iref = F.max_pool2d(nchw(a.float()), 2, return_indices=True)
    ii = nchw(idx).long()
    hh = torch.arange(2, device="cuda")[None, None, :, None] * 2 + ii // 2
    ww = torch.arange(2, device="cuda")[None, None, None, :] * 2 + ii % 2
    assert torch.equal(hh * 4 + ww, iref)
    assert torch.equal(torch.isnan(nchw(pooled.float())), torch.isnan(pref))
    assert iref[0, 2].flatten().tolist() == [0, 2, 8, 10]  # all-zero window -> first element


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 37, 37, 128)])
def test_maxpool_bwd_with_skip(ops, n, h, w, c):
    a = rand_act(n, h, w, c, 15)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device="cuda")
    ops.bn_apply_relu_maxpool2(a, None, pooled, idx, None, None)
    dp = rand_act(n, h // 2, w // 2, c, 16)
    dcat = rand_act(n, h, w, 2 * c, 17)
    dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    ops.maxpool2_bwd(dp, idx, dcat[..., :c], dx)
    af = nchw(a.float()).requires_grad_(True)
    F.max_pool2d(af, 2).backward(nchw(dp.float()))
    ref = (af.grad + nchw(dcat[..., :c].float())).to(torch.bfloat16)
    assert torch.equal(nchw(dx), ref)


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 37, 37, 128), (3, 10, 12, 512)])
def test_maxpool_bwd_fused_bn_backward_reduce(ops, n, h, w, c):
    a = rand_act(n, h, w, c, 60)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device="cuda")
    ops.bn_apply_relu_maxpool2(a, None, pooled, idx, None, None)
    dp = rand_act(n, h // 2, w // 2, c, 61)
    dcat = rand_act(n, h, w, 2 * c, 62)
    y = rand_act(n, h, w, c, 63)
    g = torch.Generator(device="cuda").manual_seed(64)
    co = [torch.rand(c, generator=g, device="cuda") + 0.5, torch.randn(c, generator=g, device="cuda") * 0.3,
          torch.randn(c, generator=g, device="cuda") * 0.2, torch.rand(c, generator=g, device="cuda") + 0.5]
    dx0 = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    dx1 = torch.empty_like(dx0)
    ops.maxpool2_bwd(dp, idx, dcat[..., :c], dx0)
    parts = torch.empty(2 * ops.bn_bwd_rows(), 2, c, device="cuda")
    ops.maxpool2_bwd(dp, idx, dcat[..., :c], dx1, bn_y=y, bn=tuple(co), bn_partials=parts)
    assert torch.equal(dx0, dx1)
    ref = torch.empty(ops.bn_bwd_rows(), 2, c, device="cuda")
    ops.bn_relu_bwd_reduce(dx0, y, co[0], co[1], co[2], co[3], ref)
    a_, b_ = parts.double().sum(0), ref.double().sum(0)
    assert float((a_ - b_).abs().max()) <= 1e-4 * float(b_.abs().max())


@pytest.mark.parametrize("n,h,w,c", [(2, 16, 16, 64), (1, 19, 23, 256), (2, 8, 8, 512)])
def test_bn_relu_backward(ops, n, h, w, c):
    y = rand_act(n, h, w, c, 18)
    da = rand_act(n, h, w, c, 19)
    g = torch.Generator(device="cuda").manual_seed(20)
    gamma = torch.rand(c, generator=g, device="cuda") + 0.5
    beta = torch.randn(c, generator=g, device="cuda") * 0.3
    yf = nchw(y.float()).requires_grad_(True)
    gam = gamma.clone().requires_grad_(True)
    bet = beta.clone().requires_grad_(True)
    out = F.relu(F.batch_norm(yf, None, None, gam, bet, True, 0.1, 1e-5))
    out.backward(nchw(da.float()))
    mean = yf.detach().mean((0, 2, 3))
    var = yf.detach().var((0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    scale = gamma * invstd
    shift = beta - mean * scale
    parts = torch.empty(ops.bn_bwd_rows(), 2, c, device="cuda")
    ops.bn_relu_bwd_reduce(da, y, scale, shift, mean, invstd, parts)
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    coef = torch.empty(2, c, device="cuda")
    ops.bn_bwd_finalize(parts, n * h * w, scale, mean, invstd, dgamma, dbeta, coef)
    dy = torch.empty_like(y)
    ops.bn_relu_bwd_apply(da, y, dy, scale, shift, coef)
    torch.cuda.synchronize()
    assert rel(dgamma, gam.grad) < 2e-3
    assert rel(dbeta, bet.grad) < 2e-3
    assert rel(nchw(dy.float()), yf.grad) < 1e-2


@pytest.mark.parametrize("n,h,w,ho,wo,c", [(2, 8, 8, 16, 16, 64), (1, 18, 18, 37, 37, 128), (1, 1, 1, 2, 2, 64),
                                           (1, 4, 5, 9, 10, 64),
                                           # several column chunks and row groups of the backward, ragged last ones
                                           (1, 20, 70, 40, 140, 64), (2, 37, 150, 75, 300, 64),
                                           # 64 channel groups (4 columns per pass), padding on every side
                                           (1, 9, 9, 21, 20, 512),
                                           # channel-group counts that do not divide the block (idle threads)
                                           (2, 11, 13, 22, 26, 24), (1, 75, 75, 151, 151, 8), (1, 6, 7, 12, 14, 192)])
def test_upsample_pad_concat(ops, n, h, w, ho, wo, c):
    x = rand_act(n, h, w, c, 21)
    cat = torch.zeros(n, ho, wo, 2 * c, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x_pad_concat_fwd(x, cat[..., c:])
    xf = nchw(x.float()).requires_grad_(True)
    up = F.interpolate(xf, scale_factor=2, mode="bilinear", align_corners=True)
    dY, dX = ho - up.shape[2], wo - up.shape[3]
    ref = F.pad(up, [dX // 2, dX - dX // 2, dY // 2, dY - dY // 2])
    assert rel(nchw(cat[..., c:].float()), ref) < 5e-3
    assert bool((cat[..., :c] == 0).all())
    dcat = rand_act(n, ho, wo, 2 * c, 22)
    dx = torch.empty_like(x)
    ops.upsample2x_pad_concat_bwd(dcat[..., c:], dx)
    ref.backward(nchw(dcat[..., c:].float()))
    assert rel(nchw(dx.float()), xf.grad) < 5e-3


def test_bilinear_align_corners_golden(ops):
    # a 4 -> 8 row ramp maps to [0, 3/7, ..., 3]
    x = torch.arange(4, dtype=torch.float32, device="cuda").view(1, 1, 4, 1).expand(1, 4, 4, 64)
    x = x.contiguous().to(torch.bfloat16)  # value = column index w
    out = torch.empty(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    ops.upsample2x_pad_concat_fwd(x, out)
    want = torch.tensor([i * 3 / 7 for i in range(8)], device="cuda")
    assert torch.allclose(out[0, 0, :, 0].float(), want, rtol=4e-3, atol=0)
    assert float(out[0, 0, 0, 0]) == 0.0 and float(out[0, 0, 7, 0]) == 3.0


@pytest.mark.parametrize("ncls,hw", [(2, (20, 24)), (3, (20, 24)), (3, (19, 23)), (4, (7, 5)), (5, (9, 11))])
def test_head_fwd_bwd(ops, ncls, hw):
    # (19, 23), (7, 5), (9, 11): pixel counts that are not multiples of the 32 / 16 pixels a warp
    # iteration covers, images that end inside a warp's pixel group
    n, (h, w), c = 2, hw, 64
    x = rand_act(n, h, w, c, 23)
    g = torch.Generator(device="cuda").manual_seed(24)
    wt = torch.randn(ncls, c, generator=g, device="cuda") / 8
    b = torch.randn(ncls, generator=g, device="cuda")
    logits = torch.empty(n, ncls, h, w, device="cuda")
    ops.head1x1_fwd(x, wt, b, logits)
    xf = nchw(x.float()).requires_grad_(True)
    wr = wt.clone().view(ncls, c, 1, 1).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    ref = F.conv2d(xf, wr, br)
    assert rel(logits, ref) < 1e-5
    dl = torch.randn(n, ncls, h, w, generator=g, device="cuda")
    ref.backward(dl)
    dx = torch.empty_like(x)
    dw = torch.empty(ncls, c, device="cuda")
    db = torch.empty(ncls, device="cuda")
    parts = torch.empty(ops.head_bwd_rows(), ncls * (c + 1), device="cuda")
    ops.head1x1_bwd(dl, x, wt, dx, dw, db, parts)
    assert rel(nchw(dx.float()), xf.grad) < 1e-2
    assert rel(dw, wr.grad.view(ncls, c)) < 1e-4
    assert rel(db, br.grad) < 1e-4


def test_head_fused_bn_apply_and_bn_bwd_reduce(ops):
    """Head kernels consuming the RAW last-conv output: forward == head(bn_apply_relu(y)) bit for
    bit; backward also emits that layer's BatchNorm-backward sums (== the separate reduce pass)."""
    n, h, w, c, ncls = 2, 20, 24, 64, 3
    y = rand_act(n, h, w, c, 40)
    g = torch.Generator(device="cuda").manual_seed(41)
    scale = torch.rand(c, generator=g, device="cuda") + 0.5
    shift = torch.randn(c, generator=g, device="cuda") * 0.3
    mean = torch.randn(c, generator=g, device="cuda") * 0.2
    invstd = torch.rand(c, generator=g, device="cuda") + 0.5
    wt = torch.randn(ncls, c, generator=g, device="cuda") / 8
    b = torch.randn(ncls, generator=g, device="cuda")
    a = torch.empty_like(y)
    ops.bn_apply_relu(y, a, scale, shift)
    l_ref = torch.empty(n, ncls, h, w, device="cuda")
    l_fused = torch.empty_like(l_ref)
    ops.head1x1_fwd(a, wt, b, l_ref)
    ops.head1x1_fwd(y, wt, b, l_fused, scale, shift)
    assert torch.equal(l_ref, l_fused)
    dl = torch.randn(n, ncls, h, w, generator=g, device="cuda") * 1e-3
    outs = []
    for fused in (False, True):
        dx = torch.empty_like(y)
        dw = torch.empty(ncls, c, device="cuda")
        db = torch.empty(ncls, device="cuda")
        parts = torch.empty(ops.head_bwd_rows(), ncls * (c + 1), device="cuda")
        bnp = torch.zeros(ops.head_bwd_rows(), 2, c, device="cuda")
        if fused:
            ops.head1x1_bwd(dl, y, wt, dx, dw, db, parts, bn=(scale, shift, mean, invstd), bn_partials=bnp)
        else:
            ops.head1x1_bwd(dl, a, wt, dx, dw, db, parts)
        outs.append((dx, dw, db, bnp))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    ref_parts = torch.empty(ops.bn_bwd_rows(), 2, c, device="cuda")
    ops.bn_relu_bwd_reduce(outs[0][0], y, scale, shift, mean, invstd, ref_parts)
    assert rel(outs[1][3].double().sum(0), ref_parts.double().sum(0)) < 1e-5


def _ce(ops, logits, target, ignore_index):
    n, ncls = logits.shape[:2]
    result = torch.empty(4, dtype=torch.float64, device="cuda")
    pred = torch.empty(target.shape, dtype=torch.int64, device="cuda")
    conf = torch.zeros(ncls, ncls, dtype=torch.int64, device="cuda")
    parts = torch.empty(ops.ce_rows(), 4, dtype=torch.float64, device="cuda")
    ops.softmax_ce_argmax_fwd(logits, target, ignore_index, result, pred, conf, parts)
    dl = torch.empty_like(logits)
    go = torch.ones((), device="cuda")
    ops.softmax_ce_bwd(logits, target, ignore_index, result, go, dl)
    torch.cuda.synchronize()
    return result, pred, conf, dl


def test_masked_ce_argmax(ops):
    g = torch.Generator(device="cuda").manual_seed(25)
    n, ncls, h, w = 3, 3, 33, 47
    logits = torch.randn(n, ncls, h, w, generator=g, device="cuda") * 3
    target = (torch.rand(n, h, w, generator=g, device="cuda") < 0.42).long()
    result, pred, conf, dl = _ce(ops, logits, target, 0)
    lr = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr, target, ignore_index=0)
    ref.backward()
    assert abs(float(result[3]) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert int(result[1]) == int((target != 0).sum())
    assert torch.equal(pred, logits.argmax(1))
    assert rel(dl, lr.grad) < 1e-5
    m = target != 0
    want = torch.zeros(ncls, ncls, dtype=torch.int64, device="cuda")
    want.index_put_((target[m], logits.argmax(1)[m]), torch.ones((), dtype=torch.int64, device="cuda"), accumulate=True)
    assert torch.equal(conf, want)


def test_ce_all_ignored_and_ties(ops):
    logits = torch.zeros(1, 3, 4, 4, device="cuda")  # ties -> class 0
    logits[0, :, 0, 0] = torch.tensor([1.0, float("nan"), 2.0])
    target = torch.zeros(1, 4, 4, dtype=torch.int64, device="cuda")
    result, pred, conf, dl = _ce(ops, logits, target, 0)
    assert math.isnan(float(result[3]))        # torch: mean over empty set -> NaN
    assert bool((dl == 0).all())               # ... and zero grads after nan_to_num
    assert torch.equal(pred, logits.argmax(1))  # ties -> lowest index, NaN maximal
    assert int(conf.sum()) == 0


def test_adam_matches_torch(ops):
    g = torch.Generator(device="cuda").manual_seed(26)
    p = torch.randn(10007, generator=g, device="cuda")
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-4)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(10007, generator=g, device="cuda")
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-8, step)
    assert rel(p, pr.detach()) < 1e-6


def test_bn_fold_eval_and_eval_conv_epilogue(ops):
    """eval-mode BatchNorm (+ conv bias) folded into the conv epilogue == F.batch_norm(training=False)."""
    n, h, w, cin, c = 2, 20, 24, 64, 128
    x = rand_act(n, h, w, cin, 30)
    wt = rand_w(c, cin, 31)
    g = torch.Generator(device="cuda").manual_seed(32)
    gamma = torch.rand(c, generator=g, device="cuda") + 0.5
    beta = torch.randn(c, generator=g, device="cuda") * 0.1
    bias = torch.randn(c, generator=g, device="cuda") * 0.1
    rm = torch.randn(c, generator=g, device="cuda") * 0.2
    rv = torch.rand(c, generator=g, device="cuda") + 0.5
    scale, shift = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_fold_eval(gamma, beta, bias, rm, rv, 1e-5, scale, shift)
    a = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3_fprop(x, ops.repack_fprop(wt, cin), a, scale=scale, shift=shift, relu=True)
    ref = F.relu(F.batch_norm(F.conv2d(nchw(x.float()), wt, bias, padding=1), rm, rv, gamma, beta, False, 0.1, 1e-5))
    assert rel(nchw(a.float()), ref) < 1e-2


def test_error_status_raises_runtime_error(ops):
    x = rand_act(1, 16, 16, 24, 33)          # 24 input channels: not a supported padding
    y = torch.empty(1, 16, 16, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="conv3x3_fprop"):
        ops.conv3x3_fprop(x, torch.empty(64, 9, 24, dtype=torch.bfloat16, device="cuda"), y)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ingest([torch.zeros(1, 4, 16, 16)], 16)


# ---------------------------------------------------------------------------------------------
# pointwise convolution + feature-level seams (late fusion, encode/decode API)
# ---------------------------------------------------------------------------------------------
PW_SHAPES = [
    # n, h, w, cin, cout
    (2, 16, 16, 64, 64),
    (1, 32, 32, 128, 64),
    (2, 20, 24, 256, 128),     # ragged patches
    (1, 37, 37, 128, 64),      # odd size
    (1, 8, 8, 1024, 512),
    (3, 18, 18, 512, 256),
    (1, 16, 16, 192, 64),      # three modalities at level 0
]


def rand_w1(cout, cin, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.randn(cout, cin, 1, 1, generator=g, device="cuda") / math.sqrt(cin)
    return w.to(torch.bfloat16).float()


@pytest.mark.parametrize("n,h,w,cin,cout", PW_SHAPES)
def test_conv1x1_fprop_dgrad_wgrad(ops, n, h, w, cin, cout):
    x = rand_act(n, h, w, cin, 11)
    wt = rand_w1(cout, cin, 12)
    bias = torch.randn(cout, device="cuda")
    # forward, written into a channel slice of a wider buffer (the decoder's concat buffer)
    wide = torch.zeros(n, h, w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    y = wide[..., :cout]
    ops.conv1x1(x, ops.repack_1x1(wt, False), y, torch.ones(cout, device="cuda"), bias)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(x.float()), wt, bias)
    assert rel(nchw(y.float()), ref) < 1e-2
    assert float(wide[..., cout:].abs().max()) == 0.0          # neighbouring channels untouched
    # data gradient
    dy = rand_act(n, h, w, cout, 13)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
    ops.conv1x1(dy, ops.repack_1x1(wt, True), dx)
    ref_dx = F.conv_transpose2d(nchw(dy.float()), wt)
    assert rel(nchw(dx.float()), ref_dx) < 2e-2
    # weight gradient (x read from a channel slice too)
    ws = torch.empty(ops.conv1x1_wgrad_workspace_bytes(n, h, w, cin, cout) // 4, device="cuda")
    dw = torch.empty(cout, cin, 1, 1, device="cuda")
    ops.conv1x1_wgrad(x, dy, dw, ws)
    ref_dw = torch.einsum("nhwo,nhwi->oi", dy.float(), x.float())
    assert rel(dw.view(cout, cin), ref_dw) < 2e-2
    dw2 = torch.empty_like(dw)
    ops.conv1x1_wgrad(x, dy, dw2, ws)
    assert torch.equal(dw, dw2)                                 # deterministic
    # bias gradient
    db = torch.empty(cout, device="cuda")
    ops.channel_sum(dy, db)
    assert rel(db, dy.float().sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("n,c,h,w", [(2, 64, 16, 16), (1, 128, 37, 37), (2, 512, 5, 7), (1, 8, 9, 9)])
def test_feature_layout_round_trip(ops, n, c, h, w):
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randn(n, c, h, w, generator=g, device="cuda")
    wide = torch.zeros(n, h, w, c + 64, dtype=torch.bfloat16, device="cuda")
    ops.nchw_f32_to_nhwc_bf16(src, wide[..., :c])
    assert torch.equal(wide[..., :c], nhwc(src).to(torch.bfloat16))     # bit-exact cast + transpose
    assert float(wide[..., c:].abs().max()) == 0.0
    back = ops.nhwc_bf16_to_nchw_f32(wide[..., :c])
    assert torch.equal(back, src.to(torch.bfloat16).float())


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes (per-GPU batch 64, 512x512): size-independent properties
# ---------------------------------------------------------------------------------------------
def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("cin,cout,hw", [(64, 64, 512), (128, 64, 512), (256, 128, 256)])
def test_full_size_conv_adjoint_identities(ops, cin, cout, hw):
    """At the bench shapes the torch reference conv would take minutes on the host, so the three
    conv kernels are checked against EACH OTHER through the identities that make them one
    operator and its adjoints (every term evaluated by a different kernel):
        <conv(x, W), dy> = <x, dgrad(dy, W)> = <W, wgrad(x, dy)>
    plus linearity of fprop and the fused BatchNorm sums against a plain reduction."""
    n = 64
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.rand(n, hw, hw, cin, generator=g, device="cuda") - 0.3).to(torch.bfloat16)
    wt = rand_w(cout, cin, 3)
    y = torch.empty(n, hw, hw, cout, dtype=torch.bfloat16, device="cuda")
    parts = torch.empty(ops.stat_rows(), 2, cout, dtype=torch.float32, device="cuda")
    ops.conv3x3_fprop(x, ops.repack_fprop(wt, cin), y, stat_partials=parts)
    # dy correlated with y, so the inner products are O(|y|^2) and not a sum of random signs
    dy = torch.randn(n, hw, hw, cout, generator=g, device="cuda").mul_(0.5).add_(y.float()).to(torch.bfloat16)
    dx = torch.empty_like(x)
    ops.conv3x3_dgrad(dy, ops.repack_dgrad(wt), dx)
    dw = torch.empty(cout, cin, 3, 3, device="cuda")
    ws = torch.empty(ops.wgrad_workspace_bytes(n, hw, hw, cin, cout) // 4, device="cuda")
    ops.conv3x3_wgrad(x, dy, dw, ws, cin)
    torch.cuda.synchronize()
    a = _dot(y, dy)          # <conv(x,W), dy>      (y rounded to bf16: ~2^-9 relative, random sign)
    b = _dot(x, dx)          # <x, dgrad(dy,W)>
    c = _dot(wt, dw)         # <W, wgrad(x,dy)>     (fp32 output)
    assert c > 0.3 * float(y.double().norm()) ** 2
    assert abs(a - c) < 1e-3 * c and abs(b - c) < 1e-3 * c, (a, b, c)
    # fused BatchNorm statistics = plain reductions of the stored tensor
    s = parts.double().sum(0)
    yd = y.double()
    assert float((s[0] - yd.sum((0, 1, 2))).abs().max()) < 1e-6 * float(yd.abs().sum((0, 1, 2)).max())
    assert torch.allclose(s[1], (yd * yd).sum((0, 1, 2)), rtol=1e-4)
    # linearity: conv(2x) = 2 conv(x) exactly (power-of-two scaling commutes with every rounding)
    y2 = torch.empty_like(y)
    ops.conv3x3_fprop((x.float() * 2).to(torch.bfloat16), ops.repack_fprop(wt, cin), y2)
    assert torch.equal(y2.float(), y.float() * 2)
    # determinism run to run (split-K wgrad included)
    dw2 = torch.empty_like(dw)
    ops.conv3x3_wgrad(x, dy, dw2, ws, cin)
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("n,hw,cin,cout", [
    (8, 512, 128, 64),      # up4.conv.double_conv.0: the FLOP-dominant Cout = 64 full-resolution layer
    (8, 64, 1024, 512),     # up1.conv.double_conv.0: deepest K (9216), 4 channel blocks, persistent wrap-around
    (16, 256, 64, 128),     # down1 first conv: 8-epilogue-warp variant, one K chunk
    (8, 512, 16, 64),       # first layer, channel-padded input (KCH = 16), MODE_X_STACK wgrad
])
def test_full_size_conv_vs_cudnn_fp32(ops, n, hw, cin, cout):
    """Full-size fprop / dgrad / wgrad (N >= 8 at the bench resolutions: persistent grids wrap around,
    split-K wgrad runs with hundreds of splits) against cuDNN in TRUE fp32 on the same GPU (TF32 off in
    conftest.py) as the checker, on the same bf16 inputs: north_star tolerances 1e-2 / 2e-2."""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.rand(n, hw, hw, cin, generator=g, device="cuda") - 0.3).to(torch.bfloat16)
    wt = rand_w(cout, cin, 12)
    dy = (torch.randn(n, hw, hw, cout, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    y = torch.empty(n, hw, hw, cout, dtype=torch.bfloat16, device="cuda")
    parts = torch.empty(ops.stat_rows(), 2, cout, dtype=torch.float32, device="cuda")
    ops.conv3x3_fprop(x, ops.repack_fprop(wt, cin), y, stat_partials=parts)
    has_dgrad = cin % 64 == 0          # the channel-padded first layer has no data gradient (image input)
    dx = torch.empty_like(x)
    if has_dgrad:
        ops.conv3x3_dgrad(dy, ops.repack_dgrad(wt), dx)
    dw = torch.empty(cout, cin, 3, 3, device="cuda")
    ws = torch.empty(ops.wgrad_workspace_bytes(n, hw, hw, cin, cout) // 4, device="cuda")
    ops.conv3x3_wgrad(x, dy, dw, ws, cin)
    torch.cuda.synchronize()
    assert not torch.backends.cudnn.allow_tf32 and not torch.backends.cuda.matmul.allow_tf32
    xf, dyf = nchw(x.float()), nchw(dy.float())
    del x, dy
    ref = F.conv2d(xf, wt, padding=1)
    e_f = rel(nchw(y.float()), ref)
    s = parts.double().sum(0)
    e_s = rel(s[1], (ref.double() ** 2).sum((0, 2, 3)))
    del ref, y
    e_d = 0.0
    if has_dgrad:
        ref = torch.nn.grad.conv2d_input(tuple(xf.shape), wt, dyf, padding=1)
        e_d = rel(nchw(dx.float()), ref)
        del ref
    del dx
    ref = torch.nn.grad.conv2d_weight(xf, tuple(wt.shape), dyf, padding=1)
    e_w = rel(dw, ref)
    print(f"\nfull-size {cin}->{cout} @{hw}^2 x{n}: fprop {e_f:.2e} (sum sq {e_s:.1e}) dgrad {e_d:.2e} wgrad {e_w:.2e}")
    assert e_f < 1e-2 and e_s < 2e-3, (e_f, e_s)
    assert e_d < 2e-2, e_d
    assert e_w < 2e-2, e_w


def test_full_size_pool_and_elementwise_properties(ops):
    """Batch 64 x 512 x 512 x 64: max-pool of the fused kernel equals pooling its own activation
    output, indices reproduce the pooled values (gather), pool-backward scatters exactly the
    incoming gradient mass."""
    n, hw, c = 64, 512, 64
    g = torch.Generator(device="cuda").manual_seed(9)
    y = torch.randn(n, hw, hw, c, generator=g, device="cuda").to(torch.bfloat16)
    scale = torch.rand(c, generator=g, device="cuda") + 0.5
    shift = torch.randn(c, generator=g, device="cuda") * 0.1
    a = torch.empty_like(y)
    pooled = torch.empty(n, hw // 2, hw // 2, c, dtype=torch.bfloat16, device="cuda")
    idx = torch.empty(n, hw // 2, hw // 2, c, dtype=torch.uint8, device="cuda")
    ops.bn_apply_relu_maxpool2(y, a, pooled, idx, scale, shift)
    win = a.view(n, hw // 2, 2, hw // 2, 2, c).permute(0, 1, 3, 5, 2, 4).reshape(n, hw // 2, hw // 2, c, 4)
    assert torch.equal(pooled, win.max(-1).values)                                   # bit-exact
    assert torch.equal(pooled, win.gather(-1, idx.long().unsqueeze(-1)).squeeze(-1))  # indices point at the max
    assert int(idx.max()) <= 3
    first = (win == pooled.unsqueeze(-1)).float().argmax(-1)                          # first maximum wins
    assert torch.equal(first, idx.long())
    del win, first
    dp = torch.randn(n, hw // 2, hw // 2, c, generator=g, device="cuda").to(torch.bfloat16)
    dxp = torch.empty_like(y)
    ops.maxpool2_bwd(dp, idx, None, dxp)
    assert int((dxp != 0).sum()) <= dp.numel()
    assert torch.equal(dxp.view(n, hw // 2, 2, hw // 2, 2, c).float().sum((2, 4)).to(torch.bfloat16), dp)
