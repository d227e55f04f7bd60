"""ORACLE -- test infrastructure only, never a product path.

A plain-torch (CPU, fp32 arithmetic, bf16 storage) re-enactment of the B200 UNet SCHEDULE that emits
the same trace records as `engine._Schedule.trace`, so that the teacher-forced walker
(oracle/teacher_forced.py) can itself be tested without a GPU: it must accept a correct trace and
REJECT traces with a planted fault (a BatchNorm-backward coefficient off by 5 %, swapped concat
halves, a missing skip-gradient add, ...).  That is what makes the walker a test that can fail.
Follows st_water_seg/models/unet.py:6-111 and the autograd of it.
"""
from types import SimpleNamespace

import torch
import torch.nn.functional as F

BN_EPS, BN_MOM = 1e-5, 0.1
ENC_CH = (64, 128, 256, 512)


def _bf(t):
    return t.to(torch.bfloat16)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def emulate_traced_step(m, specs, batch, ignore_index, cin_pad, fault=None, frozen=False):
    """m: a (CPU) module with the reference's parameter / buffer names (the product's container
    classes work); specs: engine.unet_conv_specs(n_channels).  Sets `.grad` on every parameter, updates
    the BatchNorm buffers, returns (logits, loss, trace).  `fault`: name of a planted bug.  `frozen`: eval-mode
    BatchNorm with autograd (running statistics as constants, buffers untouched, conv biases get gradients)."""
    P = dict(m.named_parameters())
    B = dict(m.named_buffers())
    trace = []
    image, target = batch["image"], batch["target"]
    n, c_in, H, W = image.shape
    x = torch.zeros((n, H, W, cin_pad), dtype=torch.bfloat16)
    x[..., :c_in] = _bf(_nhwc(image))
    grads = {k: torch.zeros_like(v) for k, v in P.items()}
    for k, p in P.items():
        p.grad = grads[k]

    saved = []

    def layer(s, xin, out_view=None, pool=False, defer=False):
        w = _bf(P[f"{s.conv}.weight"].detach()).float()
        y = _bf(_nhwc(F.conv2d(_nchw(xin[..., :s.cin].float()), w, None, padding=1)))
        yf = y.float()
        gamma, beta = P[f"{s.bn}.weight"].detach(), P[f"{s.bn}.bias"].detach()
        if frozen:
            cb = P[f"{s.conv}.bias"].detach()
            mean = B[f"{s.bn}.running_mean"].detach() - (0.0 if fault == "frozen_bias_dropped" else cb)
            invstd = torch.rsqrt(B[f"{s.bn}.running_var"].detach() + BN_EPS)
        else:
            mean = yf.mean((0, 1, 2))
            var = yf.var((0, 1, 2), unbiased=False)
            cnt = yf.numel() / yf.shape[3]
            invstd = torch.rsqrt(var + BN_EPS)
            with torch.no_grad():
                B[f"{s.bn}.running_mean"].mul_(1 - BN_MOM).add_(BN_MOM * (mean + P[f"{s.conv}.bias"].detach()))
                B[f"{s.bn}.running_var"].mul_(1 - BN_MOM).add_(BN_MOM * var * cnt / max(cnt - 1, 1))
                B[f"{s.bn}.num_batches_tracked"] += 1
        scale = gamma * invstd
        shift = beta - mean * scale
        a = pooled = idx = None
        if not defer:
            a_val = _bf(F.relu(yf * scale + shift))
            if out_view is not None:
                out_view.copy_(a_val)
                a = out_view
            else:
                a = a_val
            if pool:
                pv, pi = F.max_pool2d(_nchw(a.float()), 2, return_indices=True)
                hh, ww = a.shape[1], a.shape[2]
                ih, iw = pi // ww, pi % ww
                code = (ih % 2) * 2 + (iw % 2)
                pooled, idx = _bf(_nhwc(pv)), _nhwc(code).to(torch.uint8)
        rec = dict(op="conv_bn_relu", spec=s, x=xin, y=y, scale=scale, shift=shift, mean=mean, invstd=invstd,
                   a=a, pooled=pooled, pool_idx=idx)
        trace.append(rec)
        saved.append(rec)
        return rec

    sizes = [(H, W)]
    for _ in range(4):
        sizes.append((sizes[-1][0] // 2, sizes[-1][1] // 2))
    cat = {l: torch.zeros((n, sizes[l][0], sizes[l][1], 2 * ENC_CH[l]), dtype=torch.bfloat16) for l in range(4)}
    cur, li = x, 0
    for lvl in range(4):
        cur = layer(specs[li], cur)["a"]; li += 1
        cur = layer(specs[li], cur, out_view=cat[lvl][..., :ENC_CH[lvl]], pool=True)["pooled"]; li += 1
    cur = layer(specs[li], cur)["a"]; li += 1
    cur = layer(specs[li], cur)["a"]; li += 1
    last = None
    for lvl in (3, 2, 1, 0):
        c = ENC_CH[lvl]
        up = F.interpolate(_nchw(cur.float()), scale_factor=2, mode="bilinear", align_corners=True)
        dy_, dx_ = cat[lvl].shape[1] - up.shape[2], cat[lvl].shape[2] - up.shape[3]
        up = F.pad(up, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
        if fault == "concat_halves_swapped":
            skip = cat[lvl][..., :c].clone()
            cat[lvl][..., :c] = _bf(_nhwc(up))
            cat[lvl][..., c:] = skip
        else:
            cat[lvl][..., c:] = _bf(_nhwc(up))
        trace.append(dict(op="upsample_concat", level=lvl, x=cur, cat=cat[lvl], c=c))
        cur = layer(specs[li], cat[lvl])["a"]; li += 1
        if lvl == 0:
            last = layer(specs[li], cur, defer=True); li += 1
        else:
            cur = layer(specs[li], cur)["a"]; li += 1
    a_last = F.relu(last["y"].float() * last["scale"] + last["shift"])
    wh, bh = P["outc.conv.weight"].detach(), P["outc.conv.bias"].detach()
    logits = F.conv2d(_nchw(a_last), wh, bh)
    trace.append(dict(op="head", x=last["y"], scale=last["scale"], shift=last["shift"], logits=logits, prefix=""))

    # ---------------------------------------------------------------- backward
    lg = logits.clone().requires_grad_(True)
    loss = F.cross_entropy(lg, target, ignore_index=ignore_index)
    loss.backward()
    dlogits = lg.grad
    d_act = _bf(_nhwc(torch.nn.grad.conv2d_input(_nchw(a_last).shape, wh, dlogits)))
    grads["outc.conv.weight"].copy_(torch.nn.grad.conv2d_weight(_nchw(a_last), wh.shape, dlogits))
    grads["outc.conv.bias"].copy_(dlogits.sum((0, 2, 3)))
    trace.append(dict(op="head_bwd", dlogits=dlogits, x=last["y"], bn=None, d_act=d_act,
                      dw=grads["outc.conv.weight"], db=grads["outc.conv.bias"], prefix=""))

    def layer_bwd(i, da, need_dx=True, dx_out=None):
        r, s = saved[i], specs[i]
        yv = _nchw(r["y"].float()).requires_grad_(True)
        g_ = P[f"{s.bn}.weight"].detach().clone().requires_grad_(True)
        b_ = P[f"{s.bn}.bias"].detach().clone().requires_grad_(True)
        if frozen:
            cb_ = P[f"{s.conv}.bias"].detach().clone().requires_grad_(True)
            a = F.relu(F.batch_norm(yv + cb_[None, :, None, None], B[f"{s.bn}.running_mean"].detach().clone(),
                                    B[f"{s.bn}.running_var"].detach().clone(), g_, b_, False, BN_MOM, BN_EPS))
            a.backward(_nchw(da.float()))
            grads[f"{s.conv}.bias"].copy_(cb_.grad)
        else:
            a = F.relu(F.batch_norm(yv, None, None, g_, b_, True, BN_MOM, BN_EPS))
            a.backward(_nchw(da.float()))
        dyv = yv.grad
        if fault is not None and fault.startswith(f"bn_bwd_coef_layer{i}:"):
            off = float(fault.split(":")[1])
            # dy = s*g - P*xhat - Q with the coefficient P = scale * mean(g*xhat) off by `off` (relative)
            xhat = (yv.detach() - r["mean"][None, :, None, None]) * r["invstd"][None, :, None, None]
            gm = (_nchw(da.float()) * (a.detach() > 0)) * xhat
            pcoef = r["scale"] * gm.mean((0, 2, 3))
            dyv = dyv - off * pcoef[None, :, None, None] * xhat
        dy = _bf(_nhwc(dyv))
        grads[f"{s.bn}.weight"].copy_(g_.grad)
        grads[f"{s.bn}.bias"].copy_(b_.grad)
        xin = _nchw(r["x"][..., :s.cin].float())
        dyf = _nchw(dy.float())
        wt = P[f"{s.conv}.weight"].detach()
        grads[f"{s.conv}.weight"].copy_(torch.nn.grad.conv2d_weight(xin, wt.shape, dyf, padding=1))
        dx = None
        if need_dx:
            dxv = _bf(_nhwc(torch.nn.grad.conv2d_input(xin.shape, _bf(wt).float(), dyf, padding=1)))
            if dx_out is not None:
                dx_out.copy_(dxv)
                dx = dx_out
            else:
                dx = dxv
        trace.append(dict(op="layer_bwd", spec=s, da=da, y=r["y"], x=r["x"], dy=dy, dx=dx, coef=None,
                          dw=grads[f"{s.conv}.weight"], dgamma=grads[f"{s.bn}.weight"],
                          dbeta=grads[f"{s.bn}.bias"], fused_reduce=False))
        return dx

    li = 17
    d_cur = d_act
    dcat = {}
    for lvl in (0, 1, 2, 3):
        c = ENC_CH[lvl]
        d_mid = layer_bwd(li, d_cur); li -= 1
        dcat[lvl] = torch.zeros_like(cat[lvl])
        layer_bwd(li, d_mid, dx_out=dcat[lvl]); li -= 1
        hl, wl = sizes[lvl + 1]
        xin = torch.zeros((n, c, hl, wl), requires_grad=True)
        up = F.interpolate(xin, scale_factor=2, mode="bilinear", align_corners=True)
        dy_, dx_ = dcat[lvl].shape[1] - up.shape[2], dcat[lvl].shape[2] - up.shape[3]
        up = F.pad(up, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
        half = dcat[lvl][..., :c] if fault == "upsample_bwd_reads_skip_half" else dcat[lvl][..., c:]
        up.backward(_nchw(half.float()))
        d_cur = _bf(_nhwc(xin.grad))
        trace.append(dict(op="upsample_concat_bwd", level=lvl, dcat=dcat[lvl], c=c, dx=d_cur))
    d_mid = layer_bwd(li, d_cur); li -= 1
    d_pool = layer_bwd(li, d_mid); li -= 1
    for lvl in (3, 2, 1, 0):
        c = ENC_CH[lvl]
        a_skip = _nchw(cat[lvl][..., :c].float())
        _, i_ref = F.max_pool2d(a_skip, 2, return_indices=True)
        d = F.max_unpool2d(_nchw(d_pool.float()), i_ref, 2, output_size=a_skip.shape[2:])
        d_skip_in = dcat[lvl][..., :c]
        if fault != "skip_gradient_dropped":
            d = d + _nchw(d_skip_in.float())
        d_act_l = _bf(_nhwc(d))
        trace.append(dict(op="maxpool_bwd", level=lvl, d_pooled=d_pool, pool_idx=saved[li]["pool_idx"],
                          d_skip_in=d_skip_in, d_act=d_act_l))
        d_mid = layer_bwd(li, d_act_l); li -= 1
        d_pool = layer_bwd(li, d_mid, need_dx=lvl > 0); li -= 1
    assert li == -1
    return logits.detach(), loss.detach(), trace
