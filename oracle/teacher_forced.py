"""ORACLE -- test infrastructure only, never a product path.

Teacher-forced walk of a traced training step of the B200 UNet: every step of the real wired schedule
(18 conv3x3+BN+ReLU layers, 4 max-pools, 4 upsample+pad+concat stages, the 1x1 head, masked CE; forward
and backward) is re-computed by the fp32 torch op the reference dispatches to
(st_water_seg/models/unet.py:6-111, water_seg_model.py:40,98-107) FROM THE CUDA PATH'S OWN STORED INPUT of
that step and compared with the step's own output, and every step's input is checked (bitwise) to be the
tensor the reference graph feeds it.  Used by tests/test_teacher_forced_gpu.py and
__graft_entry__.smoke(); imports nothing from the product -- it only reads the trace records
(engine._Schedule.trace) it is handed.
"""
import torch
import torch.nn.functional as F

FWD_TOL = 1e-2        # north_star: logits / loss within 1e-2 relative of the fp32 reference
GRAD_TOL = 2e-2       # north_star: gradients within 2e-2 relative
BN_EPS, BN_MOM = 1e-5, 0.1


def rel(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def f32c(t):
    """NHWC bf16 view -> fp32 NCHW tensor (the layout/dtype the reference op sees)."""
    return nchw(t.float())


def decode_pool_idx(idx_nhwc, h, w):
    """1-byte window argmax (0..3, row-major in the 2x2 window) -> torch's flat input index."""
    ii = nchw(idx_nhwc).long()
    hp, wp = ii.shape[2], ii.shape[3]
    hh = torch.arange(hp, device=ii.device)[None, None, :, None] * 2 + ii // 2
    ww = torch.arange(wp, device=ii.device)[None, None, None, :] * 2 + ii % 2
    return hh * w + ww


def check_forward(m, sd0, batch, logits, trace, report, frozen=False):
    """Walk the forward records in the order of unet.py:100-111."""
    params = {k: v.detach() for k, v in m.named_parameters()}
    fwd = [r for r in trace if r["op"] in ("conv_bn_relu", "upsample_concat", "head")]
    assert [r["op"] for r in fwd] == ["conv_bn_relu"] * 10 + ["upsample_concat", "conv_bn_relu", "conv_bn_relu"] * 4 + ["head"]
    image = batch["image"]
    n, c_in, H, W = image.shape
    cur = None                 # the tensor the reference graph feeds to the next op (NHWC bf16)
    skips = {}
    li = 0
    by_layer = {}
    for r in fwd:
        if r["op"] == "conv_bn_relu":
            s = r["spec"]
            x, y = r["x"], r["y"]
            hh, ww = x.shape[1], x.shape[2]
            if li == 0:
                # ingest: NCHW f32 -> NHWC bf16, channels zero-padded (bit-exact)
                want = image.permute(0, 2, 3, 1).to(torch.bfloat16)
                assert torch.equal(x[..., :c_in], want), "ingest"
                assert float(x[..., c_in:].float().abs().max()) == 0.0 if x.shape[3] > c_in else True
            else:
                assert x.shape == cur.shape and torch.equal(x, cur), f"wiring: input of layer {li} ({s.conv})"
            assert s.idx == li
            wt, bias = params[f"{s.conv}.weight"], params[f"{s.conv}.bias"]
            gamma, beta = params[f"{s.bn}.weight"], params[f"{s.bn}.bias"]
            # conv: the reference op on the kernel's own input, fp32 master weights
            y_ref = F.conv2d(f32c(x[..., :s.cin]), wt, None, padding=1)
            e = rel(f32c(y), y_ref)
            report.append((f"fwd conv {s.conv}", e))
            assert e < FWD_TOL, (s.conv, e)
            # BatchNorm statistics + normalise + ReLU on the kernel's own (stored) conv output
            yf = f32c(y)
            bufs = dict(m.named_buffers())
            if frozen:
                # eval-mode BatchNorm: running statistics are constants and must be untouched by the step
                rm0, rv0 = sd0[f"{s.bn}.running_mean"], sd0[f"{s.bn}.running_var"]
                assert torch.equal(bufs[f"{s.bn}.running_mean"], rm0) and torch.equal(bufs[f"{s.bn}.running_var"], rv0)
                assert int(bufs[f"{s.bn}.num_batches_tracked"]) == int(sd0[f"{s.bn}.num_batches_tracked"])
                a_ref = F.relu(F.batch_norm(yf + bias[None, :, None, None], rm0.clone(), rv0.clone(), gamma, beta,
                                            False, BN_MOM, BN_EPS))
                assert rel(r["mean"], rm0 - bias) < 1e-5 or float((r["mean"] - (rm0 - bias)).abs().max()) < 1e-6, s.bn
                assert rel(r["invstd"], torch.rsqrt(rv0 + BN_EPS)) < 1e-5, s.bn
            else:
                rm = torch.zeros(s.cout, device=y.device)
                rv = torch.ones(s.cout, device=y.device)
                a_ref = F.relu(F.batch_norm(yf + bias[None, :, None, None], rm, rv, gamma, beta, True, BN_MOM, BN_EPS))
                mean_ref = yf.mean((0, 2, 3))
                var_ref = yf.var((0, 2, 3), unbiased=False)
                assert rel(r["mean"], mean_ref) < 1e-3 or float((r["mean"] - mean_ref).abs().max()) < 1e-5, s.bn
                assert rel(r["invstd"], torch.rsqrt(var_ref + BN_EPS)) < 1e-3, s.bn
                # running statistics after ONE step from the default (0, 1) buffers; conv bias enters the mean
                assert sd0[f"{s.bn}.running_mean"].abs().max() == 0 and (sd0[f"{s.bn}.running_var"] == 1).all()
                assert rel(bufs[f"{s.bn}.running_mean"], rm) < 1e-3 or float((bufs[f"{s.bn}.running_mean"] - rm).abs().max()) < 1e-6, s.bn
                assert rel(bufs[f"{s.bn}.running_var"], rv) < 1e-3, s.bn
                assert int(bufs[f"{s.bn}.num_batches_tracked"]) == 1
            if r["a"] is not None:
                e = rel(f32c(r["a"]), a_ref)
                report.append((f"fwd bn+relu {s.bn}", e))
                assert e < FWD_TOL, (s.bn, e)
                cur = r["a"]
            else:
                cur = None          # last layer: the head consumes the raw output
                last_raw = (y, a_ref)
            if r["pooled"] is not None:
                lvl = s.level
                p_ref, i_ref = F.max_pool2d(f32c(r["a"]), 2, return_indices=True)
                assert torch.equal(f32c(r["pooled"]), p_ref), f"max-pool values level {lvl}"
                assert torch.equal(decode_pool_idx(r["pool_idx"], hh, ww), i_ref), f"max-pool argmax level {lvl}"
                skips[lvl] = r["a"]
                cur = r["pooled"]
            by_layer[li] = r
            li += 1
        elif r["op"] == "upsample_concat":
            lvl, c, cat = r["level"], r["c"], r["cat"]
            assert torch.equal(r["x"], cur), f"wiring: upsample input level {lvl}"
            # torch.cat([x2 (skip), x1 (upsampled)], dim=1), unet.py:66: skip in the LOWER half
            assert torch.equal(cat[..., :c], skips[lvl]), f"wiring: skip half of the concat, level {lvl}"
            x1 = F.interpolate(f32c(r["x"]), scale_factor=2, mode="bilinear", align_corners=True)
            dy_, dx_ = cat.shape[1] - x1.shape[2], cat.shape[2] - x1.shape[3]
            x1 = F.pad(x1, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
            e = rel(f32c(cat[..., c:]), x1)
            report.append((f"fwd upsample level {lvl}", e))
            assert e < FWD_TOL, (lvl, e)
            cur = cat
        else:  # head
            y_last, a_last_ref = last_raw
            assert torch.equal(r["x"], y_last), "wiring: head input"
            lg_ref = F.conv2d(a_last_ref, params["outc.conv.weight"], params["outc.conv.bias"])
            assert r["logits"].data_ptr() == logits.data_ptr() or torch.equal(r["logits"], logits)
            e = rel(logits, lg_ref)
            report.append(("fwd head", e))
            assert e < FWD_TOL, e
    assert li == 18
    return by_layer, skips


def check_backward(m, batch, logits, loss, trace, fwd_layers, skips, ignore_index, report, frozen=False):
    params = {k: v.detach() for k, v in m.named_parameters()}
    grads = {k: v.grad.detach() for k, v in m.named_parameters()}
    bwd = [r for r in trace if r["op"] in ("head_bwd", "layer_bwd", "upsample_concat_bwd", "maxpool_bwd")]
    want_ops = (["head_bwd"] + ["layer_bwd", "layer_bwd", "upsample_concat_bwd"] * 4 + ["layer_bwd", "layer_bwd"]
                + ["maxpool_bwd", "layer_bwd", "layer_bwd"] * 4)
    assert [r["op"] for r in bwd] == want_ops
    # ---- loss + dlogits (water_seg_model.py:40,103): fp32 on both sides, same logits
    lg = logits.detach().clone().requires_grad_(True)
    loss_ref = F.cross_entropy(lg, batch["target"], ignore_index=ignore_index)
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref.detach())) <= 1e-4 * abs(float(loss_ref.detach())) + 1e-7
    head = bwd[0]
    e = rel(head["dlogits"], lg.grad)
    report.append(("bwd CE dlogits", e))
    assert e < 1e-4, e
    # ---- head backward
    last = fwd_layers[17]
    s17 = last["spec"]
    gamma, beta = params[f"{s17.bn}.weight"], params[f"{s17.bn}.bias"]
    bufs = dict(m.named_buffers())
    if frozen:
        a_last = F.relu(F.batch_norm(f32c(last["y"]) + params[f"{s17.conv}.bias"][None, :, None, None],
                                     bufs[f"{s17.bn}.running_mean"].clone(), bufs[f"{s17.bn}.running_var"].clone(),
                                     gamma, beta, False, BN_MOM, BN_EPS))
    else:
        a_last = F.relu(F.batch_norm(f32c(last["y"]), None, None, gamma, beta, True, BN_MOM, BN_EPS))
    wh = params["outc.conv.weight"]
    dl = head["dlogits"]
    d_act_ref = torch.nn.grad.conv2d_input(a_last.shape, wh, dl)
    dwh_ref = torch.nn.grad.conv2d_weight(a_last, wh.shape, dl)
    for name, got, ref in (("d_act", f32c(head["d_act"]), d_act_ref), ("dW", head["dw"].view_as(wh), dwh_ref),
                           ("db", head["db"], dl.sum((0, 2, 3)))):
        e = rel(got, ref)
        report.append((f"bwd head {name}", e))
        assert e < GRAD_TOL, (name, e)
    # what the walk compares IS what the optimiser sees (autograd may have copied the slab views)
    assert torch.equal(head["dw"], grads["outc.conv.weight"]) and torch.equal(head["db"], grads["outc.conv.bias"])

    state = {"da": head["d_act"], "dcat": {}, "d_pool": None}
    order = [17, 16, "up0", 15, 14, "up1", 13, 12, "up2", 11, 10, "up3", 9, 8, "pool3", 7, 6, "pool2", 5, 4,
             "pool1", 3, 2, "pool0", 1, 0]
    assert len(order) == len(bwd) - 1
    for tag, r in zip(order, bwd[1:]):
        if isinstance(tag, int):
            assert r["op"] == "layer_bwd"
            s = r["spec"]
            fr = fwd_layers[tag]
            assert s.idx == tag and r["y"].data_ptr() == fr["y"].data_ptr() and r["x"].data_ptr() == fr["x"].data_ptr()
            da = r["da"]
            assert torch.equal(da, state["da"]), f"wiring: activation gradient into layer {tag} ({s.conv})"
            gamma, beta = params[f"{s.bn}.weight"], params[f"{s.bn}.bias"]
            # BatchNorm(train) + ReLU backward through torch autograd at the kernel's own y and da
            yv = f32c(r["y"]).requires_grad_(True)
            g_, b_ = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
            cb_ = params[f"{s.conv}.bias"].clone().requires_grad_(True)
            if frozen:
                a = F.relu(F.batch_norm(yv + cb_[None, :, None, None], bufs[f"{s.bn}.running_mean"].clone(),
                                        bufs[f"{s.bn}.running_var"].clone(), g_, b_, False, BN_MOM, BN_EPS))
            else:
                a = F.relu(F.batch_norm(yv, None, None, g_, b_, True, BN_MOM, BN_EPS))
            a.backward(f32c(da))
            for name, got, ref in (("dy", f32c(r["dy"]), yv.grad), ("dgamma", r["dgamma"], g_.grad),
                                   ("dbeta", r["dbeta"], b_.grad)):
                e = rel(got, ref)
                report.append((f"bwd {s.bn} {name}{' (fused reduce)' if r['fused_reduce'] else ''}", e))
                assert e < GRAD_TOL, (s.bn, name, e)
            assert torch.equal(r["dgamma"], grads[f"{s.bn}.weight"]) and torch.equal(r["dbeta"], grads[f"{s.bn}.bias"])
            # conv weight / input gradients at the kernel's own x and dy, fp32 master weights
            wt = params[f"{s.conv}.weight"]
            xin = f32c(r["x"][..., :s.cin])
            dyf = f32c(r["dy"])
            dw_ref = torch.nn.grad.conv2d_weight(xin, wt.shape, dyf, padding=1)
            e = rel(r["dw"], dw_ref)
            report.append((f"bwd {s.conv} dW", e))
            assert e < GRAD_TOL, (s.conv, "dW", e)
            assert torch.equal(r["dw"], grads[f"{s.conv}.weight"])
            if frozen:
                # eval-mode BatchNorm does not cancel the conv bias: d bias = sum over pixels of dy
                e = rel(grads[f"{s.conv}.bias"], cb_.grad)
                report.append((f"bwd {s.conv} dbias", e))
                assert e < GRAD_TOL, (s.conv, "dbias", e)
            else:
                # conv bias feeding a training-mode BatchNorm: exactly cancelled (reference: rounding noise)
                assert float(grads[f"{s.conv}.bias"].abs().max()) == 0.0
            if tag == 0:
                assert r["dx"] is None          # no gradient into the image
            else:
                dx_ref = torch.nn.grad.conv2d_input(xin.shape, wt, dyf, padding=1)
                e = rel(f32c(r["dx"]), dx_ref)
                report.append((f"bwd {s.conv} dx", e))
                assert e < GRAD_TOL, (s.conv, "dx", e)
                if not s.second and tag >= 10:
                    state["dcat"][s.level] = r["dx"]      # first conv of an Up stage: gradient of the concat
                    assert r["dx"].shape[3] == 2 * (64, 128, 256, 512)[s.level]
                elif not s.second:
                    state["d_pool"] = r["dx"]             # first conv of a Down stage: gradient of the pooled map
                state["da"] = r["dx"]
        elif tag.startswith("up"):
            lvl = int(tag[2:])
            assert r["op"] == "upsample_concat_bwd" and r["level"] == lvl
            c = r["c"]
            dcat = state["dcat"][lvl]
            assert torch.equal(r["dcat"], dcat), f"wiring: concat gradient level {lvl}"
            hl, wl = r["dx"].shape[1], r["dx"].shape[2]
            xin = torch.zeros((dcat.shape[0], c, hl, wl), device=dcat.device, requires_grad=True)
            up = F.interpolate(xin, scale_factor=2, mode="bilinear", align_corners=True)
            dy_, dx_ = dcat.shape[1] - up.shape[2], dcat.shape[2] - up.shape[3]
            up = F.pad(up, [dx_ // 2, dx_ - dx_ // 2, dy_ // 2, dy_ - dy_ // 2])
            up.backward(f32c(dcat[..., c:]))            # UPPER half = the upsampled branch (unet.py:66)
            e = rel(f32c(r["dx"]), xin.grad)
            report.append((f"bwd upsample level {lvl}", e))
            assert e < GRAD_TOL, (lvl, e)
            state["da"] = r["dx"]
        else:
            lvl = int(tag[4:])
            assert r["op"] == "maxpool_bwd" and r["level"] == lvl
            c = (64, 128, 256, 512)[lvl]
            assert torch.equal(r["d_pooled"], state["d_pool"]), f"wiring: pooled gradient level {lvl}"
            # the skip connection's gradient = LOWER half of the concat gradient of the same level
            assert torch.equal(r["d_skip_in"], state["dcat"][lvl][..., :c]), f"wiring: skip gradient level {lvl}"
            a_skip = f32c(skips[lvl])
            _, i_ref = F.max_pool2d(a_skip, 2, return_indices=True)
            ref = F.max_unpool2d(f32c(r["d_pooled"]), i_ref, 2, output_size=a_skip.shape[2:]) + f32c(r["d_skip_in"])
            e = rel(f32c(r["d_act"]), ref)
            report.append((f"bwd maxpool+skip level {lvl}", e))
            assert e < GRAD_TOL, (lvl, e)
            state["da"] = r["d_act"]




def walk(m, sd0, batch, logits, loss, trace, ignore_index, fwd_tol=None, grad_tol=None, frozen=False):
    """Full forward + backward walk.  Returns the list of (what, relative error) comparisons made;
    raises AssertionError at the first step outside tolerance or mis-wired.  Tolerances default to
    the north_star's (1e-2 forward, 2e-2 gradients); callers may tighten them towards the bf16
    rounding floor (~1.7e-3 per stored tensor)."""
    global FWD_TOL, GRAD_TOL
    saved = (FWD_TOL, GRAD_TOL)
    FWD_TOL = saved[0] if fwd_tol is None else fwd_tol
    GRAD_TOL = saved[1] if grad_tol is None else grad_tol
    try:
        return _walk(m, sd0, batch, logits, loss, trace, ignore_index, frozen)
    finally:
        FWD_TOL, GRAD_TOL = saved


def _walk(m, sd0, batch, logits, loss, trace, ignore_index, frozen=False):
    """frozen: the step ran with the module in eval() and grad mode on -- BatchNorm uses the running statistics as
    constants (F.batch_norm(training=False)), its buffers must not change, and the conv biases get real gradients."""
    report = []
    fwd_layers, skips = check_forward(m, sd0, batch, logits, trace, report, frozen)
    check_backward(m, batch, logits, loss, trace, fwd_layers, skips, ignore_index, report, frozen)
    asserted = [k for k, _ in report if k.endswith(" dW") or "dgamma" in k or "dbeta" in k or k == "bwd head db"]
    assert len(asserted) == 18 * 3 + 2, "every trainable parameter's gradient must have been compared"
    return report
