"""One-launch refresh of all packed bf16 weight copies (engine.PackedWeights.refresh_all, the batch
repack kernel) must leave exactly the bytes the per-layer repack kernels produce, and training with the
fused Adam must stay bit-identical to a run that re-packs layer by layer."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _train(steps, use_batch):
    from floodplanet_code_b200.optim import FusedAdam
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    model = WaterSegmentationModel({"ms_image": 4}, 3, 1e-3, ignore_index=0)
    model.model.load_state_dict(O.init_state_dict(4, 3, seed=0), strict=True)
    model = model.cuda()
    opt = FusedAdam(model.model, lr=1e-3)
    packed = model.model._engine.packed
    if not use_batch:
        packed.refresh_all = lambda: 0          # fall back to lazy per-layer repacks at the next lookup
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 64, seed=1, block=8).items()}
    losses = []
    for i in range(steps):
        opt.zero_grad()
        loss = model.training_step(batch, i)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    nbt = model.model.inc.double_conv[1].num_batches_tracked.item()
    return losses, {k: v.detach().clone() for k, v in model.model.state_dict().items()}, packed, opt, nbt


def test_batch_repack_equals_per_layer_repack():
    la, sa, pa, oa, nbt_a = _train(3, True)
    lb, sb, pb, ob, nbt_b = _train(3, False)
    assert la == lb and nbt_a == nbt_b == 3
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert oa.launches == 2 and ob.launches == 1
    # after the last step the batch path has already refreshed every copy: compare with fresh packs
    from floodplanet_code_b200 import ops
    for (kind, name), (w, cin_pad) in pa._meta.items():
        ref = ops.repack_fprop(w, cin_pad) if kind == 0 else ops.repack_dgrad(w)
        got = (pa._fprop if kind == 0 else pa._dgrad)[name][1]
        assert torch.equal(ref, got), (kind, name)
    assert len(pa._meta) == 18 + 17          # 18 fprop copies, 17 dgrad copies (no dgrad into the image)


def test_eval_fold_cache_tracks_training_updates():
    """The cached eval-mode BatchNorm fold must be refreshed after anything that rewrites parameters or
    running statistics through raw pointers (training-mode BatchNorm kernels, fused Adam)."""
    from floodplanet_code_b200.optim import FusedAdam
    from floodplanet_code_b200.unet import UNet
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    from oracle import unet_oracle as O
    model = WaterSegmentationModel({"ms_image": 4}, 3, 1e-2, ignore_index=0)
    model.model.load_state_dict(O.init_state_dict(4, 3, seed=0), strict=True)
    model = model.cuda()
    opt = FusedAdam(model.model, lr=1e-2)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 64, seed=1, block=8).items()}

    def eval_logits(m):
        m.eval()
        with torch.no_grad():
            return m(batch["image"]).clone()

    e0 = eval_logits(model.model)
    launches_first = model.model.kernel_launches
    assert torch.equal(e0, eval_logits(model.model))                    # cache hit: same coefficients
    assert model.model.kernel_launches == launches_first - 18          # 18 fold kernels skipped
    for i in range(2):
        opt.zero_grad()
        model.training_step(batch, i).backward()
        opt.step()
    e1 = eval_logits(model.model)
    assert not torch.equal(e0, e1)
    fresh = UNet(4, 3)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.model.state_dict().items()}, strict=True)
    assert torch.equal(e1, eval_logits(fresh.cuda()))


def test_graph_replay_invalidates_eval_caches():
    """a CUDA-graph replay updates masters, BatchNorm affine parameters and running
    statistics through raw pointers; an eval forward after replays must see them (packed weights
    and the folded BatchNorm coefficients are rebuilt), also after MORE replays."""
    from floodplanet_code_b200.graph import GraphedTrainStep
    from floodplanet_code_b200.optim import FusedAdam
    from floodplanet_code_b200.unet import UNet
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    from oracle import unet_oracle as O
    model = WaterSegmentationModel({"ms_image": 4}, 3, 1e-2, ignore_index=0)
    model.model.load_state_dict(O.init_state_dict(4, 3, seed=0), strict=True)
    model = model.cuda()
    opt = FusedAdam(model.model, lr=1e-2)
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 64, seed=1, block=8).items()}

    def eval_logits(m):
        m.eval()
        with torch.no_grad():
            return m(batch["image"]).clone()

    def fresh_eval():
        fresh = UNet(4, 3)
        fresh.load_state_dict({k: v.detach().cpu() for k, v in model.model.state_dict().items()}, strict=True)
        return eval_logits(fresh.cuda())

    step = GraphedTrainStep(model, opt, batch, warmup_steps=3)
    e_first = eval_logits(model.model)            # fills the fold cache and the packed-weight tables
    assert torch.equal(e_first, fresh_eval())
    for _ in range(2):
        step.replay()
    e1 = eval_logits(model.model)
    assert not torch.equal(e1, e_first)
    assert torch.equal(e1, fresh_eval())
    for _ in range(2):
        step.replay()
    e2 = eval_logits(model.model)
    assert not torch.equal(e2, e1)
    assert torch.equal(e2, fresh_eval())


def test_raw_data_writes_and_fused_adam_reach_every_engine():
    """`p.data` writes (what broadcast_parameters does) announced with
    note_raw_parameter_write(), and FusedAdam steps, must invalidate the caches of ALL engines over
    the same parameters -- UNet.forward, UNet.encode and UNet.decode own separate engines."""
    from floodplanet_code_b200.engine import note_raw_parameter_write
    from floodplanet_code_b200.optim import FusedAdam
    from floodplanet_code_b200.unet import UNet
    from oracle import unet_oracle as O
    from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
    net = UNet(4, 3)
    net.load_state_dict(O.init_state_dict(4, 3, seed=0), strict=True)
    net = net.cuda()
    batch = {k: v.cuda() for k, v in O.synthetic_batch(2, 4, 64, 64, seed=1, block=8).items()}

    def all_paths(m):
        m.eval()
        with torch.no_grad():
            full = m(batch["image"]).clone()
            feats = [f.clone() for f in m.encode(batch["image"])]
            dec = m.decode(feats).clone()
        return full, feats, dec

    def fresh_paths():
        fresh = UNet(4, 3)
        fresh.load_state_dict({k: v.detach().cpu() for k, v in net.state_dict().items()}, strict=True)
        return all_paths(fresh.cuda())

    def same(a, b):
        return torch.equal(a[0], b[0]) and all(torch.equal(x, y) for x, y in zip(a[1], b[1])) and torch.equal(a[2], b[2])

    first = all_paths(net)                          # every engine now holds caches
    assert same(first, fresh_paths())
    other = O.init_state_dict(4, 3, seed=5)
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.data.copy_(other[k].cuda())           # no version counter moves
    note_raw_parameter_write()
    second = all_paths(net)
    assert not torch.equal(second[0], first[0])
    assert same(second, fresh_paths())
    # fused Adam through the main engine; the encode / decode engines must follow
    opt = FusedAdam(net, lr=1e-2)
    net.train()
    loss = MaskedCrossEntropyLoss(0)(net(batch["image"]), batch["target"])
    loss.backward()
    opt.step()
    third = all_paths(net)
    assert not torch.equal(third[0], second[0])
    assert same(third, fresh_paths())


def test_input_gradient_request_raises():
    from floodplanet_code_b200.unet import UNet
    net = UNet(4, 3).cuda().train()
    x = torch.rand(1, 4, 32, 32, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError, match="input images"):
        net(x)
