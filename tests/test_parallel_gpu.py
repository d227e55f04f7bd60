"""2-GPU data-parallel parity (needs 2 CUDA devices; skipped otherwise): the bucketed NCCL
all-reduce inside backward leaves every rank with the average of the per-shard gradients,
which equals the single-process gradients of each shard computed separately (per-GPU
BatchNorm statistics, torch-DDP default semantics) averaged on the host."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, transport):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from floodplanet_code_b200.parallel import BucketedGradAllReduce, broadcast_parameters, init_distributed
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    from oracle import unet_oracle as O
    init_distributed("nccl")
    torch.cuda.set_device(rank)
    m = WaterSegmentationModel({"ms_image": 4}, 3, 1e-4, ignore_index=0)
    m.model.load_state_dict(O.init_state_dict(4, 3, seed=0))
    m = m.cuda()
    broadcast_parameters(m)
    red = BucketedGradAllReduce(m.model, bucket_bytes=1 << 20, transport=transport, max_ctas=8 if transport == "capi" else 0)
    assert red.transport == transport and (red.comm is not None) == (transport == "capi")
    b = O.synthetic_batch(2, 4, 64, 64, seed=10 + rank, block=8, device="cuda")
    loss = m.training_step(b, 0)
    loss.backward()
    torch.cuda.synchronize()
    g = {k: p.grad.detach().cpu() for k, p in m.model.named_parameters()}
    q.put((rank, g, red.buckets_last_step))
    dist.barrier()
    if red.comm is not None:
        red.comm.destroy()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["capi", "torch"])
def test_two_gpu_allreduce_equals_mean_of_shard_gradients(transport):
    """transport 'capi': the C ABI's own ncclComm_t (fpb200_nccl_comm_create / fpb200_allreduce_f32, ncclAvg)
    on the module's communication stream; 'torch': torch.distributed all_reduce(AVG)."""
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    from oracle import unet_oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, transport)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    res = {r: g for r, g, _ in got}
    assert all(nb >= 4 for _, _, nb in got), "the 69 MB slab should leave in several buckets during backward"
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # both ranks hold identical, averaged gradients
    for k in res[0]:
        assert torch.equal(res[0][k], res[1][k]), k
    # single-process shard gradients, averaged
    shard = []
    for r in range(world):
        m = WaterSegmentationModel({"ms_image": 4}, 3, 1e-4, ignore_index=0)
        m.model.load_state_dict(O.init_state_dict(4, 3, seed=0))
        m = m.cuda()
        b = O.synthetic_batch(2, 4, 64, 64, seed=10 + r, block=8, device="cuda")
        m.training_step(b, 0).backward()
        shard.append({k: p.grad.detach().cpu() for k, p in m.model.named_parameters()})
    for k in res[0]:
        want = (shard[0][k] + shard[1][k]) / 2
        assert torch.allclose(res[0][k], want, rtol=1e-5, atol=1e-9), k
