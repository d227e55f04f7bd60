"""CPU tests of the multi-GPU host logic with world_size 2 over gloo: bucketed gradient
all-reduce hooks (order, bucket boundaries, averaging), parameter broadcast and the tile /
sample sharding used by data-parallel training and tile-sharded inference."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from floodplanet_code_b200.parallel import (BucketedGradAllReduce, broadcast_parameters,
                                                init_distributed, shard_range)
    from floodplanet_code_b200.unet import UNet
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    from floodplanet_code_b200 import engine as E
    torch.manual_seed(rank)            # different init per rank ...
    net = UNet(4, 3)
    epoch0 = E._RawWriteEpochs.param
    broadcast_parameters(net)          # ... identical after the broadcast
    # the broadcast writes through .data (no version counter moves), so it must announce itself:
    # caches made by a forward BEFORE the broadcast would otherwise survive it
    assert E._RawWriteEpochs.param > epoch0 and net._engine.packed.generation == E._RawWriteEpochs.param
    chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum()
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    same = all(torch.equal(g, gathered[0]) for g in gathered)

    engine = net._engine
    red = BucketedGradAllReduce(net, bucket_bytes=4 << 20)
    params = dict(net.named_parameters())
    layout, total = engine.grad_layout(params)
    slab = torch.full((total,), float(rank + 1))
    # replay the engine's backward notification sequence: grads become final in slab order
    names_rev = list(reversed(engine.names))
    ready = 0
    for i, name in enumerate(names_rev):
        if name.endswith("weight") and (".0." in name or ".3." in name or name.startswith("outc")):
            off, nel = layout[name]
            end = (off + nel + 3) // 4 * 4
            engine.grad_ready_hook(slab, ready, end)
            ready = end
    engine.grad_ready_hook(slab, ready, total)
    engine.grad_done_hook(slab, total)
    ok_avg = bool(torch.allclose(slab, torch.full_like(slab, (1 + world) / 2)))
    q.put((rank, same, ok_avg, red.buckets_last_step, list(shard_range(400, rank, world)),
           list(shard_range(7, rank, world))))
    dist.destroy_process_group()


def test_bucketed_allreduce_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tiles = []
    for rank, same, ok_avg, buckets, tile_ids, small in results:
        assert same, "broadcast_parameters left replicas different"
        assert ok_avg, "bucketed all-reduce did not average the whole slab"
        assert 2 <= buckets <= 40          # 69 MB of grads in >= 4 MB buckets, launched as ready
        tiles += tile_ids
    assert tiles == list(range(400))       # 400 scene tiles: contiguous, disjoint, complete
    assert results[0][5] == [0, 1, 2, 3] and results[1][5] == [4, 5, 6]


def test_grad_layout_is_reverse_forward_and_aligned():
    from floodplanet_code_b200.unet import UNet
    net = UNet(4, 3)
    layout, total = net._engine.grad_layout(dict(net.named_parameters()))
    assert layout["outc.conv.bias"][0] == 0                      # produced first in backward
    assert layout["inc.double_conv.0.weight"][0] + 64 * 4 * 9 == total
    assert all(off % 4 == 0 for off, _ in layout.values())       # 16-byte aligned slices
    assert total >= 17268099
