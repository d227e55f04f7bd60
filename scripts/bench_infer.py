"""Sliding-window inference over one synthetic 10240x10240 4-band scene (BASELINE.json configs[4]):
tiles/s and scene seconds, tiles sharded across the launched ranks.
    python scripts/bench_infer.py [--size 10240] [--stride 512]      (torchrun for N > 1)"""
import argparse, json, sys, time
import torch
sys.path.insert(0, ".")
from floodplanet_code_b200.inference import predict_scene, crop_slices
from floodplanet_code_b200.parallel import init_distributed, broadcast_parameters
from floodplanet_code_b200.unet import UNet
import torch.distributed as dist
ap = argparse.ArgumentParser(); ap.add_argument("--size", type=int, default=10240); ap.add_argument("--stride", type=int, default=512)
ap.add_argument("--tile-batch", type=int, default=40); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
rank, world, lr = init_distributed(); torch.cuda.set_device(lr)
torch.manual_seed(0)
m = UNet(4, 3).cuda(); broadcast_parameters(m); m.eval()
scene = torch.rand(4, a.size, a.size, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
n_tiles = len(crop_slices(a.size, a.size, 512, 512, a.stride))
best = 1e9
for r in range(a.reps + 1):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    mask, mine, launches = predict_scene(m, scene, 512, a.stride, a.tile_batch, rank, world)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    if r > 0: best = min(best, dt)
if rank == 0:
    flops = n_tiles * 320.21e9
    print(json.dumps({"metric": "unet_infer_tiles_per_sec", "value": n_tiles / best, "unit": "tiles/s", "n_gpus": world,
                      "scene": [4, a.size, a.size], "tiles": n_tiles, "stride": a.stride, "scene_seconds": best,
                      "tflops": flops / best / 1e12, "water_fraction": float((mask > 0).float().mean()), "kernel_launches": launches}))
if world > 1: dist.destroy_process_group()
