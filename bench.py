#!/usr/bin/env python
"""Headline benchmark: UNet training chips/sec on synthetic 4x512x512 PlanetScope-shaped chips
with masked cross-entropy (BASELINE.json configs[1]; configs[2] when launched on N>1 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # reference algorithm on host CPU cores

One "step" = ingest + UNet forward + masked CE/argmax/confusion + backward + Adam on one batch
of `--batch` chips per GPU.  `value` is measured with the batch already resident in HBM;
`e2e` is the same step driven from pinned HOST buffers through the public LightningModule
API, with the host->device copy of the step's inputs and the device->host read of its loss
inside the timed region.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "unet_train_chips_per_sec"
UNIT = "chips/s"
N_CLASSES = 3
IGNORE_INDEX = 0
LR = 1e-4


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def cpu_reference_chips_per_sec(batch: int, size: int, steps: int, warmup: int, channels: int = 4):
    """The reference algorithm (oracle port of st_water_seg's UNet + CE + Adam, fp32) on the
    host cores: one train step per 'step' on a `batch`-chip sample of the same workload."""
    from oracle import unet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_state_dict(channels, N_CLASSES, seed=0)
    keys = O.trainable_keys(sd)
    params = [sd[k] for k in keys]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=LR)
    b = O.synthetic_batch(batch, channels, size, size, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        logits = O.unet_forward(sd, b["image"], training=True)
        loss, _ = O.masked_ce(logits, b["target"], IGNORE_INDEX)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return batch / min(times), batch / (sum(times) / len(times)), cores, times


def run_reference(args):
    """Reference arm: the reference algorithm (oracle port -- /root/reference does not exist on the GPU
    box and needs pytorch_lightning etc.) on ALL host cores.  `config` is this
    repo's arm's config (the workload both arms are quoted on); each step is a bounded SAMPLE of that
    workload: `--cpu-sample-batch` chips (default 8 = SURVEY.md 8d / BASELINE configs[0]) instead of the
    per-GPU batch of 64 -- chips/s is batch-normalised.  If the first step says the whole run would not
    finish in ~6 minutes the sample is halved (stated in `reference_sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_sample_batch
    budget_s = 360.0
    while True:
        t0 = time.perf_counter()
        cpu_reference_chips_per_sec(sample, args.size, 1, 0, args.channels)
        probe = time.perf_counter() - t0
        if sample <= 1 or probe * (args.steps + args.warmup) <= budget_s:
            break
        sample = max(1, sample // 2)
    best, mean, cores, times = cpu_reference_chips_per_sec(sample, args.size, args.steps, args.warmup,
                                                            args.channels)
    ms = 1000.0 * sum(times) / len(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": mean, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": mean, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle port of the reference UNet+CE+Adam train step, fp32, {sample} chips "
                                   f"of {args.channels}x{args.size}x{args.size} per step, {args.steps} steps after "
                                   f"{args.warmup} warm-up (mean; best {best:.4f})"},
        "e2e": {"value": mean, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # the truth about what one step of THIS arm processed (config above is the shared workload)
        "reference_sample": {"chips_per_step": sample, "requested": args.cpu_sample_batch,
                             "probe_s_per_step": probe, "precision": "fp32", "host_threads": cores,
                             "note": "config.per_gpu_batch is the workload's; this arm steps on "
                                     f"{sample}-chip samples of it (chips/s is batch-normalised)"},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": (f"st_water_seg UNet ({args.channels}->64..512..64->3, 17.27M params) bf16 train step on synthetic "
                     f"{args.channels}x{args.size}x{args.size} chips, masked CE ignore_index=0, Adam lr=1e-4"
                     + ("" if args.channels == 4 else " (early fusion: PlanetScope + stacked extra sensor bands as a "
                                                      "wide-input first conv, BASELINE configs[3])")),
        "per_gpu_batch": args.batch, "global_batch": args.batch * world,
        "chip": [args.channels, args.size, args.size],
        "n_classes": N_CLASSES, "ignore_index": IGNORE_INDEX, "optimizer": "adam",
        "parallelism": f"dp{world}",
        "l2": f"inputs larger than L2: every step streams a fresh "
              f"{args.batch * args.channels * args.size * args.size * 4 / 1e6:.0f} MB image batch and >30 GB of "
              "activations, far beyond the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------
def data_parallel_check(model, reducer, dev, rank, world, args):
    """Before timing, at the launched world size: the all-reduced gradient slab must be (a) bit-identical
    on every rank and (b) the MEAN of the ranks' local gradients.  (b) uses linearity: the checksum of
    the reduced slab equals the mean of the local checksums.  One forward/backward with the hooks off
    (local), one with them on (reduced), same per-rank batch; kernels are deterministic."""
    import torch.distributed as dist
    engine = model.model._engine
    n = max(2, min(8, args.batch))
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    img = torch.rand(n, args.channels, 128, 128, generator=g, device=dev)
    tgt = (torch.rand(n, 128, 128, generator=g, device=dev) > 0.5).long()
    batch = {"image": img, "target": tgt}
    running = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}

    def slab_of():
        for p in model.parameters():
            p.grad = None
        model.training_step(batch, 0).backward()
        return torch.cat([p.grad.detach().double().flatten() for p in model.model.parameters()])

    hooks = (engine.grad_ready_hook, engine.grad_done_hook)
    engine.grad_ready_hook = engine.grad_done_hook = None
    local = slab_of()
    engine.grad_ready_hook, engine.grad_done_hook = hooks
    reduced = slab_of()
    torch.cuda.synchronize(dev)
    w = torch.arange(1, local.numel() + 1, device=dev, dtype=torch.float64).remainder_(97.0).add_(1.0)
    mine = torch.stack([local.sum(), (local * w).sum(), reduced.sum(), (reduced * w).sum(), reduced.abs().sum()])
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    allv = torch.stack(allv).cpu()
    identical = bool((allv[:, 2:] == allv[0, 2:]).all())
    scale = float(allv[0, 4])
    err = max(abs(float(allv[:, 0].mean() - allv[0, 2])), abs(float(allv[:, 1].mean() - allv[0, 3])) / 49.0) / max(scale, 1e-30)
    with torch.no_grad():                      # the check must not leave a trace in the timed model
        sd = model.state_dict()
        for k, v in running.items():
            sd[k].copy_(v)
    for p in model.parameters():
        p.grad = None
    from floodplanet_code_b200.engine import note_raw_parameter_write
    note_raw_parameter_write()
    ok = identical and err < 1e-6
    if not ok:
        raise RuntimeError(f"data-parallel gradient check failed: identical={identical} err={err:.3e}")
    return {"ranks": world, "reduced_slab_identical_on_all_ranks": identical,
            "mean_of_local_checksums_rel_err": err, "buckets": reducer.buckets_last_step}


def infer_block(model, dev, rank, world, args):
    """BASELINE.json configs[4]: sliding-window inference over ONE synthetic scene (default 10240 x 10240 x
    C fp32, 400 tiles of 512, stride = crop as infer.py:64-65), tiles sharded over the launched ranks,
    END TO END: the scene lives in pinned HOST memory, every timed pass copies each rank's row band to the
    device, runs eval-mode UNet + softmax + stitch + argmax/clip on the device and brings the uint8 mask back
    to rank 0's host memory (infer.py:112-184).  Time = max over ranks, CUDA events + host clock."""
    import torch.distributed as dist
    from floodplanet_code_b200.inference import crop_slices, predict_scene_from_host
    from floodplanet_code_b200.parallel import shard_range
    S, crop, C = args.infer_scene, args.size, args.channels
    tiles = crop_slices(S, S, crop, crop, crop)
    mine = [tiles[i] for i in shard_range(len(tiles), rank, world)]
    scene = torch.empty((C, S, S), dtype=torch.float32, pin_memory=True)
    if mine:                                   # only this rank's band is ever read: fill just that
        r0, r1 = min(t[0] for t in mine), min(S, max(t[0] + t[2] for t in mine))
        g = torch.Generator().manual_seed(4321 + rank)
        for ch in range(C):
            scene[ch, r0:r1].copy_(torch.rand(r1 - r0, S, generator=g))
    mask_host = torch.empty((S, S), dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    model._set_model_to_eval()
    unet = model.model
    passes, times, info = 3, [], None
    for it in range(1 + passes):               # 1 untimed warm-up pass
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        info = predict_scene_from_host(unet, scene, crop=crop, tile_batch=args.infer_tile_batch, rank=rank,
                                       world=world, mask_host=mask_host, device=dev)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1000.0)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(float(t.item()))
    model._set_model_to_train()
    _, n_mine, launches, h2d, d2h = info
    counts = torch.tensor([n_mine, launches, h2d], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(counts)
    best = min(times) / 1000.0
    peaks, _ = measured_peaks()
    flops = len(tiles) * 320.210e9 * (crop / 512.0) ** 2 if C == 4 else None
    water = float((mask_host != 0).float().mean()) if rank == 0 else None
    return {"workload": f"sliding-window inference over one synthetic {S}x{S}x{C} fp32 scene in pinned host memory, "
                        f"{len(tiles)} tiles of {crop} (stride = crop), eval-mode BatchNorm, uint8 water mask to rank 0's host",
            "scene_seconds": best, "scene_seconds_all_passes": [x / 1000.0 for x in times],
            "tiles_per_sec": len(tiles) / best, "n_tiles": len(tiles), "tiles_processed_all_ranks": int(counts[0]),
            "tflops": flops / best / 1e12 if flops else None,
            "frac_of_sustained_bf16_peak_per_gpu": (flops / best / 1e12 / world / float(peaks["bf16_tflops_sustained"]))
                                                   if flops else None,
            "h2d_bytes_per_scene": int(counts[2]), "d2h_bytes_per_scene": d2h, "gpu_launches_per_scene": int(counts[1]),
            "tile_batch": args.infer_tile_batch, "water_fraction": water, "n_gpus": world}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    from floodplanet_code_b200 import capi
    from floodplanet_code_b200.optim import FusedAdam
    from floodplanet_code_b200.parallel import (BucketedGradAllReduce, broadcast_parameters,
                                                init_distributed)
    from floodplanet_code_b200.water_seg_model import WaterSegmentationModel
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the B200 path has no CPU fallback "
                           "(use --impl reference for the CPU arm)")
    capi.load()
    rank, world, local_rank = init_distributed()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, S = args.batch, args.size

    torch.manual_seed(0)
    model = WaterSegmentationModel({"ms_image": args.channels}, N_CLASSES, LR, ignore_index=IGNORE_INDEX).to(dev)
    broadcast_parameters(model)
    if args.optimizer == "fused":
        opt = FusedAdam(model.model, lr=LR)
    else:                                   # the reference API's own optimiser object (water_seg_model.py:198-205)
        opt = model.configure_optimizers()
        opt.launches = 0
    # --no-allreduce: N independent replicas (no exchange step) -- NOT data-parallel training, only a probe
    # that separates GPU-to-GPU speed variance (max over ranks) from the cost of the collective
    reducer = (BucketedGradAllReduce(model.model, transport=args.transport, max_ctas=args.nccl_max_ctas,
                                     bucket_bytes=args.bucket_mb << 20)
               if world > 1 and not args.no_allreduce else None)
    dp_check = data_parallel_check(model, reducer, dev, rank, world, args) if reducer is not None else None
    engine = model.model._engine
    if os.environ.get("FPB200_WGRAD_AFTER_DGRAD") in ("0", "1"):   # A/B switch: launch order of dgrad / wgrad of a layer
        engine.wgrad_after_dgrad = os.environ["FPB200_WGRAD_AFTER_DGRAD"] == "1"
    if os.environ.get("FPB200_OVERLAP_WGRAD") in ("0", "1"):   # A/B switch: wgrads on a second stream
        engine.overlap_wgrad = os.environ["FPB200_OVERLAP_WGRAD"] == "1"

    if args.infer_only:
        # BASELINE configs[4] alone (sliding-window inference, tiles sharded over the launched ranks), same JSON shape
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        infer = infer_block(model, dev, rank, world, args)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            print(json.dumps({"metric": "scene_inference_tiles_per_sec", "value": infer["tiles_per_sec"], "unit": "tiles/s",
                              "n_gpus": world, "higher_is_better": True, "scaling": "strong", "dtype": "bf16",
                              "data": "synthetic", "config": {"workload": infer["workload"]}, "clocks": clocks,
                              "e2e": {"value": infer["tiles_per_sec"], "unit": "tiles/s",
                                      "h2d_bytes_per_step": infer["h2d_bytes_per_scene"],
                                      "d2h_bytes_per_step": infer["d2h_bytes_per_scene"]},
                              "gpu_launches": infer["gpu_launches_per_scene"], "infer": infer}), flush=True)
        if world > 1:
            dist.barrier()
            if reducer is not None and reducer.comm is not None:
                reducer.comm.destroy()
            dist.destroy_process_group()
        return

    # synthetic data: U[0,1) image, blocky int64 target ~58% ignored / 42% flood
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = []
    for _ in range(args.pool):
        img = torch.rand(B, args.channels, S, S, generator=g, device=dev)
        coarse = torch.rand(B, 1, S // 32, S // 32, generator=g, device=dev)
        tgt = (torch.nn.functional.interpolate(coarse, size=(S, S), mode="nearest")[:, 0] >= 0.58).long()
        pool.append({"image": img, "target": tgt})
    host_pool = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in pool]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_pool[0].values())

    launches = {"n": 0}
    use_graph = world == 1 and args.graph
    graphed = None
    if use_graph:
        from floodplanet_code_b200.graph import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, pool[0], warmup_steps=max(1, args.warmup))

    def graph_step(batch):
        """Replay of the captured step; the batch is copied into the graph's static inputs
        (device-to-device for `value`; the e2e path copies host->device straight into them)."""
        if batch is not graphed.batch:
            for k, v in graphed.batch.items():
                v.copy_(batch[k], non_blocking=True)
        launches["n"] += graphed.kernel_launches
        return graphed.replay()

    def train_step(batch, i):
        if graphed is not None:
            return graph_step(batch)
        opt.zero_grad()
        loss = model.training_step(batch, i)
        fwd_launches = engine.launches
        loss.backward()
        opt.step()
        # + CE fwd(2)/bwd(1) + where; torch.optim.Adam's own foreach kernels are not ours and not counted
        launches["n"] += fwd_launches + engine.launches + 3 + 1 + opt.launches
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- phase A: device-resident inputs ----------------
    for i in range(args.warmup):
        train_step(pool[i % len(pool)], i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches["n"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    marks = []
    for i in range(args.steps):
        train_step(pool[i % len(pool)], i)
        if world > 1:                       # per-step boundaries (no host sync): step-time jitter across ranks
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    gpu_launches = launches["n"]
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    per_rank_ms = None
    step_jitter = None
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank_ms = [float(x.item()) / args.steps for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-step durations of every rank: what a per-step synchronisation to the slowest rank costs
        # (mean over steps of the max over ranks) against the max over ranks of the per-rank means
        prev, mine = e0, []
        for ev in marks:
            mine.append(prev.elapsed_time(ev))
            prev = ev
        st = torch.tensor(mine, device=dev, dtype=torch.float64)
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        allst = torch.stack(allst).cpu()                       # [rank, step]
        step_jitter = {"mean_over_steps_of_max_over_ranks_ms": float(allst.max(0).values.mean()),
                       "max_over_ranks_of_mean_over_steps_ms": float(allst.mean(1).max()),
                       "per_rank_step_std_ms": [float(x) for x in allst.std(1)],
                       "per_rank_min_step_ms": [float(x) for x in allst.min(1).values],
                       "per_rank_max_step_ms": [float(x) for x in allst.max(1).values]}
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = B * world * args.steps / (elapsed_ms / 1000.0)

    # per-kernel-family roofline: the same K steps are run once more right after the timed region
    # with CUDA events on the launching stream around every conv launch (kept out of the timed
    # region so the headline number carries no instrumentation; with --graph individual launches
    # inside a replay could not be bracketed at all).
    keep = graphed
    graphed = None                          # train_step() launches eagerly
    train_step(pool[0], 0)
    engine.conv_events = []                 # (also switches the wgrad side stream off)
    barrier()
    e0.record()
    for i in range(args.steps):
        train_step(pool[i % len(pool)], i)
    e1.record()
    barrier()
    instrumented_ms_per_step = e0.elapsed_time(e1) / args.steps
    conv_events = engine.conv_events
    engine.conv_events = None
    graphed = keep
    peaks, peak_src = measured_peaks()
    fam = {}
    for tag, layer, flops, a, b_ in conv_events:
        d = fam.setdefault(tag, {"flops": 0.0, "ms": 0.0, "launches": 0})
        d["flops"] += flops
        d["ms"] += a.elapsed_time(b_)
        d["launches"] += 1
    per_layer = {}
    for tag, layer, flops, a, b_ in conv_events:
        d = per_layer.setdefault(f"{tag}:{layer}", {"flops": 0.0, "ms": 0.0})
        d["flops"] += flops
        d["ms"] += a.elapsed_time(b_)
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    roof_all = {}
    for tag, d in fam.items():
        ach = d["flops"] / (d["ms"] / 1000.0) / 1e12 if d["ms"] > 0 else 0.0
        roof_all[tag] = {"achieved": ach, "frac": ach / peak_tf, "ms_per_step": d["ms"] / args.steps,
                         "launches_per_step": d["launches"] / args.steps}
    conv_ms = sum(d["ms"] for d in fam.values())
    conv_flops = sum(d["flops"] for d in fam.values())
    dom = max(fam.items(), key=lambda kv: kv[1]["ms"])[0] if fam else None
    traffic, traffic_src = None, None
    cands = sorted((ROOT / "profiles").glob("r*_traffic.json"))
    tp = cands[-1] if cands else ROOT / "profiles" / "r01_traffic.json"      # the latest round's capture
    if tp.exists() and dom is not None:
        tj = json.loads(tp.read_text())
        key = "conv3x3_wgrad_kernel_all" if dom == "wgrad" else "conv3x3_halo_kernel_all"
        traffic = tj[key]["dram_bytes_per_launch"]
        traffic_src = ("mean dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel over one "
                       f"training step at batch 64 (profiles/{tp.name}, ncu)")
    roofline = None
    if dom is not None:
        kname = {"fprop": "conv3x3_halo_kernel (fprop launches)", "dgrad": "conv3x3_halo_kernel (dgrad launches)",
                 "wgrad": "conv3x3_wgrad_kernel (+split-K reduce)"}[dom]
        roofline = {"bound": "tensor", "kernel": kname, "achieved": roof_all[dom]["achieved"], "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": roof_all[dom]["frac"], "traffic": traffic,
                    "traffic_source": traffic_src,
                    "algorithmic_flops_per_launch": fam[dom]["flops"] / fam[dom]["launches"],
                    "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
                    "share_of_step": roof_all[dom]["ms_per_step"] / instrumented_ms_per_step,
                    "timed_with": ("CUDA events on the launching stream around every conv launch, in an "
                                   f"instrumented pass of the same {args.steps} steps "
                                   f"({instrumented_ms_per_step:.2f} ms/step) run right after the timed region "
                                   f"({ms_per_step:.2f} ms/step)"),
                    "families": roof_all,
                    "all_conv": {"achieved": conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms else 0.0,
                                 "share_of_step": conv_ms / args.steps / instrumented_ms_per_step},
                    "worst_layers": sorted(((k, v["flops"] / (v["ms"] / 1000.0) / 1e12, v["ms"] / args.steps)
                                            for k, v in per_layer.items()), key=lambda x: -x[2])[:8]}

    # ---------------- phase B: end to end from pinned host memory ----------------
    copy_stream = torch.cuda.Stream(device=dev)
    dev_slots = [{k: torch.empty_like(v, device=dev) for k, v in host_pool[0].items()} for _ in range(2)]
    # (graph mode: the step consumes slot -> static-input copies enqueued on the compute stream)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k, v in host_pool[i % len(host_pool)].items():
                dev_slots[slot][k].copy_(v, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(nsteps, base):
        for s in range(2):
            consumed[s].record(torch.cuda.current_stream())
        prefetch(base)
        losses = []
        for i in range(nsteps):
            slot = (base + i) % 2
            if i + 1 < nsteps:
                prefetch(base + i + 1)
            torch.cuda.current_stream().wait_event(ready[slot])
            loss = train_step(dev_slots[slot], i)
            consumed[slot].record(torch.cuda.current_stream())
            losses.append(float(loss.item()))  # device->host read of the step's result
        return losses

    e2e_loop(max(1, min(2, args.warmup)), 0)
    barrier()
    e0.record()
    t0 = time.perf_counter()
    losses = e2e_loop(args.steps, 0)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = B * world * args.steps / (e2e_ms / 1000.0)
    clocks = sampler.stop() if rank == 0 else None

    # per-bucket timeline of the gradient all-reduce (one extra, untimed step): when each bucket became
    # ready on the compute stream, when the collective started / ended on the communication stream
    bucket_timeline = None
    if reducer is not None and reducer.comm is not None:
        reducer.timeline = []
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        train_step(pool[0], 0)
        s1.record()
        barrier()
        bucket_timeline = {"step_ms": s0.elapsed_time(s1),
                           "buckets": [{"mb": (b - a) * 4 / 1e6, "ready_ms": s0.elapsed_time(r), "start_ms": s0.elapsed_time(t0),
                                        "end_ms": s0.elapsed_time(t1)} for a, b, r, t0, t1 in reducer.timeline]}
        reducer.timeline = None

    infer = None
    if not args.no_infer:
        del pool, host_pool, dev_slots
        torch.cuda.empty_cache()
        infer = infer_block(model, dev, rank, world, args)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        best, mean, cores, times = cpu_reference_chips_per_sec(min(args.cpu_sample_batch, 4), S, 2, 1, args.channels)
        cpu_baseline = {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle port of the reference UNet+CE+Adam train step (fp32, torch CPU), "
                                  f"{min(args.cpu_sample_batch, 4)} chips of {args.channels}x{S}x{S}, 1 warm-up + best of 2 steps "
                                  f"({min(times):.2f} s/step)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "cuda_graph": graphed is not None,
            "final_loss": losses[-1] if losses else None,
            "grad_buckets_per_step": reducer.buckets_last_step if reducer else 0,
            "optimizer_impl": "fused Adam kernel over the gradient slab" if args.optimizer == "fused"
                              else "torch.optim.Adam from configure_optimizers()",
            "allreduce": ({"transport": reducer.transport, "max_ctas": args.nccl_max_ctas,
                           "bucket_mb": args.bucket_mb, "timeline": bucket_timeline} if reducer else None),
            "dp_check": dp_check,
            "per_rank_ms_per_step": per_rank_ms,
            "step_jitter": step_jitter,
            "exchange_step": ("none (independent replicas: probe only)" if world > 1 and reducer is None
                              else ("gradient all-reduce" if world > 1 else None)),
            "infer": infer,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if reducer is not None and reducer.comm is not None:
            reducer.comm.destroy()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="chips per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--channels", type=int, default=4,
                    help="input bands: 4 = PlanetScope (headline); 16 = PS + S1 (2) + S2 (10) early fusion, configs[3]")
    ap.add_argument("--pool", type=int, default=2, help="distinct synthetic batches cycled through")
    ap.add_argument("--cpu-sample-batch", type=int, default=8,
                    help="chips per CPU step of the reference arm / cpu_baseline leg (SURVEY 8d: 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"],
                    help="fused: one Adam kernel over the flat slab; torch: stock torch.optim.Adam from the "
                         "LightningModule's configure_optimizers()")
    ap.add_argument("--transport", default=None, choices=["capi", "torch"],
                    help="gradient all-reduce: the C ABI's own ncclComm_t (default) or torch.distributed")
    ap.add_argument("--nccl-max-ctas", type=int, default=int(os.environ.get("FPB200_NCCL_MAX_CTAS", "0")),
                    help="cap of CTAs per NCCL collective for the capi transport (0 = NCCL default)")
    ap.add_argument("--bucket-mb", type=int, default=16)
    ap.add_argument("--no-allreduce", action="store_true",
                    help="probe: N independent replicas without the gradient exchange (not a training configuration)")
    ap.add_argument("--no-infer", action="store_true", help="skip the configs[4] scene-inference block")
    ap.add_argument("--infer-only", action="store_true", help="run ONLY the configs[4] scene-inference block")
    ap.add_argument("--infer-scene", type=int, default=10240, help="scene edge in pixels (configs[4]: 10240)")
    ap.add_argument("--infer-tile-batch", type=int, default=20,
                    help="tiles per forward; a multiple of the tiles per scene row (20) copies no row twice")
    ap.add_argument("--graph", action="store_true",
                    help="replay the whole step from one CUDA graph (measured: no gain at batch 64, where every "
                         "kernel is long enough to hide its launch; useful for small batches)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
