#!/usr/bin/env python
"""bench.py -- voxels/sec, fwd+bwd, SparseConvUNet m=32 / 2 cm voxels (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config NAME]

A step = InputLayer (GPU hash voxelisation + every rulebook, fresh coordinates each step) -> encoder forward ->
OutputLayer -> scalar loss (mean of the per-point outputs) -> full backward (input-feature grad + every weight grad)
-> gradient all-reduce when N > 1 -> fused Adam step (train.py:39,80-81).  voxels = active level-0 sites (N_0)
summed over the batch and over ranks.  One JSON line on stdout (rank 0).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "3d-weakly-supervised-semantic-segmentation_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# site counts differ a little every step (fresh coordinates); rounding torch's own allocations keeps block sizes repeating
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:8")
import torch  # noqa: E402

METRIC = "voxels/sec fwd+bwd SparseConvUNet m=32 2cm"
DEFAULT_CONFIG = "cfg3_unet_m32_r2_res_s50_b5"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons during the timed region, read in-process through NVML (the same counters as the
    nvidia-smi clocks line in B200_PROFILING.md).  Every NVML query takes the driver lock and stalls kernel launches
    (measured: 6 ms/step with a background thread sampling every 0.1 s, and occasional ~0.2 s stalls of a single step even
    at 0.4 s), so there is no sampling thread: the main thread takes a sample every few steps inside the timed loop, right
    after the step's loss read-back, when no launch is in flight (the SM clock is still the clock under load)."""

    def __init__(self, index):
        self.index, self.rows = index, []
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.sample()          # first-call costs (library initialisation) stay outside the timed region
            self.rows.clear()
        except Exception:
            self.h = None

    def sample(self):
        if self.h is None:
            return
        try:
            nv = self.nv
            t0 = time.perf_counter()
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.rows.append((sm, rs, time.perf_counter() - t0))
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        if self.h is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        nv = self.nv
        sm = sorted(r[0] for r in self.rows)
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, b in bits.items() if any(r[1] & b for r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sm),
                "nvml_ms_per_sample": 1e3 * sum(r[2] for r in self.rows) / len(self.rows)}


def _make_inputs(cfg, rank, n_distinct, n_points):
    from b200scn_synth import CONFIGS, make_batch
    kind, m, reps, res, scale, batch = CONFIGS[cfg]
    seeds = [rank * batch + i for i in range(batch)]
    return [make_batch(seeds, scale, n_points=n_points, step=s) for s in range(n_distinct)]


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """CPU restatement of sparseconvnet 0.2 (the oracle port; the real package is not installable) on all
    host cores.  Each step = fwd+bwd on a bounded sample of the workload: ONE scene of the config's batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from b200scn_synth import CONFIGS, build_encoder
    from oracle import scn_oracle as ref
    kind, m, reps, res, scale, batch = CONFIGS[args.config]
    torch.manual_seed(0)
    net = build_encoder(ref, kind, m, reps, res)
    from b200scn_synth import make_batch
    data = [make_batch([0], scale, n_points=args.points, step=s) for s in range(min(args.steps + args.warmup, 4))]
    times = []
    nvox = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        coords, feats, _ = data[i % len(data)]
        f = feats.clone().requires_grad_(True)
        x = net[0]([coords, f])
        n0 = x.features.shape[0]
        y = x
        for mod in list(net)[1:]:
            y = mod(y)
        y.mean().backward()
        for p in net.parameters():
            p.grad = None
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            nvox.append(n0)
    total = sum(times)
    value = sum(nvox) / total
    cores = torch.get_num_threads()
    sample = "1 scene of %s (%d pts, %d voxels) per step, fwd+bwd incl. rulebooks" % (args.config, args.points, nvox[0])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.config, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(cfg, n_points, budget_s=25.0):
    """Oracle port timed on the host cores: 1 scene of the workload, fwd+bwd, as many steps as fit the budget (>=1)."""
    from b200scn_synth import CONFIGS, build_encoder, make_batch
    from oracle import scn_oracle as ref
    kind, m, reps, res, scale, batch = CONFIGS[cfg]
    torch.manual_seed(0)
    net = build_encoder(ref, kind, m, reps, res)
    coords, feats, _ = make_batch([0], scale, n_points=n_points)
    t_all, nvox, steps = 0.0, 0, 0
    while steps < 1 or (t_all < budget_s and steps < 3):
        t0 = time.perf_counter()
        f = feats.clone().requires_grad_(True)
        x = net[0]([coords, f])
        y = x
        for mod in list(net)[1:]:
            y = mod(y)
        y.mean().backward()
        for p in net.parameters():
            p.grad = None
        dt = time.perf_counter() - t0
        if steps > 0 or dt > budget_s / 2:   # first step is the warm-up unless it already ate the budget
            t_all += dt
            nvox += x.features.shape[0]
        steps += 1
        if dt > budget_s:
            break
    return {"value": nvox / t_all, "unit": "voxels/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 scene of %s (%d pts, %d voxels/step), fwd+bwd incl. rulebooks, %.1f s of CPU work" % (
                cfg, n_points, x.features.shape[0], t_all)}


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    import sparseconvnet as scn
    from b200scn_dp import FlatGrads
    from b200scn_synth import CONFIGS, build_encoder
    from sparseconvnet import ops as scn_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kind, m, reps, res, scale, batch = CONFIGS[args.config]
    scn.set_precision(args.precision)
    torch.manual_seed(0)  # identical initial weights on every rank
    net = build_encoder(scn, kind, m, reps, res).cuda()
    flat = FlatGrads(net.parameters())
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    # every distinct batch is seen during warm-up, so the caching allocator has met every size before timing starts
    n_distinct = max(1, min(args.warmup, args.distinct))
    host = _make_inputs(args.config, rank, n_distinct, args.points)
    pinned = [(c.pin_memory(), f.pin_memory()) for c, f, _ in host]
    resident = [(c.to(dev), f.to(dev)) for c, f in pinned]
    stats = {"voxels": 0}

    debug = os.environ.get("B200SCN_BENCH_DEBUG") == "1"

    def step(i, from_host):
        t_start = time.perf_counter()
        if from_host:
            c, f = pinned[i % n_distinct]
            coords = c.to(dev, non_blocking=True)
            feats = f.to(dev, non_blocking=True)
        else:
            coords, feats = resident[i % n_distinct]
        feats = feats.detach().requires_grad_(True)
        flat.zero()
        t_in0 = time.perf_counter()
        x = net[0]([coords, feats])
        t_in1 = time.perf_counter()
        stats["voxels"] += x.features.shape[0]
        y = x
        for mod in list(net)[1:]:
            y = mod(y)
        loss = y.mean()
        loss.backward()
        if world > 1:
            flat.allreduce_mean()
        opt.step()
        if debug:
            t_end = time.perf_counter()
            ms = torch.cuda.memory_stats()
            sys.stderr.write("step %d host ms: total %.1f  input-layer (incl. its sync) %.1f  rest %.1f | cudaMalloc calls %d reserved %.2f GB allocated %.2f peak %.2f GB\n" % (
                i, 1e3 * (t_end - t_start), 1e3 * (t_in1 - t_in0), 1e3 * (t_end - t_in1), ms["num_device_alloc"], ms["reserved_bytes.all.current"] / 1e9,
                ms["allocated_bytes.all.current"] / 1e9, ms["allocated_bytes.all.peak"] / 1e9))
        return loss

    def timed(from_host, prof, sampler=None):
        for i in range(args.warmup):
            step(i, from_host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        stats["voxels"] = 0
        # no cyclic-GC pass inside the timed steps (a generation-2 sweep over a step's object graph costs ~0.2 s of host time)
        gc.collect()
        gc.disable()
        launches0 = scn.launch_count()
        if prof:
            scn_ops.profile_reserve(800 * args.steps)
            scn_ops.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(args.warmup + i, from_host)
            # D2H read of the step's result, every step in both arms: a training loop logs its loss, and without it the
            # host runs a step ahead, two steps' activations are alive at once and the caching allocator occasionally
            # grows (cudaMalloc + implicit sync, ~0.1-0.2 s) inside the timed region
            loss_host = loss.item()  # noqa: F841
            if sampler is not None and (i % 6 == 3 or (args.steps <= 3 and i == args.steps - 1)):
                sampler.sample()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gc.enable()
        ms = e0.elapsed_time(e1)
        prof_out = scn_ops.profile_end() if prof else None
        t = torch.tensor([ms, float(stats["voxels"])], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, vox = float(tmax[0]), float(t[1])
        else:
            vox = float(t[1])
        return ms, vox, scn.launch_count() - launches0, prof_out

    # allocator priming (untimed, before any warm-up): site counts differ per batch, so torch's caching allocator needs to
    # have met every distinct batch before its block pool stops growing (cudaMalloc inside a step synchronises): whole
    # cycles over the distinct batches (at least four) until a cycle passes without a new cudaMalloc, at most 8 cycles.
    # (expandable_segments was tried and made step times erratic: 46-106 ms.)  (The NVML handle
    # is opened here too: its first queries take the driver lock for ~0.2 s, which must not land in a timed step.)
    with ClockSampler(local) as clk:
        mallocs = -1
        for cycle in range(8):
            for i in range(n_distinct):
                step(i, False)
                step(i, True)
            torch.cuda.synchronize()
            now = torch.cuda.memory_stats()["num_device_alloc"]
            if cycle >= 3 and now == mallocs:
                break
            mallocs = now
        # one discarded rehearsal of the timed loop: whatever still grows on first use of this exact call sequence (caching
        # allocator blocks: observed as a 0.1-0.25 s stall of one step in the first timed loop of ~1 run in 3) happens here
        timed(False, False)
        ms, vox, launches, _ = timed(False, False, clk)         # headline: device-resident inputs, nothing but the step
    clocks = clk.summary()
    ms_e, vox_e, _, _ = timed(True, False)                 # end to end: pinned host inputs, H2D inside, loss read back
    # roofline pass: the same K steps again with CUDA events around every conv / BN launch (kept out of the headline
    # timing because recording ~700 event pairs per step costs a few ms of host time)
    ms_p, _, _, prof = timed(False, True)

    if rank == 0:
        peak, which = _peaks()
        h2d = sum(c.numel() * 8 + f.numel() * 4 for c, f in pinned) / len(pinned)
        roof = None
        # the dominant kernel: the tiled submanifold kernel ("tiled27") wherever it runs, else the gather kernel
        dom = max((k for k in ("tiled27", "gather27") if prof and prof.get(k)), key=lambda k: prof[k]["ms"], default=None)
        if dom:
            g = prof[dom]
            ach = g["bytes"] / (g["ms"] * 1e-3) / 1e9
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel from the committed ncu --set full
            # capture (profiles/r1_halo_ncu_details.txt): the level-1 64->64 SubmanifoldConvolution forward of this
            # workload (tiled kernel), 135.2 MB read + 64.5 MB written against 241.8 MB algorithmic -- no re-read beyond the
            # compulsory traffic (part of the input is still L2-resident from the producing kernel).
            traffic = {"bytes": 199.6e6, "launch": "level-1 SubM 64->64 fwd, 411829 sites", "algorithmic_bytes": 241.8e6}
            roof = {"bound": "hbm", "kernel": g["kernel"], "achieved": ach, "peak": peak, "peak_source": which, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic["bytes"], "traffic_detail": traffic, "launches": g["n"], "avg_launch_us": 1e3 * g["ms"] / g["n"],
                    "algorithmic_bytes_per_launch": g["bytes"] / g["n"], "tflops": g["flops"] / (g["ms"] * 1e-3) / 1e12,
                    "share_of_step": g["ms"] / ms_p, "profiled_ms_per_step": ms_p / args.steps, "by_kind": {k: {"ms_per_step": v["ms"] / args.steps, "n_per_step": v["n"] / args.steps,
                                                                 "GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else None,
                                                                 "TFLOPs": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] else None}
                                                             for k, v in prof.items()}}
        line = {
            "metric": METRIC, "value": vox / (ms * 1e-3), "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": args.config, "encoder": "%s m=%d block_reps=%d residual=%s" % (kind, m, reps, res),
                       "scale": scale, "batch_per_gpu": batch, "points_per_scene": args.points,
                       "voxels_per_step_per_gpu": vox / args.steps / world, "parallelism": "dp%d" % world,
                       "step": "InputLayer(hash+rulebooks, fresh coords)+fwd+loss+bwd(dI,dW)+allreduce+fused Adam",
                       "l2": "inputs larger than L2: >1 GB of activations per step, fresh coordinates each step"},
            "clocks": clocks,
            "e2e": {"value": vox_e / (ms_e * 1e-3), "unit": "voxels/s", "ms_per_step": ms_e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(args.config, args.points)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--points", type=int, default=150000)
    ap.add_argument("--distinct", type=int, default=6, help="distinct pre-generated batches cycled through")
    ap.add_argument("--precision", default=os.environ.get("B200SCN_PRECISION", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
