#!/usr/bin/env python
"""Benchmark of the 3D-WSIS scene-level hot path (BASELINE.json: "scenes/sec fwd (ScanNet-shape) ...").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path -- voxelization -> sparse-conv UNet -> superpoint pooling -> ECC -> edge
affinity -- over one batch of `--scenes` synthetic ScanNet-shaped scenes (BASELINE.json configs[1]: the ScanNet
config in inference, batch of 4 scenes of ~150k points, 2 cm voxels, ~3k superpoints each), random-init weights
of the reference architecture under torch.manual_seed(123).  One JSON line is printed by rank 0.

  value     scenes/s with the batch already resident in HBM, device time (CUDA events), max over ranks
  e2e       the same through the public API from pinned HOST buffers: H2D of the step's inputs and D2H of the
            step's results are inside the timed region
  roofline  the sparse-conv kernels of the step: algorithmic bytes (e(N_in*Cin + N_out*Cout) +
            2*idx*P + e*K*Cin*Cout per layer) / their summed CUDA-event durations, against the measured HBM peak
  cpu_baseline  the same hot path on the host cores through the reference's own CPU kernels (oracle/_ref) on a
            bounded sample (one scene), rank 0 at N=1 only
Multi-GPU: scenes shard across ranks with no data-path collective (weak scaling: every rank runs its own batches).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "3d-wsis_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

# one growing segment per stream instead of cudaMalloc/cudaFree churn when consecutive batches differ in size (every
# scene has its own voxel count, so every step allocates slightly different tensors)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "scenes/sec fwd (ScanNet-shape)"
UNIT = "scenes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=4, help="scenes per batch (configs[1]: 4)")
    ap.add_argument("--points", type=int, default=150000, help="points per scene")
    ap.add_argument("--shape", default="scannet", choices=["scannet", "s3dis"],
                    help="scannet: configs[1] (8x6x2.6 m rooms, 2 cm voxels); s3dis: configs[3] (20x15x3 m rooms, 5 cm "
                         "voxels; use --points 1000000 --scenes 1)")
    ap.add_argument("--precision", default=os.environ.get("WSIS_PRECISION", "fp32"), choices=["fp32", "bf16", "simt"])
    ap.add_argument("--cpu-sample-scenes", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, extra=None):
    if args.shape == "s3dis":
        what = ("S3DIS_Area5_3D_WSIS inference, batch of %d synthetic S3DIS-shaped rooms (%d pts/room, 20x15x3 m, 5 cm "
                "voxels, ~10k superpoints/room), UNet+pooling+ECC+affinity forward" % (args.scenes, args.points))
    else:
        what = ("ScanNet_v2_3D_WSIS inference, batch of %d synthetic ScanNet-shaped scenes "
                "(%d pts/scene, 2 cm voxels, ~3k superpoints/scene), UNet+pooling+ECC+affinity forward"
                % (args.scenes, args.points))
    cfg = {"workload": what,
           "scenes_per_step": args.scenes, "points_per_scene": args.points, "weights": "random-init, seed 123",
           "l2": "flushed between timed steps (256 MiB write)"}
    cfg.update(extra or {})
    return cfg


def make_batches(args, rank, n_batches):
    from wsis_b200 import synthetic
    out = []
    for b in range(n_batches):
        mk = synthetic.make_room_s3dis if args.shape == "s3dis" else synthetic.make_scene
        scenes = [mk(2000 + 100 * rank + 10 * b + i, n_points=args.points) for i in range(args.scenes)]
        out.append(synthetic.collate(scenes))
    return out


# ---------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region (B200_PROFILING.md "clocks line")
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_ready(self, timeout=20.0):
        """nvidia-smi start-up (NVML initialisation) takes the driver lock and stalled concurrent launches for up to
        0.5 s when it overlapped a timed step: the sampler is started before the warm-up and must be streaming before
        any timing begins; it is terminated only after every timed region."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.05)

    def terminate(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, t_begin, t_end):
        """Clocks / throttle reasons of the samples taken inside [t_begin, t_end] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for t, ln in self.lines if t_begin <= t <= t_end + 0.1]
        for ln in inside or [ln for _, ln in self.lines[-3:]]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------------------------
def cpu_pass(args, n_steps, warmup):
    """The reference's CPU path on the host cores, bounded sample = `cpu_sample_scenes` scene(s) per step."""
    from oracle import cpu_pipeline
    from wsis_b200 import pipeline, synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = pipeline.build_network(seed=123, device="cpu").eval()
    mk = synthetic.make_room_s3dis if args.shape == "s3dis" else synthetic.make_scene
    scenes = [mk(2000 + i, n_points=args.points) for i in range(args.cpu_sample_scenes)]
    batch = synthetic.collate(scenes)
    times, stages, kind = [], None, "port"
    for it in range(warmup + n_steps):
        t0 = time.perf_counter()
        _, stages, kind = cpu_pipeline.forward(net, batch)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": args.cpu_sample_scenes / best, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d scene(s) of %d pts, full hot path, best of %d after %d warm-up; threads=%d; stages(s)=%s"
                      % (args.cpu_sample_scenes, args.points, n_steps, warmup, cores,
                         {k: round(v, 3) for k, v in stages.items()}),
            "ms_per_step": 1e3 * best}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = cpu_pass(args, max(1, min(args.steps, 3)), min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"] , "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, {"note": "reference CPU path on host cores; each step is a bounded sample "
                                                      "of %d scene(s) of the workload" % args.cpu_sample_scenes}),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def conv_roofline(net, dbatch, pipeline, W, steps, peaks):
    """Times every sparse-conv launch of the forward with CUDA events on the launching stream and accumulates the
    algorithmic bytes of SURVEY.md 8(d)."""
    rec = []
    orig = W.sparse_conv
    pair_cache = {}

    def timed(src, weight3, map_, n_dst, flip, *args, **kw):
        key = map_.data_ptr()
        if key not in pair_cache:
            pair_cache[key] = int((map_ >= 0).sum().item())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = orig(src, weight3, map_, n_dst, flip, *args, **kw)
        e.record()
        K, cin, cout = weight3.shape
        has_res = kw.get("residual") is not None or (len(args) > 2 and args[2] is not None)
        rec.append((s, e, 4 * (src.shape[0] * cin + n_dst * cout) + 8 * pair_cache[key] + 4 * K * cin * cout
                    + (4 * n_dst * cout if has_res else 0),
                    2.0 * pair_cache[key] * cin * cout, (cin, cout, K, n_dst)))
        return out

    import spconv.ops as sops
    W.sparse_conv = timed
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    try:
        per_step = []
        for _ in range(steps):
            flush.zero_()
            rec.clear()
            with torch.no_grad():
                pipeline.forward_batch(net, dbatch)
            torch.cuda.synchronize()
            per_step.append([(s.elapsed_time(e) * 1e-3, b, f, shp) for s, e, b, f, shp in rec])
    finally:
        W.sparse_conv = orig
    last = per_step[-1]
    nl = len(last)
    tsum = statistics.mean(sum(t for t, _, _, _ in st) for st in per_step)
    bsum = sum(b for _, b, _, _ in last)
    fsum = sum(f for _, _, f, _ in last)
    # dominant kernel = the layer shape (Cin, Cout, K, rows) whose launches take the largest share of the step
    groups = {}
    for i, (_, b, f, shp) in enumerate(last):
        g = groups.setdefault(shp, {"idx": [], "bytes": b, "flops": f})
        g["idx"].append(i)
    for g in groups.values():
        g["t"] = statistics.mean(sum(st[i][0] for i in g["idx"]) for st in per_step)
    shp, g = max(groups.items(), key=lambda kv: kv[1]["t"])
    t_launch = g["t"] / len(g["idx"])
    ach = g["bytes"] / t_launch / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r01_ncu_conv_umma_v3.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        if list(pj.get("cin_cout_K_rows", [])) == list(shp):  # same layer shape, same rows: per-launch DRAM bytes
            traffic = pj.get("dram_bytes_per_launch")
    ach_all = bsum / tsum / 1e9
    return {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic,
            "kernel": "conv_umma_kernel, layer shape (Cin, Cout, K, rows) = %s: %d launches per step, %.1f%% of the "
                      "step's sparse-conv time" % (list(shp), len(g["idx"]), 100.0 * g["t"] / tsum),
            "algorithmic_bytes_per_launch": g["bytes"], "us_per_launch": round(t_launch * 1e6, 1),
            "useful_tflops": round(g["flops"] / t_launch / 1e12, 1),
            "all_sparse_conv": {"launches_per_step": nl, "algorithmic_bytes_per_step": bsum,
                                "ms_per_step": round(tsum * 1e3, 3), "GBps": round(ach_all, 1),
                                "frac": round(ach_all / peaks["hbm_gbs"], 4),
                                "useful_tflops": round(fsum / tsum / 1e12, 1)},
            "peak_source": peaks["source"]}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")),
                "source": "MEASURED_PEAKS.json (measured)"}
    # the driver-written file is git-ignored and can be absent in a re-created container; its round-1 values are
    # recorded in SURVEY.md 8(d) (hbm_gbs 6551.4, bf16_tflops_sustained 1386.7)
    return {"hbm_gbs": 6551.4, "bf16_tflops": 1386.7,
            "source": "MEASURED_PEAKS.json absent: its values as recorded in SURVEY.md 8(d) (measured on this pool)"}


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from wsis_b200 import ops as W
    from wsis_b200 import pipeline
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    W.set_precision(args.precision)
    peaks = load_peaks()
    clocks = ClockSampler(local_rank)
    clocks.start()
    net = pipeline.build_network(seed=123, device="cuda").eval()
    n_batches = 2
    host = [pipeline.pin_batch(b) for b in make_batches(args, rank, n_batches)]
    dev = [pipeline.to_device(b)[0] for b in host]
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # Set-up, not measurement: map the allocator's working set once (every step allocates ~1 GB of rulebook / record /
    # feature temporaries whose sizes differ from batch to batch; growing the pool inside a timed step showed up as a
    # single 0.5 s outlier) and run every distinct batch through both paths before the W official warm-up steps.
    pool = torch.empty(6 << 30, dtype=torch.uint8, device="cuda")
    del pool
    for b in range(n_batches):
        with torch.no_grad():
            pipeline.forward_batch(net, dev[b])
            pipeline.forward_batch(net, pipeline.to_device(host[b])[0])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        with torch.no_grad():
            ret, _ = pipeline.forward_batch(net, dev[i % n_batches])
        return ret

    def step_e2e(i):
        db, nb = pipeline.to_device(host[i % n_batches])
        with torch.no_grad():
            ret, _ = pipeline.forward_batch(net, db)
        outs = [ret[k].to("cpu", non_blocking=False) for k in ("edge_affinity", "sp_semantic_scores",
                                                               "sp_discriminative_feats", "pred_sp_offset_vectors")]
        return nb, sum(o.numel() * o.element_size() for o in outs)

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = W.launch_count()
        evs, extra = [], None
        for i in range(steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            extra = fn(warmup + i)
            e.record()
            evs.append((s, e))
        barrier()
        per_step = [s.elapsed_time(e) for s, e in evs]
        total = sum(per_step) * 1e-3
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, W.launch_count() - l0, extra, per_step

    clocks.wait_ready()
    t_begin = time.time()
    t_res, launches, _, ms_res = timed(step_resident, args.steps, args.warmup)
    t_end = time.time()
    t_e2e, _, io, ms_e2e = timed(step_e2e, args.steps, args.warmup)

    roof = cpu = None
    if rank == 0:
        roof = conv_roofline(net, dev[0], pipeline, W, max(2, min(args.steps, 5)), peaks)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_pass(args, 2, 1)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    clocks.terminate()
    clk = clocks.summary(t_begin, t_end)
    if rank == 0:
        scenes = args.scenes * args.steps * world
        line = {"metric": METRIC, "value": scenes / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
                "ms_per_step_min_median_max": [round(min(ms_res), 3), round(statistics.median(ms_res), 3),
                                               round(max(ms_res), 3)], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)",
                          "bf16": "bf16 operands on tcgen05, fp32 accumulate", "simt": "f32"}[args.precision],
                "data": "synthetic", "config": workload_config(args, {"parallelism": "scene-sharded x%d, no collective" % world,
                                                                       "precision": args.precision}),
                "clocks": clk,
                "e2e": {"value": scenes / t_e2e, "unit": UNIT, "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1],
                        "ms_per_step": 1e3 * t_e2e / args.steps,
                        "ms_per_step_min_median_max": [round(min(ms_e2e), 3), round(statistics.median(ms_e2e), 3),
                                                       round(max(ms_e2e), 3)]},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
