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
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer: BASELINE.json configs[1] (the headline metric); train: configs[2], one data-parallel "
                         "training step (forward + MultiTaskLoss + backward + NCCL gradient all-reduce + AdamW) per batch")
    ap.add_argument("--scenes", type=int, default=4, help="scenes per batch (configs[1]: 4)")
    ap.add_argument("--points", type=int, default=150000, help="points per scene")
    ap.add_argument("--shape", default="scannet", choices=["scannet", "s3dis"],
                    help="scannet: configs[1] (8x6x2.6 m rooms, 2 cm voxels); s3dis: configs[3] (20x15x3 m rooms, 5 cm "
                         "voxels; use --points 1000000 --scenes 1)")
    ap.add_argument("--precision", default=os.environ.get("WSIS_PRECISION", "fp32"), choices=["fp32", "bf16", "simt"])
    ap.add_argument("--no-geometry-prefetch", action="store_true",
                    help="value = steps issued one after the other on one stream, e2e = BatchStream without the geometry "
                         "prefetch (default: both through the streaming loop whose side stream builds the next batch's "
                         "voxelization maps / rulebooks / tile records under the current batch's feature compute)")
    ap.add_argument("--stream-variants", action="store_true", help="also time the streaming loop's other configurations")
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
           "l2": "flushed between timed steps (256 MiB write)",
           "value_api": ("the e2e loop over batches already resident in HBM (pipeline.BatchStream: the next step's "
                         "coordinate-only part on the side stream under the current step's compute)"
                         if not getattr(args, "no_geometry_prefetch", False) else
                         "pipeline.forward_batch on batches resident in HBM, steps issued one after the other on one stream"),
           "e2e_api": "pipeline.BatchStream (H2D of step i+1 from pinned memory%s on a side stream under step i's compute) + "
                      "pipeline.forward_batch + pipeline.ResultFetcher (async D2H into pinned buffers); the L2 flush "
                      "write is inside the e2e region"
                      % ("" if getattr(args, "no_geometry_prefetch", False) else
                         " and step i+1's coordinate-only part (voxelization maps, rulebooks, tile records)")}
    cfg.update(extra or {})
    if cfg.get("mode") == "train":
        cfg["workload"] = cfg["workload"].replace("inference", "training step").replace(
            "UNet+pooling+ECC+affinity forward", "forward + MultiTaskLoss + backward + gradient all-reduce + AdamW")
        cfg["e2e_api"] = "pipeline.to_device (H2D from pinned memory) + train.TrainStep + loss read back"
        cfg["value_api"] = "train.TrainStep on batches resident in HBM, one stream"
    return cfg


def make_batches(args, rank, n_batches, with_labels=False):
    from wsis_b200 import synthetic
    out = []
    for b in range(n_batches):
        mk = synthetic.make_room_s3dis if args.shape == "s3dis" else synthetic.make_scene
        scenes = [mk(2000 + 100 * rank + 10 * b + i, n_points=args.points) for i in range(args.scenes)]
        out.append(synthetic.collate(scenes, with_labels=with_labels))
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
def cpu_pass(args, n_steps, warmup, sample_scenes=None, want_parity=None):
    """The reference's CPU path on the host cores: `warmup` untimed + `n_steps` timed steps, each over a bounded sample
    of `sample_scenes` scene(s) of the workload's batch (seeds 2000, 2001, ...: the first scenes of the device arm's
    batch 0).  value = scenes / wall time of the timed steps (not best-of).  `want_parity` (a dict) receives the last
    step's outputs and rulebooks for the device-vs-reference comparison."""
    from oracle import cpu_pipeline
    from wsis_b200 import pipeline, synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_sc = sample_scenes or args.cpu_sample_scenes
    net = pipeline.build_network(seed=123, device="cpu").eval()
    mk = synthetic.make_room_s3dis if args.shape == "s3dis" else synthetic.make_scene
    batch = synthetic.collate([mk(2000 + i, n_points=args.points) for i in range(n_sc)])
    times, stages, kind, ret, keep = [], None, "port", None, {}
    for it in range(warmup + n_steps):
        keep = {} if want_parity is not None else None
        t0 = time.perf_counter()
        ret, stages, kind = cpu_pipeline.forward(net, batch, keep=keep)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    if want_parity is not None:
        want_parity.update(ret=ret, keep=keep, batch=batch)
    total = sum(times)
    return {"value": n_sc * n_steps / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%d scene(s) of %d pts per step (of the %d-scene batch), full hot path, %d timed steps after %d "
                      "warm-up, mean; threads=%d; stages(s)=%s"
                      % (n_sc, args.points, args.scenes, n_steps, warmup, cores, {k: round(v, 3) for k, v in stages.items()}),
            "ms_per_step": 1e3 * total / n_steps, "ms_per_step_min_max": [round(1e3 * min(times), 1), round(1e3 * max(times), 1)],
            "scenes_per_step": n_sc, "steps": n_steps, "warmup": warmup}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU kernels (oracle/_ref) on the host cores, EXACTLY `--steps` timed and
    `--warmup` untimed steps.  A step is the full batch of the device arm when ~1.2 s per scene lets the whole run end
    within about three minutes, otherwise a bounded sample of it (and the line says which)."""
    if rank != 0:
        return
    est_scene_s = 1.3 * args.points / 150000.0
    budget_s = 170.0
    n_sc = max(1, min(args.scenes, int(budget_s / max(est_scene_s * (args.steps + args.warmup), 1e-9))))
    cb = cpu_pass(args, args.steps, args.warmup, sample_scenes=n_sc)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, {"note": "reference CPU path (unmodified spconv CPU kernels, oracle/_ref) on "
                                                      "the host cores; each step covers %d of the batch's %d scene(s)"
                                                      % (n_sc, args.scenes), "scenes_per_step": n_sc,
                                             "l2": "n/a (host)", "e2e_api": "n/a (host path)", "value_api": "n/a (host path)"}),
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
        g = groups.setdefault(shp, {"idx": [], "bytes": 0, "flops": f})
        g["idx"].append(i)
        g["bytes"] += b
    for g in groups.values():
        g["bytes"] /= len(g["idx"])  # mean over the group's launches (some fuse the residual read, some do not)
    for g in groups.values():
        g["t"] = statistics.mean(sum(st[i][0] for i in g["idx"]) for st in per_step)
    shp, g = max(groups.items(), key=lambda kv: kv[1]["t"])
    t_launch = g["t"] / len(g["idx"])
    ach = g["bytes"] / t_launch / 1e9
    # `traffic` cannot be measured inside a timed run (it needs ncu counters): it is read from the committed
    # `ncu --set full` capture of this kernel on this layer shape, and `traffic_source` says so; null when the capture
    # is of another shape
    traffic, traffic_source = None, None
    prof = os.path.join(ROOT, "profiles", "r02_ncu_conv_umma_l2.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        if list(pj.get("cin_cout_K_rows", [])) == list(shp):
            traffic = pj.get("dram_bytes_per_launch")
            traffic_source = "profiles/r02_ncu_conv_umma_l2.json (ncu --set full of one launch of this layer shape, " \
                             "not measured in this run)"
    ach_all = bsum / tsum / 1e9
    return {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": traffic, "traffic_source": traffic_source,
            "kernel": "conv_umma_kernel, layer shape (Cin, Cout, K, rows) = %s: %d launches per step, %.1f%% of the "
                      "step's sparse-conv time" % (list(shp), len(g["idx"]), 100.0 * g["t"] / tsum),
            "algorithmic_bytes_per_launch": int(g["bytes"]), "us_per_launch": round(t_launch * 1e6, 1),
            "useful_tflops": round(g["flops"] / t_launch / 1e12, 1),
            "by_layer_shape": [{"cin_cout_K_rows": list(k), "launches": len(v["idx"]), "ms_per_step": round(v["t"] * 1e3, 3),
                                "frac": round(v["bytes"] * len(v["idx"]) / v["t"] / 1e9 / peaks["hbm_gbs"], 4)}
                               for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["t"])],
            "all_sparse_conv": {"launches_per_step": nl, "algorithmic_bytes_per_step": bsum,
                                "ms_per_step": round(tsum * 1e3, 3), "GBps": round(ach_all, 1),
                                "frac": round(ach_all / peaks["hbm_gbs"], 4),
                                "useful_tflops": round(fsum / tsum / 1e12, 1)},
            "peak_source": peaks["source"]}


def count_all_launches(step):
    """Every kernel launch of one step, the library's own and torch's, counted by the CUDA profiler outside the timed
    region (gpu_launches counts only this repo's kernels)."""
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        return sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
                   and not e.name.lower().startswith("memcpy") and not e.name.lower().startswith("memset"))
    except Exception as e:  # noqa: BLE001  (profiler unavailable: say so instead of guessing)
        return "unavailable: %s" % type(e).__name__


def device_vs_reference(net, got, pipeline, precision):
    """The device path against the reference's CPU kernels on the cpu_baseline's sample scenes (same seeds, same
    weights): rulebooks as sorted pair sets, voxelization maps, U-Net output and every result tensor."""
    from oracle import parity
    dbatch, _ = pipeline.to_device(got["batch"])
    with torch.no_grad():
        ret, aux = pipeline.forward_batch(net, dbatch, keep_unet_features=True)
    res = parity.compare(ret, aux, got["ret"], got["keep"])
    return {"rulebooks_equal": res["rulebooks_equal"], "rulebook_pairs": {k: v["pairs"] for k, v in res["rulebooks"].items()},
            "voxelization_equal": res["voxelization_equal"], "max_rel": res["max_rel"],
            "max_rel_per_output": {k: float("%.3g" % v) for k, v in res["outputs"].items()},
            "against": "reference spconv CPU kernels (oracle/_ref) on %d scene(s); tolerance %s"
                       % (got["batch"]["batch_size"], "1e-4" if precision == "fp32" else "1e-2 (bf16 operands)")}


def _quiesce_gc():
    """The set-up (network, synthetic batches, warm-up) leaves a few hundred thousand live Python objects; a
    generation-2 pass of the cyclic collector over them inside a timed step is a 20-40 ms CPU pause, which an 11 ms
    step cannot hide (seen as single 24-51 ms outlier steps of the e2e loop).  Collect once and move the survivors to
    the permanent generation; the collector stays enabled."""
    import gc
    gc.collect()
    gc.freeze()


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

    loader = {"stream": None, "staging": None}

    def timed_stream(source, steps, warmup, prepare, fetch_results=True):
        """The public streaming API: pipeline.BatchStream copies batch i+1 from `source` (pinned host memory for e2e) on
        a copy stream while batch i computes -- with prepare=True it also builds batch i+1's coordinate-only part
        (voxelization maps, rulebooks, tile records) there -- and pipeline.ResultFetcher reads every step's results
        back into pinned host buffers.  All K copies in and K reads out happen inside the timed region (the stream is
        created after the barrier, so the first batch is not overlapped with anything; the region ends when the last
        result has landed on the host)."""
        fetch = pipeline.ResultFetcher()
        # one loader for the whole process: its side stream (and with it the allocator pool the geometry tensors come
        # from) and its staging sets are created once and outlive every loop
        warm = pipeline.BatchStream((source[i % n_batches] for i in range(max(warmup, 6))), prepare=prepare,
                                    copy_stream=loader["stream"], staging=loader["staging"])
        loader["stream"], loader["staging"] = warm.copy_stream, warm.staging
        for db, _ in warm:                                  # a long-lived loader: its copy stream and staging buffers
            with torch.no_grad():                           # outlive the warm-up
                ret, _ = pipeline.forward_batch(net, db)
            fetch.fetch(ret)                                # pinned result buffers are allocated here, not in the region
        fetch.wait()
        barrier()
        l0 = W.launch_count()
        evs, io = [], (0, 0)
        t0 = torch.cuda.Event(enable_timing=True)
        flush.zero_()
        t0.record()
        prev = t0
        for db, nb in pipeline.BatchStream((source[(warmup + i) % n_batches] for i in range(steps)),
                                           copy_stream=warm.copy_stream, staging=warm.staging, prepare=prepare):
            with torch.no_grad():
                ret, _ = pipeline.forward_batch(net, db)
            ob = fetch.fetch(ret)[1] if fetch_results else 0
            io = (nb, ob)
            flush.zero_()                                   # L2 flush between steps (inside the region: conservative)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append((prev, e))
            prev = e
        fetch.wait()
        barrier()
        per_step = [s.elapsed_time(e) for s, e in evs]
        total = sum(per_step) * 1e-3
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total, io, per_step, W.launch_count() - l0

    # value and e2e run through the public streaming loop (pipeline.BatchStream): the loader's high-priority side stream
    # builds batch i+1's coordinate-only part (voxelization maps, rulebooks, tile records: no feature is read) under
    # batch i's feature compute; value feeds it batches that are already resident in HBM, e2e pinned host batches.
    # `single_stream` = the same K resident steps issued one after the other on one stream (what --no-geometry-prefetch
    # reports as value).
    prefetch = not args.no_geometry_prefetch
    if prefetch:
        # set-up, not measurement: the loader's side stream has its own allocator pool; an untimed pass maps it (growing
        # it inside the first timed loop showed up as single 25-75 ms steps)
        timed_stream(dev, 8, 3, prepare=True, fetch_results=False)
    _quiesce_gc()
    clocks.wait_ready()
    t_begin = time.time()
    if prefetch:
        t_res, _, ms_res, launches = timed_stream(dev, args.steps, args.warmup, prepare=True, fetch_results=False)
    else:
        t_res, launches, _, ms_res = timed(step_resident, args.steps, args.warmup)
    t_end = time.time()
    t_seq, ms_seq = t_res, ms_res
    if prefetch:
        t_seq, _, _, ms_seq = timed(step_resident, args.steps, args.warmup)
    t_e2e, io, ms_e2e, _ = timed_stream(host, args.steps, args.warmup, prepare=prefetch)
    extra_streams = None
    if args.stream_variants:             # experiment: the same loop without the geometry prefetch, and from resident inputs
        extra_streams = {}
        for name, src, prep, fr in (("e2e_no_geometry_prefetch", host, False, True), ("resident_stream_no_geometry_prefetch", dev, False, False)):
            tt, _, ms, _ = timed_stream(src, args.steps, args.warmup, prepare=prep, fetch_results=fr)
            extra_streams[name] = {"value": args.scenes * args.steps * world / tt,
                                   "ms_per_step_min_median_max": [round(min(ms), 3), round(statistics.median(ms), 3), round(max(ms), 3)]}

    roof = cpu = parity_obj = None
    launches_total = None
    if rank == 0:
        roof = conv_roofline(net, dev[0], pipeline, W, max(2, min(args.steps, 5)), peaks)
        launches_total = count_all_launches(lambda: step_resident(0))
        if world == 1 and not args.no_cpu_baseline:
            got = {}
            cb = cpu_pass(args, 2, 1, want_parity=got)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            parity_obj = device_vs_reference(net, got, pipeline, args.precision)
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
                "single_stream": None if not prefetch else {"value": scenes / t_seq, "ms_per_step": 1e3 * t_seq / args.steps,
                                  "ms_per_step_min_median_max": [round(min(ms_seq), 3), round(statistics.median(ms_seq), 3),
                                                                 round(max(ms_seq), 3)],
                                  "note": "the same K steps issued one after the other on one stream (resident inputs)"},
                "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
                "launches_total_per_step": launches_total, "roofline": roof, "cpu_baseline": cpu, "parity": parity_obj}
        if extra_streams:
            line["stream_variants"] = extra_streams
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# training step (BASELINE.json configs[2])
# ---------------------------------------------------------------------------------------------------------
TRAIN_METRIC = "scenes/sec training step (ScanNet-shape, fwd+loss+bwd+gradient all-reduce+AdamW)"


def cpu_train_pass(args, n_steps, warmup, want=None):
    """The reference's CPU kernels under autograd (indiceConv + indiceConvBackward, torch BatchNorm in training mode,
    the MultiTaskLoss) on ONE scene of the batch per step."""
    from oracle import cpu_pipeline
    from wsis_b200 import pipeline, synthetic, train as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = pipeline.build_network(seed=123, device="cpu").train()
    batch = synthetic.collate([synthetic.make_scene(2000, n_points=args.points)], with_labels=True)
    crit = T.MultiTaskLoss()
    times, stages, loss = [], None, None
    for it in range(warmup + n_steps):
        net.zero_grad(set_to_none=True)
        if it == 0 and want is not None:                # parity sample: the very first step from the seed-123 weights
            t0 = time.perf_counter()
            loss, parts, _, stages = cpu_pipeline.train_step(net, batch, crit)
            want.update(loss=float(loss), grads={n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None},
                        batch=batch)
        else:
            t0 = time.perf_counter()
            loss, parts, _, stages = cpu_pipeline.train_step(net, batch, crit)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": n_steps / total, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "1 scene of %d pts per step (of the %d-scene batch): forward + MultiTaskLoss + backward on the "
                      "reference's CPU kernels (oracle/_ref indiceConv/indiceConvBackward, torch BatchNorm), no optimizer; "
                      "%d timed steps after %d warm-up; threads=%d; stages(s)=%s"
                      % (args.points, args.scenes, n_steps, warmup, cores, {k: round(v, 3) for k, v in stages.items()}),
            "ms_per_step": 1e3 * total / n_steps, "steps": n_steps, "warmup": warmup}


def run_train_reference(args, rank):
    if rank != 0:
        return
    n = max(1, min(args.steps, 6))
    cb = cpu_train_pass(args, n, min(args.warmup, 1))
    line = {"impl": "reference", "metric": TRAIN_METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": cb["warmup"], "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, {"mode": "train", "scenes_per_step": 1, "l2": "n/a (host)"}),
                           e2e_api="n/a (host path)", value_api="n/a (host path)"),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_train(args, rank, world, local_rank):
    import torch.distributed as dist
    from wsis_b200 import ops as W
    from wsis_b200 import pipeline
    from wsis_b200 import train as T
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    W.set_precision(args.precision)
    clocks = ClockSampler(local_rank)
    clocks.start()
    net = pipeline.build_network(seed=123, device="cuda").train()
    step = T.TrainStep(net)
    n_batches = 2
    host = [pipeline.pin_batch(b) for b in make_batches(args, rank, n_batches, with_labels=True)]
    dev = [pipeline.to_device(b)[0] for b in host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pool = torch.empty(16 << 30, dtype=torch.uint8, device="cuda")
    del pool
    for _ in range(2):                       # set-up: both batches twice (allocator pool, packed-weight buffers)
        for b in range(n_batches):
            step(dev[b])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        step(dev[i % n_batches])
        return None

    def step_e2e(i):
        db, nb = pipeline.to_device(host[i % n_batches])
        loss, _ = step(db)
        out = loss.to("cpu")
        return nb, out.numel() * out.element_size()

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

    _quiesce_gc()
    clocks.wait_ready()
    t_begin = time.time()
    t_res, launches, _, ms_res = timed(step_resident, args.steps, args.warmup)
    t_end = time.time()
    t_e2e, _, io, ms_e2e = timed(step_e2e, args.steps, args.warmup)

    # stage split of one step (CUDA events on the launching stream; outside the timed regions).  EVERY rank runs it:
    # the synchronised BatchNorm statistics and the gradient bucket are collectives.
    from wsis_b200.train import loss_inputs
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    flush.zero_()
    step.opt.zero_grad()
    evs[0].record()
    with torch.enable_grad():
        ret, _ = pipeline.forward_batch(net, dev[0])
        evs[1].record()
        loss, _ = step.loss(loss_inputs(ret, dev[0]), step.epoch)
        evs[2].record()
        loss.backward()
    evs[3].record()
    if world > 1:
        dist.all_reduce(step.opt.bucket.flat)
    evs[4].record()
    step.opt.step(grad_scale=1.0 / world)
    evs[5].record()
    torch.cuda.synchronize()
    names = ["forward", "loss", "backward", "grad_allreduce", "adamw"]
    stages = {n: round(evs[i].elapsed_time(evs[i + 1]), 3) for i, n in enumerate(names)}
    launches_total = count_all_launches(lambda: step_resident(0))       # a full step: all ranks (collectives inside)
    cpu = parity_obj = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        want = {}
        cb = cpu_train_pass(args, 2, 1, want=want)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        net2 = pipeline.build_network(seed=123, device="cuda").train()
        chk = T.TrainStep(net2)
        loss_d, _ = chk(pipeline.to_device(want["batch"])[0], optimize=False)
        worst, wname = 0.0, None
        gmax = max(float(r.abs().max()) for r in want["grads"].values())
        for name, p in net2.named_parameters():
            if name in want["grads"]:
                r = want["grads"][name]     # relative to max(own largest gradient, 1e-4 of the network's largest)
                d = float((p.grad.cpu() - r).abs().max()) / max(float(r.abs().max()), 1e-4 * gmax)
                if d > worst:
                    worst, wname = d, name
        parity_obj = {"loss_device": float(loss_d), "loss_reference_cpu": want["loss"],
                      "loss_rel": abs(float(loss_d) - want["loss"]) / abs(want["loss"]),
                      "max_param_grad_rel": float("%.3g" % worst), "worst_param": wname,
                      "against": "the reference's CPU kernels under autograd on 1 scene of %d pts, same weights" % args.points}
    clocks.terminate()
    clk = clocks.summary(t_begin, t_end)
    if rank == 0:
        scenes = args.scenes * args.steps * world
        line = {"metric": TRAIN_METRIC, "value": scenes / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps,
                "ms_per_step_min_median_max": [round(min(ms_res), 3), round(statistics.median(ms_res), 3), round(max(ms_res), 3)],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32 (bf16x3 split operands on tcgen05, fp32 accumulate)",
                          "bf16": "bf16 operands on tcgen05, fp32 accumulate", "simt": "f32"}[args.precision],
                "data": "synthetic",
                "config": workload_config(args, {"mode": "train", "parallelism": "data parallel x%d: synchronised BatchNorm "
                                                 "statistics + ONE NCCL all-reduce of the flat gradient bucket (%d floats) per step"
                                                 % (world, step.opt.flat_p.numel()), "precision": args.precision,
                                                 "optimizer": "AdamW lr 1e-3 wd 1e-4, ECC gradient clamp"}),
                "clocks": clk,
                "e2e": {"value": scenes / t_e2e, "unit": UNIT, "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1],
                        "ms_per_step": 1e3 * t_e2e / args.steps,
                        "ms_per_step_min_median_max": [round(min(ms_e2e), 3), round(statistics.median(ms_e2e), 3), round(max(ms_e2e), 3)]},
                "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
                "launches_total_per_step": launches_total, "stages_ms": stages, "roofline": None,
                "cpu_baseline": cpu, "parity": parity_obj}
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.mode == "train":
            run_train_reference(args, rank)
        else:
            run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        (run_train if args.mode == "train" else run_ours)(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
