#!/usr/bin/env python
"""bench.py — NDT scan-matches/sec on the workload BASELINE.json's metric is quoted on.

Workload (N = 1 and per rank for N > 1): BASELINE.json configs[1], "scan-to-map NDT align, 1080-beam scans vs
200 x 200 m map at 0.25 m cells". One step = one batched align pass over `--scans` synthetic 1080-beam scans
(SURVEY.md 8(d) generator) against the same map; every scan runs the full Levenberg-Marquardt loop of SPEC.md 5.

  value   whole-job matches/s with scans, offsets and initial poses already resident in HBM
  e2e     the same through the host-buffer C-ABI call (ndt2d_align_batch / _ranges): pinned host -> device copy
          of the step's scans, kernel, device -> host copy of the results, all inside the timed region
  roofline  Newton-step evaluations x algorithmic bytes (N*(8+32K)+92) / kernel time,
          against the measured HBM copy bandwidth (MEASURED_PEAKS.json). It is a gather-traffic convention:
          the cell table is L2-resident, so the fraction is not DRAM utilisation.
  cpu_baseline  the CPU spec oracle (a port of SPEC.md; the reference mount holds no source) on the host cores

--impl reference times that same CPU oracle as the reference arm (the reference's own matcher does not exist
in /root/reference). Multi-GPU: independent scans are sharded per rank, no data-path collective.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NDT scan-matches/sec (1080-pt 2D scans)"
UNIT = "matches/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 10; 50 for --workload sweep, whose steps are 0.3-2 ms)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="scan2map", choices=["scan2map", "pyramid", "sweep", "build", "newton", "odometry"],
                    help="scan2map: BASELINE configs[1] (default, the metric's config); pyramid: configs[2] (2.0/1.0/0.5 m, "
                         "10k scans, 0.2 m / 3 deg prior error); sweep: configs[3] (1M hypotheses x one 1080-pt scan)")
    ap.add_argument("--hyps", type=int, default=1000000, help="sweep: total hypotheses (sharded across GPUs)")
    ap.add_argument("--combine", default="p2p", choices=["p2p", "nccl"],
                    help="sweep on N > 1 GPUs: best-hypothesis combine by peer-memory stores from the arg-max kernel (p2p) "
                         "or by an NCCL all-gather on a side stream (nccl)")
    ap.add_argument("--scans", type=int, default=None, help="scans per step per GPU (65536 x 1080 x 8 B = 566 MB, far above the 126 MB L2)")
    ap.add_argument("--map-scans", type=int, default=2048, help="scans fused into the target map")
    ap.add_argument("--res", type=float, nargs="+", default=None)
    ap.add_argument("--overlap", type=int, default=0)
    ap.add_argument("--perturb", type=float, nargs=2, default=None, help="initial guess error: metres, degrees")
    ap.add_argument("--input", default="ranges_f32", choices=["xy", "ranges_f32", "ranges_u16"],
                    help="host input of the e2e leg: LaserScan ranges (what the sensor delivers; f32, SPEC.md section 8) or already "
                         "converted float2 points; every format is timed and listed under e2e.by_input. The CPU arms always get points.")
    ap.add_argument("--cpu-sample", type=int, default=0, help="scans in the cpu_baseline sample (0: auto, about 10-20 s)")
    ap.add_argument("--ref-scans", type=int, default=0, help="scans per step of the reference arm (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    dflt = {"scan2map": (65536, [0.25], [0.03, 0.3]), "pyramid": (10000, [2.0, 1.0, 0.5], [0.2, 3.0]), "sweep": (1, [0.25], [0.0, 0.0]),
            "build": (1, [0.25], [0.0, 0.0]), "newton": (1, [0.25], [0.0, 0.0]),
            "odometry": (16384, [0.5], [0.02, 0.2])}[a.workload]
    a.scans = a.scans or dflt[0]
    if a.steps is None:
        a.steps = 50 if a.workload == "sweep" else 10
    a.res = a.res or dflt[1]
    a.perturb = a.perturb or dflt[2]
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(args, rank, with_map=True, count=None):
    """Synthetic scans for this rank (distinct trajectory slots per rank), the shared map, and initial guesses.
    `count` < args.scans takes the first scans of the same workload (bounded CPU samples)."""
    from gtsam_ndt_b200 import synth
    world = max(1, args.gpus)
    traj = args.scans * world
    sc = synth.SCAN_1080
    count = count or args.scans
    ranges, poses = synth.scans(count, traj_len=traj, first=rank, step=world, **sc)
    pert = synth.uniform3(count, first=rank * args.scans) * np.array([args.perturb[0], args.perturb[0], math.radians(args.perturb[1])])
    init = poses + pert
    map_xy = synth.make_map(args.map_scans, traj_len=args.map_scans, **sc) if with_map else None
    return ranges, poses, init, map_xy


def to_points(ranges):
    """SPEC 8 conversion on the host; all 1080 beams return in this world, so the batch is rectangular."""
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    cb, sb = synth.beam_table(sc["nbeams"], sc["angle_min"], sc["angle_inc"])
    assert np.all(ranges > 0)
    xy = np.stack([ranges * cb[None, :], ranges * sb[None, :]], axis=-1).astype(np.float32)
    offsets = np.arange(ranges.shape[0] + 1, dtype=np.int64) * ranges.shape[1]
    return xy.reshape(-1, 2), offsets


def ncu_traffic(kernel, args, scans):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture of this exact
    configuration (profiles/traffic.json); None when no capture matches."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s/%s/scans=%d/res=%s/K=%d" % (kernel, args.workload, scans, "-".join(str(r) for r in args.res), 4 if args.overlap else 1)
        return tab.get(key, {}).get("dram_bytes")
    except Exception:
        return None


def gather_pipe(kernel_key, kernel_ms, clocks, sm_count):
    """The path's real ceiling next to the HBM figure: an SM's L1TEX data pipe takes one cycle per distinct 32 B sector of
    a gather. Sector count per launch from the committed ncu capture of this configuration."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        sectors = tab[kernel_key]["l1tex_sectors"]
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak = sm_count * mhz * 1e6
        ach = sectors / (kernel_ms / 1e3)
        return {"sectors_per_launch": sectors, "achieved_sectors_per_s": ach, "peak_sectors_per_s": peak, "frac": ach / peak,
                "note": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld from the ncu capture; peak = SMs x SM clock (one sector per cycle per SM)"}
    except Exception:
        return None


def ncu_traffic_sweep(args, hyps):
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "k_eval_poses/sweep/hyps=%d/res=%s/K=%d" % (hyps, "-".join(str(r) for r in args.res), 4 if args.overlap else 1)
        return tab.get(key, {}).get("dram_bytes")
    except Exception:
        return None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def eval_bytes(npts, K):
    return npts * (8 + 32 * K) + 12 + 80


def run_reference(args):
    """Reference arm: the CPU spec oracle (no upstream matcher exists in the mount), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from gtsam_ndt_b200 import synth
    o = oracle.Oracle(args.res, overlap=args.overlap)
    o.set_grid(-100.0, -100.0, 200.0, 200.0)
    cores = host_cores()      # torchrun exports OMP_NUM_THREADS=1; the arm uses every core the process may run on
    nref = args.ref_scans or max(64, min(args.scans, 1024 * cores))   # ~0.3 s of work per step: start-up costs are amortised
    ranges, poses, init, map_xy = make_workload(args, 0, count=nref)
    o.set_target(map_xy)
    xy, off = to_points(ranges)
    for _ in range(max(1, min(args.warmup, 1))):
        o.align_batch(xy, off, init, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):         # the CPU arm gets already converted points: its polar-to-point conversion is not charged
        res = o.align_batch(xy, off, init, nthreads=cores)
    dt = time.perf_counter() - t0
    v = nref * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 per point, f64 sums", "data": "synthetic",
            "config": workload_config(args, nref, extra={"sample": f"{nref} scans per step"}),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{nref} scans x {args.steps} steps, CPU spec oracle (SPEC.md port; reference mount has no source), host input: float2 points (conversion from ranges not charged)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "mean_iterations": float(res["iterations"].mean()), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, scans, extra=None):
    K = 4 if args.overlap else 1
    name = {"scan2map": "configs[1]: scan-to-map NDT align", "pyramid": "configs[2]: multi-resolution NDT align (pyramid)",
            "sweep": "configs[3]"}[args.workload]
    c = {"workload": "%s, 1080-beam scans vs 200x200 m map at %s m cells (K=%d), batch of %d scans per GPU per step"
         % (name, "/".join(str(r) for r in args.res), K, scans),
         "scans_per_step_per_gpu": scans, "points_per_scan": 1080, "cell_res_m": list(args.res), "K": K,
         "map_points": args.map_scans * 1080, "init_error": {"trans_m": args.perturb[0], "rot_deg": args.perturb[1]},
         "l2": "per-step scan input %.0f MB > 126 MB L2 (no flush needed); the 200x200 m cell table is L2-resident by design"
               % (scans * 1080 * 8 / 1e6),
         "parallelism": "independent scans sharded per GPU, no collective"}
    if extra:
        c.update(extra)
    return c


def run_native(args):
    import torch
    import torch.distributed as dist
    import gtsam_ndt_b200 as g

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(1, args.gpus):
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDT path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    K = 4 if args.overlap else 1

    ranges, poses, init, map_xy = make_workload(args, rank)
    xy, offsets = to_points(ranges)
    B, npts = args.scans, 1080

    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(args.res, device=local, stream=stream.cuda_stream, overlap=args.overlap)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    t0 = time.perf_counter()
    m.set_target(map_xy)
    build_ms = (time.perf_counter() - t0) * 1e3

    # device-resident inputs for `value`
    d_xy = torch.from_numpy(xy).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_init = torch.from_numpy(np.ascontiguousarray(init)).to(dev)
    d_res = torch.zeros(B * 144, dtype=torch.uint8, device=dev)

    def step_device():
        m.align_batch_device(d_xy, d_off, B, npts, d_init, d_res)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = m.kernel_launches - l0
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=g.RESULT_DTYPE)
    iters = res["iterations"].astype(np.int64)

    # e2e: host buffers through the public C-ABI call, pinned memory, copies inside the timed region.
    # The primary figure uses --input (default: float2 points, the same input the CPU arm gets); the LaserScan
    # formats (SPEC.md section 8) are timed as well and reported under e2e.by_input.
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    h_init = torch.from_numpy(np.ascontiguousarray(init)).pin_memory()
    h_res = torch.zeros(B * 144, dtype=torch.uint8).pin_memory()
    res_view = h_res.numpy().view(g.RESULT_DTYPE)
    U16_SCALE = 0.004  # 4 mm quantisation, 262 m maximum range

    def e2e_run(mode):
        if mode == "xy":
            h_in = torch.from_numpy(xy).pin_memory()
            h_off = torch.from_numpy(offsets).pin_memory()
            nbytes = h_in.numel() * 4 + h_off.numel() * 8 + h_init.numel() * 8
            fn = lambda: m.align_batch(h_in.numpy(), h_off.numpy(), h_init.numpy(), out=res_view)
        elif mode == "ranges_u16":
            h_in = torch.from_numpy(np.round(ranges / U16_SCALE).clip(1, 65535).astype(np.uint16)).pin_memory()
            nbytes = h_in.numel() * 2 + h_init.numel() * 8
            fn = lambda: m.align_batch_ranges(h_in.numpy(), sc["angle_min"], sc["angle_inc"], h_init.numpy(), range_scale=U16_SCALE, out=res_view)
        else:
            h_in = torch.from_numpy(ranges).pin_memory()
            nbytes = h_in.numel() * 4 + h_init.numel() * 8
            fn = lambda: m.align_batch_ranges(h_in.numpy(), sc["angle_min"], sc["angle_inc"], h_init.numpy(), range_scale=1.0, out=res_view)
        for _ in range(args.warmup):
            fn()
        sync_all()
        t0 = time.perf_counter()          # the call is host-synchronous: wall clock covers copies, kernels and the result read-back
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt_ms = (time.perf_counter() - t0) * 1e3
        same = bool(np.array_equal(res_view["iterations"], res["iterations"])) if mode != "ranges_u16" else None
        return dt_ms, int(nbytes), same

    e2e_all = {}
    for mode in dict.fromkeys([args.input, "xy", "ranges_f32", "ranges_u16"]):
        e2e_all[mode] = e2e_run(mode)
    sync_all()
    e2e_ms, in_bytes, e2e_iters_equal = e2e_all[args.input]
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        cnt = torch.tensor([float(iters.sum()), float(B)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_evals_per_step, total_scans = float(cnt[0]), float(cnt[1])
    else:
        total_evals_per_step, total_scans = float(iters.sum()), float(B)

    if rank == 0:
        hbm, peak_src = peaks()
        value = total_scans * args.steps / (ms / 1e3)
        e2e_value = total_scans * args.steps / (e2e_ms / 1e3)
        # roofline of the dominant kernel (k_align), per launch on this rank
        kernel_ms = ms / args.steps
        alg_bytes = float(iters.sum()) * eval_bytes(npts, K)
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 per point, f64 sums and solver", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(B * 144),
                    "ms_per_step": e2e_ms / args.steps, "input": args.input, "api": "ndt2d_align_batch" + ("" if args.input == "xy" else "_ranges"),
                    "iterations_equal_device_run": e2e_iters_equal,
                    "by_input": {k: {"value": B * args.steps / (v[0] / 1e3) * world, "h2d_bytes_per_step": v[1], "note": "rank 0 timing"} for k, v in e2e_all.items()}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": ncu_traffic("k_align", args, B),
                         "kernel": "k_align (whole LM loop per scan, one warp per scan)", "peak_source": peak_src,
                         "convention": "gather traffic: every point read and 32 B cell gather counts, bytes/eval = N*(8+32K)+92; "
                                       "cells are served by L1/L2, so this is not DRAM utilisation",
                         "evals_per_launch": float(iters.sum()), "bytes_per_eval": eval_bytes(npts, K),
                         "evals_per_s": total_evals_per_step * args.steps / (ms / 1e3),
                         "gather_pipe": gather_pipe("k_align/%s/scans=%d/res=%s/K=%d" % (args.workload, B, "-".join(str(r) for r in args.res), K),
                                                    kernel_ms, clocks, torch.cuda.get_device_properties(local).multi_processor_count)},
            "mean_iterations": float(iters.mean()), "status_counts": np.bincount(res["status"], minlength=4).tolist(),
            "map_build_ms": build_ms, "clocks": clocks,
        }
        # single-scan latency (configs[0]/[1] at batch 1): one warp runs the whole LM loop, so this is a latency
        # figure, not a throughput or roofline figure (hard part 4)
        lat = []
        one = np.ascontiguousarray(xy[:npts])
        for i in range(60):
            t0 = time.perf_counter()
            r1 = m.align(one, init[0])
            lat.append((time.perf_counter() - t0) * 1e6)
        line["single_align_latency_us"] = {"median": float(np.median(lat[10:])), "min": float(np.min(lat[10:])),
                                           "iterations": int(r1["iterations"]), "api": "ndt2d_align (host buffers, synchronous)"}
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = cpu_baseline(args, xy, offsets, init, map_xy, res)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_build(args):
    """north_star stage (1), the cell-grid build: `--map-scans` x 1080 map points binned into the 200 x 200 m lattice
    (integer accumulation with warp-level segmented reduction, then per-cell finalisation). One step = one
    ndt2d_set_target_device on device-resident points; single GPU (every rank builds its own replica of the map)."""
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDT path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K = 4 if args.overlap else 1
    sc = synth.SCAN_1080
    map_xy = synth.make_map(args.map_scans, traj_len=args.map_scans, **sc)
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(args.res, device=local, stream=stream.cuda_stream, overlap=args.overlap)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    # rotate through copies of the point cloud totalling more than the 126 MB L2 (timing rule: inputs not L2-resident)
    nrot = max(2, int(math.ceil(160e6 / (map_xy.nbytes))))
    d_maps = [torch.from_numpy(map_xy).to(dev) for _ in range(nrot)]
    for i in range(args.warmup):
        m.set_target_device(d_maps[i % nrot], len(map_xy))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        m.set_target_device(d_maps[i % nrot], len(map_xy))
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = m.kernel_launches - l0
    # incremental update: one scan's 1080 points added to the finished map, device-resident
    d_scan = d_maps[0][:1080].contiguous()
    for _ in range(5):
        m.add_target_device(d_scan, 1080)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(stream)
    for _ in range(50):
        m.add_target_device(d_scan, 1080)
    eb.record(stream)
    torch.cuda.synchronize()
    add_us = ea.elapsed_time(eb) * 1e3 / 50
    # e2e: host points through ndt2d_set_target (pinned memory, copy in the timed region, host-synchronous)
    h_map = torch.from_numpy(map_xy).pin_memory()
    m.set_target(h_map.numpy())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.set_target(h_map.numpy())
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    ncells = sum(m.geometry(lv)["njx"] * m.geometry(lv)["njy"] for lv in range(len(args.res)))
    valid = int((m.cells(len(args.res) - 1)[..., 7] != 0).sum())
    # algorithmic bytes of one build: every point read once per level; every cell's count and five sums cleared, then
    # read and written once by the atomics' read-modify-write at least, read by the finalisation; every record written
    per_build = len(args.res) * len(map_xy) * 8 + ncells * (44 + 2 * 44 + 44 + 32)
    hbm, peak_src = peaks()
    achieved = per_build / (ms / 1e3) / 1e9
    line = {"metric": "NDT map points binned/sec (cell-grid build)", "value": len(map_xy) / (ms / 1e3), "unit": "points/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i64 fixed-point sums, f64 finalisation, f32 records", "data": "synthetic",
            "config": {"workload": "north_star stage 1: cell-grid build, %d map points into the 200x200 m lattice at %s m cells (K=%d): %d cells, %d valid on the finest level"
                                   % (len(map_xy), "/".join(str(r) for r in args.res), K, ncells, valid),
                       "l2": "steps rotate through %d copies of the point cloud (%.0f MB > 126 MB L2)" % (nrot, nrot * map_xy.nbytes / 1e6),
                       "parallelism": "replicas only (each rank builds its own copy of the map)"},
            "e2e": {"value": len(map_xy) / (e2e_ms / 1e3), "unit": "points/s", "h2d_bytes_per_step": int(map_xy.nbytes), "d2h_bytes_per_step": 0,
                    "ms_per_step": e2e_ms, "api": "ndt2d_set_target"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                         "kernel": "k_accumulate + k_finalize (+ 3 clears) per level", "peak_source": peak_src, "bytes_per_build": per_build,
                         "convention": "points 8 B per level; per cell: clear 44 B, atomic read-modify-write 88 B, finalise read 44 B, record 32 B"},
            "incremental_update_us": {"value": add_us, "what": "ndt2d_add_target_device of one 1080-point scan into the finished map (accumulate + finalise), per call"},
            "clocks": clocks}
    print(json.dumps(line), flush=True)


def run_odometry(args):
    """Batched scan-to-scan odometry (ndt2d_align_pairs): `--scans` consecutive 1080-beam scans of the trajectory, pair
    k = (scan k-1 as target, scan k as source), prior = true relative motion + `--perturb`. One step = every target's
    grid built (one warp per target and level, hash tables) + every pair aligned. BASELINE configs[0] is the single
    CPU case of this align (scan-to-scan, 0.5 m cells); the batched form is what a log or a set of loop-closure
    candidates needs. Single GPU (pairs would shard like scans)."""
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDT path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K = 4 if args.overlap else 1
    sc = synth.SCAN_1080
    n = args.scans
    ranges, poses = synth.scans(n, traj_len=3770, first=0, step=1, **sc)      # the 377 m loop in 0.1 m steps (laps repeat with new noise)
    xy, offsets = to_points(ranges)
    pairs = np.stack([np.arange(n - 1), np.arange(1, n)], 1).astype(np.int32)
    c, s_ = np.cos(poses[:-1, 2]), np.sin(poses[:-1, 2])
    d = poses[1:] - poses[:-1]
    rel = np.stack([c * d[:, 0] + s_ * d[:, 1], -s_ * d[:, 0] + c * d[:, 1], d[:, 2]], 1)
    init = rel + synth.uniform3(n - 1) * np.array([args.perturb[0], args.perturb[0], math.radians(args.perturb[1])])
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(args.res, device=local, stream=stream.cuda_stream, overlap=args.overlap)
    d_xy = torch.from_numpy(xy).to(dev)
    d_off = torch.from_numpy(offsets).to(dev)
    d_init = torch.from_numpy(np.ascontiguousarray(init)).to(dev)
    d_res = torch.zeros((n - 1) * 144, dtype=torch.uint8, device=dev)
    for _ in range(args.warmup):
        m.align_pairs_device(d_xy, d_off, offsets, pairs, d_init, d_res)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        m.align_pairs_device(d_xy, d_off, offsets, pairs, d_init, d_res)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = m.kernel_launches - l0
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=g.RESULT_DTYPE)
    # e2e: pinned host scans through ndt2d_align_pairs, copies and result read-back in the timed region
    h_xy = torch.from_numpy(xy).pin_memory()
    m.align_pairs(h_xy.numpy(), offsets, pairs, init)
    t0 = time.perf_counter()
    reps = max(1, args.steps // 2)
    for _ in range(reps):
        rh = m.align_pairs(h_xy.numpy(), offsets, pairs, init)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / reps
    clocks = sampler.stop()
    # the same pairs one at a time through set_target + align (what a front end without the batched call does)
    t0 = time.perf_counter()
    nseq = min(256, n - 1)
    scans_list = xy.reshape(n, -1, 2)
    same = True
    for p in range(nseq):
        m.set_target(scans_list[pairs[p, 0]])
        one = m.align(scans_list[pairs[p, 1]], init[p])
        same = same and one.tobytes() == res[p].tobytes()
    seq_ms = (time.perf_counter() - t0) * 1e3 / nseq
    err = res["pose"] - rel
    good = res["status"] == 0
    iters = res["iterations"].astype(np.int64)
    hbm, peak_src = peaks()
    alg_bytes = float(iters.sum()) * eval_bytes(1080, K)
    achieved = alg_bytes / (ms / 1e3) / 1e9
    line = {"metric": "NDT scan-to-scan matches/sec (1080-pt 2D scans, batched pairs)", "value": (n - 1) / (ms / 1e3), "unit": "matches/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 per point, f64 sums and solver", "data": "synthetic",
            "config": {"workload": "batched scan-to-scan odometry: %d consecutive 1080-beam scans 0.1 m apart, pair k = (scan k-1 -> scan k), %s m cells (K=%d), "
                                   "per-target grids in hash tables, prior error %.2f m / %.1f deg" % (n, "/".join(str(r) for r in args.res), K, args.perturb[0], args.perturb[1]),
                       "l2": "per-step scan input %.0f MB; per-target tables %.1f GB are rebuilt every step" % (xy.nbytes / 1e6, (n - 1) * len(args.res) * 4097 * 76 / 1e9),
                       "parallelism": "replicas only in this bench (pairs shard like scans)"},
            "e2e": {"value": (n - 1) / (e2e_ms / 1e3), "unit": "matches/s", "h2d_bytes_per_step": int(xy.nbytes + offsets.nbytes + init.nbytes),
                    "d2h_bytes_per_step": int((n - 1) * 144), "ms_per_step": e2e_ms, "api": "ndt2d_align_pairs",
                    "equals_device_run": bool(rh.tobytes() == res.tobytes())},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                         "kernel": "k_align<PAIRS> (hash-table gathers) after k_pairs_build", "peak_source": peak_src,
                         "bytes_per_eval": eval_bytes(1080, K), "evals_per_launch": float(iters.sum()),
                         "convention": "gather traffic (DESIGN.md section 4); the build's traffic is not counted"},
            "mean_iterations": float(iters.mean()), "status_counts": np.bincount(res["status"], minlength=4).tolist(),
            "median_abs_err_vs_truth_m": float(np.median(np.abs(err[good, :2]))) if good.any() else None,
            "sequential_set_target_plus_align": {"ms_per_pair": seq_ms, "pairs": nseq, "bit_identical_to_batched": bool(same)},
            "clocks": clocks}
    print(json.dumps(line), flush=True)


def run_newton(args):
    """north_star stage (2) in isolation: the Newton-step evaluation (score, 3-gradient, 3x3 Hessian) of one 1080-point
    scan at `--hyps` poses scattered like LM iterates around the true pose (k_eval_poses<FULL>, ndt2d_evaluate_device).
    This is SURVEY.md 8(d)'s unit of work with no solver and no iteration control around it. Single GPU."""
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDT path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K = 4 if args.overlap else 1
    sc = synth.SCAN_1080
    map_xy = synth.make_map(args.map_scans, traj_len=args.map_scans, **sc)
    ranges, poses = synth.scans(1, traj_len=1000, first=137, **sc)
    xy = synth.polar_to_points(ranges[0], sc["angle_min"], sc["angle_inc"])
    npose = args.hyps
    rng = np.random.default_rng(7)
    P = poses[0] + rng.normal(size=(npose, 3)) * np.array([0.03, 0.03, math.radians(0.3)])
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(args.res, device=local, stream=stream.cuda_stream, overlap=args.overlap)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    m.set_target(map_xy)
    d_xy = torch.from_numpy(xy).to(dev)
    nrot = max(2, int(math.ceil(160e6 / (npose * (24 + 80 + 4)))))      # poses in, sums and counts out: rotate beyond the L2
    d_P = [torch.from_numpy(P).to(dev) for _ in range(nrot)]
    d_out = [torch.zeros(npose * 10, dtype=torch.float64, device=dev) for _ in range(nrot)]
    d_cnt = [torch.zeros(npose, dtype=torch.int32, device=dev) for _ in range(nrot)]
    for i in range(args.warmup):
        m.evaluate_device(d_xy, len(xy), d_P[i % nrot], npose, d_out[i % nrot], d_cnt[i % nrot])
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(args.steps):
        m.evaluate_device(d_xy, len(xy), d_P[i % nrot], npose, d_out[i % nrot], d_cnt[i % nrot])
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = m.kernel_launches - l0
    # e2e: host poses in, host sums out, through ndt2d_evaluate (pinned memory, copies in the timed region)
    h_P = torch.from_numpy(P).pin_memory()
    m.evaluate(xy, h_P.numpy())
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 5)):
        out, cnt = m.evaluate(xy, h_P.numpy())
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, args.steps // 5)
    clocks = sampler.stop()
    dev_out = d_out[(args.steps - 1) % nrot].cpu().numpy().reshape(npose, 10)
    per_eval = eval_bytes(len(xy), K)
    hbm, peak_src = peaks()
    achieved = npose * per_eval / (ms / 1e3) / 1e9
    line = {"metric": "NDT Newton-step evaluations/sec (1080-pt scan: score, gradient, Hessian)", "value": npose / (ms / 1e3), "unit": "evaluations/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 per point, f64 sums", "data": "synthetic",
            "config": {"workload": "north_star stage 2: Newton-step kernel alone, one %d-pt scan at %d poses (sigma 3 cm / 0.3 deg around the truth) vs 200x200 m map at %s m cells (K=%d)"
                                   % (len(xy), npose, args.res[0], K),
                       "l2": "steps rotate through %d copies of the pose and result buffers (%.0f MB > 126 MB L2); the scan and the cell table are cache-resident by design"
                             % (nrot, nrot * npose * 108 / 1e6),
                       "parallelism": "replicas only"},
            "e2e": {"value": npose / (e2e_ms / 1e3), "unit": "evaluations/s", "h2d_bytes_per_step": int(npose * 24 + len(xy) * 8),
                    "d2h_bytes_per_step": int(npose * 84), "ms_per_step": e2e_ms, "api": "ndt2d_evaluate",
                    "equals_device_run": bool(np.array_equal(out, dev_out))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                         "kernel": "k_eval_poses<FULL> (one warp per pose, scan staged per block)", "peak_source": peak_src, "bytes_per_eval": per_eval,
                         "convention": "gather traffic (DESIGN.md section 4): every point read and 32 B record gather counts; scan and cells are cache-resident"},
            "clocks": clocks}
    print(json.dumps(line), flush=True)


def run_sweep(args):
    """BASELINE configs[3]: relocalisation, `--hyps` pose hypotheses x one 1080-pt scan vs the global map, hypotheses
    sharded across the GPUs, best-hypothesis combine as the only collective. One step = one full sweep + combine."""
    import torch
    import torch.distributed as dist
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth, distributed as D

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the NDT path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    K = 4 if args.overlap else 1
    sc = synth.SCAN_1080
    map_xy = synth.make_map(args.map_scans, traj_len=args.map_scans, **sc)
    ranges, poses = synth.scans(1, traj_len=1000, first=137, **sc)
    xy = synth.polar_to_points(ranges[0], sc["angle_min"], sc["angle_inc"])
    # regular lattice, 0.2 m in x and y, 3 deg in theta, truncated to --hyps
    nth = 120
    side = int(math.ceil(math.sqrt(args.hyps / nth)))
    gx = (np.arange(side) - side // 2) * 0.2
    lat = np.stack(np.meshgrid(gx, gx, np.radians(np.arange(nth) * 3.0 - 180.0), indexing="ij"), -1).reshape(-1, 3)[: args.hyps]
    hyp = (poses[0] + lat).astype(np.float32)
    lo, hi = D.shard_range(len(hyp), rank, world)
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(args.res, device=local, stream=stream.cuda_stream, overlap=args.overlap)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    m.set_target(map_xy)
    d_xy = torch.from_numpy(xy).to(dev)
    # timing rule: per-step inputs must not be L2-resident from the previous step. A shard's hypotheses (12 B each) and
    # scores (8 B) are far smaller than the 126 MB L2, so the steps rotate through copies totalling more than the L2.
    nrot = max(2, int(math.ceil(160e6 / max(1, (hi - lo) * 20))))
    d_hyps = [torch.from_numpy(hyp[lo:hi].copy()).to(dev) for _ in range(nrot)]
    d_scores_rot = [torch.zeros(hi - lo, dtype=torch.float64, device=dev) for _ in range(nrot)]
    # Successive sweeps are independent queries. Combine of the per-GPU bests, two implementations:
    #  p2p  - ndt2d_sweep_publish: the arg-max kernel's final block stores this rank's best (global index, score) into
    #         every rank's table over NVLink; no collective per query; the host polls its own table LAG queries later.
    #  nccl - the library writes (local best index, score bits) into one 16-byte buffer that is all-gathered on a side
    #         stream while the next query's kernel runs on the main stream (two buffers, events both ways).
    LAG, NSLOTS = 8, 64
    ex, combine = None, "none (1 GPU)"
    if world > 1 and args.combine == "p2p":
        try:
            ex = D.PeerExchange(m, nslots=NSLOTS)
            combine = "p2p: peer-memory stores from the arg-max kernel (ndt2d_sweep_publish), host poll %d queries behind" % LAG
        except Exception as e:      # e.g. CUDA IPC not permitted in this container: say so and use the NCCL path
            print(f"[bench] peer-memory exchange unavailable ({e}); using the NCCL combine", file=sys.stderr)
            ex = None
    if world > 1 and ex is None:
        combine = "nccl: all-gather of one 16 B pair per rank on a side stream, overlapping the next sweep"
    pair = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(2)]
    gathered = [torch.zeros(2 * world, dtype=torch.int64, device=dev) for _ in range(2)]
    comm = torch.cuda.Stream(device=dev) if world > 1 and ex is None else None
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "pending": [False, False], "waited": 0, "best": None}

    def step():
        i = state["i"]
        state["i"] += 1
        r = i % nrot
        if ex is not None:
            ex.publish(d_xy, len(xy), d_hyps[r], hi - lo, d_scores_rot[r], lo, i)
            if i >= LAG:
                state["best"] = ex.wait(i - LAG)
                state["waited"] = i - LAG + 1
            return
        b = i & 1
        if world > 1 and state["pending"][b]:
            stream.wait_event(consumed[b])        # the combine of query i-2 has read pair[b]
        m.sweep_device(d_xy, len(xy), d_hyps[r], hi - lo, d_scores_rot[r], 1, pair[b].data_ptr(), pair[b].data_ptr() + 8)
        if world > 1:   # best-hypothesis combine: 16 B per rank
            ready[b].record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ready[b])
                dist.all_gather_into_tensor(gathered[b], pair[b])
                consumed[b].record(comm)
            state["pending"][b] = True

    def drain():
        """every outstanding combine has delivered its result (called before the closing timing event)"""
        if ex is not None:
            for q in range(state["waited"], state["i"]):
                state["best"] = ex.wait(q)
            state["waited"] = state["i"]
        elif world > 1:
            for b in range(2):
                if state["pending"][b]:
                    stream.wait_event(consumed[b])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    drain()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = m.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    drain()
    ev1.record(stream)
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = m.kernel_launches - l0
    # e2e: host hypotheses in, best pose out, through the host-buffer call + the gloo/nccl combine
    h_hyp = torch.from_numpy(hyp[lo:hi].copy()).pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gi, gs = D.sweep_sharded(lambda a, b, k: m.sweep(xy, h_hyp.numpy(), k=k, want_scores=False)[1:], len(hyp), k=1) if world == 1 else \
            D.combine_topk(*(lambda r: (r[1] + lo, r[2]))(m.sweep(xy, h_hyp.numpy(), k=1, want_scores=False)), 1)
    sync_all()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        hbm, peak_src = peaks()
        per_hyp = len(xy) * (8 + 32 * K) + 12 + 8
        kernel_ms = ms / args.steps
        achieved = (hi - lo) * per_hyp / (kernel_ms / 1e3) / 1e9
        truth_err = float(np.abs(hyp[int(gi[0])] - poses[0]).max())
        line = {"metric": "NDT pose hypotheses/sec (1080-pt scan vs global map)", "value": len(hyp) * args.steps / (ms / 1e3),
                "unit": "hypotheses/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 per point, f64 sums", "data": "synthetic",
                "config": {"workload": "configs[3]: relocalisation, %d pose hypotheses (0.2 m x 0.2 m x 3 deg lattice) x one %d-pt scan vs 200x200 m map at %s m cells (K=%d), hypotheses sharded over %d GPU(s)"
                                       % (len(hyp), len(xy), args.res[0], K, world),
                           "l2": "steps rotate through %d copies of the shard's hypotheses and score buffers (%.0f MB in total > 126 MB L2); the scan and the cell table are cache-resident by design" % (nrot, nrot * (hi - lo) * 20 / 1e6),
                           "parallelism": "hypotheses sharded per GPU; combine of one (score, index) pair per rank: " + combine},
                "e2e": {"value": len(hyp) * args.steps / (e2e_ms / 1e3), "unit": "hypotheses/s", "h2d_bytes_per_step": int((hi - lo) * 12 + len(xy) * 8),
                        "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / args.steps, "api": "ndt2d_sweep + combine_topk"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": ncu_traffic_sweep(args, len(hyp)) if world == 1 else None,
                             "kernel": "k_eval_poses (score only)", "peak_source": peak_src, "bytes_per_hypothesis": per_hyp,
                             "gather_pipe": gather_pipe("k_eval_poses/sweep/hyps=%d/res=%s/K=%d" % (len(hyp), "-".join(str(r) for r in args.res), K),
                                                        kernel_ms, clocks, torch.cuda.get_device_properties(local).multi_processor_count) if world == 1 else None,
                             "convention": "gather traffic (see DESIGN.md section 4); scan and cells are cache-resident"},
                "best_hypothesis_abs_err_vs_truth": truth_err, "clocks": clocks}
        if ex is not None and state["best"] is not None:
            line["combine_equals_host_api_result"] = bool(int(state["best"][0]) == int(gi[0]) and float(state["best"][1]) == float(gs[0]))
        print(json.dumps(line), flush=True)
    if ex is not None:
        ex.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def cpu_baseline(args, xy, offsets, init, map_xy, res_gpu):
    """The CPU spec oracle on a bounded sample of the same workload, all host threads (reported, not the target)."""
    import oracle
    o = oracle.Oracle(args.res, overlap=args.overlap)
    o.set_grid(-100.0, -100.0, 200.0, 200.0)
    o.set_target(map_xy)
    cores = host_cores()
    probe = min(len(offsets) - 1, 8 * cores)
    t0 = time.perf_counter()
    o.align_batch(xy, offsets[: probe + 1], init[:probe], nthreads=cores)
    rate = probe / (time.perf_counter() - t0)
    n = args.cpu_sample or int(min(len(offsets) - 1, max(probe, rate * 12.0)))
    t0 = time.perf_counter()
    reps = 0
    while True:
        r = o.align_batch(xy, offsets[: n + 1], init[:n], nthreads=cores)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= 10.0 or reps >= 64:
            break
    dp = np.abs(r["pose"] - res_gpu["pose"][:n]).max()
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} scans of the step x {reps} passes, CPU spec oracle (SPEC.md port), {dt:.1f} s",
            "max_abs_pose_diff_vs_gpu": float(dp), "iterations_equal": bool(np.array_equal(r["iterations"], res_gpu["iterations"][:n]))}


if __name__ == "__main__":
    # stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "sweep":
        run_sweep(a)
    elif a.workload == "build":
        run_build(a)
    elif a.workload == "newton":
        run_newton(a)
    elif a.workload == "odometry":
        run_odometry(a)
    else:
        run_native(a)
