#!/usr/bin/env python
"""bench.py -- throughput of the MaD local-feature hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c5]

One *step* = one pass of the hot path over one synthetic map of BASELINE.json configs[1] (C2):
a 256^3 assembly map (8 A, 2 A/voxel, 6 random-walk components) through
    zero-pad -> 2x spline upsample + presmooth -> LoG/Gauss/gradient (both octaves)
    -> 3x3x3 maxima + Newton refinement -> EQSP orientations -> int16[1024] descriptors
    -> cosine matching (threshold 0.6) of every component's descriptors against the map's
all in libmad_b200.so (hand-written sm_100a CUDA) through the C ABI.  `value` = input voxels
(256^3 per map) per second with the map already resident in HBM; `e2e` = the same through the
host-buffer API (pinned host grid in, host descriptor/keypoint/pair tables out, copies timed).

N > 1 (torchrun, one rank per GPU): maps are independent units, rank r describes
its own map (seed offset r), no data-path collective; value = N maps' voxels / max-over-ranks time
("weak" scaling).  `--workload c5` runs the all-pairs matching config instead (reference axis
sharded over ranks, NCCL all-gather + top-k merge).

`--impl reference` times the CPU restatement of the reference (oracle/mad_oracle.py, NumPy/SciPy --
the reference is pure Python and cannot travel to the GPU box) on a bounded sample of the same
map, one process per host core.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

import synth  # noqa: E402  (oracle/synth.py: input generators only)

C2 = synth.C2
CPU_SAMPLE_SIDE = 64
METRIC = "voxels/sec scale-space+detect+describe(+match)"
UNIT = "voxels/s"


def workload_name():
    return ("C2: 256^3 synthetic assembly map, 8 A, 2 A/voxel, 6 components x 40000 atoms; "
            "describe map + threshold-match each component's descriptors against it")


# ---------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference) -- used by cpu_baseline and --impl reference only
# ---------------------------------------------------------------------------------------------
def _cpu_one(args):
    crop, voxelsp = args
    import mad_oracle as mo
    t0 = time.perf_counter()
    sp, kp, ori, dsc = mo.describe_struct(crop, voxelsp)
    if len(dsc):
        mo.match_threshold(dsc[: max(1, len(dsc) // 6)], dsc, 0.6)
    return time.perf_counter() - t0, len(kp["oct"]), len(ori["kp"])


def cpu_crops(grid, n_crops, side=CPU_SAMPLE_SIDE):
    """n_crops sub-cubes with occupancy closest to the whole map's (first = the most representative)."""
    import synth
    occ = grid > 0.05
    target = float(occ.mean())
    cands = []
    step = side // 2
    for x in range(0, grid.shape[0] - side + 1, step):
        for y in range(0, grid.shape[1] - side + 1, step):
            for z in range(0, grid.shape[2] - side + 1, step):
                cands.append((abs(float(occ[x:x + side, y:y + side, z:z + side].mean()) - target), x, y, z))
    cands.sort()
    return [np.ascontiguousarray(grid[x:x + side, y:y + side, z:z + side]) for _, x, y, z in cands[:n_crops]]


def cpu_run(grid, voxelsp, procs, steps, warmup, side=CPU_SAMPLE_SIDE):
    """Each step: `procs` processes describe one side^3 crop each (the reference is single-threaded
    Python; one map per process is its many-core form).  Returns (voxels/s, seconds per step, info)."""
    import multiprocessing as mp
    crops = cpu_crops(grid, procs, side)
    while len(crops) < procs:
        crops.append(crops[len(crops) % max(len(crops), 1)])
    jobs = [(c, voxelsp) for c in crops]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        for _ in range(warmup):
            pool.map(_cpu_one, jobs)
        t0 = time.perf_counter()
        info = None
        for _ in range(steps):
            info = pool.map(_cpu_one, jobs)
        dt = time.perf_counter() - t0
    vox = steps * procs * side ** 3
    return vox / dt, dt / steps, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    grid = synth.assembly_map(**C2)
    procs = max(1, min(os.cpu_count() or 1, 32))
    # the sample is sized so that the whole run (steps + warm-up) stays near two minutes: a 64^3 crop per
    # process takes ~4.6 s on a 16-core host, and the cost scales with the crop volume
    budget = 120.0 / max(1, args.steps + args.warmup)
    side = 64
    for cand in (64, 56, 48, 40, 32):
        side = cand
        if 4.6 * (cand / 64.0) ** 3 <= budget:
            break
    value, s_per_step, info = cpu_run(grid, C2["voxelsp"], procs, args.steps, args.warmup, side)
    sample = ("%d x %d^3 occupancy-matched crops of the C2 map per step, one process each "
              "(oracle/mad_oracle.py: NumPy/SciPy port of the reference, incl. matching)" % (procs, side))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# algorithmic bytes per launch (DESIGN.md "Kernels")
# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(name, ctx):
    """Bytes one launch of kernel `name` must move by the data-flow convention of SURVEY 8(d):
    every logical array read once by its consumer and written once by its producer."""
    v = ctx["V_cur"]          # voxels of the octave this launch worked on (set by the caller)
    table = {
        "log_gauss_fused_kernel": 12 * v,                 # reads f32 grid, writes LoG + Gauss
        "log_pass_x_kernel": 12 * v, "log_pass_y_kernel": 20 * v, "log_pass_z_kernel": 20 * v,
        "gradient_kernel": 16 * v,                        # reads Gauss, writes 3 components
        "detect_kernel": 4 * v,
        "pad3d_kernel": 4 * ctx["V_in"] + 4 * ctx["V1"],
    }
    return table.get(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--match-impl", type=int, default=None, help="default: product (uint8 tcgen05 one-pass); 1 = SIMT check, 2 = fp16 tcgen05")
    ap.add_argument("--exact", type=int, default=1, help="1 = float64 line accumulation (bit-exact with SciPy)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        import bench_match
        return bench_match.main(args)
    if args.workload == "c4":
        import bench_batch
        return bench_batch.main(args)

    import torch
    import torch.distributed as dist
    import synth
    from mad_b200 import pipeline as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's log (NCCL_DEBUG is left to the caller) goes to a file or stderr, never to stdout: rank 0 prints ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: each rank its own map (independent units) ---------------------------------
    grid_h, comps_h = synth.c2_inputs(rank)
    n_vox = int(grid_h.size)
    grid_pin = torch.from_numpy(grid_h).pin_memory()
    grid_d = grid_pin.to(dev)
    exact = bool(args.exact)

    # component descriptor sets: resident in HBM before the timed region (MaD caches them: dsc_db/).
    # All components are stacked into ONE hi set, so one matching launch per step serves the six
    # subunits (pairs carry the global hi row; comp_offs maps rows back to components).
    comp_sets = []
    for c in comps_h:
        _, _, _, dsc = P.describe_struct(c, exact_f64=exact)
        comp_sets.append(P.DescriptorSet(dsc))
    hi_all, comp_offs = P.concat_sets(comp_sets)
    torch.cuda.synchronize()
    stage = P.HostStage()

    def step_device():
        sp, kp, ori, dsc = P.describe_struct(grid_d, exact_f64=exact)
        lo = P.DescriptorSet(dsc)
        pairs = P.match_threshold(hi_all, lo, 0.6, impl=args.match_impl)
        return sp, kp, ori, dsc, [pairs]

    # e2e goes through the streaming API a user of a batch of maps calls (pipeline.MapStream): every step uploads ITS map
    # from pinned host memory and downloads ITS results (descriptors, keypoints, orientations, pair lists); the upload of
    # step i+1 and the download of step i-1 overlap the kernels of step i.
    stream_api = P.MapStream(hi=hi_all, cc=0.6, exact_f64=exact, match_impl=args.match_impl)

    def run_e2e(n_steps):
        prev, out = None, None
        nxt = stream_api.upload(grid_pin)
        for s_i in range(n_steps):
            cur = nxt
            nxt = stream_api.upload(grid_pin) if s_i + 1 < n_steps else None
            ticket = stream_api.submit(cur)
            if prev is not None:
                out = stream_api.result(prev)
            prev = ticket
        out = stream_api.result(prev)
        return list(out.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step_device()
    sp, kp, ori, dsc, pairs = res
    K, D = len(kp), len(ori)
    n_pairs = int(sum(p[0].numel() for p in pairs))
    dims = sp.dims
    V0 = dims[0][0] * dims[0][1] * dims[0][2]
    V1 = dims[1][0] * dims[1][1] * dims[1][2]
    del res, sp, kp, ori, dsc, pairs

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)

    # ---- timed region: K steps, device-resident input ------------------------------------------
    barrier()
    P.profile_enable(True)
    l0 = P.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    launches = P.launch_count() - l0
    recs = P.profile_records()
    P.profile_enable(False)

    # ---- e2e: host buffers in and out --------------------------------------------------------------
    out = run_e2e(2)
    d2h = int(sum(t.numel() * t.element_size() for t in out))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.summary(t_wall0, t_wall1) if rank == 0 else None
    if rank == 0:
        sampler.stop()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel accounting + roofline of the dominant kernel ---------------------------------
    peaks = {}
    pk_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
    tc_src = ("2 x measured bf16 burst (MEASURED_PEAKS.json bf16_tflops): uint8 x uint8 -> int32 tcgen05.mma.kind::i8 "
              "runs at twice the 16-bit rate" if "bf16_tflops" in peaks else "2 x fallback bf16 1.59 PFLOP/s")

    by_name = {}
    for nm, t in recs:
        by_name.setdefault(nm, []).append(t)
    total_kernel_ms = sum(sum(v) for v in by_name.values())
    kernels = sorted(((nm, sum(v), len(v)) for nm, v in by_name.items()), key=lambda x: -x[1])

    # algorithmic bytes / ops of each kernel PER STEP (SURVEY 8d convention): kernels that run
    # once per octave (or, for the slab-pipelined Y / Z passes, once per slab of x planes) are accounted over all
    # their launches of a step: bytes of both octaves / summed duration.
    M_hi = hi_all.rows
    VV = V0 + V1
    alg = {
        "log_pass_x_kernel": ("hbm", 12 * VV), "log_pass_y_kernel": ("hbm", 20 * VV), "log_pass_z_kernel": ("hbm", 20 * VV),
        "log_pass_yz_kernel": ("hbm", 16 * VV),          # fused Y + Z: reads P0, Q0, writes LoG, Gauss
        "gradient_kernel": ("hbm", 16 * VV + 4 * VV), "detect_peaks_kernel": ("hbm", 4 * VV),
        "spline_up_z_kernel": ("hbm", 4 * V1 + 8 * 2 * V1), "spline_up_x_kernel": ("hbm", 8 * 2 * V1 + 8 * 4 * V1),
        "spline_up_y_kernel": ("hbm", 8 * 4 * V1 + 4 * V0), "pad3d_kernel": ("hbm", 4 * n_vox + 4 * V1),
        "orient_kernel": ("hbm", 58956 * K), "describe_kernel": ("hbm", 51200 * D),
        "match_u8_pairs_kernel": ("tensor", 2.0 * M_hi * D * 1024),
    }
    traffic = {}
    tr_path = os.path.join(REPO, "profiles", "ncu_traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path))

    def kernel_roofline(name):
        if name not in alg or name not in by_name:
            return None
        bound, amount = alg[name]
        step_ms_k = float(sum(by_name[name])) / args.steps
        if bound == "hbm":
            achieved, pk, unit, src = amount / (step_ms_k * 1e-3) / 1e9, hbm_peak, "GB/s", hbm_src
        else:
            achieved, pk, unit, src = amount / (step_ms_k * 1e-3) / 1e12, 2.0 * bf16_peak, "TFLOP/s", tc_src
        return {"bound": bound, "kernel": name, "achieved": achieved, "peak": pk, "unit": unit, "frac": achieved / pk,
                "traffic": traffic.get(name), "peak_source": src, "algorithmic_per_step": amount,
                "launches_per_step": len(by_name[name]) // args.steps, "avg_launch_ms": step_ms_k / max(1, len(by_name[name]) // args.steps),
                "ms_per_step": step_ms_k, "share_of_kernel_time": sum(by_name[name]) / total_kernel_ms}

    roofline = None
    for nm, _, _ in kernels:                                  # dominant kernel = largest share with a model
        roofline = kernel_roofline(nm)
        if roofline:
            break
    if roofline and roofline["kernel"] == "log_pass_yz_kernel":
        # This kernel moves exactly its algorithmic bytes (ncu traffic == 16 B/voxel) but is bound by FP64 issue:
        # SciPy's float64 line accumulation costs 43 (Y, computed for 128 columns per 112 kept) + 60 (Z) FP64
        # operations per voxel.  Peak = DFMA / DADD / DMUL issue rate measured on this part
        # (scripts/ubench/fp64_rate.cu, profiles/r01_fp64_issue_rate.txt: 63 lanes/clk/SM = 18.3 T lane-ops/s).
        ops = (43.0 * 128.0 / 112.0 + 60.0) * VV
        t_s = roofline["ms_per_step"] * 1e-3
        roofline["fp64_issue"] = {"lane_ops_per_voxel": round(ops / VV, 1), "achieved": ops / t_s / 1e12, "peak": 18.3,
                                  "unit": "T lane-ops/s", "frac": ops / t_s / 1e12 / 18.3}
    per_kernel = {}
    for nm, _, _ in kernels:
        r = kernel_roofline(nm)
        if r:
            per_kernel[nm] = {"ms": round(r["ms_per_step"], 4), "achieved": round(r["achieved"], 1), "unit": r["unit"],
                              "frac": round(r["frac"], 4)}
    map_bytes = 36 * (V0 + V1) + 58956 * K + 51200 * D
    step_ms = ms / args.steps
    describe_ms = sum(sum(v) for k, v in by_name.items() if "match" not in k and "pairs" not in k and "dsc_prepare" not in k
                      and k != "cub_radix_sort_pairs") / args.steps
    match_ms = total_kernel_ms / args.steps - describe_ms
    value = world * n_vox / (step_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 grids, f64 line accumulation, u8 x u8 -> s32 matching" if exact else "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(), "grids": {"up": list(dims[0]), "base": list(dims[1])},
                   "keypoints": K, "oriented_features": D, "component_descriptors": M_hi, "pairs": n_pairs,
                   "parallelism": "1 map per GPU, no collective" if world > 1 else "1 GPU",
                   "l2": "no explicit flush: per-step working set %.1f GB >> 126 MB L2" % (40.0 * (V0 + V1) / 1e9),
                   "exact_f64": exact},
        "e2e": {"value": world * n_vox / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(grid_pin.numel() * 4), "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "path_roofline": {"algorithmic_bytes_per_map": map_bytes, "describe_kernels_ms": describe_ms,
                          "achieved": map_bytes / (describe_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": map_bytes / (describe_ms * 1e-3) / 1e9 / hbm_peak},
        "describe_voxels_per_s": world * n_vox / (describe_ms * 1e-3),
        "matches_per_s": world * M_hi * D / (match_ms * 1e-3 + 1e-12),
        "per_kernel": per_kernel,
        "kernels_ms_per_step": {nm: round(t / args.steps, 4) for nm, t, _ in kernels},
    }
    if not args.no_cpu_baseline and world == 1:
        procs = 1
        v, s_per, info = cpu_run(grid_h, C2["voxelsp"], procs, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
                                "sample": "one %d^3 occupancy-matched crop of the same map (K=%d, D=%d) through describe + match, "
                                          "oracle/mad_oracle.py (vectorised NumPy/SciPy port, bit-exact with the reference on the "
                                          "fixtures; the reference itself is ~10x slower, BASELINE.md)"
                                          % (CPU_SAMPLE_SIDE, info[0][1], info[0][2]),
                                "seconds": s_per}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
