"""bench.py --workload c4: a batch of 64 conformational-snapshot maps (BASELINE.json configs[3]).

64 snapshots of the C1 component (9 000-atom random walk, 4 A, 1 A/voxel -> ~96^3 maps, every atom
displaced by N(0, 1.5 A), seeds 100..163) are independent units (the loops of mad/MaD.py:143-162,
178-189): snapshot i belongs to rank i mod G, every rank runs the whole describe chain on its maps,
no data-path collective; the per-map descriptor counts are all-gathered at the end of the step.
`value` = voxels of the whole batch / max-over-ranks time; total work is fixed ("strong" scaling).
"""
import json
import os

import numpy as np


def main(args):
    import torch
    import torch.distributed as dist
    import synth
    from mad_b200 import pipeline as P
    from mad_b200 import parallel as par

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
    n_maps = int(os.environ.get("MAD_C4_MAPS", "64"))
    base = synth.random_walk_atoms(9000, 85.0, 1)
    mine = par.assign_units(n_maps, rank, world)
    grids = []
    for i in mine:
        grids.append(synth.c4_snapshot(i, base))
    pins = [torch.from_numpy(g).pin_memory() for g in grids]
    devs = [p.to(dev) for p in pins]
    n_vox_total = n_maps * 96 ** 3
    P.describe_struct(devs[0])            # one-time table / mask initialisation before the worker threads start
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Small maps are launch- and host-latency-bound (a 96^3 map is ~30 launches and 3 count read-backs
    # for ~0.3 ms of HBM traffic): NW host threads, each with its own CUDA stream, work on different
    # maps so that one map's host round trips hide behind another map's kernels.
    from concurrent.futures import ThreadPoolExecutor
    NW = int(os.environ.get("MAD_C4_STREAMS", "4"))
    streams = [torch.cuda.Stream() for _ in range(NW)]
    stages = [P.HostStage() for _ in range(NW)]
    pool = ThreadPoolExecutor(NW)

    def work(w, host):
        torch.cuda.set_device(local)
        out = []
        with torch.cuda.stream(streams[w]):
            for j in range(w, len(pins), NW):
                g = pins[j].to(dev, non_blocking=True) if host else devs[j]
                sp, kp, ori, dsc = P.describe_struct(g)
                out.append((j, len(ori)))
                if host:
                    stages[w].fetch("dsc%d" % j, dsc)
                    stages[w].fetch("kp%d" % j, kp.table[:len(kp)])
            streams[w].synchronize()
        return out

    def step(host):
        main = torch.cuda.current_stream()
        for s_ in streams:
            s_.wait_stream(main)
        res_ = [f.result() for f in [pool.submit(work, w, host) for w in range(NW)]]
        for s_ in streams:
            main.wait_stream(s_)
        counts = [c for _, c in sorted(x for r_ in res_ for x in r_)]
        c = torch.tensor(counts, dtype=torch.int64, device=dev)
        return par.gather_varlen(c)

    for _ in range(max(args.warmup, 3)):
        res = step(False)
    barrier()
    l0 = P.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(False)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = P.launch_count() - l0
    step(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        d_total = int(sum(int(x.sum()) for x in res))
        step_ms = ms / args.steps
        line = {
            "metric": "voxels/sec scale-space+detect+describe (batch of maps)", "value": n_vox_total / (step_ms * 1e-3),
            "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 grids, f64 line accumulation",
            "data": "synthetic",
            "config": {"workload": "C4: %d snapshot maps of 96^3 (9000 atoms, 4 A, sigma 1.5 A displacements), map i -> rank i mod %d, "
                                   "no collective on the data path; %d host threads / CUDA streams per rank" % (n_maps, world, NW),
                       "oriented_features_total": d_total,
                       "l2": "small maps: a map's working set (~0.6 GB) exceeds the 126 MB L2"},
            "e2e": {"value": n_vox_total / (ms_e2e / args.steps * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": int(sum(p.numel() * 4 for p in pins)), "d2h_bytes_per_step": int(d_total * 2048 // max(world, 1)),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "roofline": None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
