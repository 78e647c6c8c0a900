"""bench.py --workload c5: all-pairs descriptor matching (BASELINE.json configs[4]).

M = N = 100 000 int16[1024] descriptors (multinomial/Dirichlet blocks, half of the hi
rows are noisy copies of lo rows), per-row top-8 by (score desc, index asc).  The lo (reference)
axis is cut into one contiguous shard per GPU; every rank computes its local top-k on the uint8
tcgen05 kernel with global indices, then NCCL all_gather + k-way merge (mad_b200/parallel.py).
`value` = M*N scored pairs / second, whole job; total work is fixed as N grows ("strong").
"""
import json
import os
import sys
import time

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
    M = N = int(os.environ.get("MAD_C5_ROWS", "100000"))
    k = 8
    hi_h, lo_h = synth.c5_descriptor_sets(M, N)          # identical inputs on every rank (same seeds)
    s, e = par.shard_bounds(N, world)[rank]
    hi_pin = torch.from_numpy(hi_h).pin_memory()
    lo_pin = torch.from_numpy(np.ascontiguousarray(lo_h[s:e])).pin_memory()
    hi = P.DescriptorSet(hi_pin.to(dev))
    lo = P.DescriptorSet(lo_pin.to(dev))
    stage = P.HostStage()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return par.match_topk_sharded(hi, lo, k, s)

    def step_e2e():
        h = P.DescriptorSet(hi_pin.to(dev, non_blocking=True))
        l = P.DescriptorSet(lo_pin.to(dev, non_blocking=True))
        idx, sc = par.match_topk_sharded(h, l, k, s)
        out = [stage.fetch("idx", idx), stage.fetch("sc", sc)]
        stage.sync()
        return out

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    P.profile_enable(True)
    l0 = P.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = P.launch_count() - l0
    recs = P.profile_records()
    P.profile_enable(False)
    out = step_e2e()
    d2h = int(sum(t.numel() * t.element_size() for t in out))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        peaks = {}
        pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        bf16 = float(peaks.get("bf16_tflops", 1590.0))
        kms = [t for nm, t in recs if nm == "match_u8_topk_kernel"]
        avg = float(np.mean(kms)) if kms else float("nan")
        ops = 2.0 * M * (e - s) * 1024
        step_ms = ms / args.steps
        line = {
            "metric": "descriptor matches/sec (all-pairs cosine + top-8)", "value": M * N / (step_ms * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 x u8 -> s32 (exact), f64 scores", "data": "synthetic",
            "config": {"workload": "C5: %d x %d int16[1024] descriptors, top-%d, lo axis sharded over %d GPU(s), "
                                   "NCCL all_gather + merge" % (M, N, k, world),
                       "l2": "operands %.0f MB per rank > 126 MB L2" % ((M + e - s) * 1024 / 1e6)},
            "e2e": {"value": M * N / (ms_e2e / args.steps * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": int(hi_pin.numel() * 2 + lo_pin.numel() * 2), "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "match_u8_topk_kernel", "achieved": ops / (avg * 1e-3) / 1e12,
                         "peak": 2 * bf16, "unit": "TFLOP/s", "frac": ops / (avg * 1e-3) / 1e12 / (2 * bf16), "traffic": None,
                         "peak_source": "2 x measured bf16 burst (uint8 tcgen05.mma.kind::i8 runs at twice the 16-bit rate)",
                         "avg_launch_ms": avg},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
