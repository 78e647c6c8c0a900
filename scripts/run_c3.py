"""BASELINE config 3: one 512^3 map at 10 A (2.5 A/voxel, 20 random-walk subunits) through a1-a12 on one GPU; prints
keypoints / features, the peak memory and the per-kernel times.  python scripts/run_c3.py [n=512]"""
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t0 = time.perf_counter()
grid = synth.assembly_map(n, 10.0, 2.5, 20, 60000, 300)
print("input %s built in %.1f s, occupancy %.3f" % (grid.shape, time.perf_counter() - t0, float((grid > 0.05).mean())), flush=True)
g = torch.from_numpy(grid).cuda()
for it in range(3):
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    if it == 2:
        P.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sp, kp, ori, dsc = P.describe_struct(g)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("run %d: %.2f ms  %.3f G voxels/s  K=%d D=%d  peak %.1f GB" % (it, ms, n ** 3 / ms / 1e6, len(kp), len(ori),
                                                                       torch.cuda.max_memory_allocated() / 1e9), flush=True)
    del sp, kp, ori, dsc
recs = P.profile_records()
tot = {}
for nm, t in recs:
    tot[nm] = tot.get(nm, 0.0) + t
print(json.dumps({k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}))
