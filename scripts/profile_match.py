"""Matching only, for ncu: python scripts/profile_match.py [M] [N] [mode]  (mode: pairs | topk)"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 6784
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40320
mode = sys.argv[3] if len(sys.argv) > 3 else "pairs"
cc = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
base = synth.synthetic_descriptors(min(N, 8192), 7)
rng = np.random.default_rng(11)
reps = (N + len(base) - 1) // len(base)
lo = np.concatenate([np.roll(base, int(rng.integers(0, 1024)), axis=1) if r else base for r in range(reps)])[:N]
hi = synth.synthetic_descriptors(M, 8, noisy_copy_of=lo[:8192])
dl, dh = P.DescriptorSet(lo), P.DescriptorSet(hi)
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if mode == "pairs":
        r = P.match_threshold(dh, dl, cc)
        n = r[0].numel()
    else:
        r = P.match_topk(dh, dl, 8)
        n = r[0].numel()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s M=%d N=%d out=%d  %.3f ms  %.1f TOP/s" % (mode, M, N, n, dt * 1e3, 2.0 * M * N * 1024 / dt / 1e12))
