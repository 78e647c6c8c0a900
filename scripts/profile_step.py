"""One warm step of the C2 hot path inside a cudaProfilerStart/Stop range (for ncu
--profile-from-start off).  Usage: python scripts/profile_step.py [n=256] [match=1]"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
do_match = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if n == 256:
    grid = synth.c2_inputs(0)[0]                      # the bench's C2 map
else:
    grid = synth.assembly_map(n, 8.0, 2.0, 6, 5000, 10)
g = torch.from_numpy(grid).cuda()


def step():
    sp, kp, ori, dsc = P.describe_struct(g)
    if do_match:
        lo = P.DescriptorSet(dsc)
        hi = P.DescriptorSet(dsc[: max(1, dsc.shape[0] // 6)].contiguous())
        P.match_threshold(hi, lo, 0.6)
    return len(kp), len(ori)


for _ in range(2):
    print(step())
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
