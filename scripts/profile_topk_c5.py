"""The top-8 launch of config C5 (100 000 x 100 000) inside a cudaProfilerStart/Stop range, for ncu --profile-from-start off."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

M = int(os.environ.get("MAD_C5_ROWS", "100000"))
hi_h, lo_h = synth.c5_descriptor_sets(M, M)
hi, lo = P.DescriptorSet(hi_h), P.DescriptorSet(lo_h)
for _ in range(2):
    P.match_topk(hi, lo, 8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
P.match_topk(hi, lo, 8)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
