"""Times mad_match_topk at C5 size (or MAD_C5_ROWS) on one GPU: kernel milliseconds from the library's per-launch events."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import numpy as np, torch, synth
from mad_b200 import pipeline as P
M = int(os.environ.get("MAD_C5_ROWS", "100000")); N = int(os.environ.get("MAD_C5_COLS", str(M)))
hi_h, lo_h = synth.c5_descriptor_sets(M, N)
hi, lo = P.DescriptorSet(hi_h), P.DescriptorSet(lo_h)
for _ in range(3): P.match_topk(hi, lo, 8)
torch.cuda.synchronize()
P.profile_enable(True)
for _ in range(5): P.match_topk(hi, lo, 8)
torch.cuda.synchronize()
r = {}
for nm, t in P.profile_records(): r.setdefault(nm, []).append(t)
print("M=%d N=%d dbg=%s slack=%s " % (M, N, os.environ.get("MAD_TOPK_DBG"), os.environ.get("MAD_TOPK_SLACK")) + "  ".join("%s %.3f ms per call (%d launches)" % (k, np.sum(v) / 5, len(v) // 5) for k, v in r.items()))
