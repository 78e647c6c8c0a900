"""Per-kernel milliseconds of describe_struct on one map, run alone on one stream: `c4` (a 96^3 snapshot) or `c2` (256^3)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import numpy as np, torch, synth
from mad_b200 import pipeline as P
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
grid = synth.c4_snapshot(0) if which == "c4" else synth.c2_inputs(0)[0]
g = torch.from_numpy(grid).cuda()
for _ in range(3): P.describe_struct(g)
torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): sp, kp, ori, dsc = P.describe_struct(g)
e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / n
P.profile_enable(True)
for _ in range(n): P.describe_struct(g)
torch.cuda.synchronize()
r = {}
for nm, t in P.profile_records(): r.setdefault(nm, []).append(t)
tot = sum(sum(v) for v in r.values()) / n
print("%s: K=%d D=%d  wall %.3f ms per map, kernel sum %.3f ms, %d launches" % (which, len(kp), len(ori), wall, tot, sum(len(v) for v in r.values()) // n))
for k, v in sorted(r.items(), key=lambda kv: -sum(kv[1])): print("  %-28s %7.4f ms  x%d" % (k, sum(v) / n, len(v) // n))
