import os, sys, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "scripts")]
os.environ.setdefault("E2E_MODE", "overlap")
import e2e_breakdown as E  # runs its own warm-up + one recorded step
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    E.step(False)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    d = e.time_range.end - e.time_range.start
    if d > 30:
        print("%9.1f us  +%8.1f us  %s" % (e.time_range.start - t0, d, e.name[:70]))
