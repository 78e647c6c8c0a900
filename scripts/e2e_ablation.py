"""Where the end-to-end time of the C2 step goes: MapStream with / without the upload and the download.
python scripts/e2e_ablation.py [steps=20]"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
import bench  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
grid_h, comps_h = synth.assembly_with_components(**dict(bench.C2))
sets = []
for c in comps_h:
    _, _, _, dsc = P.describe_struct(c)
    sets.append(P.DescriptorSet(dsc))
hi_all, _ = P.concat_sets(sets)
pin = torch.from_numpy(grid_h).pin_memory()
dev_grid = pin.cuda()


def run(download, upload, n):
    ms = P.MapStream(hi=hi_all, cc=0.6, download=download)
    ready = torch.cuda.Event()
    ready.record()
    prev = None
    nxt = ms.upload(pin) if upload else (dev_grid, ready, None)
    for i in range(n):
        cur = nxt
        nxt = (ms.upload(pin) if upload else (dev_grid, ready, None)) if i + 1 < n else None
        t = ms.submit(cur)
        if prev is not None:
            ms.result(prev)
        prev = t
    ms.result(prev)


for name, dl, ul in (("device only", False, False), ("upload", False, True), ("download", True, False), ("both", True, True)):
    run(dl, ul, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(dl, ul, steps)
    e1.record()
    torch.cuda.synchronize()
    print("%-12s %.3f ms per map" % (name, e0.elapsed_time(e1) / steps), flush=True)
