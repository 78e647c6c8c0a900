"""Where the end-to-end step spends its time (CUDA events between the phases of bench.py's step_e2e)."""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle")]
import synth  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402
import bench  # noqa: E402

cfg = dict(bench.C2)
grid_h, comps_h = synth.assembly_with_components(**cfg)
dev = torch.device("cuda", 0)
grid_pin = torch.from_numpy(grid_h).pin_memory()
comp_sets = [P.DescriptorSet(P.describe_struct(c)[3]) for c in comps_h]
hi_all, offs = P.concat_sets(comp_sets)
stage = P.HostStage()


def step(rec):
    ev = []

    def mark(name):
        if rec:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append((name, e, time.perf_counter()))
    mark("start")
    g = grid_pin.to(dev, non_blocking=True)
    mark("h2d")
    sp = P.build_space(g, keep_gauss=False)
    mark("build_space")
    kp = P.detect(sp)
    mark("detect")
    ori = P.orient(sp, kp)
    mark("orient")
    dsc = P.describe(sp, kp, ori)
    mark("describe")
    mode = os.environ.get("E2E_MODE", "overlap")
    out = []
    if mode == "overlap":
        out = [stage.fetch("dsc", dsc, overlap=True), stage.fetch("kp", kp.table[:len(kp)], overlap=True),
               stage.fetch("ori", ori.table[:len(ori)], overlap=True)]
    lo = P.DescriptorSet(dsc)
    mark("prepare")
    ph, pl, sc = P.match_threshold(hi_all, lo, 0.6)
    mark("match")
    if mode == "after":
        out = [stage.fetch("dsc", dsc), stage.fetch("kp", kp.table[:len(kp)]), stage.fetch("ori", ori.table[:len(ori)])]
    out += [stage.fetch("ph", ph), stage.fetch("pl", pl), stage.fetch("sc", sc)]
    mark("d2h_pairs_issued")
    stage.sync()
    mark("sync")
    return ev


for _ in range(4):
    step(False)
torch.cuda.synchronize()
t0 = time.perf_counter()
ev = step(True)
torch.cuda.synchronize()
print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3))
for (n0, e0, h0), (n1, e1, h1) in zip(ev, ev[1:]):
    print("%-18s gpu %.3f ms   host %.3f ms" % (n1, e0.elapsed_time(e1), (h1 - h0) * 1e3))
print("total gpu %.3f ms" % ev[0][1].elapsed_time(ev[-1][1]))
