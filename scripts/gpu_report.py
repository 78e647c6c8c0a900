"""Stage-by-stage parity report of the CUDA path against the oracle (run on a B200 box).

    python scripts/gpu_report.py [case ...]  > gpurun_out/report.txt
"""
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")]
import helpers as H  # noqa: E402
import mad_oracle as mo  # noqa: E402
from mad_b200 import pipeline as P  # noqa: E402


def sync_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


def report(case):
    print("=" * 30, case)
    g = H.golden(case)
    grid, osp, okp, oori, odsc, tab_o = H.oracle_case(case)
    sp, dt = sync_time(lambda: P.build_space(grid))
    print("build_space %.2f ms" % (dt * 1e3))
    print(H.compare_dense("up_grid", sp.grids[0].cpu().numpy(), osp["grid_list"][0]))
    for o in range(2):
        print(H.compare_dense("log%d" % o, sp.logs[o].cpu().numpy(), osp["map_space"][o]))
        print(H.compare_dense("gauss%d" % o, sp.gauss[o].cpu().numpy(), osp["gauss_list"][o]))
        print(H.compare_dense("grad%d" % o, sp.grad4[o].cpu().numpy()[..., :3], osp["grad_list"][o]))
    # stage-isolated: feed the oracle's up grid to the LoG stage
    sp2 = P.Space()
    # detect on our own space
    kp, dt = sync_time(lambda: P.detect(sp))
    hk = kp.host()
    print("detect %.2f ms  K=%d (oracle %d)" % (dt * 1e3, len(kp), len(okp["oct"])))
    same_n = len(kp) == len(okp["oct"])
    if same_n:
        print("  coords equal:", np.array_equal(hk["vox"], okp["coords"]), " oct equal:", np.array_equal(hk["oct"], okp["oct"]))
        v = float(g["voxelsp"])
        vs = np.where(hk["oct"] == 0, v / 2, v)[:, None]
        org = np.asarray(g["origin"], dtype=np.float64) - 9 * v
        sub = (hk["vox"].astype(np.float64) + hk["off"].astype(np.float64)) * vs + org
        print("  subvoxel max abs diff (A):", np.abs(sub - okp["subv_map_coords"]).max(),
              " val equal:", np.array_equal(hk["val"], okp["val"]))
    else:
        a = set(H.keypoint_keys(hk["oct"], hk["vox"]))
        b = set(H.keypoint_keys(okp["oct"], okp["coords"]))
        print("  only gpu:", sorted(a - b)[:10], " only oracle:", sorted(b - a)[:10])
    ori, dt = sync_time(lambda: P.orient(sp, kp))
    ho = ori.host()
    print("orient %.2f ms  D=%d (oracle %d)" % (dt * 1e3, len(ori), len(oori["kp"])))
    if len(ori) == len(oori["kp"]):
        print("  kp equal:", np.array_equal(ho["kp"], oori["kp"]), " main equal:", np.array_equal(ho["main"], oori["main"]),
              " sec equal:", np.array_equal(ho["sec"], oori["sec"]))
    else:
        a = set(zip(ho["kp"].tolist(), ho["main"].tolist(), ho["sec"].tolist()))
        b = set(zip(oori["kp"].tolist(), oori["main"].tolist(), oori["sec"].tolist()))
        print("  only gpu:", sorted(a - b)[:10], " only oracle:", sorted(b - a)[:10], "common", len(a & b))
    dsc, dt = sync_time(lambda: P.describe(sp, kp, ori))
    hd = dsc.cpu().numpy()
    print("describe %.2f ms" % (dt * 1e3))
    if hd.shape == odsc.shape:
        bad = np.nonzero((hd != odsc).any(1))[0]
        print("  descriptors differing: %d of %d; total abs diff %d" % (len(bad), len(hd), np.abs(hd.astype(int) - odsc).sum()))
        print("  golden crc rows equal:", int((H.crc_rows(hd) == g["dsc_crc32"]).sum()), "of", len(g["dsc_crc32"]))
    # stage-isolated orient/describe on the ORACLE's keypoints (robust to upstream flips)
    karr = np.zeros(len(okp["oct"]), dtype=P.KEYPOINT_DTYPE)
    karr["vox"] = okp["coords"]; karr["oct"] = okp["oct"]; karr["accepted"] = 1
    kp_o = P.keypoints_from_host(karr, sp.grad4[0].device)
    ori2 = P.orient(sp, kp_o)
    ho2 = ori2.host()
    a = set(zip(ho2["kp"].tolist(), ho2["main"].tolist(), ho2["sec"].tolist()))
    b = set(zip(oori["kp"].tolist(), oori["main"].tolist(), oori["sec"].tolist()))
    print("isolated orient: gpu %d oracle %d common %d; in-order equal: %s" % (len(a), len(b), len(a & b),
          len(ho2) == len(oori["kp"]) and np.array_equal(ho2["main"], oori["main"]) and np.array_equal(ho2["sec"], oori["sec"])))
    oarr = np.zeros(len(oori["kp"]), dtype=P.ORIENTED_DTYPE)
    oarr["kp"] = oori["kp"]; oarr["main"] = oori["main"]; oarr["sec"] = oori["sec"]
    ori_o = P.oriented_from_host(oarr, sp.grad4[0].device)
    hd2 = P.describe(sp, kp_o, ori_o).cpu().numpy()
    bad = np.nonzero((hd2 != odsc).any(1))[0]
    print("isolated describe: differing %d of %d; total abs diff %d" % (len(bad), len(hd2), np.abs(hd2.astype(int) - odsc).sum()))
    return hd


def report_match():
    print("=" * 30, "match")
    ghi, glo, gm = H.golden("pair_hi"), H.golden("pair_lo"), H.golden("pair_match")
    hi, lo = ghi["dsc"], glo["dsc"]
    for impl in (1, 0):
        try:
            (ph, pl, sc), dt = sync_time(lambda: P.match_threshold(hi, lo, 0.6, impl=impl))
        except Exception as e:  # noqa
            print("impl", impl, "failed:", e)
            continue
        pairs = np.stack([ph.cpu().numpy(), pl.cpu().numpy()], 1)
        print("impl %d: %.2f ms  pairs %d (golden %d) equal: %s  max score diff %.3g" % (
            impl, dt * 1e3, len(pairs), len(gm["pairs"]), np.array_equal(pairs, gm["pairs"]),
            np.abs(sc.cpu().numpy() - gm["scores"]).max() if len(pairs) == len(gm["pairs"]) else float("nan")))
        idx, val = P.match_topk(hi, lo, 8, impl=impl)
        oi, ov = mo.match_topk(hi, lo, 8)
        print("   topk idx equal to oracle(stable argsort on reference preds): %.4f" % (idx.cpu().numpy() == oi).mean())


if __name__ == "__main__":
    cases = sys.argv[1:] or ["tiny", "small"]
    print(torch.cuda.get_device_name(0))
    for c in cases:
        report(c)
    report_match()
