"""Stage-isolated diagnosis of the scale-space kernels against the oracle (run on a B200 box).

    python scripts/diag_space.py > gpurun_out/diag_space.txt
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")]
import mad_oracle as mo  # noqa: E402
import synth  # noqa: E402
from scipy import ndimage as ndi  # noqa: E402
from mad_b200 import pipeline as P, tables, _lib  # noqa: E402
from mad_b200._lib import call  # noqa: E402


def where(name, got, ref):
    neq = got != ref
    n = int(neq.sum())
    if n == 0:
        print("   %-8s equal" % name)
        return
    idx = np.argwhere(neq)
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    print("   %-8s MISMATCH %d of %d (%.4f%%) max abs %.3e (max ref %.3e)" % (name, n, neq.size, 100.0 * n / neq.size, d.max(), np.abs(ref).max()))
    print("            index range per axis: min %s max %s ; first %s" % (idx.min(0).tolist(), idx.max(0).tolist(), idx[:4].tolist()))
    for ax in range(idx.shape[1] if idx.shape[1] <= 3 else 3):
        vals, cnts = np.unique(idx[:, ax], return_counts=True)
        top = np.argsort(-cnts)[:6]
        print("            axis %d hot indices: %s" % (ax, [(int(vals[t]), int(cnts[t])) for t in top]))


def log_gauss_direct(grid):
    dev = torch.device("cuda")
    g = torch.from_numpy(np.ascontiguousarray(grid)).to(dev)
    gx, gy, gz = g.shape
    rad = tables.gaussian_radius(2)
    w0 = np.ascontiguousarray(tables.gaussian_weights(2, 0, rad))
    w2 = np.ascontiguousarray(tables.gaussian_weights(2, 2, rad))
    ws_bytes = _lib.lib.mad_log_gauss_workspace_bytes(gx, gy, gz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    lg = torch.empty_like(g)
    gs = torch.empty_like(g)
    call("mad_log_gauss", C.c_void_p(g.data_ptr()), gx, gy, gz, w0.ctypes.data_as(C.c_void_p), w2.ctypes.data_as(C.c_void_p), rad,
         C.c_float(4.0), C.c_void_p(lg.data_ptr()), C.c_void_p(gs.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, 1,
         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    gr = torch.empty((gx, gy, gz, 4), dtype=torch.float32, device=dev)
    call("mad_gradient", C.c_void_p(gs.data_ptr()), gx, gy, gz, C.c_void_p(gr.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return lg.cpu().numpy(), gs.cpu().numpy(), gr.cpu().numpy()[..., :3]


def case(name, grid):
    print("=" * 20, name, grid.shape)
    osp = mo.build_space(grid)
    sp = P.build_space(grid, keep_gauss=True)
    where("up_grid", sp.grids[0].cpu().numpy(), osp["grid_list"][0])
    for o in range(2):
        print("  octave %d, whole chain:" % o)
        where("log", sp.logs[o].cpu().numpy(), osp["map_space"][o])
        where("gauss", sp.gauss[o].cpu().numpy(), osp["gauss_list"][o])
        where("grad", sp.grad4[o].cpu().numpy()[..., :3], osp["grad_list"][o])
        print("  octave %d, LoG/Gauss/grad fed with the ORACLE's grid:" % o)
        lg, gs, gr = log_gauss_direct(osp["grid_list"][o])
        where("log", lg, osp["map_space"][o])
        where("gauss", gs, osp["gauss_list"][o])
        where("grad", gr, osp["grad_list"][o])
    # zero map
    z = np.zeros_like(grid)
    spz = P.build_space(z)
    print("  zero map: up any=%s log any=%s/%s nan=%s" % (bool(spz.grids[0].any()), bool(spz.logs[0].any()), bool(spz.logs[1].any()),
                                                          bool(torch.isnan(spz.logs[0]).any())))


def blob(shape, seed):
    rng = np.random.default_rng(seed)
    g = np.zeros(shape, dtype=np.float32)
    x, y, z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    for _ in range(8):
        p = [rng.integers(3, s - 3) for s in shape]
        g += np.exp(-((x - p[0]) ** 2 + (y - p[1]) ** 2 + (z - p[2]) ** 2) / 6.0).astype(np.float32)
    return g / g.max()


if __name__ == "__main__":
    case("ragged 20x22x24", blob((20, 22, 24), 5))
    case("ragged 31x17x23", blob((31, 17, 23), 6))
    case("cube 52", blob((52, 52, 52), 7))
    g = np.load(os.path.join(REPO, "tests", "golden", "pair_lo.npz"))
    case("pair_lo", synth.dequantise_u16(g["input_q"]))
    g = np.load(os.path.join(REPO, "tests", "golden", "c1.npz"))
    case("c1", synth.dequantise_u16(g["input_q"]))
