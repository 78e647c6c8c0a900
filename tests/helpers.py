"""Shared helpers of the parity tests (oracle side + comparison utilities)."""
import functools
import hashlib
import os
import zlib

import numpy as np

import mad_oracle as mo
import synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(REPO, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


FLUSH = 1e-10


def flushed(a):
    """|v| < FLUSH -> +0.  The spline's decaying tails inside the zero padding (|v| ~ 1e-20,
    against values of order 1) depend on the banded solver's rounding order (LAPACK gbsv in
    SciPy, a streaming Thomas recurrence on the GPU) and are outside the parity contract;
    everything at or above FLUSH is compared bit for bit."""
    a = np.ascontiguousarray(a).copy()
    a[np.abs(a) < FLUSH] = 0
    return a


def sha_flushed(a):
    return hashlib.sha256(flushed(a).tobytes()).hexdigest()


def equal_flushed(a, b):
    return np.array_equal(flushed(a), flushed(b))


def plane_crcs(a):
    """CRC32 of every x plane after flushing |v| < FLUSH (twin of oracle/gen_golden_bench.py:plane_crcs)."""
    a = np.ascontiguousarray(a)
    out = np.zeros(a.shape[0], dtype=np.uint32)
    for x in range(a.shape[0]):
        out[x] = zlib.crc32(flushed(a[x]).tobytes())
    return out


def crc_rows(dsc):
    return np.array([zlib.crc32(np.ascontiguousarray(r).tobytes()) for r in dsc], dtype=np.uint32)


@functools.lru_cache(maxsize=None)
def golden(name):
    with np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=None)
def oracle_case(name):
    """Runs the oracle on a golden case's input; returns (grid, space, kp, ori, dsc, tab_o)."""
    g = golden(name)
    grid = synth.dequantise_u16(g["input_q"])
    sp = mo.build_space(grid)
    org = np.asarray(g["origin"], dtype=np.float64) - 9 * float(g["voxelsp"])
    v = float(g["voxelsp"])
    kp = mo.detect(sp["map_space"], [v / 2, v], org)
    ori, tab_o = mo.orient(sp["grad_list"], kp)
    dsc = mo.describe(sp["grad_list"], kp, ori, tab_o)
    return grid, sp, kp, ori, dsc, tab_o


def compare_dense(name, got, ref):
    """Returns dict(mismatch fraction, max abs diff, max abs diff / max|ref|)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    diff = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    neq = (got != ref) & ~((got == 0) & (ref == 0))
    return dict(name=name, frac_neq=float(neq.mean()), max_abs=float(diff.max()),
                rel=float(diff.max() / max(np.abs(ref).max(), 1e-30)))


def keypoint_keys(octs, coords):
    return [(int(o), int(c[0]), int(c[1]), int(c[2])) for o, c in zip(octs, coords)]
