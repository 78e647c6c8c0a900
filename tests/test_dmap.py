"""Row a0 (Dmap container): the oracle restatement against the reference-generated fixture (CPU), and the
device implementation (mad_b200.Dmap, through the C ABI) against both (GPU)."""
import os

import numpy as np
import pytest

import helpers as H

CASES = (("a", dict(isovalue=0.3)), ("b", dict(isovalue=0.0, normalize=False, pad=3)), ("c", dict(isovalue=50.0)))


def _raw_grid(g):
    txt = bytes(g["sit_text"]).decode()
    lines = txt.split("\n")
    header = lines[0].replace("  ", "").split(" ")
    voxsp, xi, yi, zi = [float(x) for x in header[:4]]
    xb, yb, zb = [int(x) for x in header[4:]]
    vals = np.array(" ".join(lines[2:]).split(), dtype=np.float64).astype(np.float32)
    return np.reshape(vals, (xb, yb, zb), order="F"), voxsp, (xi, yi, zi)


@pytest.mark.parametrize("tag,kw", CASES)
def test_oracle_dmap_equals_reference(tag, kw):
    import dmap_oracle as do
    g = H.golden("dmap")
    grid, voxsp, origin = _raw_grid(g)
    d = do.construct(grid, voxsp, origin, **kw)
    assert np.array_equal(d.grid3d, g[tag + "_ctor"])
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_ctor_meta"])
    do.reduce_void(d)
    assert np.array_equal(d.grid3d, g[tag + "_void"])
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_void_meta"])
    do.pad_grid(d, 2)
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_pad_meta"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,kw", CASES)
def test_device_dmap_equals_reference(tmp_path, tag, kw):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mad_b200.Dmap import Dmap
    g = H.golden("dmap")
    path = os.path.join(str(tmp_path), "case.sit")
    open(path, "wb").write(bytes(g["sit_text"]))
    d = Dmap(path, **kw)
    assert d.device_grid().is_cuda
    assert np.array_equal(d.grid3d, g[tag + "_ctor"])
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_ctor_meta"])
    d.reduce_void()
    assert np.array_equal(d.grid3d, g[tag + "_void"])
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_void_meta"])
    d.pad_grid(2)
    assert np.array_equal([d.voxsp, d.xi, d.yi, d.zi, d.xb, d.yb, d.zb], g[tag + "_pad_meta"])
    assert tuple(d.grid3d.shape) == (d.xb, d.yb, d.zb)
    # MRC round trip with the reference's conventions (int-truncated origin on read, mad/Dmap.py:38)
    out = os.path.join(str(tmp_path), "rt.mrc")
    d.write_to_mrc(out)
    e = Dmap(out, isovalue=0.0, normalize=False)
    assert np.array_equal(e.grid3d, d.grid3d)
    assert (e.xb, e.yb, e.zb) == (d.xb, d.yb, d.zb) and abs(e.voxsp - d.voxsp) < 1e-6
    assert (e.xi, e.yi, e.zi) == (int(d.xi), int(d.yi), int(d.zi))


@pytest.mark.gpu
def test_dmap_feeds_mapspace_on_device():
    """Dmap.device_grid() -> pipeline.build_space without a host round trip equals the array path."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import synth
    from mad_b200.Dmap import Dmap
    from mad_b200 import pipeline as P
    g = H.golden("tiny")
    grid = synth.dequantise_u16(g["input_q"])
    d = Dmap.from_array(grid, float(g["voxelsp"]), tuple(g["origin"]), isovalue=0.0, normalize=True)
    ref = grid / np.amax(grid)
    assert np.array_equal(d.grid3d, ref)
    a = P.build_space(d.device_grid())
    b = P.build_space(ref)
    assert torch.equal(a.logs[0], b.logs[0]) and torch.equal(a.grad4[1], b.grad4[1])
