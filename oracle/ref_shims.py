"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mad_b200/).

Import shims that let the *unmodified* reference (``/root/reference/mad``) be imported in
this container, where four of its third-party imports are absent:

* ``skimage.feature.peak_local_max``  (arithmetic!  call site ``mad/Detector.py:29``;
  pinned scikit-image 0.17.2 in ``requirements.txt:5``).  Restated here from the published
  0.17.x algorithm: 3x3x3 ``maximum_filter(mode='constant')``, ``image == image_max``,
  ``image > threshold_abs``, border exclusion, ``nonzero`` and "highest first" ordering.
  The *set* of peaks is certain; the order among exactly-equal intensities is not pinned by
  any reference test, so a stable sort (raster order on ties) is used -> "parity unpinned"
  for tie ORDER only.
* ``h5py``      (I/O only, ``mad/MaD.py:2``, ``mad/Descriptor.py:6``)   -> .npz-backed stub
* ``mrcfile``   (I/O only, ``mad/MapSpace.py:3``, ``mad/Dmap.py:4``)    -> .npz-backed stub
* ``matplotlib`` (plots only, ``mad/MaD.py:5``)                          -> empty stub

Usage (always in a subprocess whose cwd holds a ``mad -> /root/reference/mad`` symlink,
because ``mad/eqsp/eqsp.py:16,26`` opens its tables relative to the cwd)::

    import ref_shims; ref_shims.install()
    from mad.MapSpace import MapSpace
"""
import sys
import types

import numpy as np
from scipy import ndimage as ndi


# --------------------------------------------------------------------------- skimage
def peak_local_max(image, min_distance=1, threshold_abs=None, threshold_rel=None,
                   exclude_border=True, **_unused):
    nd = image.ndim
    if threshold_abs is None:
        threshold_abs = image.min()
    if isinstance(exclude_border, bool):
        exclude_border = min_distance if exclude_border else 0
    if image.size == 0 or np.all(image == image.flat[0]):
        return np.empty((0, nd), dtype=int)
    neighbourhood_max = ndi.maximum_filter(image, size=2 * min_distance + 1, mode="constant")
    is_peak = image == neighbourhood_max
    level = threshold_abs
    if threshold_rel is not None:
        level = max(threshold_abs, threshold_rel * image.max())
    is_peak &= image > level
    if exclude_border:
        for ax in range(nd):
            head = [slice(None)] * nd
            tail = [slice(None)] * nd
            head[ax] = slice(None, exclude_border)
            tail[ax] = slice(-exclude_border, None)
            is_peak[tuple(head)] = False
            is_peak[tuple(tail)] = False
    where = np.nonzero(is_peak)
    order = np.argsort(-image[where], kind="stable")
    return np.transpose(where)[order]


# --------------------------------------------------------------------------- mrcfile
class _Rec(object):
    pass


class _MrcHandle(object):
    """Minimal stand-in for an ``mrcfile`` object, stored as <path> (an .npz under the hood)."""

    def __init__(self, path, mode):
        self._path = path
        self._mode = mode
        self.header = _Rec()
        self.header.origin = _Rec()
        self.header.cella = _Rec()
        self.voxel_size = _Rec()
        self.data = None
        self.mode = 2
        if mode == "r":
            with np.load(path, allow_pickle=False) as z:
                self.data = z["data"]
                h = z["header"]
            (self.header.mapc, self.header.mapr, self.header.maps,
             self.header.nxstart, self.header.nystart, self.header.nzstart,
             self.header.mx, self.header.my, self.header.mz) = [int(v) for v in h[:9]]
            self.header.origin.x, self.header.origin.y, self.header.origin.z = [float(v) for v in h[9:12]]
            self.header.cella.x, self.header.cella.y, self.header.cella.z = [float(v) for v in h[12:15]]
            self.voxel_size.x = np.float32(self.header.cella.x / self.header.mx)
            self.voxel_size.y = np.float32(self.header.cella.y / self.header.my)
            self.voxel_size.z = np.float32(self.header.cella.z / self.header.mz)
        else:
            for k in ("mapc", "mapr", "maps"):
                setattr(self.header, k, {"mapc": 1, "mapr": 2, "maps": 3}[k])
            for k in ("nxstart", "nystart", "nzstart", "mx", "my", "mz"):
                setattr(self.header, k, 0)
            for k in ("x", "y", "z"):
                setattr(self.header.origin, k, 0.0)
                setattr(self.header.cella, k, 0.0)

    def set_data(self, data):
        self.data = np.asarray(data)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._mode != "r" and self.data is not None:
            hd = self.header
            h = np.array([hd.mapc, hd.mapr, hd.maps, hd.nxstart, hd.nystart, hd.nzstart,
                          hd.mx, hd.my, hd.mz, hd.origin.x, hd.origin.y, hd.origin.z,
                          hd.cella.x, hd.cella.y, hd.cella.z], dtype=np.float64)
            with open(self._path, "wb") as f:
                np.savez(f, data=self.data, header=h)


def _mrc_open(path, mode="r", **_kw):
    return _MrcHandle(path, "r" if mode.startswith("r") and "+" not in mode else "w")


def _mrc_new(path, overwrite=False, **_kw):
    return _MrcHandle(path, "w")


def write_mrc_stub(path, grid_xyz, voxelsp, origin=(0.0, 0.0, 0.0)):
    """Write ``grid[x][y][z]`` in the stub format, with the conventions of ``PDB.py:181-206``."""
    with _mrc_new(path) as m:
        m.set_data(np.ascontiguousarray(np.transpose(grid_xyz, (2, 1, 0))).astype(np.float32))
        xb, yb, zb = grid_xyz.shape
        m.header.mx, m.header.my, m.header.mz = xb, yb, zb
        m.header.origin.x, m.header.origin.y, m.header.origin.z = origin
        m.header.cella.x, m.header.cella.y, m.header.cella.z = xb * voxelsp, yb * voxelsp, zb * voxelsp


# --------------------------------------------------------------------------- h5py
class _H5File(object):
    def __init__(self, path, mode="r"):
        self._path, self._mode, self._d = path, mode, {}
        if mode.startswith("r"):
            with np.load(path, allow_pickle=False) as z:
                self._d = {k: z[k] for k in z.files}

    def create_dataset(self, name, data=None, **_kw):
        self._d[name] = np.asarray(data)

    def get(self, name):
        return self._d.get(name)

    def __getitem__(self, name):
        return self._d[name]

    def close(self):
        if not self._mode.startswith("r"):
            with open(self._path, "wb") as f:
                np.savez(f, **self._d)


def install():
    """Inject the stand-in modules into ``sys.modules`` (idempotent)."""
    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        skf = types.ModuleType("skimage.feature")
        skf.peak_local_max = peak_local_max
        sk.feature = skf
        sys.modules["skimage"] = sk
        sys.modules["skimage.feature"] = skf
    if "mrcfile" not in sys.modules:
        m = types.ModuleType("mrcfile")
        m.open = _mrc_open
        m.new = _mrc_new
        sys.modules["mrcfile"] = m
    if "h5py" not in sys.modules:
        h = types.ModuleType("h5py")
        h.File = _H5File
        sys.modules["h5py"] = h
    if "matplotlib" not in sys.modules:
        mp = types.ModuleType("matplotlib")
        mpp = types.ModuleType("matplotlib.pyplot")
        mp.pyplot = mpp
        sys.modules["matplotlib"] = mp
        sys.modules["matplotlib.pyplot"] = mpp
