"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy/SciPy) of the MaD local-feature hot path.

This is the *oracle* the CUDA path is checked against.  It is NOT part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it; ``mad_b200`` never does (and fails loudly without its CUDA library).

Each function follows the cited lines of the reference (paths relative to /root/reference).
The third-party arithmetic the reference calls is SciPy 1.5.2 / NumPy 1.19.2 / scikit-image
0.17.2 (``requirements.txt``); here SciPy 1.18 / NumPy 2.3 are installed (same algorithms for
the calls used: ``gaussian_filter``, ``gaussian_laplace``, ``interp1d(kind='cubic')``,
``np.gradient``) and are called exactly where the reference calls them; scikit-image is absent,
so ``peak_local_max`` is restated (``local_maxima`` below; tie ORDER is "parity unpinned",
see ref_shims.py).  The reference ships no tests or golden vectors of its own, so the pin is:
the UNMODIFIED reference executed in the build container on seeded synthetic inputs
(oracle/gen_goldens.py -> tests/golden/*.npz); tests/test_oracle_golden.py requires this file
to reproduce those fixtures bit-for-bit (dense arrays by SHA-256).

Unlike the reference this restatement is vectorised over patch voxels (not over Python
objects), so it is also a much *faster* CPU baseline than the reference itself.
"""
import math

import numpy as np
from scipy import ndimage as ndi
from scipy.interpolate import interp1d

TWO_PI = 2.0 * np.pi


# ----------------------------------------------------------------------------------------------
# EQSP tables (mad/eqsp/eqsp.py:14-59 + sphere_*.txt / centers_*.txt)
# ----------------------------------------------------------------------------------------------
class Eqsp(object):
    def __init__(self, size):
        from mad_b200.eqsp import tables  # pure data (integers in 1e-4 rad)
        self.size = size
        self.bounds = np.array(getattr(tables, "BOUNDS_%d" % size), dtype=np.float64) / 10000.0
        self.p_centers = np.array(getattr(tables, "CENTERS_%d" % size), dtype=np.float64) / 10000.0
        # mad/eqsp/eqsp.py:30-32 uses math.sin / math.cos
        self.c_centers = np.array([[math.sin(p) * math.cos(t), math.sin(p) * math.sin(t), math.cos(p)]
                                   for t, p in self.p_centers])
        # belts = runs of equal phi_min (mad/eqsp/eqsp.py:38-49)
        self.belts = []
        last = None
        for i in range(size):
            if self.bounds[i, 1] != last:
                self.belts.append([])
                last = self.bounds[i, 1]
            self.belts[-1].append(i)
        self.belt_of = np.zeros(size, dtype=int)
        for b, members in enumerate(self.belts):
            self.belt_of[members] = b


def _unit(v):
    v = np.asarray(v)
    return v / np.sqrt(np.dot(v, v))            # mad/math_utils.py:9


def rodrigues(axis, angle):
    """mad/math_utils.py:15-27 (note the minus sign on the axis terms)."""
    a = np.cos(angle / 2.0)
    b, c, d = -np.asarray(axis) * np.sin(angle / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


def to_pole_matrix(eq, a):
    """mad/Orientator.py:198-213: rotation taking the centre of zone ``a`` to +z."""
    if a == 0:
        return np.identity(3)
    c = _unit(eq.c_centers[a])
    angle = np.arccos(np.clip(np.dot(c, [0, 0, 1]), -1.0, 1.0))
    axis = _unit(np.cross(c, [0, 0, 1]))
    return rodrigues(axis, angle)


def about_z_matrix(eq, b):
    """mad/Orientator.py:253-263: rotation about z putting zone ``b`` on its belt's first zone."""
    first = eq.belts[eq.belt_of[b]][0]
    ftheta = -1 * (eq.p_centers[b][0] - eq.p_centers[first][0])
    return rodrigues([0, 0, 1], ftheta)


def rfinal(eq, a, b):
    return np.dot(about_z_matrix(eq, b), to_pole_matrix(eq, a))     # mad/Orientator.py:105


# ----------------------------------------------------------------------------------------------
# a1-a4  MapSpace.build_space  (mad/MapSpace.py:116-189, 191-214)
# ----------------------------------------------------------------------------------------------
def upsample2(grid):
    """mad/MapSpace.py:137-142,191-214: separable not-a-knot cubic interpolation at half steps."""
    a = grid
    for ax in range(3):
        n = a.shape[ax]
        a = interp1d(np.arange(0, n, 1), a, axis=ax, kind="cubic")(np.arange(0, n - 0.5, 0.5))
    return a


def build_space(grid, map_padding=9, sig_init=2, sig_presmooth=1, oct_mode="both"):
    """Returns dict(grid_list, map_space, gauss_list, grad_list); oct_mode as mad/MapSpace.py:149-163 ("up" / "base": that
    grid alone, as octave 0)."""
    grid = np.asarray(grid, dtype=np.float32)
    if map_padding:
        grid = np.pad(grid, map_padding, mode="constant")                      # :118
    up = upsample2(grid)
    if sig_presmooth:
        up = ndi.gaussian_filter(up, sigma=sig_presmooth)                      # :144
    up = up.astype(np.float32)
    grids = {"both": [up, grid], "up": [up], "base": [grid]}[oct_mode]
    out = dict(grid_list=grids, map_space=[], gauss_list=[], grad_list=[])
    for g in grids:
        log_g = -1 * ndi.gaussian_laplace(g, sigma=sig_init) * sig_init ** 2   # :171
        log_g[log_g < 0] = 0.0
        out["map_space"].append(log_g)
    for g in grids:
        gs = ndi.gaussian_filter(g, sig_init)                                  # :182
        out["gauss_list"].append(gs)
        out["grad_list"].append(np.moveaxis(np.array(np.gradient(gs)), 0, -1))  # :187
    return out


# ----------------------------------------------------------------------------------------------
# a5-a6  Detector  (mad/Detector.py:26-45, 53-128; skimage.feature.peak_local_max 0.17.x)
# ----------------------------------------------------------------------------------------------
def local_maxima(image, border=12, threshold=5e-2):
    if image.size == 0 or np.all(image == image.flat[0]):
        return np.empty((0, 3), dtype=np.int64)
    peak = image == ndi.maximum_filter(image, size=3, mode="constant")
    peak &= image > threshold
    for ax in range(3):
        s = [slice(None)] * 3
        s[ax] = slice(None, border)
        peak[tuple(s)] = False
        s[ax] = slice(-border, None)
        peak[tuple(s)] = False
    idx = np.nonzero(peak)
    order = np.argsort(-image[idx], kind="stable")
    return np.transpose(idx)[order]


def newton_localise(L, p, max_offset=0.6, max_iter=5):
    """mad/Detector.py:53-123.  Returns (ok, voxel[3], offset[3])."""
    x, y, z = p
    nx, ny, nz = L.shape
    off = None
    H = None
    done = False
    for _ in range(max_iter):
        c2 = 2 * L[x, y, z]
        xx = L[x - 1, y, z] + L[x + 1, y, z] - c2
        yy = L[x, y - 1, z] + L[x, y + 1, z] - c2
        zz = L[x, y, z - 1] + L[x, y, z + 1] - c2
        xy = 0.25 * ((L[x + 1, y + 1, z] - L[x + 1, y - 1, z]) - (L[x - 1, y + 1, z] - L[x - 1, y - 1, z]))
        xz = 0.25 * ((L[x + 1, y, z + 1] - L[x + 1, y, z - 1]) - (L[x - 1, y, z + 1] - L[x - 1, y, z - 1]))
        yz = 0.25 * ((L[x, y + 1, z + 1] - L[x, y + 1, z - 1]) - (L[x, y - 1, z + 1] - L[x, y - 1, z - 1]))
        H = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
        G = np.array([0.5 * (L[x + 1, y, z] - L[x - 1, y, z]),
                      0.5 * (L[x, y + 1, z] - L[x, y - 1, z]),
                      0.5 * (L[x, y, z + 1] - L[x, y, z - 1])])
        try:
            Hinv = np.linalg.inv(H)
        except Exception:
            return False, p, None
        off = -np.dot(Hinv, G)
        if np.all(np.abs(off) < max_offset):
            done = True
            break
        if off[0] < -max_offset and x - 1 > 0:
            x -= 1
        elif off[0] > max_offset and x + 1 < nx - 1:
            x += 1
        if off[1] < -max_offset and y - 1 > 0:
            y -= 1
        elif off[1] > max_offset and y + 1 < ny - 1:
            y += 1
        if off[2] < -max_offset and z - 1 > 0:
            z -= 1
        elif off[2] > max_offset and z + 1 < nz - 1:
            z += 1
    if not done:
        return False, p, None
    if np.any(np.linalg.eigvals(H) > 0):
        return False, p, None
    return True, (x, y, z), off


def detect(map_space, voxelsp_list, origin):
    """Returns dict of arrays: index, oct, coords (int), map_coords, subv_map_coords, val."""
    oc, co, mc, sc, va = [], [], [], [], []
    org = np.asarray(origin, dtype=np.float64)
    for o, L in enumerate(map_space):
        for p in local_maxima(L):
            ok, vox, off = newton_localise(L, p)
            if not ok:
                continue
            vox = np.array(vox, dtype=np.int64)
            sub = vox + off                                     # int64 + f32 -> f64 (numpy 2)
            oc.append(o)
            co.append(vox)
            mc.append(vox * voxelsp_list[o] + org)              # :126-128
            sc.append(sub * voxelsp_list[o] + org)
            va.append(L[tuple(p)])
    k = len(oc)
    return dict(index=np.arange(k, dtype=np.int32), oct=np.array(oc, dtype=np.int32),
                coords=np.array(co, dtype=np.int32).reshape(-1, 3),
                map_coords=np.array(mc, dtype=np.float64).reshape(-1, 3),
                subv_map_coords=np.array(sc, dtype=np.float64).reshape(-1, 3),
                val=np.array(va, dtype=np.float32))


# ----------------------------------------------------------------------------------------------
# a7-a10  Orientator  (mad/Orientator.py:35-54, 116-169, 171-270, 290-343)
# ----------------------------------------------------------------------------------------------
def zone_membership(eq, th, sth, ph):
    """Strict, per-zone independent test of mad/Orientator.py:324-331 -> bool[n, zones]."""
    b = eq.bounds
    th = th.astype(np.float64)[:, None]
    sth = sth.astype(np.float64)[:, None]
    ph = ph.astype(np.float64)[:, None]
    in_th = ((th < b[None, :, 2]) & (th > b[None, :, 0])) | ((sth < b[None, :, 2]) & (sth > b[None, :, 0]))
    return in_th & (ph < b[None, :, 3]) & (ph > b[None, :, 1])


def spherical_angles(v):
    """theta in [0,2pi), theta+2pi, phi -- in the dtype of ``v`` (mad/Orientator.py:307-321)."""
    th = np.arctan2(v[:, 1], v[:, 0])
    th[th < 0] += 2 * np.pi
    sth = th + 2 * np.pi
    ph = np.arccos(np.clip(v[:, 2], -1, 1))
    return th, sth, ph


def zone_histogram(eq, v, w):
    th, sth, ph = spherical_angles(v)
    member = zone_membership(eq, th, sth, ph)
    return (member * w[:, None]).sum(0).astype(np.int32)


def _norm50(h):
    return np.array(h / np.amax(h) * 50, dtype=np.int32)      # :340


class OrientTables(object):
    def __init__(self, patch_size=16, eqsp_size=112):
        r = patch_size - (patch_size % 2)
        self.r = r // 2                                          # :26-29
        self.eq = Eqsp(eqsp_size)
        d = np.mgrid[-self.r:self.r + 1, -self.r:self.r + 1, -self.r:self.r + 1]
        dist = np.sqrt(np.sum(d * d, 0))
        self.mask = (dist <= self.r * 1.05).astype(np.int64).reshape(-1)   # :44-47
        self._r1 = {}
        self._rf = {}

    def r1(self, a):
        if a not in self._r1:
            self._r1[a] = to_pole_matrix(self.eq, a)
        return self._r1[a]

    def rf(self, a, b):
        if (a, b) not in self._rf:
            self._rf[(a, b)] = np.dot(about_z_matrix(self.eq, b), self.r1(a))
        return self._rf[(a, b)]


def orient_keypoint(tab, grad, c, oct_scale, lim_main=6, lim_sec=6, cutoff=1e-5):
    """One keypoint -> list of (main_bin, sec_bin).  mad/Orientator.py:80-108."""
    r = tab.r
    s = 1 if oct_scale == 1 else 2
    lo = np.asarray(c) - s * r
    hi = np.asarray(c) + s * r + 1
    if np.any(lo < 0) or np.any(hi > np.array(grad.shape[:3]) - 1):        # :129-135,149-155
        return []
    p = grad[lo[0]:hi[0]:s, lo[1]:hi[1]:s, lo[2]:hi[2]:s, :].reshape(-1, 3).copy()
    m = np.sqrt(np.sum(np.square(p), -1))
    nz = m > cutoff
    p[nz] = p[nz] / m[nz][:, None]
    w = tab.mask.copy()
    w[m < cutoff] = 0
    h = zone_histogram(tab.eq, p, w)
    if not np.amax(h):
        return []
    hn = _norm50(h)
    mains = np.where(hn > max(hn) * 0.8)[0]
    if len(mains) > lim_main:
        return []
    out = []
    for a in mains:
        if a != 0:
            pr = np.matmul(p, tab.r1(int(a)).T)
            h2 = zone_histogram(tab.eq, pr, w)
            cur = _norm50(h2) if np.amax(h2) else h2
        else:
            cur = hn
        q = cur[1:-1]
        if np.amax(q) == 0:
            continue
        qn = np.array(q / np.amax(q) * 50, dtype=np.int32)
        secs = np.where(qn > max(qn) * 0.8)[0] + 1
        if len(secs) > lim_sec:
            continue
        out.extend((int(a), int(b)) for b in secs)
    return out


def orient(grad_list, kp, patch_size=16, tab=None):
    """Returns dict(kp (row in the keypoint table), main, sec) in emission order."""
    tab = tab or OrientTables(patch_size)
    rows, mains, secs = [], [], []
    for i in range(len(kp["oct"])):
        o = int(kp["oct"][i])
        for a, b in orient_keypoint(tab, grad_list[o], kp["coords"][i], o):
            rows.append(i)
            mains.append(a)
            secs.append(b)
    return dict(kp=np.array(rows, dtype=np.int32), main=np.array(mains, dtype=np.int32),
                sec=np.array(secs, dtype=np.int32)), tab


# ----------------------------------------------------------------------------------------------
# a11-a12  Descriptor  (mad/Descriptor.py:31-64, 123-202)
# ----------------------------------------------------------------------------------------------
class DescribeTables(object):
    def __init__(self, patch_size=16, subeqsp_size=16):
        r = patch_size - (patch_size % 2)
        self.r = dr = r // 2
        self.eq = Eqsp(subeqsp_size)
        self.layout = {
            0: np.rollaxis(np.array(np.mgrid[-2 * dr + 1:2 * dr + 1:2, -2 * dr + 1:2 * dr + 1:2, -2 * dr + 1:2 * dr + 1:2]), 0, 4),
            1: np.rollaxis(np.array(np.mgrid[-dr + 0.5:dr + 0.5, -dr + 0.5:dr + 0.5, -dr + 0.5:dr + 0.5]), 0, 4),
        }
        cuts = [0, dr // 2, dr, 3 * dr // 2, 2 * dr]
        blk = np.zeros(2 * dr, dtype=np.int64)
        for k in range(4):
            blk[cuts[k]:cuts[k + 1]] = k
        bx, by, bz = np.meshgrid(blk, blk, blk, indexing="ij")
        self.block = (16 * by + 4 * bx + bz).reshape(-1)         # list order of :44-64
        self.n = (2 * dr) ** 3


def nearest_index(P, n):
    """scipy RegularGridInterpolator(method='nearest') on integer grids: half rounds DOWN."""
    i = np.clip(np.floor(P).astype(np.int64), 0, n - 2)
    t = P - i
    return np.where(t <= 0.5, i, i + 1)


def describe_one(tab, grad, c, oct_scale, R):
    lay = tab.layout[oct_scale]
    P = np.add(np.matmul(lay, np.linalg.inv(R).T), c)            # :132-133
    shape = np.array(grad.shape[:3])
    flat = P.reshape(-1, 3)
    if np.any(flat < 0) or np.any(flat > shape - 1):             # bounds_error -> zero descriptor
        return np.zeros(64 * tab.eq.size, dtype=np.int16)
    ix = nearest_index(flat[:, 0], shape[0])
    iy = nearest_index(flat[:, 1], shape[1])
    iz = nearest_index(flat[:, 2], shape[2])
    v = grad[ix, iy, iz, :].reshape(lay.shape)
    m = np.sqrt(np.sum(np.square(v), -1))
    nz = np.where(m > 1e-12)
    v[nz] = v[nz] / np.expand_dims(m[nz], -1)
    v = np.matmul(v, R.T).reshape(-1, 3)                         # :155
    m = m.reshape(-1)
    th, sth, ph = spherical_angles(v)
    member = zone_membership(tab.eq, th, sth, ph)
    zone = np.zeros(tab.n, dtype=np.int16)
    for a in range(tab.eq.size):                                 # ascending assignment :176-187
        zone[member[:, a]] = a
    zone[m < 1e-5] = -1
    ok = zone >= 0
    counts = np.bincount(tab.block[ok] * tab.eq.size + zone[ok], minlength=64 * tab.eq.size)
    return counts.astype(np.int16)


def describe(grad_list, kp, ori, tab_o, patch_size=16, tab_d=None):
    tab_d = tab_d or DescribeTables(patch_size)
    out = np.zeros((len(ori["kp"]), 64 * tab_d.eq.size), dtype=np.int16)
    for j in range(len(ori["kp"])):
        i = int(ori["kp"][j])
        o = int(kp["oct"][i])
        R = tab_o.rf(int(ori["main"][j]), int(ori["sec"][j]))
        out[j] = describe_one(tab_d, grad_list[o], kp["coords"][i], o, R)
    return out


# ----------------------------------------------------------------------------------------------
# a15  descriptor matching  (mad/MaD.py:414-424) + the top-k extension
# ----------------------------------------------------------------------------------------------
def unit_rows(dsc):
    d = np.asarray(dsc).astype(np.float64)
    n = np.sqrt((d * d).sum(1))                                  # == np.linalg.norm on integers
    out = d.copy()
    nzr = n > 0
    out[nzr] = d[nzr] / n[nzr][:, None]
    return out


def match_scores(hi, lo):
    return np.dot(unit_rows(hi), unit_rows(lo).T)                # :420


def match_threshold(hi, lo, cc=0.6, block=4096):
    """Row-major (hi-major) list of (i, j) with preds > cc, and the scores; row-blocked."""
    hi_u, lo_u = unit_rows(hi), unit_rows(lo)
    pairs, scores = [], []
    for s in range(0, hi_u.shape[0], block):
        pr = np.dot(hi_u[s:s + block], lo_u.T)
        i, j = np.where(pr > cc)
        pairs.append(np.stack([i + s, j], 1))
        scores.append(pr[i, j])
    if not pairs:
        return np.empty((0, 2), dtype=np.int32), np.empty(0)
    return np.concatenate(pairs).astype(np.int32), np.concatenate(scores)


def match_topk(hi, lo, k=8, block=4096):
    hi_u, lo_u = unit_rows(hi), unit_rows(lo)
    idx, val = [], []
    for s in range(0, hi_u.shape[0], block):
        pr = np.dot(hi_u[s:s + block], lo_u.T)
        o = np.argsort(-pr, axis=1, kind="stable")[:, :k]
        idx.append(o)
        val.append(np.take_along_axis(pr, o, 1))
    return np.concatenate(idx).astype(np.int32), np.concatenate(val)


# ----------------------------------------------------------------------------------------------
# whole path (order of mad/MaD.py:358-368)
# ----------------------------------------------------------------------------------------------
def describe_struct(grid, voxelsp, origin=(0.0, 0.0, 0.0), patch_size=16, map_padding=9, timings=None):
    import time
    t0 = time.perf_counter()
    sp = build_space(grid, map_padding=map_padding)
    t1 = time.perf_counter()
    org = np.asarray(origin, dtype=np.float64) - map_padding * voxelsp
    kp = detect(sp["map_space"], [voxelsp / 2, voxelsp], org)
    t2 = time.perf_counter()
    ori, tab_o = orient(sp["grad_list"], kp, patch_size)
    t3 = time.perf_counter()
    dsc = describe(sp["grad_list"], kp, ori, tab_o, patch_size)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(build_space=t1 - t0, detect=t2 - t1, orient=t3 - t2, describe=t4 - t3)
    return sp, kp, ori, dsc


# ---------------------------------------------------------------------------------------------
# next component: the per-pair loop of MaD._match_dsc, mad/MaD.py:426-453
# ---------------------------------------------------------------------------------------------
def match_dsc(lo, hi, anchor_dist_thresh=4, cc_threshold=0.65):
    """``lo`` / ``hi``: dicts with dsc int16[D,1024], subv float64[D,3], index, oct, main, sec, Rfinal
    float64[D,3,3].  Returns (results float64[P,23], lo_mapcoords, hi_mapcoords) exactly as
    mad/MaD.py:416-453 builds them."""
    from scipy.spatial import cKDTree
    preds = match_scores(hi["dsc"], lo["dsc"])                                   # :416-420
    ph, pl = np.where(preds > cc_threshold)                                      # :423-424
    hi_mapcoords = np.unique(hi["subv"][ph], axis=0)                             # :427
    lo_mapcoords = np.unique(lo["subv"][pl], axis=0)                             # :428
    lo_tree = cKDTree(lo_mapcoords)                                              # :431
    results = []
    for phi, plo in zip(ph, pl):
        R = np.dot(np.linalg.inv(lo["Rfinal"][plo]), hi["Rfinal"][phi])          # :438
        l = hi_mapcoords.shape[0]
        cur = hi_mapcoords - hi["subv"][phi]
        cur = np.dot(cur, R.T)                                                   # :443
        cur = cur + lo["subv"][plo]
        distances, _ = lo_tree.query(cur, distance_upper_bound=anchor_dist_thresh)   # :447
        repeatability = 100 * np.count_nonzero(distances < anchor_dist_thresh) / l
        results.append(np.concatenate([[preds[phi, plo], repeatability, lo["index"][plo], lo["oct"][plo], lo["main"][plo],
                                        hi["index"][phi], hi["oct"][phi], hi["main"][phi]],
                                       hi["subv"][phi], lo["subv"][plo], R.flatten()]))   # :451
    return np.array(results, dtype=np.float64).reshape(-1, 23), lo_mapcoords, hi_mapcoords
