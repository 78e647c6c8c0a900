"""Host-side constant tables for the CUDA kernels (filter weights, EQSP zones, rotations).

These are parameters of the kernels, computed once in float64 with NumPy using the same
operations the reference's libraries use, so that the device sees bit-identical constants:

* ``gaussian_weights``  -- SciPy's 1-D Gaussian (derivative) kernel, ``scipy/ndimage/_filters.py``
  ``_gaussian_kernel1d`` as called by ``gaussian_filter`` / ``gaussian_laplace``
  (reference call sites ``mad/MapSpace.py:144,171,182``);
* ``ZoneTables``        -- the EQSP partitions (``mad/eqsp/eqsp.py:14-59``);
* ``OrientationTables`` -- ``to_dom_mat`` per main zone and ``Rfinal`` / ``inv(Rfinal)`` per
  (main, sec) pair (``mad/Orientator.py:198-213,253-263,105``, ``mad/math_utils.py:5-27``,
  ``mad/Descriptor.py:132``): Rfinal depends only on the two zone indices.
"""
import functools
import math

import numpy as np

from .eqsp import tables as _eqsp_tables


def gaussian_radius(sigma, truncate=4.0):
    return int(truncate * float(sigma) + 0.5)


def gaussian_weights(sigma, order=0, radius=None):
    """float64[2*radius+1], SciPy's normalised Gaussian kernel or its ``order``-th derivative."""
    sigma = float(sigma)
    if radius is None:
        radius = gaussian_radius(sigma)
    s2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    g = np.exp(-0.5 / s2 * x ** 2)
    g = g / g.sum()
    if order == 0:
        return g
    # derivative of q(x) exp(p(x)): coefficients of q advance by (d/dx + p'(x)) per order
    powers = np.arange(order + 1)
    coeff = np.zeros(order + 1)
    coeff[0] = 1
    step = np.diag(powers[1:], 1) + np.diag(np.ones(order) / -s2, -1)
    for _ in range(order):
        coeff = step.dot(coeff)
    poly = (x[:, None] ** powers).dot(coeff)
    return poly * g


class ZoneTables(object):
    """EQSP zone bounds / centres and their belt structure, as float64 arrays."""

    def __init__(self, size):
        self.size = int(size)
        try:
            b = getattr(_eqsp_tables, "BOUNDS_%d" % self.size)
            c = getattr(_eqsp_tables, "CENTERS_%d" % self.size)
        except AttributeError:
            raise ValueError("no EQSP table for %d zones (available: 16, 112)" % self.size)
        self.bounds = np.array(b, dtype=np.float64) / 10000.0          # theta_min, phi_min, theta_max, phi_max
        self.p_centers = np.array(c, dtype=np.float64) / 10000.0       # theta, phi
        self.c_centers = np.array([[math.sin(p) * math.cos(t), math.sin(p) * math.sin(t), math.cos(p)]
                                   for t, p in self.p_centers])
        first = [0]
        for i in range(1, self.size):
            if self.bounds[i, 1] != self.bounds[i - 1, 1]:
                first.append(i)
        first.append(self.size)
        self.belt_first = np.array(first, dtype=np.int32)
        self.n_belts = len(first) - 1
        self.belt_of = np.zeros(self.size, dtype=np.int32)
        phi = []
        for b_i in range(self.n_belts):
            lo, hi = first[b_i], first[b_i + 1]
            self.belt_of[lo:hi] = b_i
            # the kernels rely on: zones of a belt share the phi range, belts are contiguous in phi
            assert np.all(self.bounds[lo:hi, 1] == self.bounds[lo, 1])
            assert np.all(self.bounds[lo:hi, 3] == self.bounds[lo, 3])
            phi.append(self.bounds[lo, 1])
            if b_i:
                assert self.bounds[lo, 1] == self.bounds[first[b_i - 1], 3]
        phi.append(self.bounds[-1, 3])
        self.belt_phi = np.array(phi, dtype=np.float64)

    def belt_members(self, b):
        return list(range(self.belt_first[b], self.belt_first[b + 1]))


@functools.lru_cache(maxsize=None)
def zone_tables(size):
    return ZoneTables(size)


def unit_vector(vec):
    vec = np.asarray(vec)
    return vec / np.sqrt(np.dot(vec, vec))


def euler_rodrigues(axis, angle):
    """Rotation matrix of the reference's convention (``mad/math_utils.py:15-27``): Euler-Rodrigues
    parameters (a; b, c, d) = (cos(angle/2); -axis * sin(angle/2))."""
    a = np.cos(angle / 2.0)
    b, c, d = -np.asarray(axis) * np.sin(angle / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


class OrientationTables(object):
    """r1[a] (zone centre a -> +z), rf[a, b] = R2(b) @ r1[a], rf_inv[a, b] = inv(rf[a, b])."""

    def __init__(self, eqsp_size=112):
        zt = zone_tables(eqsp_size)
        n = zt.size
        self.n = n
        self.r1 = np.zeros((n, 3, 3))
        zhat = [0, 0, 1]
        for a in range(n):
            if a == 0:
                self.r1[a] = np.identity(3)
                continue
            c = unit_vector(zt.c_centers[a])
            angle = np.arccos(np.clip(np.dot(c, zhat), -1.0, 1.0))
            self.r1[a] = euler_rodrigues(unit_vector(np.cross(c, zhat)), angle)
        self.r2 = np.zeros((n, 3, 3))
        for b in range(n):
            first = zt.belt_first[zt.belt_of[b]]
            ftheta = -1 * (zt.p_centers[b][0] - zt.p_centers[first][0])
            self.r2[b] = euler_rodrigues(zhat, ftheta)
        self.rf = np.zeros((n, n, 3, 3))
        self.rf_inv = np.zeros((n, n, 3, 3))
        for a in range(n):
            for b in range(n):
                m = np.dot(self.r2[b], self.r1[a])
                self.rf[a, b] = m
                self.rf_inv[a, b] = np.linalg.inv(m)


@functools.lru_cache(maxsize=None)
def orientation_tables(eqsp_size=112):
    return OrientationTables(eqsp_size)
