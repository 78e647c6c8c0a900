"""Host-side constant tables (filter weights, EQSP zones, Rfinal) against SciPy and the reference's
golden Rfinal matrices.  CPU only."""
import numpy as np
import pytest
from scipy import ndimage as ndi

import helpers as H
from mad_b200 import tables


@pytest.mark.parametrize("sigma,order", [(1, 0), (2, 0), (2, 2), (1.5, 2)])
def test_gaussian_weights_equal_scipy(sigma, order):
    r = tables.gaussian_radius(sigma)
    w = tables.gaussian_weights(sigma, order, r)
    imp = np.zeros(4 * r + 1)
    imp[2 * r] = 1.0
    ref = ndi.gaussian_filter1d(imp, sigma, order=order, mode="constant")[r:3 * r + 1]
    # correlate1d with a symmetric kernel: the response to an impulse is the kernel itself
    assert np.array_equal(w, ref)


def test_zone_tables_structure():
    z112, z16 = tables.zone_tables(112), tables.zone_tables(16)
    assert np.diff(z112.belt_first).tolist() == [1, 7, 12, 17, 19, 19, 17, 12, 7, 1]
    assert np.diff(z16.belt_first).tolist() == [1, 7, 7, 1]
    assert z112.bounds[7, 2] == 6.2832 and z112.bounds[8].tolist() == [6.1336, 0.5411, 6.6572, 0.8726]
    assert z16.belt_phi.tolist() == [0.0, 0.5054, 1.5708, 2.6362, 3.1416]


@pytest.mark.parametrize("case", ["tiny", "small", "c1", "pair_lo"])
def test_rfinal_table_equals_reference(case):
    g = H.golden(case)
    ot = tables.orientation_tables(112)
    for (a, b), m in zip(g["rfinal_ab"], g["rfinal_mat"]):
        assert np.array_equal(ot.rf[a, b], m)
        assert np.array_equal(ot.rf_inv[a, b], np.linalg.inv(m))


def test_eqsp_sphere_interface():
    from mad_b200.eqsp.eqsp import EQSP_Sphere
    s = EQSP_Sphere(112)
    assert s.size == 112 and len(s.belt_l) == 10 and s.belt_of_idx(8) == 2
    assert s.belt_indices(1) == list(range(1, 8))
    assert np.allclose(np.linalg.norm(s.c_center(40)), 1.0)
    assert s.p_center(0).tolist() == [0.0, 0.0]
