"""EQSP_Sphere: Leopardi equal-area sphere partition tables with the reference's interface
(mad/eqsp/eqsp.py:12-87).  The tables are embedded (eqsp/tables.py) instead of being opened by a
cwd-relative path."""
import numpy as np

from ..tables import zone_tables


class EQSP_Sphere(object):
    def __init__(self, size=112):
        zt = zone_tables(size)
        self.size = size
        self.sphere_eqsp = zt.bounds.copy()            # theta_min, phi_min, theta_max, phi_max
        self.p_centers_eqsp = zt.p_centers.copy()      # theta, phi
        self.c_centers_eqsp = zt.c_centers.copy()
        self.belt_l = [zt.belt_members(b) for b in range(zt.n_belts)]
        self.max_belt_l = max(len(b) for b in self.belt_l)
        # index of the first zone of the first longest belt (reference: equator_idx)
        longest = [len(b) for b in self.belt_l].index(self.max_belt_l)
        self.equator_idx = int(zt.belt_first[longest])
        # 0.1 x mean nearest-neighbour distance between zone centres
        c = self.c_centers_eqsp
        d = np.sqrt(((c[:, None, :] - c[None, :, :]) ** 2).sum(-1))
        np.fill_diagonal(d, np.inf)
        self.feature_dist_thresh = float(np.average(d.min(1)) * 0.1)

    def p_center(self, idx):
        return self.p_centers_eqsp[idx]

    def c_center(self, idx):
        return self.c_centers_eqsp[idx]

    def area(self, idx):
        return self.sphere_eqsp[idx]

    def dist_thresh(self):
        return self.feature_dist_thresh

    def belt_indices(self, idx):
        return self.belt_l[idx]

    def belt_of_idx(self, idx):
        for i, members in enumerate(self.belt_l):
            if idx in members:
                return i

    def belt_equator_bounds(self):
        return self.sphere_eqsp[self.equator_idx]
