"""Orientator with the reference's interface (mad/Orientator.py:12-412): dominant orientations
from EQSP gradient-direction histograms (a7-a10) on the device."""
import copy

import numpy as np

from . import pipeline as _P
from . import tables as _tables
from .DensityFeature import DensityFeature, FeatureList
from .eqsp.eqsp import EQSP_Sphere


def keypoints_of(ms, df_list):
    """Device keypoint table of a feature list: reused when the list comes unchanged from this
    package's Detector, rebuilt from (coords, oct_scale) otherwise."""
    if isinstance(df_list, FeatureList) and df_list.device_keypoints is not None and df_list.unchanged():
        return df_list.device_keypoints
    k = np.zeros(len(df_list), dtype=_P.KEYPOINT_DTYPE)
    for i, df in enumerate(df_list):
        k["vox"][i] = df.coords
        k["peak"][i] = df.coords
        k["oct"][i] = df.oct_scale
        k["val"][i] = df.voxel_val
    k["accepted"] = 1
    return _P.keypoints_from_host(k, ms.space.grad4[0].device)


class Orientator(object):
    def __init__(self, eqsp_size=112, main_ori=6, sec_ori=6, ori_radius=16, gw_sig=0, magn_weighted=False):
        self.magn_weighted = magn_weighted
        self.eqsp_size = eqsp_size
        self.main_ori_lim = main_ori
        self.sec_ori_lim = sec_ori
        self.ori_radius = ori_radius
        self.cutoff = 1e-5
        if self.ori_radius % 2:
            print("MaD> ERROR: ori_radius is uneven (%i). Decreasing by 1" % self.ori_radius)
            self.ori_radius -= 1
        self.ori_radius = self.ori_radius // 2                          # mad/Orientator.py:26-31
        if eqsp_size != 112 or gw_sig or magn_weighted:
            raise NotImplementedError("the CUDA path implements the configuration MaD.run uses: eqsp_size=112, "
                                      "gw_sig=0, magn_weighted=False (mad/MaD.py:360)")
        self.eqsp = EQSP_Sphere(eqsp_size)
        self.time1 = self.time2 = self.time3 = self.time4 = self.time5 = 0

    def assign_orientations(self, ms, df_list):
        print("MaD> Orienting %i anchors..." % (len(df_list)))
        kp = keypoints_of(ms, df_list)
        ori = _P.orient(ms.space, kp, self.ori_radius, self.main_ori_lim, self.sec_ori_lim)
        ho = ori.host()
        rf = _tables.orientation_tables(self.eqsp_size).rf
        out = FeatureList()
        for j in range(len(ho)):
            src = df_list[int(ho["kp"][j])]
            df = copy.copy(src)                     # the reference deep-copies; scratch patches are not kept here
            df.set_orientator_info(self.eqsp_size, self.ori_radius)
            df.main_bin = int(ho["main"][j])
            df.sec_bin = int(ho["sec"][j])
            df.Rfinal = rf[df.main_bin, df.sec_bin].copy()
            out.append(df)
        out.device_keypoints = kp
        out.device_oriented = ori
        out.stamp()
        return out

    def show_timing(self):
        print("Orientator> timing is recorded per kernel on the device (pipeline.profile_records)")
