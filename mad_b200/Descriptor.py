"""Descriptor with the reference's interface (mad/Descriptor.py:13-255): int16[1024] EQSP
sub-block descriptors (a11/a12) on the device."""
import numpy as np

from . import pipeline as _P
from .DensityFeature import FeatureList
from .Orientator import keypoints_of
from .eqsp.eqsp import EQSP_Sphere


class Descriptor(object):
    def __init__(self, subeqsp_size=16, dsc_radius=16, dsc_size=64):
        self.subeqsp_size = subeqsp_size
        self.dsc_radius = dsc_radius
        self.dsc_size = dsc_size
        if self.dsc_radius % 2:
            print("MaD> ERROR: dsc_radius is uneven (%i). Decreasing by 1" % self.dsc_radius)
            self.dsc_radius -= 1
        self.dsc_radius = self.dsc_radius // 2                          # mad/Descriptor.py:23-28
        if subeqsp_size != 16 or dsc_size != 64:
            raise NotImplementedError("the CUDA path implements the configuration MaD.run uses: subeqsp_size=16, "
                                      "dsc_size=64 (mad/MaD.py:361)")
        self.eqsp = EQSP_Sphere(subeqsp_size)
        self.time6 = 0

    def generate_descriptors(self, ms, df_list):
        print("MaD> Generating descriptors from %i oriented anchors..." % len(df_list))
        if isinstance(df_list, FeatureList) and df_list.device_oriented is not None and df_list.unchanged():
            kp, ori = df_list.device_keypoints, df_list.device_oriented
        else:
            # features from elsewhere (e.g. a cache): one keypoint row per feature
            kp = keypoints_of(ms, list(df_list))
            o = np.zeros(len(df_list), dtype=_P.ORIENTED_DTYPE)
            o["kp"] = np.arange(len(df_list))
            o["main"] = [df.main_bin for df in df_list]
            o["sec"] = [df.sec_bin for df in df_list]
            ori = _P.oriented_from_host(o, ms.space.grad4[0].device)
        dsc = _P.describe(ms.space, kp, ori, self.dsc_radius)
        host = dsc.cpu().numpy()
        for j, df in enumerate(df_list):
            df.set_descriptor_info(self.subeqsp_size, self.dsc_radius)
            df.lin_ar_subeqsp = host[j]
        if isinstance(df_list, FeatureList):
            df_list.device_descriptors = dsc
        return df_list

    def show_timing(self):
        print("Descriptor> timing is recorded per kernel on the device (pipeline.profile_records)")
