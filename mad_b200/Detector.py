"""Detector with the reference's interface (mad/Detector.py:6-189): 3x3x3 LoG maxima + sub-voxel
Newton localisation (a5/a6) on the device, returned as DensityFeature objects."""
import os

import numpy as np

from . import pipeline as _P
from .DensityFeature import DensityFeature, FeatureList


class Detector(object):
    # rejection statistics the reference keeps as attributes (mad/Detector.py:9-15; never read by MaD)
    COUNTERS = ("lowdensity", "lowcontrast", "saddlepoint", "lowratio", "largeoffset", "badhessian")

    def __init__(self):
        for name in self.COUNTERS:
            setattr(self, name, 0)

    def find_anchors(self, ms, outname=""):
        if outname != "" and not os.path.exists(os.path.split(outname)[0]):
            print("Detector> WARNING: if writing files, a valid outname must be specified.")
            print("          Specified: %s" % outname)
            outname = ""
        print("MaD> Finding anchors in %s... " % ms.name)
        kp = _P.detect(ms.space, border=12, threshold=5e-2)
        hk = kp.host()
        org = np.array([ms.xi, ms.yi, ms.zi], dtype=np.float64)
        vs = np.asarray(ms.voxelsp_list, dtype=np.float64)[hk["oct"]][:, None]
        vox = hk["vox"].astype(np.int64)
        map_coords = vox * vs + org                                      # mad/Detector.py:126-128
        sub = (vox + hk["off"]) * vs + org                               # int64 + f32 -> f64 (NumPy 2)
        df_list = FeatureList()
        for i in range(len(hk)):
            df = DensityFeature()
            df.set_detector_info(i, int(hk["oct"][i]), [int(c) for c in vox[i]], map_coords[i], sub[i], hk["val"][i])
            df_list.append(df)
        df_list.device_keypoints = kp
        df_list.stamp()
        if outname:
            self.write_df_to_file(df_list, outname + "_data.txt")
            self.write_df_to_pdb(df_list, outname + ".pdb")
        return df_list

    def get_coord_in_ref_map(self, x, y, z, xi, yi, zi, voxsp):
        return np.array([x * voxsp + xi, y * voxsp + yi, z * voxsp + zi])

    def write_df_to_file(self, df_list, outname):
        np.save(outname, list(df_list))

    def load_df_from_file(self, df_file):
        df_list = np.load(df_file, allow_pickle=True)
        print("Det> Loaded %i anchors." % len(df_list))
        return df_list

    def write_df_to_pdb(self, df_list, outname, save_regular=False):
        rows = [(i, "SUB", "A", df.subv_map_coords) for i, df in enumerate(df_list)]
        if save_regular:
            rows += [(i, "ORI", "B", df.map_coords) for i, df in enumerate(df_list)]
        with open(outname, "w") as f:
            for i, res, chain, c in rows:
                f.write("ATOM%7i%5s%5s%1s%4i    %8.3f%8.3f%8.3f\n" % (i, "O", res, chain, i, c[0], c[1], c[2]))
