"""Descriptor cache round trip (mad/MaD.py:846-873): the four datasets with the reference's dtypes."""
import os

import numpy as np

import helpers as H


def test_descriptor_cache_round_trip(tmp_path):
    from mad_b200 import cache
    from mad_b200.DensityFeature import DensityFeature
    g = H.golden("tiny")
    rf = {(int(a), int(b)): m for (a, b), m in zip(g["rfinal_ab"], g["rfinal_mat"])}
    feats = []
    for i in range(len(g["of_index"])):
        df = DensityFeature()
        df.set_detector_info(int(g["of_index"][i]), int(g["of_oct"][i]), g["of_coords"][i].astype(np.float64),
                             g["of_subv_map_coords"][i] * 0 + 1.5, g["of_subv_map_coords"][i], 0.0)
        df.eqsp_size, df.subeqsp_size = 112, 16
        df.main_bin, df.sec_bin = int(g["of_main"][i]), int(g["of_sec"][i])
        df.Rfinal = rf[(df.main_bin, df.sec_bin)]
        df.lin_ar_subeqsp = g["dsc"][i]
        feats.append(df)
    path = os.path.join(str(tmp_path), "dsc_cache.h5")
    cache.save_descriptors(feats, path)
    with np.load(path) as z:
        assert sorted(z.files) == ["coords", "dsc", "info", "rot"]
        assert z["info"].dtype == np.uint16 and z["info"].shape == (len(feats), 6)
        assert z["dsc"].dtype == np.int16 and z["coords"].shape == (len(feats), 3, 3) and z["rot"].shape == (len(feats), 3, 3)
    back = cache.load_descriptors(path)
    assert len(back) == len(feats)
    for a, b in zip(feats, back):
        assert (a.index, a.main_bin, a.sec_bin, a.oct_scale) == (b.index, b.main_bin, b.sec_bin, b.oct_scale)
        assert (b.eqsp_size, b.subeqsp_size) == (112, 16)
        assert np.array_equal(a.lin_ar_subeqsp, b.lin_ar_subeqsp) and np.array_equal(a.Rfinal, b.Rfinal)
        assert np.array_equal(a.subv_map_coords, b.subv_map_coords) and np.array_equal(a.coords, b.coords)
