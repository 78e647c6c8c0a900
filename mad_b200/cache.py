"""Descriptor cache with the reference's four datasets (``MaD._save_descriptors`` / ``_load_descriptors``,
mad/MaD.py:846-873): 'dsc' (descriptors), 'info' (index, main_bin, sec_bin, oct_scale, eqsp_size, subeqsp_size as
uint16), 'coords' (coords, map_coords, subv_map_coords) and 'rot' (Rfinal).

The reference writes HDF5 through h5py; when h5py is importable the same file layout is used, otherwise the four
datasets go into an .npz container with the same names, dtypes and shapes.  Host glue only.
"""
import numpy as np

from .DensityFeature import DensityFeature, FeatureList

try:                                   # pragma: no cover - not in this image
    import h5py
except Exception:                      # noqa: BLE001
    h5py = None


def _arrays(df_list):
    dsc_ar = np.array([df.lin_ar_subeqsp for df in df_list])
    info_ar = np.array([[df.index, df.main_bin, df.sec_bin, df.oct_scale, df.eqsp_size, df.subeqsp_size]
                        for df in df_list]).astype(np.uint16)
    coords_ar = np.array([[df.coords, df.map_coords, df.subv_map_coords] for df in df_list], dtype=np.float64)
    rot_ar = np.array([df.Rfinal for df in df_list])
    return dsc_ar, info_ar, coords_ar, rot_ar


def save_descriptors(df_list, outname):
    dsc_ar, info_ar, coords_ar, rot_ar = _arrays(df_list)
    if h5py is not None and not str(outname).endswith(".npz"):
        with h5py.File(outname, "w") as hf:
            hf.create_dataset("dsc", data=dsc_ar)
            hf.create_dataset("info", data=info_ar)
            hf.create_dataset("coords", data=coords_ar)
            hf.create_dataset("rot", data=rot_ar)
        return outname
    with open(outname, "wb") as f:      # an explicit handle: np.savez would append ".npz" to a ".h5" name
        np.savez(f, dsc=dsc_ar, info=info_ar, coords=coords_ar, rot=rot_ar)
    return outname


def load_descriptors(input_name):
    if h5py is not None and h5py.is_hdf5(input_name):
        with h5py.File(input_name, "r") as hf:
            dsc, coords, info, rot = hf["dsc"][...], hf["coords"][...], hf["info"][...], hf["rot"][...]
    else:
        with np.load(input_name, allow_pickle=False) as z:
            dsc, coords, info, rot = z["dsc"], z["coords"], z["info"], z["rot"]
    df_list = FeatureList()
    for d, c, i, r in zip(dsc, coords, info, rot):
        df = DensityFeature()
        df.set_from_file_dsc(i[0], i[1], i[2], i[3], i[4], i[5], c[0], c[1], c[2], r, d)
        df_list.append(df)
    return df_list
