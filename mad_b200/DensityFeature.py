"""Per-feature record with the reference's attribute names (mad/DensityFeature.py:5-84).

In this implementation the features live in device tables (pipeline.Keypoints / Oriented / the
descriptor matrix); DensityFeature objects are the host-side view the reference's downstream
Python (matching bookkeeping, clustering, refinement, HDF5 cache) consumes.
"""
import numpy as np


class DensityFeature(object):
    def __init__(self):
        # detector level
        self.voxel_val = 0
        self.oct_scale = -1
        self.coords = []
        self.map_coords = []
        self.subv_map_coords = []
        self.ratio = 0
        # orientation level
        self.eqsp_size = -1
        self.main_bin = -1
        self.sec_bin = -1
        self.list_bins = []
        self.list_sec_bins = []
        self.to_dom_mat = []
        self.adj_sec_mat = []
        self.Rfinal = []
        # descriptor level
        self.lin_ar_subeqsp = []

    def set_detector_info(self, index, oct_scale, coords, map_coords, subv_map_coords, voxel_val):
        self.index = index
        self.oct_scale = oct_scale
        self.coords = coords
        self.map_coords = map_coords
        self.subv_map_coords = subv_map_coords
        self.voxel_val = voxel_val

    def set_orientator_info(self, eqsp_size, radius):
        self.eqsp_size = eqsp_size
        self.box_size = radius * 2 + 1
        self.box_side = radius

    def set_descriptor_info(self, subeqsp_size, radius):
        self.subeqsp_size = subeqsp_size
        self.box_size = radius * 2 + 1
        self.box_side = radius

    def set_from_file_ori(self, index, main_bin, sec_bin, oct_scale, eqsp_size,
                          coord, map_coord, subv_map_coord, Rfinal, ar_count):
        self.set_detector_info(index, oct_scale, coord, map_coord, subv_map_coord, self.voxel_val)
        self.eqsp_size = eqsp_size
        self.main_bin = main_bin
        self.sec_bin = sec_bin
        self.Rfinal = Rfinal
        self.ar_count = ar_count

    def set_from_file_dsc(self, index, main_bin, sec_bin, oct_scale, eqsp_size, subeqsp_size,
                          coord, map_coord, subv_map_coord, Rfinal, descr):
        self.set_detector_info(index, oct_scale, coord, map_coord, subv_map_coord, self.voxel_val)
        self.eqsp_size = eqsp_size
        self.subeqsp_size = subeqsp_size
        self.main_bin = main_bin
        self.sec_bin = sec_bin
        self.Rfinal = Rfinal
        self.lin_ar_subeqsp = descr

    def show(self):
        print("DF @o=%i: idx=%i main_bin=%i sec_bin=%i (EQSP %i)" % (self.oct_scale, self.index, self.main_bin,
                                                                    self.sec_bin, self.eqsp_size))
        for label, v in (("Coords", self.coords), ("Map coords", self.map_coords), ("Subv coords", self.subv_map_coords)):
            print("> %s: %.3f %.3f %.3f" % (label, v[0], v[1], v[2]))


class FeatureList(list):
    """A list of DensityFeature that remembers the device tables it was made from, so the next
    stage can skip the host -> device rebuild when the list is passed on unchanged."""
    device_keypoints = None     # pipeline.Keypoints
    device_oriented = None      # pipeline.Oriented
    device_descriptors = None   # torch int16 [D, 1024]
    _signature = None

    def stamp(self):
        self._signature = tuple(id(x) for x in self)          # every element: an edit anywhere in the list is seen
        return self

    def unchanged(self):
        return self._signature == tuple(id(x) for x in self)
