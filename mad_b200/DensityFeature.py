"""Per-feature record with the reference's attribute names (mad/DensityFeature.py:5-84).

In this implementation the features live in device tables (pipeline.Keypoints / Oriented / the
descriptor matrix); DensityFeature objects are the host-side view the reference's downstream
Python (matching bookkeeping, clustering, refinement, HDF5 cache) consumes.  The attribute and
setter names are the interface (``MaD._match_dsc`` and the cache read them, mad/MaD.py:416-451,
849-852); the scratch arrays the reference parks on every feature (gradient patches, per-step
histograms, VMD helpers) have no counterpart here -- those intermediates never leave the GPU.
"""

# attribute -> value of a fresh record (the reference's defaults, mad/DensityFeature.py:6-33); `list` = a new empty list
_FRESH = {
    # detector level
    "voxel_val": 0, "oct_scale": -1, "coords": list, "map_coords": list, "subv_map_coords": list, "ratio": 0,
    # orientation level
    "eqsp_size": -1, "main_bin": -1, "sec_bin": -1, "list_bins": list, "list_sec_bins": list,
    "to_dom_mat": list, "adj_sec_mat": list, "Rfinal": list,
    # descriptor level
    "lin_ar_subeqsp": list,
}


class DensityFeature(object):
    def __init__(self):
        for name, fresh in _FRESH.items():
            setattr(self, name, fresh() if fresh is list else fresh)

    def _take(self, **fields):
        self.__dict__.update(fields)

    def _patch_geometry(self, radius):
        self._take(box_size=2 * radius + 1, box_side=radius)

    def set_detector_info(self, index, oct_scale, coords, map_coords, subv_map_coords, voxel_val):
        """mad/DensityFeature.py:35-41 (called by Detector.find_anchors, mad/Detector.py:126-128)."""
        self._take(index=index, oct_scale=oct_scale, coords=coords, map_coords=map_coords,
                   subv_map_coords=subv_map_coords, voxel_val=voxel_val)

    def set_orientator_info(self, eqsp_size, radius):
        """mad/DensityFeature.py:43-52 without the per-feature patch / histogram scratch arrays."""
        self._take(eqsp_size=eqsp_size)
        self._patch_geometry(radius)

    def set_descriptor_info(self, subeqsp_size, radius):
        """mad/DensityFeature.py:54-57."""
        self._take(subeqsp_size=subeqsp_size)
        self._patch_geometry(radius)

    def _from_file(self, index, main_bin, sec_bin, oct_scale, eqsp_size, coord, map_coord, subv_map_coord, Rfinal):
        self._take(index=index, main_bin=main_bin, sec_bin=sec_bin, oct_scale=oct_scale, eqsp_size=eqsp_size,
                   coords=coord, map_coords=map_coord, subv_map_coords=subv_map_coord, Rfinal=Rfinal)

    def set_from_file_ori(self, index, main_bin, sec_bin, oct_scale, eqsp_size,
                          coord, map_coord, subv_map_coord, Rfinal, ar_count):
        """A cached oriented feature (mad/DensityFeature.py:59-70)."""
        self._from_file(index, main_bin, sec_bin, oct_scale, eqsp_size, coord, map_coord, subv_map_coord, Rfinal)
        self._take(ar_count=ar_count)

    def set_from_file_dsc(self, index, main_bin, sec_bin, oct_scale, eqsp_size, subeqsp_size,
                          coord, map_coord, subv_map_coord, Rfinal, descr):
        """A cached described feature (mad/DensityFeature.py:72-84; the descriptor cache of mad/MaD.py:857-875)."""
        self._from_file(index, main_bin, sec_bin, oct_scale, eqsp_size, coord, map_coord, subv_map_coord, Rfinal)
        self._take(subeqsp_size=subeqsp_size, lin_ar_subeqsp=descr)

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
