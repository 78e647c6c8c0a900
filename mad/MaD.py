"""``mad.MaD``: loads the REFERENCE's own orchestrator file on top of this package.

Nothing of ``mad/MaD.py`` (1 089 lines of run / clustering / refinement / scoring logic, out of the hot path's scope) is
re-implemented or shipped here.  The user points at their copy of the reference's file:

    export MAD_REFERENCE_MAD_PY=/path/to/LBM-EPFL-MaD/mad/MaD.py      # or drop it next to this file as _reference_MaD.py
    python run_MaD.py                                                 # unchanged (run_MaD.py:63-76)

The file is executed AS this module, so its relative imports (``from .MapSpace import MapSpace`` ... ``mad/MaD.py:13-22``)
resolve to this package, i.e. to the CUDA path: ``MaD._describe_struct`` (``mad/MaD.py:358-368``) then runs a1-a12 on the
GPU without a changed line.  ``MaD._match_dsc`` (``mad/MaD.py:414-453``: NumPy dgemm + np.where + a cKDTree query per pair)
is rebound to ``mad_b200.pipeline.match_dsc_lists`` -- same arguments, same return value, computed by the tcgen05 matcher
and the repeatability kernel; set MAD_B200_KEEP_REFERENCE_MATCH=1 to keep the reference's own lines.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_path = _os.environ.get("MAD_REFERENCE_MAD_PY") or _os.path.join(_here, "_reference_MaD.py")
if not _os.path.isfile(_path):
    raise ImportError("mad.MaD: the reference's orchestrator is not shipped with mad_b200 -- set MAD_REFERENCE_MAD_PY to the "
                      "reference's mad/MaD.py (or copy it to %s)" % _os.path.join(_here, "_reference_MaD.py"))
with open(_path) as _f:
    exec(compile(_f.read(), _path, "exec"), globals())          # relative imports resolve inside this package

if not _os.environ.get("MAD_B200_KEEP_REFERENCE_MATCH"):
    def _match_dsc(self, lo_dsc_list, hi_dsc_list, anchor_dist_thresh=4, cc_threshold=0.65):
        """mad/MaD.py:414-453 on the device (mad_b200.pipeline.match_dsc_lists)."""
        from mad_b200.pipeline import match_dsc_lists
        return match_dsc_lists(lo_dsc_list, hi_dsc_list, anchor_dist_thresh, cc_threshold)
    MaD._match_dsc = _match_dsc                                   # noqa: F821  (defined by the executed file)
