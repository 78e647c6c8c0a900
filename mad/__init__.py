"""Drop-in ``mad`` package: the reference's module names (LBM-EPFL/MaD ``mad/*.py``) bound to the B200 implementation.

``run_MaD.py`` and the notebooks do ``from mad import MaD`` / ``import mad.MaD as MaD`` and only ever touch ``MaD.MaD``
(``run_MaD.py:63-76``).  This package supplies every module ``mad/MaD.py`` imports (``mad/MaD.py:13-22``) from
``mad_b200`` -- same class names, constructor arguments, attributes and error behaviour -- so the ONE file a user brings
is the reference's own orchestrator ``mad/MaD.py`` (see ``mad/MaD.py`` here: it loads that file; no orchestrator is
re-implemented).  There is no CPU fallback behind these names: without libmad_b200.so / a CUDA device they raise.
"""
