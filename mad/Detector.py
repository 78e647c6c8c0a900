"""``mad.Detector`` of the reference -> the B200 implementation (mad_b200/Detector.py)."""
from mad_b200.Detector import Detector  # noqa: F401
