"""``mad.MapSpace`` of the reference -> the B200 implementation (mad_b200/MapSpace.py)."""
from mad_b200.MapSpace import MapSpace  # noqa: F401
