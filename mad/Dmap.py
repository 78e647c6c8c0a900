"""``mad.Dmap`` of the reference -> the B200 implementation (mad_b200/Dmap.py)."""
from mad_b200.Dmap import Dmap  # noqa: F401
