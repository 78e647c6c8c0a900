"""``mad.Orientator`` of the reference -> the B200 implementation (mad_b200/Orientator.py)."""
from mad_b200.Orientator import Orientator  # noqa: F401
