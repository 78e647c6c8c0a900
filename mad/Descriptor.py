"""``mad.Descriptor`` of the reference -> the B200 implementation (mad_b200/Descriptor.py)."""
from mad_b200.Descriptor import Descriptor  # noqa: F401
