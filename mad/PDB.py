"""``mad.PDB`` of the reference -> the B200 implementation (mad_b200/PDB.py)."""
from mad_b200.PDB import PDB  # noqa: F401
