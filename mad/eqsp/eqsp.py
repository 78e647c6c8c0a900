"""``mad.eqsp.eqsp`` of the reference (mad/eqsp/eqsp.py:12-87) -> mad_b200/eqsp/eqsp.py (tables embedded, no cwd-relative files)."""
from mad_b200.eqsp.eqsp import EQSP_Sphere  # noqa: F401
