"""``mad.math_utils`` of the reference (mad/math_utils.py:5-56) -> mad_b200/math_utils.py."""
from mad_b200.math_utils import unit_vector, euler_rod_mat, get_rototrans_SVD, polar_to_cart  # noqa: F401
