"""Small vector / rotation helpers with the reference's names (mad/math_utils.py:5-53)."""
import numpy as np

from .tables import euler_rodrigues, unit_vector  # noqa: F401  (re-exported)


def euler_rod_mat(axis, angle):
    """Reference name of tables.euler_rodrigues (mad/math_utils.py:15)."""
    return euler_rodrigues(axis, angle)


def polar_to_cart(theta, phi):
    return np.array([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)])


def get_rototrans_SVD(mobile, reference):
    """Kabsch superposition: returns (R, T) with reference ~= mobile @ R + T (mad/math_utils.py:29-53)."""
    mobile = np.asarray(mobile)
    reference = np.asarray(reference)
    if mobile.shape != reference.shape or mobile.shape[1] != 3:
        raise Exception("Descript> ERROR: Coordinates mismatch for SVD")
    cm, cr = mobile.mean(0), reference.mean(0)
    u, _, vt = np.linalg.svd(np.dot((mobile - cm).T, reference - cr))
    rot = np.dot(vt.T, u.T).T
    if np.linalg.det(rot) < 0:
        vt[2] = -vt[2]
        rot = np.dot(vt.T, u.T).T
    return rot, cr - np.dot(cm, rot)
