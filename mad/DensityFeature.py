"""``mad.DensityFeature`` of the reference -> mad_b200/DensityFeature.py (record + the device-backed FeatureList)."""
from mad_b200.DensityFeature import DensityFeature, FeatureList  # noqa: F401
