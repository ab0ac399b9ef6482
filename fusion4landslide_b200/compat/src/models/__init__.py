"""`from src.models import PointNetFeature, FilteringNetwork` (main_f2s3.py:20, main_fusion.py:15).
FilteringNetwork: same parameters / state_dict keys, plus the all-supervoxels forward and the kernel tail.
PointNetFeature (DIPs descriptor network) stays the reference's module."""
from .outlier_classifier import FilteringNetwork  # noqa: F401
from fusion4landslide_b200.compat import _upstream

try:
    _lfd = _upstream.load("src.models.local_feature_descriptor")
except ImportError:
    _lfd = None
if _lfd is not None:
    PointNetFeature = _lfd.PointNetFeature
else:
    class PointNetFeature(object):
        def __init__(self, *a, **k):
            raise NotImplementedError("PointNetFeature is the reference's descriptor network "
                                      "(src/models/local_feature_descriptor.py); put the reference tree on sys.path")
