"""`src.models.outlier_classifier` of the drop-in (src/models/outlier_classifier.py:10-105)."""
from fusion4landslide_b200.nets import FilteringNetwork, PointCN  # noqa: F401
