"""`scripts.weighted_svd` of the drop-in (scripts/weighted_svd.py:10-159)."""
from fusion4landslide_b200.weighted_svd import (refine_local_rigid_correspondences, weighted_procrustes,  # noqa: F401
                                                weighted_svd)
