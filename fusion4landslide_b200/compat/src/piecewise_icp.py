"""`from src.piecewise_icp import Piecewise_ICP` (main_piecewise_icp.py:14,93) on the B200 kernels."""
from fusion4landslide_b200.piecewise_icp import Piecewise_ICP, piecewise_icp  # noqa: F401
