"""`src.functions` of the drop-in: the Kabsch / transform / C2C functions run on the kernels, `point_cloud_tiling`
(native PCL tiling, src/functions.py:147-177) and anything else is the reference's own."""
from fusion4landslide_b200.functions import (compute_c2c, kabsch_transformation_estimation,  # noqa: F401
                                             transform_point_cloud, transformation_residuals)
from fusion4landslide_b200.compat import _upstream

_OURS = ("kabsch_transformation_estimation", "transformation_residuals", "transform_point_cloud", "compute_c2c")
try:
    _upstream.reexport("src.functions", globals(), skip=_OURS)
except ImportError:                              # the reference module needs its native tiling extension
    pass

if "point_cloud_tiling" not in globals():
    def point_cloud_tiling(config):
        raise NotImplementedError("point_cloud_tiling is the reference's native PCL tiler (cpp_core/pcd_tiling); put the "
                                  "reference tree on sys.path behind compat/ to use it")
