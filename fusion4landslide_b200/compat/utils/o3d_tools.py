"""`utils.o3d_tools` of the drop-in: `icp_registration` (utils/o3d_tools.py:12-71) runs as one persistent CUDA
kernel; the converters and visualisation helpers are re-exported from the reference's module when it is importable."""
from fusion4landslide_b200.o3d_tools import icp_registration  # noqa: F401
from fusion4landslide_b200.compat import _upstream

try:
    _upstream.reexport("utils.o3d_tools", globals(), skip=("icp_registration",))
except ImportError:                              # open3d missing: only the hot function is available
    pass
