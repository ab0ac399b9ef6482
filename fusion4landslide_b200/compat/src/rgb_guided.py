"""`from src.rgb_guided import Image_DVFs` (main_rgb_guided.py:13,109-111): the reference class with its per-patch
`local_rigid_refinement` loop (src/rgb_guided.py:981-1062) on the B200 kernels; every other name of the module (image
projection, lifting, 2D matching glue) is re-exported from the reference file.  Needs the reference tree behind compat/
on sys.path (the class is built on image networks and Open3D readers that stay in the reference)."""
from fusion4landslide_b200 import rgb_guided as _hot
from fusion4landslide_b200.compat import _upstream
from fusion4landslide_b200.rgb_guided import refine_local_rigid_correspondences  # noqa: F401

_up = _upstream.reexport("src.rgb_guided", globals(), skip=("refine_local_rigid_correspondences", "Image_DVFs"))
if _up is None:
    raise ImportError("src.rgb_guided: put the reference tree on sys.path behind compat/ (Image_DVFs is the reference's "
                      "class with one method replaced)")
UPSTREAM_BASE = _up.Image_DVFs
Image_DVFs = _hot.bind(UPSTREAM_BASE)
