"""fusion4landslide_b200 -- B200 (sm_100a) implementation of the patch-wise 3D correspondence and
rigid-estimation hot path of gseg-ethz/fusion4landslide, behind the reference's Python call
signatures.  All computation runs in libf4l_b200.so (hand-written CUDA, C ABI in
include/f4l_b200.h); there is no CPU fallback."""
__version__ = "0.1.0"
