"""`from src.f2s3 import Deformation_Analyze` (main_f2s3.py:12,72-81): the reference class with
`_compute_median_resolution`, `correspondence_searching`, `correspondence_pruning` on the B200 kernels."""
import warnings

from fusion4landslide_b200 import entry_f2s3
from fusion4landslide_b200.compat import _upstream

_base = None
if _upstream.find("src.f2s3") is not None:
    try:
        _base = _upstream.load("src.f2s3").Deformation_Analyze
    except ImportError as e:
        warnings.warn("reference Deformation_Analyze not importable (%s); using the stand-alone base" % (e,))
UPSTREAM_BASE = _base
Deformation_Analyze = entry_f2s3.bind(_base) if _base is not None else entry_f2s3.Deformation_Analyze
