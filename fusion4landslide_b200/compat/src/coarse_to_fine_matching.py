"""`from src.coarse_to_fine_matching import Coarse2Fine` (main_fusion.py:8,147-148).

With the reference tree importable, `Coarse2Fine` = HotPathMixin over the reference's own class: image matching,
lifting, partitioning, descriptor networks, I/O stay the reference's code; the hot methods run on libf4l_b200.so.
Otherwise the stand-alone base of fusion4landslide_b200.entry_c2f is used."""
import warnings

from fusion4landslide_b200 import entry_c2f
from fusion4landslide_b200.coarse_to_fine import merge_correspondences_by_priority_with_distance_threshold  # noqa: F401
from fusion4landslide_b200.compat import _upstream

_base = None
if _upstream.find("src.coarse_to_fine_matching") is not None:
    try:
        _base = _upstream.load("src.coarse_to_fine_matching").Coarse2Fine
    except ImportError as e:                     # e.g. open3d / faiss missing
        warnings.warn("reference Coarse2Fine not importable (%s); using the stand-alone base" % (e,))
UPSTREAM_BASE = _base
Coarse2Fine = entry_c2f.bind(_base) if _base is not None else entry_c2f.Coarse2Fine
