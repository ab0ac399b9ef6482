"""CPU tests of the class-level drop-in boundary (compat/): with compat/ ahead of the reference tree on sys.path the
reference's own import lines resolve to classes whose hot methods are this repository's, while everything else is
still the reference's code.  Needs /root/reference (skipped on the GPU box, where the stand-alone bases are used);
no kernel is launched here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("F4L_REFERENCE_ROOT", "/root/reference")
COMPAT = os.path.join(ROOT, "fusion4landslide_b200", "compat")

HOT_C2F = ["_compute_median_resolution", "_voxel_subsampling", "prepare_pts2spt_dict", "global_matches_from_3d",
           "coarse_matching_with_different_types", "fine_matching_with_different_types", "implement_c2f_matching",
           "_compute_spt_feat_and_coord_with_fused_feats"]
HOT_F2S3 = ["_compute_median_resolution", "correspondence_searching", "correspondence_pruning"]

_PROBE = r"""
import json, sys, inspect
sys.path.insert(0, %(root)r)
from oracle import ref_shim
ref_shim.load()                                   # stubs for open3d / faiss / hnswlib ...; puts the reference on sys.path
sys.path.insert(0, %(compat)r)                    # the drop-in goes FIRST
import src.coarse_to_fine_matching as c2f_mod
import src.f2s3 as f2s3_mod
import src.functions as fn
import src.piecewise_icp as pw
import scripts.weighted_svd as wsvd
import utils.o3d_tools as o3t
import utils.common as common                     # not shadowed: must be the reference's file
import src.data_loader as dl                      # not shadowed either
from src.models import PointNetFeature, FilteringNetwork
out = {}
C = c2f_mod.Coarse2Fine
out["c2f_mro"] = [k.__module__ + "." + k.__name__ for k in C.__mro__]
out["c2f_hot"] = {m: getattr(C, m).__module__ for m in %(hot_c2f)r}
out["c2f_inherited"] = {m: getattr(C, m).__module__ for m in ("save_process_dvf", "load_partition", "_read_data",
                                                              "global_matches_from_2d_with_different_types", "start_debugging")}
D = f2s3_mod.Deformation_Analyze
out["f2s3_mro"] = [k.__module__ + "." + k.__name__ for k in D.__mro__]
out["f2s3_hot"] = {m: getattr(D, m).__module__ for m in %(hot_f2s3)r}
out["f2s3_inherited"] = {m: getattr(D, m).__module__ for m in ("compute_features", "implement_segmentation")}
out["fn"] = {k: getattr(fn, k).__module__ for k in ("kabsch_transformation_estimation", "transform_point_cloud", "compute_c2c",
                                                    "point_cloud_tiling")}
out["pw"] = pw.Piecewise_ICP.__module__
out["wsvd"] = wsvd.refine_local_rigid_correspondences.__module__
out["o3t"] = {k: getattr(o3t, k).__module__ for k in ("icp_registration", "tensor2pcd", "pcd2tensor")}
out["common_file"] = common.__file__
out["dl_file"] = dl.__file__
out["models"] = [PointNetFeature.__module__, FilteringNetwork.__module__]
ref_c2f = sys.modules["_f4l_upstream.src.coarse_to_fine_matching"].Coarse2Fine
ref_f2s3 = sys.modules["_f4l_upstream.src.f2s3"].Deformation_Analyze
sig = lambda f: str(inspect.signature(f))
out["sig_equal"] = {m: sig(getattr(C, m)) == sig(getattr(ref_c2f, m)) for m in %(hot_c2f)r if hasattr(ref_c2f, m)}
out["sig_equal_f2s3"] = {m: sig(getattr(D, m)) == sig(getattr(ref_f2s3, m)) for m in %(hot_f2s3)r}
out["ctor"] = [sig(ref_c2f.__init__), sig(C.__init__), sig(ref_f2s3.__init__), sig(D.__init__)]
import scripts.weighted_svd as _w
ref_w = ref_shim  # noqa
print("JSON" + json.dumps(out))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_compat_shadows_only_the_hot_path():
    code = _PROBE % dict(root=ROOT, compat=COMPAT, hot_c2f=HOT_C2F, hot_f2s3=HOT_F2S3)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("JSON")][-1][4:])
    # Coarse2Fine = our mixin over the REFERENCE's class
    assert out["c2f_mro"][1] == "fusion4landslide_b200.entry_c2f.HotPathMixin"
    assert "_f4l_upstream.src.coarse_to_fine_matching.Coarse2Fine" in out["c2f_mro"]
    assert "_f4l_upstream.src.coarse_to_fine_matching_base.Coarse2Fine_Base" in out["c2f_mro"] or \
        any(k.endswith("Coarse2Fine_Base") for k in out["c2f_mro"])
    assert all(v == "fusion4landslide_b200.entry_c2f" for v in out["c2f_hot"].values()), out["c2f_hot"]
    assert all("fusion4landslide_b200" not in v for v in out["c2f_inherited"].values()), out["c2f_inherited"]
    assert out["f2s3_mro"][1] == "fusion4landslide_b200.entry_f2s3.HotPathMixin"
    assert all(v == "fusion4landslide_b200.entry_f2s3" for v in out["f2s3_hot"].values())
    assert all("fusion4landslide_b200" not in v for v in out["f2s3_inherited"].values())
    # functions: hot ones ours, the native tiler the reference's
    assert out["fn"]["kabsch_transformation_estimation"] == "fusion4landslide_b200.functions"
    assert out["fn"]["compute_c2c"] == "fusion4landslide_b200.functions"
    assert out["fn"]["point_cloud_tiling"].startswith("_f4l_upstream")
    assert out["pw"] == "fusion4landslide_b200.piecewise_icp"
    assert out["wsvd"] == "fusion4landslide_b200.weighted_svd"
    assert out["o3t"]["icp_registration"] == "fusion4landslide_b200.o3d_tools"
    assert out["o3t"]["tensor2pcd"].startswith("_f4l_upstream")
    assert out["common_file"].startswith(REF) and out["dl_file"].startswith(REF)
    assert out["models"][0].startswith("_f4l_upstream") and out["models"][1] == "fusion4landslide_b200.nets"
    # same call signatures as the reference methods they replace
    assert all(out["sig_equal"].values()), out["sig_equal"]
    assert all(out["sig_equal_f2s3"].values()), out["sig_equal_f2s3"]
    assert out["ctor"][0] == out["ctor"][1] and out["ctor"][2] == out["ctor"][3], out["ctor"]


def test_standalone_classes_import_without_reference():
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import src.coarse_to_fine_matching as a, src.f2s3 as b, src.functions as f\n"
            "from src.models import FilteringNetwork\n"
            "assert a.UPSTREAM_BASE is None and b.UPSTREAM_BASE is None\n"
            "assert a.Coarse2Fine.__mro__[2].__name__ == 'StandaloneBase'\n"
            "n = FilteringNetwork(); assert len(n.state_dict()) == 52\n"
            "print('ok')") % (ROOT, COMPAT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
