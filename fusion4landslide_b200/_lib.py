"""ctypes binding of libf4l_b200.so (the C ABI declared in include/f4l_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("F4L_LIB") or os.path.join(_HERE, "libf4l_b200.so")   # F4L_LIB: A/B experiments only

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_f64 = ctypes.c_double
c_size = ctypes.c_size_t


MAX_PEERS = 7          # F4L_MAX_PEERS
PEER_HANDLE_BYTES = 64  # F4L_PEER_HANDLE_BYTES


class FineParams(ctypes.Structure):
    """f4l_fine_params (include/f4l_b200.h)."""
    _fields_ = [
        ("mode", c_i32), ("remove_low_quality", c_i32), ("num_min_quality", c_i32),
        ("thres_dist_diff", c_f32), ("thres_inlier_ratio", c_f32), ("num_min_fine_match", c_i32),
        ("icp_refine", c_i32), ("assign_type", c_i32), ("output_tgt2src", c_i32),
        ("icp_threshold", c_f64), ("median_max_resolution", c_f64), ("icp_max_iter", c_i32),
    ]


class FineBuffers(ctypes.Structure):
    """f4l_fine_buffers (include/f4l_b200.h)."""
    _fields_ = [
        ("src_pts", c_void_p), ("n_src", c_i32), ("tgt_pts", c_void_p), ("n_tgt", c_i32),
        ("corr3d", c_void_p), ("corr2d", c_void_p),
        ("sp_idx", c_void_p), ("sp_ptr", c_void_p), ("tp_idx", c_void_p), ("tp_ptr", c_void_p),
        ("tgt_patch_of_point", c_void_p), ("pair_tgt_patch", c_void_p), ("Q", c_i32),
        ("n_src_items", c_i32), ("n_tgt_items", c_i32), ("d_median_resolution", c_void_p),
        ("T", c_void_p), ("T64", c_void_p), ("status", c_void_p), ("K", c_void_p),
        ("fitness", c_void_p), ("rmse", c_void_p), ("iters", c_void_p),
        ("ratio_inlier", c_void_p), ("dist_mean", c_void_p),
        ("dense", c_void_p), ("sparse", c_void_p), ("tgt2src", c_void_p), ("counts", c_void_p),
        ("n_peers", c_i32), ("peer_dense", c_void_p * MAX_PEERS),
        ("sparse_pair_rows", c_void_p), ("median_ready_event", c_void_p), ("phases", c_i32),
        ("icp_fragile", c_void_p), ("corr3d_tgt", c_void_p), ("corr2d_tgt", c_void_p),
    ]


# name -> (restype, argtypes).  Must list every symbol include/f4l_b200.h declares
# (tests/test_abi.py checks this against the header).
P = c_void_p
SIGNATURES = {
    "f4l_abi_version": (c_int, []),
    "f4l_last_error": (ctypes.c_char_p, []),
    "f4l_launch_count": (ctypes.c_longlong, []),
    "f4l_launch_count_reset": (None, []),
    "f4l_profile_enable": (None, [c_int]),
    "f4l_profile_collect": (c_int, []),
    "f4l_profile_get": (c_int, [c_int, ctypes.c_char_p, c_int, ctypes.POINTER(c_f64), ctypes.POINTER(ctypes.c_longlong)]),
    "f4l_profile_reset": (None, []),
    "f4l_segmented_kabsch": (c_int, [P, P, P, P, P, P, P, c_i32, c_f32, c_f32, c_int, P, P, P, P, P, P]),
    "f4l_apply_transforms": (c_int, [P, P, P, P, P, P, c_i32, P, c_int, P, P, P]),
    "f4l_rigidity_check": (c_int, [P, P, P, P, P, P, c_i32, c_f32, P, P, P]),
    "f4l_segmented_median": (c_int, [P, P, P, c_i32, P, P]),
    "f4l_f2s3_prune_tail": (c_int, [P, P, P, P, c_i32, c_f32, P, P, P, P, P, P]),
    "f4l_knn_grid_workspace_bytes": (c_size, [c_i32, c_i32]),
    "f4l_knn_grid": (c_int, [P, c_i32, P, c_i32, c_i32, c_f32, c_f32, P, P, P, c_size, P]),
    "f4l_knn_grid_ties": (c_int, [P, c_i32, P, c_i32, c_i32, c_f32, c_f32, c_f32, P, P, c_size, P]),
    "f4l_select_kth_workspace_bytes": (c_size, [c_i32]),
    "f4l_select_kth": (c_int, [P, c_i32, c_i32, c_i32, c_i32, c_i32, P, P, c_size, P]),
    "f4l_median_resolution_workspace_bytes": (c_size, [c_i32, c_i32]),
    "f4l_median_resolution": (c_int, [P, c_i32, P, c_i32, P, P, c_size, P]),
    "f4l_segmented_nn": (c_int, [P, P, P, P, P, P, P, P, c_i32, P, P, P, P, P]),
    "f4l_patch_icp": (c_int, [P, P, P, P, P, P, P, P, P, c_i32, P, c_f64, c_i32, c_f64, c_f64,
                              P, P, P, P, P, P]),
    "f4l_patch_icp_ex": (c_int, [P, P, P, P, P, P, P, P, P, c_i32, P, c_f64, c_i32, c_f64, c_f64,
                                 P, P, P, P, P, P, c_f64, P]),
    "f4l_peer_push": (c_int, [P, P, c_i32, ctypes.c_int64, P, c_i32, c_i32, P]),
    "f4l_fine_fit_tiles": (c_int, [P, P, P, c_i32, c_i32, P, P]),
    "f4l_desc_nn_workspace_bytes": (c_size, [c_i32, c_i32, c_i32, c_int]),
    "f4l_desc_nn": (c_int, [P, c_i32, P, c_i32, c_i32, P, P, c_f32, c_int, c_int, P, P, P, P, P, c_size, P]),
    "f4l_desc_nn_ex": (c_int, [P, c_i32, P, c_i32, c_i32, P, P, c_f32, c_int, c_int, P, P, P, P, P, P, c_f64, P, c_size, P]),
    "f4l_scatter_global_matches_workspace_bytes": (c_size, [c_i32]),
    "f4l_scatter_global_matches": (c_int, [P, P, P, c_i32, P, P, c_f32, P, c_i32, P, c_size, P]),
    "f4l_vote_tgt_patch": (c_int, [P, P, P, c_i32, P, c_i32, P, c_i32, P, P, P, P]),
    "f4l_magnitude_mask": (c_int, [P, c_i32, c_i32, c_f32, P, c_f32, c_int, P, P, P]),
    "f4l_labels_to_csr_workspace_bytes": (c_size, [c_i32]),
    "f4l_labels_to_csr": (c_int, [P, c_i32, c_i32, P, P, P, P, P, P, c_size, P]),
    "f4l_gather_pairs_csr_workspace_bytes": (c_size, [c_i32]),
    "f4l_gather_pairs_csr": (c_int, [P, P, P, c_i32, P, P, c_i32, P, c_size, P]),
    "f4l_piecewise_icp_workspace_bytes": (c_size, [c_i32, c_i32]),
    "f4l_piecewise_icp": (c_int, [P, c_i32, P, c_i32, c_f64, c_i32, c_i32, P, P, P, P, P, P, P, P, c_size, P]),
    "f4l_peer_alloc": (c_int, [c_size, ctypes.POINTER(c_void_p), ctypes.c_char_p]),
    "f4l_peer_open": (c_int, [ctypes.c_char_p, ctypes.POINTER(c_void_p)]),
    "f4l_peer_close": (c_int, [P]),
    "f4l_peer_free": (c_int, [P]),
    "f4l_peer_enable_access": (c_int, [c_i32]),
    "f4l_dips_workspace_bytes": (c_size, [c_i32]),
    "f4l_dips_build": (c_int, [P, c_i32, c_f64, P, c_size, P]),
    "f4l_dips_patches": (c_int, [P, c_i32, c_i32, c_f64, c_i32, P, ctypes.c_uint64, P, P, P, P, c_size, P]),
    "f4l_dips_patches_large": (c_int, [P, c_i32, c_i32, c_f64, c_i32, ctypes.c_uint64, P, P, P, P, c_size, P]),
    "f4l_voxel_downsample_workspace_bytes": (c_size, [c_i32]),
    "f4l_voxel_downsample": (c_int, [P, c_i32, c_f64, P, P, P, P, c_size, P]),
    "f4l_fine_matching_workspace_bytes": (c_size, [c_i32, c_i32, c_i32, c_i32]),
    "f4l_fine_matching": (c_int, [ctypes.POINTER(FineParams), ctypes.POINTER(FineBuffers), P, c_size, P]),
    "f4l_segment_scale_maxabs": (c_int, [P, c_i32, P, c_i32, c_i32, P, P]),
    "f4l_segment_norm2_relu": (c_int, [P, P, c_i32, c_i32, c_f32, P, P, P]),
    "f4l_segment_attention_pool": (c_int, [P, P, P, P, c_i32, c_i32, c_f32, c_i32, P, P]),
    "f4l_segment_mean": (c_int, [P, P, c_i32, c_i32, P, P]),
    "f4l_host_expand_sparse": (ctypes.c_longlong, [P, P, c_i32, P, c_i32]),
    "f4l_host_pack_corr_targets": (None, [P, ctypes.c_int64, P, c_i32]),
}


class F4LError(RuntimeError):
    pass


def lib():
    """Load the shared library once.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise F4LError(
                "libf4l_b200.so is not built (%s). Run `python -m fusion4landslide_b200.build`; "
                "there is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            # a symbol the library does not export stays unbound: calling it raises
            # AttributeError (tests/test_abi.py checks that none is missing)
            fn = getattr(L, name, None)
            if fn is not None:
                fn.restype = res
                fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().f4l_last_error().decode("utf-8", "replace")
        raise F4LError("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t, dtype=None, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (dtype checked)."""
    if t is None:
        if allow_none:
            return None
        raise F4LError("required tensor is None")
    if not t.is_cuda:
        raise F4LError("tensor must live on a CUDA device (no CPU fallback)")
    if not t.is_contiguous():
        raise F4LError("tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise F4LError("expected dtype %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


def stream_ptr(device=None):
    if device is not None and torch.device(device).type != "cuda":
        raise F4LError("tensor must live on a CUDA device (no CPU fallback)")
    return torch.cuda.current_stream(device).cuda_stream


def profile_table():
    """Collect the in-stream kernel timings: {kernel name: (total ms, launches)}."""
    L = lib()
    n = L.f4l_profile_collect()
    out = {}
    buf = ctypes.create_string_buffer(128)
    ms = c_f64()
    cnt = ctypes.c_longlong()
    for i in range(n):
        if L.f4l_profile_get(i, buf, 128, ctypes.byref(ms), ctypes.byref(cnt)) == 0:
            out[buf.value.decode()] = (ms.value, cnt.value)
    return out
