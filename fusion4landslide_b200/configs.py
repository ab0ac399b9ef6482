"""Config objects with the reference's key names (configs/landslide/fusion_3d_brienz.yaml, f2s3_brienz.yaml) for driving
the class-level entry points on IN-MEMORY tiles (tests, benchmarks, callers that already hold a tile): the same EasyDict
tree main_fusion.py:134-148 / main_f2s3.py:60-81 build per tile, with `tile_tensors` in place of the file paths."""
from .entry_c2f import edict


def fusion_config(tile_tensors, mode="only_3d", levels=(1, 2, 3), partition_type="superpoint", device="cuda",
                  feat_aggregate_model=None, output_root="/tmp/f4l_b200_out", tile_id=0, max_magnitude=5.0,
                  icp_threshold=0.1, output_tgt2src=False, write_results=False, **method_overrides):
    """mode: 'only_3d' (fusion_3d_brienz.yaml) or 'fusion' (coarse + fine matching on 2D-lifted and 3D matches;
    tile_tensors then carries corres_3d_from_2d_idx).  levels: a list -> multi-level + merge, an int -> one level."""
    fusion = mode == "fusion"
    method = edict(
        use_2d_matches=fusion, partition=False, partition_type=partition_type,
        level_of_superpoint=list(levels) if isinstance(levels, (list, tuple)) else int(levels),
        small_patch_removal=True, num_min_matches_for_small_patch=10, point_feat_compute=False, feat_type="DIPs",
        feat_dim=64, global_matching_from_3d_type="faiss",
        coarse_matching_fusion=fusion, coarse_matching_only_3d=not fusion, coarse_matching_only_2d=False,
        fine_matching_fusion=fusion, fine_matching_only_3d=not fusion, fine_matching_only_2d=False,
        feat_aggregate_type="learning_based", use_normal_3d_aggregation=True,
        use_img_patch_enhanced_3d_aggregation=False, use_img_pixel_enhanced_3d_aggregation=False,
        remove_low_quality_patch_matches=True, num_min_matches_for_quality_check=10, thres_dist_diff=0.5,
        thres_inlier_ratio=0.15, coarse_refinement_3d_type="nn_mutual", num_min_fine_match=10, weighting_svd=False,
        icp_refine=True, icp_register_type="only_matches", output_tgt2src=output_tgt2src, assign_type="assign_then_nn")
    for k, v in method_overrides.items():
        method[k] = v
    return edict(
        verbose=False, save_interim=False, device=device, tile_id=tile_id, tile_tensors=tile_tensors,
        feat_aggregate_model=feat_aggregate_model, write_interim_files=False, write_results=write_results,
        path_name=edict(input_root="", output_root=output_root, weight_dir="weights/",
                        pretrained_feature_aggregation_weight="feat_aggregation_3d.pth"),
        data=edict(dataset="brienz_tls", src_pcd="", tgt_pcd="", multiple_case=True),
        method=method,
        parameter_setting=edict(batch_size=1, num_workers=0, points_per_batch=1000, icp_threshold=icp_threshold,
                                max_magnitude=max_magnitude),
        visualization=edict(visualize_patch=False), debugging=edict(use_debugging=False))


def f2s3_config(tile_tensors, outlier_removal_nn, device="cuda", output_dir="/tmp/f4l_b200_out", tile_id=0,
                data_dir="synthetic", max_disp_magnitude=5.0, filter_median_magnitude=True, fill_gaps_c2c=False,
                refine_results=True, write_results=False):
    """The per-tile config main_f2s3.py:60-71 hands to Deformation_Analyze (flat EasyDict)."""
    return edict(
        verbose=False, save_interim=False, device=device, tile_id=tile_id, tile_tensors=tile_tensors,
        output_dir=output_dir, output_folder="demo_run", data_dir=data_dir, voxel_size=0.1, points_per_batch=1000,
        batch_size=1, num_workers=0, max_disp_magnitude=max_disp_magnitude,
        filter_median_magnitude=filter_median_magnitude, fill_gaps_c2c=fill_gaps_c2c, refine_results=refine_results,
        feat_compute=False, feat_type="DIPs", pcd_segment=False, segment_type="supervoxel", small_patch_removal=True,
        outlier_removal=True, removal_type="binary_classifier", correspondence_searching=True,
        correspondence_pruning=True, outlier_removal_nn=outlier_removal_nn, feat_desc_nn=None,
        write_interim_files=False, write_results=write_results)
