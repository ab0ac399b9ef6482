/*
 * f4l_b200.h -- C ABI of libf4l_b200.so: the B200 (sm_100a) implementation of the patch-wise 3D
 * correspondence and rigid-estimation hot path of gseg-ethz/fusion4landslide.
 *
 * The reference has no FFI for this path (it is Python calling torch / Open3D / scipy), so each
 * entry point names the reference Python function or code range it replaces
 * (paths relative to the reference tree; "base.py" = src/coarse_to_fine_matching_base.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name starts with
 *     `h_`; buffers are caller-owned, inputs are never written, outputs are fully written
 *   - point arrays are row-major (n,3); f32 unless the name ends in 64
 *   - segments ("patches") are CSR: ptr[Q+1] int32 offsets into a packed item array.  Where a
 *     pair (start,count) is taken instead, segment q owns items [start[q], start[q]+count[q])
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the device, no host-visible scalars are produced unless documented
 *   - workspaces are caller-provided; f4l_*_workspace_bytes() gives the size (host function)
 *   - return value: 0 = ok, <0 = error (F4L_E_*); f4l_last_error() gives the message of the
 *     last failing call on this thread
 *   - there is NO CPU fallback: every function fails with F4L_E_CUDA if no sm_100 device/kernel
 *     image is usable
 */
#ifndef F4L_B200_H
#define F4L_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F4L_ABI_VERSION 4

#if defined(__GNUC__)
#define F4L_API __attribute__((visibility("default")))
#else
#define F4L_API
#endif

#define F4L_OK 0
#define F4L_E_ARG (-1)   /* bad argument (null pointer, negative size, unsupported k / D ...) */
#define F4L_E_CUDA (-2)  /* CUDA runtime error; message in f4l_last_error() */
#define F4L_E_WORKSPACE (-3) /* workspace too small */

F4L_API int f4l_abi_version(void);
F4L_API const char* f4l_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench bookkeeping) */
F4L_API long long f4l_launch_count(void);
F4L_API void f4l_launch_count_reset(void);
/* Optional in-stream timing of every kernel the library launches (CUDA events recorded on the
 * launching stream around each launch).  Off by default; bench.py turns it on for its roofline
 * leg only.  f4l_profile_collect() synchronises the events and returns the number of table rows;
 * f4l_profile_get(i) reads row i: kernel name, accumulated milliseconds, launches. */
F4L_API void f4l_profile_enable(int on);
F4L_API int f4l_profile_collect(void);
F4L_API int f4l_profile_get(int index, char* h_name, int name_cap, double* h_total_ms, long long* h_launches);
F4L_API void f4l_profile_reset(void);

/* ------------------------------------------------------------------------------------------
 * (d) Weighted Kabsch / Procrustes, one fit per segment.                           kernel K-d
 * Replaces scripts/weighted_svd.py:58-129 weighted_procrustes (variant 0: w/(sum w+eps),
 * reflection fix sign(det)) and src/functions.py:12-85 kabsch_transformation_estimation
 * (variant 1: w/(sum w+eps), means divided again by (sum+eps), raw det).
 *   src,tgt : packed (K,3) pairs, or base arrays gathered through src_idx/tgt_idx (K) when
 *             those are non-null
 *   w       : (K) weights or NULL (all ones);  weight_thresh: w < thresh -> 0 (variant 0)
 *   seg_start,seg_count : (Q) ; seg_count may be NULL -> CSR: count = seg_start[q+1]-seg_start[q]
 *                         (seg_start then has Q+1 entries)
 *   R (Q,9), t (Q,3) f32; T64 (Q,16) f64 row-major 4x4 or NULL; res (K) residual
 *   ||R s + t - tgt|| or NULL; flag (Q) uint8 or NULL: 1 = degenerate (K<1 or non-finite) ->
 *   identity, mirroring functions.py:62-71.
 *   Packed inputs are staged with 16-byte-granular bulk copies: up to 12 bytes before the first and after the last
 *   element of src / tgt / w are read (never used) -- always inside a 16-byte line that holds valid bytes of the array.
 */
#define F4L_KABSCH_PROCRUSTES 0
#define F4L_KABSCH_F2S3 1
F4L_API int f4l_segmented_kabsch(const float* src, const float* tgt, const int32_t* src_idx,
                         const int32_t* tgt_idx, const float* w, const int32_t* seg_start,
                         const int32_t* seg_count, int32_t Q, float eps, float weight_thresh,
                         int variant, float* R, float* t, double* T64, float* res, uint8_t* flag,
                         void* stream);

/* (d)/apply: rows [p | R p + t] (or [R^T (p - t) | p] when inverse != 0) for every item of
 * every segment, plus the displacement magnitude.                                   kernel K-f
 * Replaces src/functions.py:107-124 transform_point_cloud and base.py:3373-3380, 3389-3390,
 * 3463-3472.  pts (n,3) gathered through idx (items) when idx != NULL.  T (Q,16) f32 row-major.
 * seg_skip (Q) uint8 or NULL: segments with seg_skip != 0 emit nothing; out_start (Q) gives the
 * first output row of each segment.  dvf (rows,6), mag (rows) or NULL. */
F4L_API int f4l_apply_transforms(const float* pts, const int32_t* idx, const int32_t* seg_start,
                         const int32_t* seg_count, const int32_t* out_start,
                         const uint8_t* seg_skip, int32_t Q, const float* T, int inverse,
                         float* dvf, float* mag, void* stream);

/* (c) rigidity / isometry check of each segment's correspondences.                  kernel K-c
 * Replaces base.py:3308-3317: ratio_inlier = (#(|dS-dT| <= thres) - K)/(K(K-1)),
 * dist_mean = sum triu(|dS-dT|,1)/(K(K-1)/2).  Outputs (Q) f32 each. */
F4L_API int f4l_rigidity_check(const float* src, const float* tgt, const int32_t* src_idx,
                       const int32_t* tgt_idx, const int32_t* seg_start,
                       const int32_t* seg_count, int32_t Q, float thres_dist_diff,
                       float* ratio_inlier, float* dist_mean, void* stream);

/* (c) lower median (torch.median semantics) of each segment of x.  med (Q) f32.
 * Replaces torch.median at src/models/outlier_classifier.py:80,91. */
F4L_API int f4l_segmented_median(const float* x, const int32_t* seg_start, const int32_t* seg_count,
                         int32_t Q, float* med, void* stream);

/* (c)+(d) F2S3 pruning tail per supervoxel: Kabsch(w) -> residuals -> res < coeff*median ->
 * (>=5 inliers and median < 0.5) -> refit with 0/1 weights.
 * Replaces src/models/outlier_classifier.py:71-105 (everything after the network forward).
 * corr (K,6) rows [src|tgt] f32, scores (K).  Outputs R (Q,9), t (Q,3), robust (Q) uint8,
 * res (K) residuals of the final fit (REQUIRED: also scratch), median (Q) of the first fit or NULL. */
F4L_API int f4l_f2s3_prune_tail(const float* corr, const float* scores, const int32_t* seg_start,
                        const int32_t* seg_count, int32_t Q, float coeff, float* R, float* t,
                        uint8_t* robust, float* res, float* median, void* stream);

/* ------------------------------------------------------------------------------------------
 * (a) exact xyz k-nearest neighbours on a uniform grid.                             kernel K-a
 * Replaces sklearn NearestNeighbors(kd_tree) at base.py:2727-2736, src/f2s3.py:492-501,
 * src/functions.py:139-142 and scipy cKDTree.query at base.py:1038-1042.
 *   q (N,3), r (M,3); k in [1,8]; max_radius <= 0 -> unbounded.
 *   idx (N,k) int32 (-1 = none within radius), d2 (N,k) f32 squared distances, ascending;
 *   exact ties are broken towards the lower reference index.
 *   Set r == q (same pointer) for a self query (the point itself is neighbour 0).
 * cell <= 0 lets the library choose the cell size from the bounding box and M. */
F4L_API size_t f4l_knn_grid_workspace_bytes(int32_t N, int32_t M);
F4L_API int f4l_knn_grid(const float* q, int32_t N, const float* r, int32_t M, int32_t k,
                 float max_radius, float cell, int32_t* idx, float* d2, void* workspace,
                 size_t workspace_bytes, void* stream);
/* Tie flags of the same query ("indices identical except rows flagged as distance ties"): tie (N)
 * uint8 = 1 when two adjacent squared distances among the first k + 1 neighbours differ by <= eps_rel * (the larger one)
 * + 1e-12, i.e. where another exact search may order the indices differently.  k in [1,7]; same workspace size. */
F4L_API int f4l_knn_grid_ties(const float* q, int32_t N, const float* r, int32_t M, int32_t k, float max_radius,
                      float cell, float eps_rel, uint8_t* tie, void* workspace, size_t workspace_bytes, void* stream);

/* k-th smallest (0-based) of x (n) f32, written to out[0] (device).  With k2 >= 0 also the k2-th
 * to out[1] (np.median of an even count averages the two middle elements: base.py:2732). */
F4L_API size_t f4l_select_kth_workspace_bytes(int32_t n);
F4L_API int f4l_select_kth(const float* x, int32_t n, int32_t stride, int32_t offset, int32_t k,
                   int32_t k2, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* (a) A1: median point-cloud resolution = max over the two epochs of the median distance to the
 * nearest OTHER point (k=2 self query).  Replaces base.py:2716-2754 and src/f2s3.py:481-508
 * (_compute_median_resolution).  out[0] (device f32). */
F4L_API size_t f4l_median_resolution_workspace_bytes(int32_t n_src, int32_t n_tgt);
F4L_API int f4l_median_resolution(const float* src, int32_t n_src, const float* tgt, int32_t n_tgt,
                          float* out, void* workspace, size_t workspace_bytes, void* stream);

/* (a) segmented 1-NN: for every query item of segment q, the nearest reference item of the SAME
 * segment pair, optionally after applying T[q] to the query, kept iff d2 < thr[q]^2.
 * Replaces base.py:48-97 refine_dvfs_with_threshold (Open3D KDTreeFlann per point).
 *   nn (items) int32 = position in the reference segment's item list (or -1), d2 (items). */
F4L_API int f4l_segmented_nn(const float* qpts, const int32_t* qidx, const int32_t* q_start,
                     const int32_t* q_count, const float* rpts, const int32_t* ridx,
                     const int32_t* r_start, const int32_t* r_count, int32_t Q, const float* T,
                     const float* thr, int32_t* nn, float* d2, void* stream);

/* ------------------------------------------------------------------------------------------
 * (e) per-patch point-to-point ICP, the kNN <-> Kabsch iterate loop in one persistent kernel.
 * Replaces utils/o3d_tools.py:12-71 icp_registration -> Open3D registration_icp
 * (TransformationEstimationPointToPoint(False), ICPConvergenceCriteria(1e-6,1e-6,30)).
 *   source segment q = src items, target segment q = tgt items (gathered through *_idx if given);
 *   T0 (Q,16) f64 row-major initial transforms (NULL = identity); fp64 arithmetic throughout.
 *   Outputs: T (Q,16) f64, fitness (Q) f64, rmse (Q) f64, iters (Q) int32,
 *   corr (src items) int32 = matched target item position or -1 (final correspondence set), or NULL.
 *   seg_skip (Q) or NULL: skipped segments get T = T0, fitness = rmse = 0, iters = 0. */
F4L_API int f4l_patch_icp(const float* src, const int32_t* src_idx, const int32_t* s_start,
                  const int32_t* s_count, const float* tgt, const int32_t* tgt_idx,
                  const int32_t* t_start, const int32_t* t_count, const uint8_t* seg_skip,
                  int32_t Q, const double* T0, double max_corr_dist, int32_t max_iter,
                  double rel_fitness, double rel_rmse, double* T, double* fitness, double* rmse,
                  int32_t* iters, int32_t* corr, void* stream);

/* The same with the tie[Q] flag of the boundary schema: fragile (Q) u8 or NULL = OR of
 *   F4L_ICP_FRAGILE_NN     a matched source point had a DISTINCT target within tie_eps (relative, squared distance) of
 *                          its nearest one in some iteration (Open3D's KD-tree may return either);
 *   F4L_ICP_FRAGILE_INLIER a nearest distance within tie_eps of max_corr_dist^2 (strict `<`, o3d hybrid search);
 *   F4L_ICP_FRAGILE_STOP   |delta fitness| or |delta rmse| within tie_eps of the convergence thresholds.
 * An implementation with another fp64 operation order can leave the path of this one only at such a decision: parity
 * tests assert "same iteration count and transform except for flagged pairs".  tie_eps <= 0: 1e-9. */
#define F4L_ICP_FRAGILE_NN 1
#define F4L_ICP_FRAGILE_INLIER 2
#define F4L_ICP_FRAGILE_STOP 4
F4L_API int f4l_patch_icp_ex(const float* src, const int32_t* src_idx, const int32_t* s_start,
                     const int32_t* s_count, const float* tgt, const int32_t* tgt_idx,
                     const int32_t* t_start, const int32_t* t_count, const uint8_t* seg_skip,
                     int32_t Q, const double* T0, double max_corr_dist, int32_t max_iter,
                     double rel_fitness, double rel_rmse, double* T, double* fitness, double* rmse,
                     int32_t* iters, int32_t* corr, uint8_t* fragile, double tie_eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * The fused fine-matching stage for one tile (all patch pairs in one launch sequence).
 * Replaces base.py:3236-3457 fine_matching_with_different_types:
 *   F2 select correspondences of each pair, F3 rigidity check, D2 Procrustes, E1 ICP on the
 *   matched points, D5 apply to all src patch points (+ inverse for tgt2src), A4 assign_then_nn.
 * See f4l_fine_params and the buffer list below. */
#define F4L_MAX_PEERS 7
typedef struct f4l_fine_params {
    int32_t mode;                 /* 0 only_3d, 1 only_2d, 2 fusion (3D rows then 2D rows) */
    int32_t remove_low_quality;   /* method.remove_low_quality_patch_matches */
    int32_t num_min_quality;      /* method.num_min_matches_for_quality_check */
    float thres_dist_diff;        /* method.thres_dist_diff */
    float thres_inlier_ratio;     /* method.thres_inlier_ratio */
    int32_t num_min_fine_match;   /* method.num_min_fine_match */
    int32_t icp_refine;           /* method.icp_refine */
    int32_t assign_type;          /* 0 assign_all_src, 1 assign_then_nn, 2 assign_then_nn with every kept row
                                     emitted ONCE (the host layer repeats them, f4l_host_expand_sparse) */
    int32_t output_tgt2src;       /* method.output_tgt2src */
    double icp_threshold;         /* parameter_setting.icp_threshold */
    double median_max_resolution; /* para.median_max_resolution (base.py:2751) */
    int32_t icp_max_iter;         /* 30 */
} f4l_fine_params;

typedef struct f4l_fine_buffers {
    /* inputs */
    const float* src_pts;  int32_t n_src;      /* (n_src,3) */
    const float* tgt_pts;  int32_t n_tgt;      /* (n_tgt,3) */
    const int64_t* corr3d;                      /* (n_src,2) or NULL; col1 = tgt index or -1 */
    const int64_t* corr2d;                      /* (n_src,2) or NULL */
    const int32_t* sp_idx;  const int32_t* sp_ptr;   /* src patch point lists, CSR over pairs (Q+1) */
    const int32_t* tp_idx;  const int32_t* tp_ptr;   /* tgt patch point lists, CSR over pairs (Q+1) */
    const int32_t* tgt_patch_of_point;          /* (n_tgt) id of the tgt patch owning each point, -1 none */
    const int32_t* pair_tgt_patch;              /* (Q) tgt patch id of each pair */
    int32_t Q;
    int32_t n_src_items;                        /* sp_ptr[Q] (host copy, sizes the workspace) */
    int32_t n_tgt_items;                        /* tp_ptr[Q] */
    const float* d_median_resolution;           /* device scalar overriding params.median_max_resolution
                                                   (output of f4l_median_resolution), or NULL */
    /* outputs (caller-allocated upper bounds) */
    float* T;              /* (Q,16) f32 row-major, identity when not fitted */
    double* T64;           /* (Q,16) */
    int8_t* status;        /* (Q) 0 fitted, 1 rejected by quality check, 2 too few matches */
    int32_t* K;            /* (Q) matched correspondences per pair */
    double* fitness; double* rmse; int32_t* iters;   /* (Q) */
    float* ratio_inlier; float* dist_mean;           /* (Q) */
    float* dense;          /* (sum n_s, 6) rows [p | T p] of fitted pairs, pair order */
    float* sparse;         /* (2*sum n_s, 6) */
    float* tgt2src;        /* (sum n_t, 6) or NULL */
    int32_t* counts;       /* (4): dense rows, sparse rows, tgt2src rows, fitted pairs */
    /* multi-GPU exchange fused into the D5 kernel: every dense row is also stored, with the
       same row offset, into the buffers of n_peers other GPUs (peer-mapped device pointers, see
       f4l_peer_*): the all-gather of the displacement field happens inside the producing kernel over
       NVLink, tile by tile.  n_peers == 0: no exchange. */
    int32_t n_peers;
    float* peer_dense[F4L_MAX_PEERS];
    int32_t* sparse_pair_rows;   /* (Q) or NULL: sparse rows emitted per pair (what f4l_host_expand_sparse needs) */
    void* median_ready_event;    /* cudaEvent_t or NULL: d_median_resolution is produced on ANOTHER stream (A1 is
                                    independent of correspondence selection and the rigid fits); the library makes
                                    `stream` wait for this event right before the first kernel that reads it */
    int32_t phases;              /* 0 = the whole stage; else an OR of F4L_FINE_*: a caller that fits the small pairs of
                                    many tiles in one launch (f4l_fine_fit_tiles) runs SELECT, then that, then
                                    FIT_LARGE | FINISH on the same buffers and workspace */
    uint8_t* icp_fragile;        /* (Q) or NULL: OR of F4L_ICP_FRAGILE_* per fitted pair (0 for the others), tie_eps 1e-9 --
                                    the tie[Q] flag of f4l_patch_icp_ex for the fused stage */
    const int32_t* corr3d_tgt;   /* (n_src) or NULL: column 1 of corr3d as int32 (target index of source point p, -1 = none),
                                    read when corr3d is NULL.  The stage only ever reads that column; a host caller ships 4
                                    instead of 16 bytes per source point over PCIe (f4l_host_pack_corr_targets) */
    const int32_t* corr2d_tgt;   /* the same for corr2d */
} f4l_fine_buffers;

#define F4L_FINE_SELECT 1      /* F2: correspondence selection                       (k_select_corr) */
#define F4L_FINE_FIT_SMALL 2   /* F3 D2 E1 of pairs with <= 224 matches, warp per pair (k_patch_fit_warp) */
#define F4L_FINE_FIT_LARGE 4   /* F3 D2 E1 of the larger pairs, CTA per pair           (k_patch_fit) */
#define F4L_FINE_FINISH 8      /* D5 A4: row offsets, apply + assign, sparse rows */
#define F4L_FINE_ALL 15

F4L_API size_t f4l_fine_matching_workspace_bytes(int32_t n_src_items, int32_t n_tgt_items, int32_t Q,
                                         int32_t mode);
F4L_API int f4l_fine_matching(const f4l_fine_params* h_params, const f4l_fine_buffers* h_buffers,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The per-patch loop of base.py:3254-3368 (rigidity check, Procrustes, ICP) for the small pairs of n_tiles (<= 128)
 * tiles in ONE persistent launch: every resident warp draws the next patch pair from a device-side queue, so there is
 * no per-tile wave quantisation and only one tail.  h_buffers: n_tiles structs (host array) whose F4L_FINE_SELECT
 * phase has been enqueued before this call in stream order; workspaces[i]: the workspace tile i uses in all its
 * phases (one per tile -- they are live at the same time); queue: one int32 of device memory (zeroed here).
 * ctas_per_sm: 0 = as many as fit (4); 1-3 leave registers / shared memory for the other phases' kernels of other
 * tiles to run concurrently on other streams. */
F4L_API int f4l_fine_fit_tiles(const f4l_fine_params* h_params, const f4l_fine_buffers* h_buffers,
                       void* const* workspaces, int32_t n_tiles, int32_t ctas_per_sm, int32_t* queue, void* stream);

/* HOST function (host pointers, no CUDA): the reference appends the sparse rows of a pair twice
 * (base.py:3430,3436).  A caller that moves results over PCIe runs the path with
 * assign_type 2 -- every kept row once, h_pair_rows[q] rows for pair q, pairs back to back -- and restores the
 * reference layout on the host: out = [rows of pair 0][rows of pair 0][rows of pair 1][rows of pair 1]...
 * h_out holds 2 * sum(h_pair_rows) rows of 6 floats.  Uses up to n_threads host threads.  Returns the rows written. */
/* HOST function: h_out[i] = column 1 of the (n,2) int64 correspondence table h_corr as int32; entries outside
 * [0, 2^31) become -1 (the stage treats every negative target as "no correspondence").  Up to n_threads host threads. */
F4L_API void f4l_host_pack_corr_targets(const int64_t* h_corr, int64_t n, int32_t* h_out, int32_t n_threads);

F4L_API long long f4l_host_expand_sparse(const float* h_once, const int32_t* h_pair_rows, int32_t Q, float* h_out,
                                 int32_t n_threads);

/* ------------------------------------------------------------------------------------------
 * (b) exact descriptor-space nearest neighbour, D in {32, 64}, on tensor cores.      kernel K-b
 * Replaces torch.cdist + min at base.py:2783-2815 (global_matches_from_3d exact branches),
 * the hnswlib query at src/f2s3.py:273-281 (exact instead of approximate) and, with
 * both_dirs + the xyz gate, the coarse mutual matching at base.py:2966-2995.
 *   a (N,D), b (M,D) f32, finite.  a_xyz/b_xyz (.,3) + max_mag > 0: pairs with
 *   ||xyz_a - xyz_b|| > max_mag are excluded (base.py:2969).  row_idx (N) int32 argmin over b
 *   (-1 = all excluded), row_d2 (N) f32 squared L2; col_idx (M)/col_d2 (M) likewise for b over a
 *   when both_dirs (else NULL).
 *   algo: F4L_DESC_AUTO picks the tensor-core path (fp16 tcgen05.mma candidates within a proven
 *   error margin, re-ranked in fp64) for >= 2^27 pairs without a gate, else the fp64 brute-force
 *   kernel; F4L_DESC_TENSOR / F4L_DESC_EXACT force one.  Either way the result is the fp64 argmin
 *   of the f32 inputs, exact ties resolved towards the lower index (torch.min semantics). */
#define F4L_DESC_AUTO 0
#define F4L_DESC_TENSOR 1
#define F4L_DESC_EXACT 2
F4L_API size_t f4l_desc_nn_workspace_bytes(int32_t N, int32_t M, int32_t D, int both_dirs);
F4L_API int f4l_desc_nn(const float* a, int32_t N, const float* b, int32_t M, int32_t D,
                const float* a_xyz, const float* b_xyz, float max_mag, int both_dirs, int algo,
                int32_t* row_idx, float* row_d2, int32_t* col_idx, float* col_d2,
                void* workspace, size_t workspace_bytes, void* stream);
/* The same search with tie flags (eps_desc = 1e-6 absolute on the squared distance): row_tie (N) /
 * col_tie (M) uint8 or NULL = 1 when another reference row lies within tie_eps of the reported minimum, i.e. where
 * torch.min / an approximate index may report a different index.  f4l_desc_nn == this with NULL flags. */
F4L_API int f4l_desc_nn_ex(const float* a, int32_t N, const float* b, int32_t M, int32_t D,
                   const float* a_xyz, const float* b_xyz, float max_mag, int both_dirs, int algo,
                   int32_t* row_idx, float* row_d2, int32_t* col_idx, float* col_d2,
                   uint8_t* row_tie, uint8_t* col_tie, double tie_eps,
                   void* workspace, size_t workspace_bytes, void* stream);

/* (b)+(c) scatter of global 3D matches: base.py:2872-2889.  labels (n_sub) int32 from f4l_desc_nn,
 * src_sub/tgt_sub (n_sub.,3) voxel points, voxel2pts_* int64 maps, corres (n_raw,2) int64 out:
 * col0 = arange, col1 = matched raw target index or -1.  Several voxels mapping to one raw point
 * (quirk q5, non-deterministic in the reference): the largest voxel index wins. */
F4L_API size_t f4l_scatter_global_matches_workspace_bytes(int32_t n_raw);
F4L_API int f4l_scatter_global_matches(const int32_t* labels, const float* src_sub, const float* tgt_sub,
                               int32_t n_sub, const int64_t* voxel2pts_src,
                               const int64_t* voxel2pts_tgt, float max_magnitude,
                               int64_t* corres, int32_t n_raw, void* workspace,
                               size_t workspace_bytes, void* stream);

/* (b') B4: 2D-vote coarse matching, base.py:3016-3070.  For each of the P source patches (CSR sp_ptr /
 * sp_idx over source points) the target-patch label that most of its points' 2D-lifted matches
 * (corr2d (n_src,2) int64, col1 = target point or -1) fall into: label_tgt (n_tgt) int32 is
 * idx_pts2spt_tgt.  best (P) = that label, mapped through label_to_local (n_labels) when given
 * (position of the label in idx_spt_tgt, -1 if the patch was removed: base.py:3062-3064), or -1 when the
 * patch has no 2D match; best_count (P); flag (P): 1 = several labels share the top count (the
 * reference's argsort order is unspecified; the smallest label is returned), 255 = not resolved
 * (> 512 distinct labels in one patch). */
F4L_API int f4l_vote_tgt_patch(const int64_t* corr2d, const int32_t* sp_idx, const int32_t* sp_ptr, int32_t P,
                       const int32_t* label_tgt, int32_t n_tgt, const int32_t* label_to_local,
                       int32_t n_labels, int32_t* best, int32_t* best_count, uint8_t* flag, void* stream);

/* (c) F1: displacement-magnitude gates, base.py:2875-2876, :1635-1636, src/f2s3.py:392,419-441.
 * rows (K,stride>=6) f32 [src | tgt | ...]; mag (K) f32 or NULL; mask (K) uint8 = mag <= limit (strict 0)
 * or mag < limit (strict 1); limit = max_mag, or d_max[0]*factor when d_max (device scalar, e.g. the
 * median from f4l_select_kth for the 30 x median gate) is given. */
F4L_API int f4l_magnitude_mask(const float* rows, int32_t K, int32_t stride, float max_mag, const float* d_max,
                       float factor, int strict, float* mag, uint8_t* mask, void* stream);

/* ------------------------------------------------------------------------------------------
 * Label -> CSR tables, the step right before the path.
 * f4l_labels_to_csr replaces prepare_pts2spt_dict, base.py:1301-1351 (collections.Counter + one boolean
 * mask per patch): labels (n) int64 -> patches with count > min_pts in ascending label order, the points
 * of a patch in ascending index order.  Outputs sized for the worst case: patch_label (n) int64,
 * ptr (n+1) int32, idx (n) int32, patch_of_point (n) int32 (-1 = point of a removed patch),
 * counts (2) int32 = {patches P, items ptr[P]}.
 * f4l_gather_pairs_csr builds the concatenated point lists of matched patch pairs (spt_corres_src/tgt,
 * base.py:3156-3157): sel (Q) patch positions -> out_ptr (Q+1), out_idx (out_ptr[Q]); pass out_idx = NULL
 * to only compute out_ptr. */
F4L_API size_t f4l_labels_to_csr_workspace_bytes(int32_t n);
F4L_API int f4l_labels_to_csr(const int64_t* labels, int32_t n, int32_t min_pts, int64_t* patch_label,
                      int32_t* ptr, int32_t* idx, int32_t* patch_of_point, int32_t* counts,
                      void* workspace, size_t workspace_bytes, void* stream);
F4L_API size_t f4l_gather_pairs_csr_workspace_bytes(int32_t Q);
F4L_API int f4l_gather_pairs_csr(const int32_t* ptr, const int32_t* idx, const int32_t* sel, int32_t Q,
                         int32_t* out_ptr, int32_t* out_idx, int32_t out_capacity, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Piecewise "ICP" (rows G1, A5, F5): the whole of src/piecewise_icp.py:89-202 in one launch
 * sequence, fp64 like Open3D / numpy.                                                kernel K-g
 *   union bounding box, its 8 corners appended to both clouds (:96-105), depth =
 *   ceil(log2(max_extent/smax)) (:108-109), Open3D octree leaf of every point (points on the max
 *   faces are dropped), traversal with the early stop on internal nodes holding fewer than
 *   internal_min_points points (hard-coded 250 at :52) and leaves with >= number_points_min (:55),
 *   per-cell centroid (:58-61), 1-NN of every source centroid among the target centroids (:134-149),
 *   thr = mean + std of the centroid distances, stable = d <= thr (:152-161), rows [p | p] for the
 *   stable cells in lexicographic centroid order followed by rows [p | p + (c_t - c_s)] for the
 *   unstable cells in traversal order (:166-202).
 *   src64 (n_src,3), tgt64 (n_tgt,3) f64.  dvfs ((n_src+8),6) f64 and mag (n_src+8) f64 (or NULL) are
 *   upper bounds; counts (6) int32 = {rows, stable rows, source cells, target cells, depth,
 *   unstable cells}; thr_out (1) f64 or NULL; cent_src ((n_src+8),3) / cent_tgt ((n_tgt+8),3) /
 *   nn_out (n_src+8) optional dumps of the cell tables (first counts[2] / counts[3] entries valid).
 *   The reference raises when no cell is unstable (np.vstack of an empty list, :197); here
 *   counts[5] == 0 reports that case and the host wrapper raises. */
F4L_API size_t f4l_piecewise_icp_workspace_bytes(int32_t n_src, int32_t n_tgt);
F4L_API int f4l_piecewise_icp(const double* src64, int32_t n_src, const double* tgt64, int32_t n_tgt,
                      double smax, int32_t number_points_min, int32_t internal_min_points,
                      double* dvfs, double* mag, int32_t* counts, double* thr_out, double* cent_src,
                      double* cent_tgt, int32_t* nn_out, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Peer-visible device buffers for the fused displacement-field exchange (the reference
 * has no multi-GPU path: main_fusion.py:134-148 walks the tiles serially).  One process per GPU:
 *   f4l_peer_alloc   cudaMalloc on the current device + export a 64-byte CUDA IPC handle
 *   f4l_peer_open    map another process's buffer (handle received out of band, e.g. all_gather)
 *                    into this process; the pointer is usable by kernels on the current device
 *   f4l_peer_close   unmap a pointer obtained from f4l_peer_open
 *   f4l_peer_free    free a buffer obtained from f4l_peer_alloc (after every peer closed it)
 *   f4l_peer_enable_access   same-process multi-device use: let the current device store into
 *                    memory of `peer_device` (cudaDeviceEnablePeerAccess; already-enabled is ok)
 * All are host-synchronous set-up calls, never on the per-step path. */
#define F4L_PEER_HANDLE_BYTES 64
F4L_API int f4l_peer_alloc(size_t bytes, void** d_ptr, unsigned char* h_handle /* 64 bytes */);
F4L_API int f4l_peer_open(const unsigned char* h_handle, void** d_ptr);
F4L_API int f4l_peer_close(void* d_ptr);
F4L_API int f4l_peer_free(void* d_ptr);
F4L_API int f4l_peer_enable_access(int32_t peer_device);

/* The displacement-field all-gather as a copy kernel next to the compute (per-step path): rows [0, d_rows[0]) of
 * `src` (device scalar row count, e.g. f4l_fine_buffers.counts; at most max_rows rows of row_bytes bytes) into the same
 * slot of every peer's field (peer_dst[i]: mapped pointers from f4l_peer_open; when one of them does not share src's
 * alignment mod 16 a plain-store kernel is used instead of the bulk copies).
 * A few CTAs (n_ctas, 0 = 32) stream local HBM -> shared memory -> all peers with TMA bulk copies; enqueue it on a side
 * stream after the tile's f4l_fine_matching so the transfer overlaps the next tile's fits. */
F4L_API int f4l_peer_push(const void* src, const int32_t* d_rows, int32_t row_bytes, int64_t max_rows,
                  void* const* peer_dst, int32_t n_peers, int32_t n_ctas, void* stream);

/* ------------------------------------------------------------------------------------------
 * 8(f) rank 1 -- DIPs patch front-end.  Replaces src/data_loader.py:16-109
 * (Preprocess_Dataset.__init__ / extract_patch / __getitem__), called from src/f2s3.py:104-134 and
 * base.py:1981-2034.
 *
 * f4l_dips_build bins the reference cloud ("data_overlap", (n_ref,3) f64) for ball queries of the given
 * radius into the workspace (the counterpart of o3d.geometry.KDTreeFlann(data_overlap), data_loader.py:26).
 * f4l_dips_patches then produces, for every query point (the rows of "data", (n_query,3) f64), the
 * (3, num_points) f32 patch of data_loader.py:37-105: neighbours with d^2 < radius^2 (fp64, nanoflann's
 * summation order), local reference frame of :46-80 when more than 10 neighbours, coordinates divided by the
 * radius, zero rows when fewer than num_points neighbours.
 *   ranks == NULL: slot t holds a uniform random sample without replacement (keyed by seed, query index) --
 *                  the reference draws np.random.choice from the global generator inside DataLoader workers;
 *   ranks != NULL: (n_query, num_points) i32, slot t holds the neighbour whose distance rank is ranks[q][t]
 *                  (ties by original index; a rank >= the neighbour count selects a zero row), i.e. the
 *                  reference's `ptall[inds]`.
 * lrf (n_query,9) f64 or NULL: rows xp, yp, zp of the frame (zeros when no frame was estimated);
 * count (n_query) i32: neighbours found.  f4l_dips_patches keeps up to 1408 neighbours of a query on chip (the
 * reference's radius rule yields ~940); a query with more gets a zero patch and count reports the size -- the caller
 * re-runs those queries through f4l_dips_patches_large (up to 8192 neighbours, one CTA per SM, no ranked mode). */
F4L_API size_t f4l_dips_workspace_bytes(int32_t n_ref);
F4L_API int f4l_dips_build(const double* ref64, int32_t n_ref, double radius, void* workspace,
                   size_t workspace_bytes, void* stream);
F4L_API int f4l_dips_patches(const double* query64, int32_t n_query, int32_t n_ref, double radius,
                   int32_t num_points, const int32_t* ranks, uint64_t seed, float* patches, double* lrf,
                   int32_t* count, void* workspace, size_t workspace_bytes, void* stream);
F4L_API int f4l_dips_patches_large(const double* query64, int32_t n_query, int32_t n_ref, double radius,
                   int32_t num_points, uint64_t seed, float* patches, double* lrf, int32_t* count,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * 8(f) rank 3 -- voxel subsampling.  Replaces Open3D PointCloud.voxel_down_sample as called at
 * base.py:1024-1025 (_voxel_subsampling) and :906-907: voxel_min_bound = min_bound - voxel_size/2, voxel index
 * floor((p - voxel_min_bound)/voxel_size), output = fp64 mean of the points of every occupied voxel (summed in point
 * order).  pts64 (n,3) f64 -> centroids (>= number of voxels, 3) f64 in ascending (ix,iy,iz) voxel order (Open3D:
 * unordered_map order, i.e. the same rows permuted), voxel_of_point (n) i32 or NULL = output row of every input
 * point, counts[0] (device) = number of voxels, or -1 when the cloud spans more than 2^21 voxels along an axis. */
F4L_API size_t f4l_voxel_downsample_workspace_bytes(int32_t n);
F4L_API int f4l_voxel_downsample(const double* pts64, int32_t n, double voxel_size, double* centroids,
                         int32_t* voxel_of_point, int32_t* counts, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * 8(f) rank 2 -- the per-segment parts of the two small networks inside the reference's per-patch loops,
 * for all segments (CSR seg_ptr (Q+1) int32 over rows) of a tile at once.  The per-point dense products in
 * between (1x1 convolutions, linear layers) are ordinary GEMMs and stay with the caller's BLAS.
 *   f4l_segment_scale_maxabs    src/f2s3.py:343: rows (K,C) f32 or f64 (x_is_f64) of every supervoxel divided by its
 *                               max |value| in the input precision, stored as f32 (:346)
 *   f4l_segment_norm2_relu      src/models/outlier_classifier.py:15-23,29-33 (PointCN): InstanceNorm2d(eps) ->
 *                               BatchNorm2d(eps, batch statistics of the single-sample batch) -> ReLU
 *                               [-> + residual] per segment and channel; y, residual, out (K,C) f32
 *   f4l_segment_attention_pool  src/feature_aggregation/cluster_feature_net_self_attention.py:18-33,:91:
 *                               out[p] = mean_i sum_j softmax_j(Q_i.K_j * scale) V_j over the rows of segment p;
 *                               Q,K,V (rows,hidden) f32, hidden in {32,64}; out (P,hidden).  Empty segment -> NaN
 *                               (torch.mean of an empty tensor)
 *   f4l_segment_mean            :97 per-segment mean of x (rows,C) f32 (fp64 accumulation) -> (P,C) */
F4L_API int f4l_segment_scale_maxabs(const void* x, int32_t x_is_f64, const int32_t* seg_ptr, int32_t Q, int32_t C,
                             float* out, void* stream);
F4L_API int f4l_segment_norm2_relu(const float* y, const int32_t* seg_ptr, int32_t Q, int32_t C, float eps,
                           const float* residual, float* out, void* stream);
/* max_seg_rows: the longest segment (host value).  With it (and <= 2392 rows) the tensor-core kernel runs: mean_i sum_j
 * a_ij V_j = sum_j (mean_i a_ij) V_j, so only the scores Q K^T are formed -- mma.sync TF32 with the 3xTF32 split (fp32-level
 * accuracy), scores of 64 queries at a time in shared memory; <= 0 (unknown) or longer segments: the flash-style
 * CUDA-core kernel (Q K^T and the product with V per query). */
F4L_API int f4l_segment_attention_pool(const float* Qm, const float* Km, const float* Vm, const int32_t* seg_ptr,
                               int32_t P, int32_t hidden, float scale, int32_t max_seg_rows, float* out, void* stream);
F4L_API int f4l_segment_mean(const float* x, const int32_t* seg_ptr, int32_t P, int32_t C, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* F4L_B200_H */
