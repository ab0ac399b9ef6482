// The fused fine-matching stage for one tile: every patch pair of the tile in six launches.
// Replaces src/coarse_to_fine_matching_base.py:3236-3457.
//
//   k_select_corr   F2  warp per pair: ordered compaction of the pair's correspondences
//   k_patch_fit     F3 + D2 + E1  CTA per pair: rigidity check, Procrustes, on-chip ICP loop
//   k_row_offsets   scan of the per-pair output row counts (dense, tgt2src)
//   k_apply_assign  D5 + A4  CTA per pair: dense DVF rows, inverse rows, 1-NN assignment
//   k_sparse_offsets, k_emit_sparse   ordered compaction of the kept sparse rows (twice, q4)
//
// No host round trip happens between the stages: per-pair decisions (quality reject, too few
// matches) are status bytes consumed by the later kernels, row counts are device scalars.
#include "icp_warp.cuh"
#include "patch_grid.cuh"

#define FINE_MODE_3D 0
#define FINE_MODE_2D 1
#define FINE_MODE_FUSION 2

// ------------------------------------------------------------------------------------------
// corr: the reference's (n_src,2) int64 table, or NULL with corr32 = its column 1 as int32
__device__ __forceinline__ int warp_compact_rows(const int64_t* __restrict__ corr, const int32_t* __restrict__ corr32,
                                                 const int32_t* __restrict__ sp_idx,
                                                 int s0, int ns, const int32_t* __restrict__ tgt_patch_of_point,
                                                 int want_patch, int n_tgt, int lane, int32_t* __restrict__ cs,
                                                 int32_t* __restrict__ ct, int base) {
    // rows of corr[sp_idx[s0..s0+ns)] whose target belongs to the pair's target patch
    // (torch.isin at base.py:3260 == label test because target patches are disjoint)
    int written = base;
    for (int i0 = 0; i0 < ns; i0 += 32) {
        const int i = i0 + lane;
        int p = -1;
        long long t = -1;
        bool ok = false;
        if (i < ns) {
            p = sp_idx[s0 + i];
            t = corr ? corr[2 * (size_t)p + 1] : (long long)corr32[p];
            ok = t >= 0 && t < n_tgt && tgt_patch_of_point[t] == want_patch;
        }
        const unsigned m = __ballot_sync(F4L_FULL, ok);
        if (ok) {
            const int pos = written + __popc(m & ((1u << lane) - 1u));
            cs[pos] = p;
            ct[pos] = (int)t;
        }
        written += __popc(m);
    }
    return written;
}

__global__ void __launch_bounds__(128)
k_select_corr(const int64_t* __restrict__ corr3d, const int64_t* __restrict__ corr2d,
              const int32_t* __restrict__ sp_idx, const int32_t* __restrict__ sp_ptr,
              const int32_t* __restrict__ tgt_patch_of_point, const int32_t* __restrict__ pair_tgt_patch,
              int n_tgt, int Q, int mode, int32_t* __restrict__ cs, int32_t* __restrict__ ct,
              int32_t* __restrict__ kstart, int32_t* __restrict__ K, const int32_t* __restrict__ corr3d_tgt,
              const int32_t* __restrict__ corr2d_tgt) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int s0 = sp_ptr[q], ns = sp_ptr[q + 1] - s0;
    const int slot = (mode == FINE_MODE_FUSION ? 2 : 1) * s0;
    const int want = pair_tgt_patch[q];
    int w = slot;
    if (mode != FINE_MODE_2D) w = warp_compact_rows(corr3d, corr3d_tgt, sp_idx, s0, ns, tgt_patch_of_point, want, n_tgt, lane, cs, ct, w);
    if (mode != FINE_MODE_3D) w = warp_compact_rows(corr2d, corr2d_tgt, sp_idx, s0, ns, tgt_patch_of_point, want, n_tgt, lane, cs, ct, w);
    if (lane == 0) {
        kstart[q] = slot;
        K[q] = w - slot;
    }
}

// ------------------------------------------------------------------------------------------
struct FitShared {
    IcpShared icp;
    double red_sum[ICP_WARPS];
    unsigned long long red_cnt[ICP_WARPS];
    double Tsvd[16];
    int decision;
};

template <bool FRAG>
__global__ void __launch_bounds__(ICP_THREADS)
k_patch_fit(const float* __restrict__ src_pts, const float* __restrict__ tgt_pts,
            const int32_t* __restrict__ cs, const int32_t* __restrict__ ct,
            const int32_t* __restrict__ kstart, const int32_t* __restrict__ K, int Q, f4l_fine_params prm,
            float* __restrict__ T32, double* __restrict__ T64, int8_t* __restrict__ status,
            double* __restrict__ fitness, double* __restrict__ rmse, int32_t* __restrict__ iters,
            float* __restrict__ ratio_inlier, float* __restrict__ dist_mean, uint8_t* __restrict__ fragile) {
    extern __shared__ float dyn[];
    __shared__ FitShared sh;
    const int tid = threadIdx.x;
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        __syncthreads();
        const int k0 = kstart[q], k = K[q];
        if (k <= WICP_CAP) continue;          // small pairs are fitted by k_patch_fit_warp
        int st = 0;
        float ratio = 0.f, dmean = 0.f;
        if (prm.remove_low_quality && k >= prm.num_min_quality) {
            double sum;
            unsigned long long cnt;
            block_rigidity<2048>(src_pts, tgt_pts, cs, ct, k0, k, prm.thres_dist_diff, dyn, sh.red_sum, sh.red_cnt, sum, cnt);
            if (tid == 0) {
                const double ne = 0.5 * (double)k * (double)(k - 1);
                const float dm = (float)(sum / ne);
                const float ra = (float)((double)(2ull * cnt) / (ne * 2.0));
                ratio_inlier[q] = ra;
                dist_mean[q] = dm;
                sh.decision = (ra <= prm.thres_inlier_ratio || dm >= prm.thres_dist_diff) ? 1 : 0;   // base.py:3320
            }
            __syncthreads();
            st = sh.decision;
        } else if (tid == 0) {
            ratio_inlier[q] = ratio;
            dist_mean[q] = dmean;
        }
        if (st == 0 && (k < prm.num_min_fine_match || k < 1)) st = 2;   // (no fit without a match)                                          // base.py:3338
        double* T64q = T64 + (size_t)q * 16;
        if (st != 0) {
            if (tid < 16) {
                const double v = (tid % 5 == 0) ? 1.0 : 0.0;
                T64q[tid] = v;
                T32[(size_t)q * 16 + tid] = (float)v;
            }
            if (tid == 0) { status[q] = (int8_t)st; fitness[q] = 0; rmse[q] = 0; iters[q] = 0; if (fragile) fragile[q] = 0; }
            continue;
        }
        // D2: Procrustes on the matched pairs (weights None, eps 1e-6: weighted_svd.py:134-142)
        if (tid < 32) {
            double R[9], t[3];
            warp_fit_segment(src_pts, tgt_pts, cs, ct, nullptr, k0, k, 1e-6, 0.f, 0, tid, R, t);
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    sh.Tsvd[r * 4 + 0] = R[r * 3 + 0]; sh.Tsvd[r * 4 + 1] = R[r * 3 + 1];
                    sh.Tsvd[r * 4 + 2] = R[r * 3 + 2]; sh.Tsvd[r * 4 + 3] = t[r];
                }
                sh.Tsvd[12] = 0; sh.Tsvd[13] = 0; sh.Tsvd[14] = 0; sh.Tsvd[15] = 1;
            }
        }
        __syncthreads();
        IcpResult r;
        r.fitness = 0; r.rmse = 0; r.iters = 0; r.fragile = 0;
        if (prm.icp_refine) {
            // E1: ICP between the MATCHED points (base.py:3353-3358), init = T_svd
            r = block_icp<FRAG>(src_pts, cs, k0, k, tgt_pts, ct, k0, k, sh.Tsvd, prm.icp_threshold, prm.icp_max_iter,
                          1e-6, 1e-6, T64q, nullptr, dyn, sh.icp);
        } else if (tid < 16) {
            T64q[tid] = sh.Tsvd[tid];
        }
        __syncthreads();
        if (tid < 16) T32[(size_t)q * 16 + tid] = (float)T64q[tid];                                   // base.py:3366
        if (tid == 0) { status[q] = 0; fitness[q] = r.fitness; rmse[q] = r.rmse; iters[q] = r.iters; if (fragile) fragile[q] = (uint8_t)r.fragile; }
    }
}


// ------------------------------------------------------------------------------------------
// Small pairs (K <= WICP_CAP matched points): the whole fit by ONE warp -- stage the matched pairs
// in the warp's shared-memory slice, rigidity check, Procrustes, ICP loop (icp_warp.cuh).
#ifndef FITW_WARPS
#define FITW_WARPS 4
#endif
#ifndef FITW_MIN_BLOCKS
#define FITW_MIN_BLOCKS (16 / FITW_WARPS)
#endif
// The per-pair outputs / inputs of one tile, as the fit kernels see them.
struct FitTile {
    const float* src_pts; const float* tgt_pts;
    const int32_t* cs; const int32_t* ct; const int32_t* kstart; const int32_t* K;
    float* T32; double* T64; int8_t* status; double* fitness; double* rmse; int32_t* iters;
    float* ratio_inlier; float* dist_mean;
    uint8_t* fragile;
    int32_t Q;
};

// One small pair (K <= WICP_CAP) fitted by one warp.
__device__ __forceinline__ void fit_pair_warp(const FitTile& tl, int q, const f4l_fine_params& prm, WarpIcpSmem& sm, int lane) {
    const float* __restrict__ src_pts = tl.src_pts;
    const float* __restrict__ tgt_pts = tl.tgt_pts;
    const int32_t* __restrict__ cs = tl.cs;
    const int32_t* __restrict__ ct = tl.ct;
    const int32_t* __restrict__ kstart = tl.kstart;
    const int32_t* __restrict__ K = tl.K;
    float* __restrict__ T32 = tl.T32;
    double* __restrict__ T64 = tl.T64;
    int8_t* __restrict__ status = tl.status;
    double* __restrict__ fitness = tl.fitness;
    double* __restrict__ rmse = tl.rmse;
    int32_t* __restrict__ iters = tl.iters;
    float* __restrict__ ratio_inlier = tl.ratio_inlier;
    float* __restrict__ dist_mean = tl.dist_mean;
    uint8_t* __restrict__ fragile = tl.fragile;
    const int k0 = kstart[q], k = K[q];
    if (k > WICP_CAP) return;             // fitted by the CTA kernel
    __syncwarp();
    DBG_T0
#ifdef F4L_DEBUG_SCANS
    const long long dbg_start = dbg_t;
#endif
    int st = 0;
    double* T64q = T64 + (size_t)q * 16;
    float ra = 0.f, dm = 0.f;
    const bool staged = prm.remove_low_quality && k >= prm.num_min_quality;
    if (staged) {
        float* arena = reinterpret_cast<float*>(&sm);          // aliases the ICP staging area (filled later)
        warp_rigidity_stage(arena, src_pts, tgt_pts, cs, ct, k0, k, lane);
        __syncwarp();
        DBG_T(17)
        double sum;
        unsigned cnt;
        warp_rigidity(arena, k, prm.thres_dist_diff, lane, sum, cnt);
        const double ne = 0.5 * (double)k * (double)(k - 1);
        dm = (float)(sum / ne);
        ra = (float)((double)(2ull * cnt) / (ne * 2.0));
        if (ra <= prm.thres_inlier_ratio || dm >= prm.thres_dist_diff) st = 1;                   // base.py:3320
        __syncwarp();
    }
    DBG_T(8)
    if (st == 0 && (k < prm.num_min_fine_match || k < 1)) st = 2;   // (no fit without a match)                                            // base.py:3338
    if (lane == 0) { ratio_inlier[q] = ra; dist_mean[q] = dm; }
    if (st != 0) {
        if (lane < 16) {
            const double v = (lane % 5 == 0) ? 1.0 : 0.0;
            T64q[lane] = v;
            T32[(size_t)q * 16 + lane] = (float)v;
        }
        if (lane == 0) { status[q] = (int8_t)st; fitness[q] = 0; rmse[q] = 0; iters[q] = 0; if (fragile) fragile[q] = 0; }
        return;
    }
    // D2: Procrustes (weights None, eps 1e-6)
    double R[9], t[3], Tsvd[16];
    __syncwarp();
    if (lane < 9) sm.Vw[lane] = (lane % 4 == 0) ? 1.0 : 0.0;          // cold start; later fits of this pair start warm
    __syncwarp();
    if (staged) warp_fit_arena(reinterpret_cast<const float*>(&sm), k, 1e-6, 0, lane, R, t, sm.Vw);
    else warp_fit_segment(src_pts, tgt_pts, cs, ct, nullptr, k0, k, 1e-6, 0.f, 0, lane, R, t, sm.Vw);
    __syncwarp();                  // the arena is dead from here: warp_icp restages over it
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Tsvd[r * 4 + 0] = R[r * 3 + 0]; Tsvd[r * 4 + 1] = R[r * 3 + 1];
        Tsvd[r * 4 + 2] = R[r * 3 + 2]; Tsvd[r * 4 + 3] = t[r];
    }
    Tsvd[12] = 0; Tsvd[13] = 0; Tsvd[14] = 0; Tsvd[15] = 1;
    DBG_T(9)
    IcpResult r;
    r.fitness = 0; r.rmse = 0; r.iters = 0; r.fragile = 0;
    if (prm.icp_refine) {
        r = warp_icp(src_pts, cs, k0, k, tgt_pts, ct, k0, k, Tsvd, prm.icp_threshold, prm.icp_max_iter, 1e-6, 1e-6,
                     T64q, nullptr, sm, lane);
    } else {
#pragma unroll
        for (int a = 0; a < 16; ++a)
            if (lane == a) T64q[a] = Tsvd[a];
    }
    __syncwarp();
    if (lane < 16) T32[(size_t)q * 16 + lane] = (float)T64q[lane];                                // base.py:3366
    if (lane == 0) { status[q] = 0; fitness[q] = r.fitness; rmse[q] = r.rmse; iters[q] = r.iters; if (fragile) fragile[q] = (uint8_t)r.fragile; }
#ifdef F4L_DEBUG_SCANS
    if (lane == 0) atomicAdd(&g_dbg[16], (unsigned long long)(clock64() - dbg_start));
#endif
}

__global__ void __launch_bounds__(FITW_WARPS * 32, FITW_MIN_BLOCKS)
k_patch_fit_warp(FitTile tl, f4l_fine_params prm) {
    extern __shared__ __align__(16) unsigned char fitw_raw[];
    WarpIcpSmem* smem = reinterpret_cast<WarpIcpSmem*>(fitw_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = blockIdx.x * FITW_WARPS + wid; q < tl.Q; q += gridDim.x * FITW_WARPS) fit_pair_warp(tl, q, prm, smem[wid], lane);
}

// Persistent form over the pairs of MANY tiles (f4l_fine_fit_tiles): every resident warp draws the next pair from a
// global queue, so the launch has no wave quantisation (784 CTAs on 592 slots = 1.32 waves per tile before) and its
// tail is one pair long instead of one pair per tile.  Pair i of the concatenation belongs to tile t with
// prefix[t] <= i < prefix[t+1].  The table travels as a kernel parameter (CUDA >= 12.1: up to 32 KB), so the launch
// captures into a CUDA graph by value.
#define FIT_MAX_TILES 128
struct FitTable {
    int32_t n;
    int32_t prefix[FIT_MAX_TILES + 1];
    FitTile t[FIT_MAX_TILES];
};

__global__ void __launch_bounds__(FITW_WARPS * 32, FITW_MIN_BLOCKS)
k_patch_fit_warp_tiles(const __grid_constant__ FitTable tab, f4l_fine_params prm, int32_t* __restrict__ queue) {
    extern __shared__ __align__(16) unsigned char fitw_raw[];
    WarpIcpSmem* smem = reinterpret_cast<WarpIcpSmem*>(fitw_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int total = tab.prefix[tab.n];
    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(queue, 1);
        i = __shfl_sync(F4L_FULL, i, 0);
        if (i >= total) break;
        int lo = 0, hi = tab.n;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (tab.prefix[mid] <= i) lo = mid; else hi = mid;
        }
        fit_pair_warp(tab.t[lo], i - tab.prefix[lo], prm, smem[wid], lane);
    }
}

// ------------------------------------------------------------------------------------------
// single-CTA exclusive scans over the pairs (Q is 10^3..10^4)
__global__ void __launch_bounds__(1024)
k_row_offsets(const int8_t* __restrict__ status, const int32_t* __restrict__ sp_ptr,
              const int32_t* __restrict__ tp_ptr, int Q, int emit, int32_t* __restrict__ dense_off,
              int32_t* __restrict__ t2s_off, int32_t* __restrict__ counts) {
    __shared__ int wsum[2][32];
    __shared__ int carry[3];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 3) carry[tid] = 0;
    __syncthreads();
    for (int base = 0; base < Q; base += 1024) {
        const int q = base + tid;
        int a = 0, b = 0, f = 0;
        if (q < Q && status[q] == 0) {
            f = 1;
            if (emit) { a = sp_ptr[q + 1] - sp_ptr[q]; b = tp_ptr[q + 1] - tp_ptr[q]; }
        }
        int ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int va = __shfl_up_sync(F4L_FULL, ia, o), vb = __shfl_up_sync(F4L_FULL, ib, o);
            if (lane >= o) { ia += va; ib += vb; }
        }
        if (lane == 31) { wsum[0][wid] = ia; wsum[1][wid] = ib; }
        const int fc = __syncthreads_count(f);
        if (wid == 0) {
            int sa = wsum[0][lane], sb = wsum[1][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int va = __shfl_up_sync(F4L_FULL, sa, o), vb = __shfl_up_sync(F4L_FULL, sb, o);
                if (lane >= o) { sa += va; sb += vb; }
            }
            wsum[0][lane] = sa; wsum[1][lane] = sb;
        }
        __syncthreads();
        const int pa = (wid ? wsum[0][wid - 1] : 0) + carry[0];
        const int pb = (wid ? wsum[1][wid - 1] : 0) + carry[1];
        if (q < Q) { dense_off[q] = pa + ia - a; t2s_off[q] = pb + ib - b; }
        __syncthreads();
        if (tid == 0) { carry[0] += wsum[0][31]; carry[1] += wsum[1][31]; carry[2] += fc; }
        __syncthreads();
    }
    if (tid == 0) { counts[0] = carry[0]; counts[2] = carry[1]; counts[3] = carry[2]; }
}

__global__ void __launch_bounds__(1024)
k_sparse_offsets(const int32_t* __restrict__ sparse_cnt, int Q, int32_t* __restrict__ sparse_off,
                 int32_t* __restrict__ counts) {
    __shared__ int wsum[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < Q; base += 1024) {
        const int q = base + tid;
        const int a = q < Q ? sparse_cnt[q] : 0;
        int ia = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int va = __shfl_up_sync(F4L_FULL, ia, o);
            if (lane >= o) ia += va;
        }
        if (lane == 31) wsum[wid] = ia;
        __syncthreads();
        if (wid == 0) {
            int sa = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int va = __shfl_up_sync(F4L_FULL, sa, o);
                if (lane >= o) sa += va;
            }
            wsum[lane] = sa;
        }
        __syncthreads();
        if (q < Q) sparse_off[q] = (wid ? wsum[wid - 1] : 0) + carry + ia - a;
        __syncthreads();
        if (tid == 0) carry += wsum[31];
        __syncthreads();
    }
    if (tid == 0) counts[1] = carry;
}

// ------------------------------------------------------------------------------------------
#ifndef AA_THREADS
#define AA_THREADS 128     // a pair has ~250 source points: 2 per thread; small CTAs interleave better across pairs
#endif
#ifndef AA_MIN_BLOCKS
#define AA_MIN_BLOCKS 8      // 64 registers, 8 CTAs of 128 threads per SM (measured +4 % on the step over 4 x 256)
#endif
struct PeerDense { float* p[F4L_MAX_PEERS]; };
#ifndef AA_SMEM_PTS
#define AA_SMEM_PTS 1024    // staged target patch (16 KB); larger patches are scanned from global memory
#endif

// D5 + A4.  CTA per pair.  The target patch is staged in shared memory as pivot-local float4 (|coordinate|
// of a patch is metres, so f32 keeps ~1e-7 m), binned on a small uniform grid (patch_grid.cuh: counting sort
// in shared memory, three passes over the patch's targets); every source point is transformed, written to the
// dense DVF and looks for its two nearest targets in f32, ring by ring of grid cells.  When the two are
// separated by more than the f32 error bound the f32 argmin IS the fp64 argmin and only that one distance
// is re-evaluated in fp64 from the original coordinates (threshold test d^2 < thr^2 of base.py:82 stays
// exact); otherwise the point takes the exact fp64 scan (first minimal index).
__global__ void __launch_bounds__(AA_THREADS, AA_MIN_BLOCKS)
k_apply_assign(const float* __restrict__ src_pts, const float* __restrict__ tgt_pts,
               const int32_t* __restrict__ sp_idx, const int32_t* __restrict__ sp_ptr,
               const int32_t* __restrict__ tp_idx, const int32_t* __restrict__ tp_ptr,
               const int32_t* __restrict__ cs, const int32_t* __restrict__ kstart, const int32_t* __restrict__ K,
               const int8_t* __restrict__ status, const float* __restrict__ T32, const double* __restrict__ rmse,
               const int32_t* __restrict__ dense_off, const int32_t* __restrict__ t2s_off, int Q,
               f4l_fine_params prm, const float* __restrict__ d_median_res, float* __restrict__ dense,
               float* __restrict__ tgt2src, int32_t* __restrict__ nn, int32_t* __restrict__ sparse_cnt,
               int n_peers, PeerDense peers) {
    extern __shared__ float4 sref[];
    // dense rows of one chunk of AA_THREADS source points: written out as contiguous float2 runs to the local
    // arena AND to the same rows of every peer GPU's arena (the displacement-field all-gather, fused)
    // (+4 floats: the staging is shifted by 8 bytes when the peers' destination is 8 mod 16, so that the bulk copy's
    // source and destination are both 16-byte aligned)
    __shared__ __align__(16) float srow_buf[AA_THREADS * 6 + 4];
    __shared__ int s_cnt;
    __shared__ unsigned s_maxabs;
    __shared__ unsigned s_bb[6];
    __shared__ int s_start[260], s_fill[256], s_wsum[AA_THREADS / 32];
    const int tid = threadIdx.x;
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        __syncthreads();
        if (status[q] != 0 || !prm.icp_refine) {
            if (tid == 0) sparse_cnt[q] = 0;
            continue;
        }
        const int s0 = sp_ptr[q], ns = sp_ptr[q + 1] - s0;
        const int t0 = tp_ptr[q], nt = tp_ptr[q + 1] - t0;
        const float* Tq = T32 + (size_t)q * 16;
        double R[9], tv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            R[i * 3] = Tq[i * 4]; R[i * 3 + 1] = Tq[i * 4 + 1]; R[i * 3 + 2] = Tq[i * 4 + 2];
            tv[i] = Tq[i * 4 + 3];
        }
        const bool assign_nn = prm.assign_type >= 1;
        const bool staged = nt <= AA_SMEM_PTS;
        if (tid == 0) { s_cnt = 0; s_maxabs = 0u; }
        if (tid < 6) s_bb[tid] = tid < 3 ? 0xffffffffu : 0u;
        double cB[3] = {0, 0, 0};
        if (nt > 0) {
            float x, y, z;
            load_ptf(tgt_pts, tp_idx, t0, x, y, z);
            cB[0] = x; cB[1] = y; cB[2] = z;
        }
        __syncthreads();
        const bool binned = assign_nn && staged && nt > 0;
        if (binned || prm.output_tgt2src) {
            // pass 1 over the targets: bounding box + largest |coordinate| in the pivot-local frame (and the
            // inverse rows, which need every target once anyway)
            float mabs = 0.f;
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int j = tid; j < nt; j += AA_THREADS) {
                float x, y, z;
                load_ptf(tgt_pts, tp_idx, t0 + j, x, y, z);
                if (binned) {
                    const float lx = (float)((double)x - cB[0]), ly = (float)((double)y - cB[1]), lz = (float)((double)z - cB[2]);
                    mn[0] = fminf(mn[0], lx); mx[0] = fmaxf(mx[0], lx);
                    mn[1] = fminf(mn[1], ly); mx[1] = fmaxf(mx[1], ly);
                    mn[2] = fminf(mn[2], lz); mx[2] = fmaxf(mx[2], lz);
                    mabs = fmaxf(mabs, fmaxf(fabsf(lx), fmaxf(fabsf(ly), fabsf(lz))));
                }
                if (prm.output_tgt2src) {
                    // base.py:3389-3390: R^T (q - t), the subtraction in f32
                    const double dx = (double)(x - Tq[3]), dy = (double)(y - Tq[7]), dz = (double)(z - Tq[11]);
                    float2* row = reinterpret_cast<float2*>(tgt2src + (size_t)(t2s_off[q] + j) * 6);
                    row[0] = make_float2((float)(R[0] * dx + R[3] * dy + R[6] * dz), (float)(R[1] * dx + R[4] * dy + R[7] * dz));
                    row[1] = make_float2((float)(R[2] * dx + R[5] * dy + R[8] * dz), x);
                    row[2] = make_float2(y, z);
                }
            }
            if (binned) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mabs = fmaxf(mabs, __shfl_xor_sync(F4L_FULL, mabs, o));
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        mn[a] = fminf(mn[a], __shfl_xor_sync(F4L_FULL, mn[a], o));
                        mx[a] = fmaxf(mx[a], __shfl_xor_sync(F4L_FULL, mx[a], o));
                    }
                }
                if ((tid & 31) == 0) {
                    atomicMax(&s_maxabs, __float_as_uint(mabs));
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        atomicMin(&s_bb[a], f2ord(mn[a]));
                        atomicMax(&s_bb[3 + a], f2ord(mx[a]));
                    }
                }
            }
        }
        __syncthreads();
        PatchGrid grid;
        if (binned) {
            // pass 2: cell histogram;  scan;  pass 3: counting-sort scatter into sref (x fastest cell order)
            const float bmn[3] = {ord2f(s_bb[0]), ord2f(s_bb[1]), ord2f(s_bb[2])};
            const float bmx[3] = {ord2f(s_bb[3]), ord2f(s_bb[4]), ord2f(s_bb[5])};
            grid = make_patch_grid(bmn, bmx, nt, 16);
            const int ncells = grid.ncu * grid.ncv;
            // the grid has at most 256 cells; a thread owns AA_CPT consecutive ones
            constexpr int AA_CPT = 256 / AA_THREADS;
            static_assert(AA_CPT * AA_THREADS == 256 && AA_CPT >= 1, "AA_THREADS must divide 256");
#pragma unroll
            for (int u = 0; u < AA_CPT; ++u) s_fill[tid * AA_CPT + u] = 0;
            __syncthreads();
            for (int j = tid; j < nt; j += AA_THREADS) {
                float x, y, z;
                load_ptf(tgt_pts, tp_idx, t0 + j, x, y, z);
                atomicAdd(&s_fill[pg_cell(grid, (float)((double)x - cB[0]), (float)((double)y - cB[1]), (float)((double)z - cB[2]))], 1);
            }
            __syncthreads();
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < AA_CPT; ++u) cnt += s_fill[tid * AA_CPT + u];
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(F4L_FULL, inc, o);
                if ((tid & 31) >= o) inc += v;
            }
            if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
            __syncthreads();
            int before = 0;
#pragma unroll
            for (int w = 0; w < AA_THREADS / 32; ++w) before += (w < (tid >> 5)) ? s_wsum[w] : 0;
            int run = before + inc - cnt;
#pragma unroll
            for (int u = 0; u < AA_CPT; ++u) {
                const int c = s_fill[tid * AA_CPT + u];
                s_start[tid * AA_CPT + u] = run;
                s_fill[tid * AA_CPT + u] = run;
                run += c;
            }
            if (tid == 0) s_start[256] = nt;
            __syncthreads();
            for (int j = tid; j < nt; j += AA_THREADS) {
                float x, y, z;
                load_ptf(tgt_pts, tp_idx, t0 + j, x, y, z);
                const float lx = (float)((double)x - cB[0]), ly = (float)((double)y - cB[1]), lz = (float)((double)z - cB[2]);
                const int pos = atomicAdd(&s_fill[pg_cell(grid, lx, ly, lz)], 1);
                sref[pos] = make_float4(lx, ly, lz, __int_as_float(j));
            }
            (void)ncells;
        }
        __syncthreads();
        const float maxabs_t = __uint_as_float(s_maxabs);
        // adaptive threshold, base.py:3420-3423
        double thr = rmse[q] * 2.0;
        const double mres = d_median_res ? (double)d_median_res[0] : prm.median_max_resolution;
        if (isnan(thr) || isinf(thr)) thr = mres;
        thr = fmax(thr, mres);
        const double thr2 = thr * thr;
        int kept = 0;
        const size_t row0 = (size_t)dense_off[q];
        for (int base = 0; base < ns; base += AA_THREADS) {
          const int i = base + tid;
          const unsigned mis = n_peers > 0 ? (unsigned)((uintptr_t)(peers.p[0] + (row0 + base) * 6) & 15u) : 0u;   // 0 or 8
          float* srow = srow_buf + (mis >> 2);
          if (i < ns) {
            double x, y, z;
            load_pt(src_pts, sp_idx, s0 + i, x, y, z);
            const float mx = (float)(R[0] * x + R[1] * y + R[2] * z + tv[0]);
            const float my = (float)(R[3] * x + R[4] * y + R[5] * z + tv[1]);
            const float mz = (float)(R[6] * x + R[7] * y + R[8] * z + tv[2]);
            float2* row = reinterpret_cast<float2*>(srow + tid * 6);
            row[0] = make_float2((float)x, (float)y);
            row[1] = make_float2((float)z, mx);
            row[2] = make_float2(my, mz);
            if (assign_nn) {
                double best = INFINITY;
                int bj = -1;
                bool exact_scan = !staged;
                if (staged && nt > 0) {
                    const float qx = (float)((double)mx - cB[0]), qy = (float)((double)my - cB[1]), qz = (float)((double)mz - cB[2]);
                    // |sqrt(dd) - true distance| <= eta for every candidate (coordinate rounding + f32 arithmetic)
                    const float mabs = fmaxf(maxabs_t, fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
                    auto certain = [mabs](float a1, float a2) {
                        const float eta = 1e-6f * (mabs + sqrtf(a2));
                        const float r1 = sqrtf(a1);
                        return a2 > a1 + 4.f * eta * r1 + 4.f * eta * eta && a2 > a1;
                    };
                    float d1, d2;
                    int j1;
                    if (pg_top2(grid, sref, s_start, qx, qy, qz, certain, d1, j1, d2)) {
                        float gx, gy, gz;
                        load_ptf(tgt_pts, tp_idx, t0 + j1, gx, gy, gz);
                        const double dx = (double)mx - (double)gx, dy = (double)my - (double)gy, dz = (double)mz - (double)gz;
                        best = dx * dx + dy * dy + dz * dz;
                        bj = j1;
                    } else {
                        exact_scan = true;
                    }
                }
                if (exact_scan) {
                    for (int j = 0; j < nt; ++j) {
                        float gx, gy, gz;
                        load_ptf(tgt_pts, tp_idx, t0 + j, gx, gy, gz);
                        const double dx = (double)mx - (double)gx, dy = (double)my - (double)gy, dz = (double)mz - (double)gz;
                        const double d2 = dx * dx + dy * dy + dz * dz;
                        if (d2 < best) { best = d2; bj = j; }
                    }
                }
                const bool ok = best < thr2;
                nn[s0 + i] = ok ? bj : -1;
                kept += ok ? 1 : 0;
            }
          }
          __syncthreads();
          {   // flush the chunk: rows [base, base + rows) are one contiguous run of rows * 3 float2
              const int n2 = min(AA_THREADS, ns - base) * 3;
              const float2* s2 = reinterpret_cast<const float2*>(srow);
              float2* d2 = reinterpret_cast<float2*>(dense + (row0 + base) * 6);
              for (int t = tid; t < n2; t += AA_THREADS) d2[t] = s2[t];
              if (n_peers > 0) {
                  // the same rows into every peer GPU's field over NVLink: one bulk copy (TMA, shared -> peer global)
                  // per peer, issued by one thread -- the CTA does not wait for the link (7 unicast float2 store
                  // loops stalled it before: 0.74 scaling efficiency at 8 GPUs); the 8-byte head / tail that the
                  // 16-byte granularity of the bulk copy leaves over go as plain stores
                  const unsigned bytes = (unsigned)n2 * 8u;
                  const unsigned head = mis ? 8u : 0u;
                  const unsigned body = (bytes - head) & ~15u;
                  const unsigned tail = bytes - head - body;
                  if (tid == 0 && body) {
                      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                      const uint32_t src = (uint32_t)__cvta_generic_to_shared(reinterpret_cast<const char*>(srow) + head);
                      for (int p = 0; p < n_peers; ++p) {
                          char* dst = reinterpret_cast<char*>(peers.p[p] + (row0 + base) * 6) + head;
                          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(body)
                                       : "memory");
                      }
                      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                  }
                  if (tid >= 32 && tid < 32 + n_peers) {
                      float2* r2 = reinterpret_cast<float2*>(peers.p[tid - 32] + (row0 + base) * 6);
                      if (head) r2[0] = s2[0];
                      if (tail) r2[n2 - 1] = s2[n2 - 1];
                  }
                  if (tid == 0 && body) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // srow reusable
              }
          }
          __syncthreads();
        }
        if (assign_nn) {
            kept = warp_sum(kept);
            if ((tid & 31) == 0 && kept) atomicAdd(&s_cnt, kept);
            __syncthreads();
            if (tid == 0) sparse_cnt[q] = (prm.assign_type == 2 ? 1 : 2) * s_cnt;   // appended twice (q4) unless the host repeats them
        } else if (tid == 0) {
            sparse_cnt[q] = K[q];                              // assign_all_src: the matched points
        }
    }
    if (n_peers > 0 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // pushed rows are complete
}

__global__ void __launch_bounds__(128)
k_emit_sparse(const float* __restrict__ src_pts, const float* __restrict__ tgt_pts,
              const int32_t* __restrict__ sp_idx, const int32_t* __restrict__ sp_ptr,
              const int32_t* __restrict__ tp_idx, const int32_t* __restrict__ tp_ptr,
              const int32_t* __restrict__ cs, const int32_t* __restrict__ kstart, const int32_t* __restrict__ K,
              const int8_t* __restrict__ status, const float* __restrict__ T32, const int32_t* __restrict__ nn,
              const int32_t* __restrict__ sparse_cnt, const int32_t* __restrict__ sparse_off, int Q,
              f4l_fine_params prm, float* __restrict__ sparse) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q || status[q] != 0 || !prm.icp_refine) return;
    const int total = sparse_cnt[q];
    if (total == 0) return;
    const int o0 = sparse_off[q];
    if (prm.assign_type >= 1) {
        const int s0 = sp_ptr[q], ns = sp_ptr[q + 1] - s0, t0 = tp_ptr[q];
        const int reps = prm.assign_type == 2 ? 1 : 2;
        const int half = total / reps;
        int written = 0;
        for (int i0 = 0; i0 < ns; i0 += 32) {
            const int i = i0 + lane;
            const int j = i < ns ? nn[s0 + i] : -1;
            const unsigned m = __ballot_sync(F4L_FULL, j >= 0);
            if (j >= 0) {
                const int r = written + __popc(m & ((1u << lane) - 1u));
                float x, y, z, gx, gy, gz;
                load_ptf(src_pts, sp_idx, s0 + i, x, y, z);
                load_ptf(tgt_pts, tp_idx, t0 + j, gx, gy, gz);
#pragma unroll
                for (int rep = 0; rep < reps; ++rep) {
                    float2* row = reinterpret_cast<float2*>(sparse + (size_t)(o0 + rep * half + r) * 6);
                    row[0] = make_float2(x, y);
                    row[1] = make_float2(z, gx);
                    row[2] = make_float2(gy, gz);
                }
            }
            written += __popc(m);
        }
    } else {
        // assign_all_src: [A | T A] over the matched source points (base.py:3381-3384, 3412-3413)
        const float* Tq = T32 + (size_t)q * 16;
        const int k0 = kstart[q], k = K[q];
        for (int i = lane; i < k; i += 32) {
            double x, y, z;
            load_pt(src_pts, cs, k0 + i, x, y, z);
            float2* row = reinterpret_cast<float2*>(sparse + (size_t)(o0 + i) * 6);
            row[0] = make_float2((float)x, (float)y);
            row[1] = make_float2((float)z, (float)((double)Tq[0] * x + (double)Tq[1] * y + (double)Tq[2] * z + (double)Tq[3]));
            row[2] = make_float2((float)((double)Tq[4] * x + (double)Tq[5] * y + (double)Tq[6] * z + (double)Tq[7]),
                                 (float)((double)Tq[8] * x + (double)Tq[9] * y + (double)Tq[10] * z + (double)Tq[11]));
        }
    }
}

// ------------------------------------------------------------------------------------------
static inline size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct FineWs {
    int32_t *cs, *ct, *kstart, *dense_off, *t2s_off, *nn, *sparse_cnt, *sparse_off;
    size_t total;
};

static FineWs fine_layout(void* base, int n_src_items, int Q, int mode) {
    FineWs w;
    char* b = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = b + off; off += al(bytes); return (int32_t*)p; };
    const size_t cap = (size_t)(mode == FINE_MODE_FUSION ? 2 : 1) * n_src_items;
    w.cs = take(cap * 4);
    w.ct = take(cap * 4);
    w.kstart = take((size_t)Q * 4);
    w.dense_off = take((size_t)Q * 4);
    w.t2s_off = take((size_t)Q * 4);
    w.nn = take((size_t)n_src_items * 4);
    w.sparse_cnt = take((size_t)Q * 4);
    w.sparse_off = take((size_t)Q * 4);
    w.total = off;
    return w;
}

extern "C" size_t f4l_fine_matching_workspace_bytes(int32_t n_src_items, int32_t n_tgt_items, int32_t Q, int32_t mode) {
    (void)n_tgt_items;
    return fine_layout(nullptr, n_src_items < 0 ? 0 : n_src_items, Q < 0 ? 0 : Q, mode).total;
}

static bool fine_optin() {
    static F4lPerDevice once;
    if (once.done()) return true;
    if (!f4l_optin_smem(k_patch_fit<false>, (size_t)ICP_SMEM_PTS * 3 * sizeof(float), "k_patch_fit") ||
        !f4l_optin_smem(k_patch_fit<true>, (size_t)ICP_SMEM_PTS * 3 * sizeof(float), "k_patch_fit") ||
        !f4l_optin_smem(k_apply_assign, (size_t)AA_SMEM_PTS * sizeof(float4), "k_apply_assign") ||
        !f4l_optin_smem(k_patch_fit_warp, FITW_WARPS * sizeof(WarpIcpSmem), "k_patch_fit_warp") ||
        !f4l_optin_smem(k_patch_fit_warp_tiles, FITW_WARPS * sizeof(WarpIcpSmem), "k_patch_fit_warp_tiles"))
        return false;
    once.mark();
    return true;
}

static FitTile fit_tile_of(const f4l_fine_buffers* bf, const FineWs& w) {
    FitTile t;
    t.src_pts = bf->src_pts; t.tgt_pts = bf->tgt_pts;
    t.cs = w.cs; t.ct = w.ct; t.kstart = w.kstart; t.K = bf->K;
    t.T32 = bf->T; t.T64 = bf->T64; t.status = bf->status; t.fitness = bf->fitness; t.rmse = bf->rmse; t.iters = bf->iters;
    t.ratio_inlier = bf->ratio_inlier; t.dist_mean = bf->dist_mean;
    t.fragile = bf->icp_fragile;
    t.Q = bf->Q;
    return t;
}

// The small-pair fits of MANY tiles in one persistent launch (k_patch_fit_warp_tiles): between the
// F4L_FINE_SELECT phase of every tile and their F4L_FINE_FIT_LARGE | F4L_FINE_FINISH phases.
extern "C" int f4l_fine_fit_tiles(const f4l_fine_params* prm, const f4l_fine_buffers* bufs, void* const* workspaces,
                                  int32_t n_tiles, int32_t ctas_per_sm, int32_t* queue, void* stream) {
    F4L_REQUIRE(prm && bufs && workspaces && queue, "null argument");
    F4L_REQUIRE(n_tiles >= 0 && n_tiles <= FIT_MAX_TILES, "n_tiles out of range (at most 128 per call)");
    if (n_tiles == 0) return F4L_OK;
    if (!fine_optin()) return F4L_E_CUDA;
    static_assert(sizeof(FitTable) <= 32000, "the table must fit the kernel parameter space");
    FitTable tab;
    tab.n = n_tiles;
    tab.prefix[0] = 0;
    for (int i = 0; i < n_tiles; ++i) {
        const f4l_fine_buffers* bf = bufs + i;
        F4L_REQUIRE(bf->Q >= 0 && bf->n_src_items >= 0, "negative size");
        F4L_REQUIRE(bf->Q == 0 || (workspaces[i] && bf->src_pts && bf->tgt_pts && bf->K && bf->T && bf->T64 && bf->status &&
                                   bf->fitness && bf->rmse && bf->iters && bf->ratio_inlier && bf->dist_mean), "null pointer");
        tab.t[i] = fit_tile_of(bf, fine_layout(workspaces[i], bf->n_src_items, bf->Q, prm->mode));
        tab.prefix[i + 1] = tab.prefix[i] + bf->Q;
    }
    for (int i = n_tiles; i < FIT_MAX_TILES; ++i) tab.prefix[i + 1] = tab.prefix[n_tiles];
    if (tab.prefix[n_tiles] == 0) return F4L_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(queue, 0, sizeof(int32_t), st);
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // one resident wave, every warp loops on the queue; fewer CTAs per SM than fit (4: registers and shared memory
    // are then exhausted) leave room for the kernels of other tiles' phases to run next to it
    const int per_sm = ctas_per_sm > 0 && ctas_per_sm < FITW_MIN_BLOCKS ? ctas_per_sm : FITW_MIN_BLOCKS;
    const int slots = sms * per_sm;
    const int need = f4l_div_up(tab.prefix[n_tiles], FITW_WARPS);
    f4l_mark("k_patch_fit_warp_tiles", st);
    k_patch_fit_warp_tiles<<<need < slots ? need : slots, FITW_WARPS * 32, FITW_WARPS * sizeof(WarpIcpSmem), st>>>(tab, *prm, queue);
    return f4l_finish("f4l_fine_fit_tiles", stream);
}

extern "C" int f4l_fine_matching(const f4l_fine_params* prm, const f4l_fine_buffers* bf, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(prm && bf, "null params");
    const int Q = bf->Q;
    const int n_src_items = bf->n_src_items;
    F4L_REQUIRE(Q >= 0 && n_src_items >= 0, "negative size");
    F4L_REQUIRE(prm->mode >= 0 && prm->mode <= 2, "unknown mode");
    F4L_REQUIRE(bf->counts, "counts is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (Q == 0) {
        cudaMemsetAsync(bf->counts, 0, 4 * sizeof(int32_t), st);
        return F4L_OK;
    }
    F4L_REQUIRE(bf->src_pts && bf->tgt_pts && bf->sp_idx && bf->sp_ptr && bf->tp_idx && bf->tp_ptr &&
                    bf->tgt_patch_of_point && bf->pair_tgt_patch, "null input");
    F4L_REQUIRE(prm->mode == FINE_MODE_2D || bf->corr3d || bf->corr3d_tgt, "corr3d and corr3d_tgt are null");
    F4L_REQUIRE(prm->mode == FINE_MODE_3D || bf->corr2d || bf->corr2d_tgt, "corr2d and corr2d_tgt are null");
    F4L_REQUIRE(bf->T && bf->T64 && bf->status && bf->K && bf->fitness && bf->rmse && bf->iters &&
                    bf->ratio_inlier && bf->dist_mean && bf->dense && bf->sparse, "null output");
    F4L_REQUIRE(!prm->output_tgt2src || bf->tgt2src, "tgt2src is null");
    F4L_REQUIRE(!prm->icp_refine || prm->icp_threshold > 0.0, "icp_threshold must be > 0");
    F4L_REQUIRE(bf->n_peers >= 0 && bf->n_peers <= F4L_MAX_PEERS, "n_peers out of range");
    PeerDense peers;
    for (int p = 0; p < F4L_MAX_PEERS; ++p) {
        peers.p[p] = p < bf->n_peers ? bf->peer_dense[p] : nullptr;
        F4L_REQUIRE(p >= bf->n_peers || peers.p[p], "peer_dense pointer is null");
    }
    FineWs w = fine_layout(workspace, n_src_items, Q, prm->mode);
    if (!workspace || workspace_bytes < w.total) {
        f4l_set_error("f4l_fine_matching: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    if (!fine_optin()) return F4L_E_CUDA;
    const size_t smem_fit = (size_t)ICP_SMEM_PTS * 3 * sizeof(float);
    const size_t smem_aa = (size_t)AA_SMEM_PTS * sizeof(float4);
    const int phases = bf->phases ? bf->phases : F4L_FINE_ALL;
    const FitTile tl = fit_tile_of(bf, w);
    if (phases & F4L_FINE_SELECT) {
        f4l_mark("k_select_corr", st);
        k_select_corr<<<f4l_div_up(Q, 4), 128, 0, st>>>(bf->corr3d, bf->corr2d, bf->sp_idx, bf->sp_ptr,
                                                       bf->tgt_patch_of_point, bf->pair_tgt_patch, bf->n_tgt, Q,
                                                       prm->mode, w.cs, w.ct, w.kstart, bf->K, bf->corr3d_tgt, bf->corr2d_tgt);
    }
    if (phases & F4L_FINE_FIT_SMALL) {
        const int grid_w = f4l_div_up(Q, FITW_WARPS) < 148 * 32 / FITW_WARPS ? f4l_div_up(Q, FITW_WARPS) : 148 * 32 / FITW_WARPS;
        f4l_mark("k_patch_fit_warp", st);
        k_patch_fit_warp<<<grid_w, FITW_WARPS * 32, FITW_WARPS * sizeof(WarpIcpSmem), st>>>(tl, *prm);
    }
    if (phases & F4L_FINE_FIT_LARGE) {
        const int grid_fit = Q < 148 * 16 ? Q : 148 * 16;
        f4l_mark("k_patch_fit", st);
        if (bf->icp_fragile)
            k_patch_fit<true><<<grid_fit, ICP_THREADS, smem_fit, st>>>(bf->src_pts, bf->tgt_pts, w.cs, w.ct, w.kstart, bf->K, Q,
                                                                      *prm, bf->T, bf->T64, bf->status, bf->fitness, bf->rmse,
                                                                      bf->iters, bf->ratio_inlier, bf->dist_mean, bf->icp_fragile);
        else
            k_patch_fit<false><<<grid_fit, ICP_THREADS, smem_fit, st>>>(bf->src_pts, bf->tgt_pts, w.cs, w.ct, w.kstart, bf->K, Q,
                                                                       *prm, bf->T, bf->T64, bf->status, bf->fitness, bf->rmse,
                                                                       bf->iters, bf->ratio_inlier, bf->dist_mean, nullptr);
    }
    if (phases & F4L_FINE_FINISH) {
        f4l_mark("k_row_offsets", st);
        k_row_offsets<<<1, 1024, 0, st>>>(bf->status, bf->sp_ptr, bf->tp_ptr, Q, prm->icp_refine ? 1 : 0, w.dense_off,
                                          w.t2s_off, bf->counts);
        const int grid_aa = Q < 148 * 8 ? Q : 148 * 8;
        if (bf->median_ready_event) cudaStreamWaitEvent(st, (cudaEvent_t)bf->median_ready_event, 0);
        f4l_mark("k_apply_assign", st);
        k_apply_assign<<<grid_aa, AA_THREADS, smem_aa, st>>>(bf->src_pts, bf->tgt_pts, bf->sp_idx, bf->sp_ptr, bf->tp_idx,
                                                            bf->tp_ptr, w.cs, w.kstart, bf->K, bf->status, bf->T,
                                                            bf->rmse, w.dense_off, w.t2s_off, Q, *prm,
                                                            bf->d_median_resolution, bf->dense, bf->tgt2src, w.nn,
                                                            w.sparse_cnt, bf->n_peers, peers);
        if (bf->sparse_pair_rows)
            cudaMemcpyAsync(bf->sparse_pair_rows, w.sparse_cnt, (size_t)Q * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
        f4l_mark("k_sparse_offsets", st);
        k_sparse_offsets<<<1, 1024, 0, st>>>(w.sparse_cnt, Q, w.sparse_off, bf->counts);
        f4l_mark("k_emit_sparse", st);
        k_emit_sparse<<<f4l_div_up(Q, 4), 128, 0, st>>>(bf->src_pts, bf->tgt_pts, bf->sp_idx, bf->sp_ptr, bf->tp_idx,
                                                       bf->tp_ptr, w.cs, w.kstart, bf->K, bf->status, bf->T, w.nn,
                                                       w.sparse_cnt, w.sparse_off, Q, *prm, bf->sparse);
    }
    return f4l_finish("f4l_fine_matching", stream);
}

#ifdef F4L_DEBUG_SCANS
extern "C" __attribute__((visibility("default"))) void f4l_debug_counters(unsigned long long* h_out, int reset) {
    cudaMemcpyFromSymbol(h_out, g_dbg, sizeof(unsigned long long) * 24);
    if (reset) { unsigned long long z[24] = {0}; cudaMemcpyToSymbol(g_dbg, z, sizeof(z)); }
}
#endif
