// K-e per-patch ICP (standalone entry point) and the segmented 1-NN used by assign_then_nn.
#include "icp_warp.cuh"

template <bool FRAG>
__global__ void __launch_bounds__(ICP_THREADS)
k_patch_icp(const float* __restrict__ src, const int32_t* __restrict__ src_idx,
            const int32_t* __restrict__ s_start, const int32_t* __restrict__ s_count,
            const float* __restrict__ tgt, const int32_t* __restrict__ tgt_idx,
            const int32_t* __restrict__ t_start, const int32_t* __restrict__ t_count,
            const uint8_t* __restrict__ seg_skip, int Q, const double* __restrict__ T0, double max_dist,
            int max_iter, double rel_fit, double rel_rmse, double* __restrict__ T,
            double* __restrict__ fitness, double* __restrict__ rmse, int32_t* __restrict__ iters,
            int32_t* __restrict__ corr, uint8_t* __restrict__ fragile, double tie_eps) {
    extern __shared__ float pts[];
    __shared__ IcpShared sh;
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        int s0, ns, t0, nt;
        seg_bounds(s_start, s_count, q, s0, ns);
        seg_bounds(t_start, t_count, q, t0, nt);
        const double* T0q = T0 ? T0 + (size_t)q * 16 : nullptr;
        const bool skip = seg_skip && seg_skip[q];
        if (!skip && ns >= 1 && nt >= 1 && ns <= WICP_CAP && nt <= WICP_CAP) continue;   // k_patch_icp_warp's
        if (skip) {
            if (threadIdx.x < 16) {
                double v = (threadIdx.x % 5 == 0) ? 1.0 : 0.0;
                if (T0q) v = T0q[threadIdx.x];
                T[(size_t)q * 16 + threadIdx.x] = v;
            }
            if (threadIdx.x == 0) { fitness[q] = 0; rmse[q] = 0; iters[q] = 0; if (fragile) fragile[q] = 0; }
            for (int i = threadIdx.x; corr && i < ns; i += ICP_THREADS) corr[s0 + i] = -1;
            continue;
        }
        IcpResult r = block_icp<FRAG>(src, src_idx, s0, ns, tgt, tgt_idx, t0, nt, T0q, max_dist, max_iter, rel_fit,
                                rel_rmse, T + (size_t)q * 16, corr, pts, sh, tie_eps);
        if (threadIdx.x == 0) { fitness[q] = r.fitness; rmse[q] = r.rmse; iters[q] = r.iters; if (fragile) fragile[q] = (uint8_t)r.fragile; }
        __syncthreads();
    }
}

// small segments: one warp per segment (icp_warp.cuh)
#define ICPW_WARPS 4
__global__ void __launch_bounds__(ICPW_WARPS * 32)
k_patch_icp_warp(const float* __restrict__ src, const int32_t* __restrict__ src_idx,
                 const int32_t* __restrict__ s_start, const int32_t* __restrict__ s_count,
                 const float* __restrict__ tgt, const int32_t* __restrict__ tgt_idx,
                 const int32_t* __restrict__ t_start, const int32_t* __restrict__ t_count,
                 const uint8_t* __restrict__ seg_skip, int Q, const double* __restrict__ T0, double max_dist,
                 int max_iter, double rel_fit, double rel_rmse, double* __restrict__ T,
                 double* __restrict__ fitness, double* __restrict__ rmse, int32_t* __restrict__ iters,
                 int32_t* __restrict__ corr, uint8_t* __restrict__ fragile, double tie_eps) {
    extern __shared__ __align__(16) unsigned char icpw_raw[];
    WarpIcpSmem* smem = reinterpret_cast<WarpIcpSmem*>(icpw_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int q = blockIdx.x * ICPW_WARPS + wid; q < Q; q += gridDim.x * ICPW_WARPS) {
        int s0, ns, t0, nt;
        seg_bounds(s_start, s_count, q, s0, ns);
        seg_bounds(t_start, t_count, q, t0, nt);
        const bool skip = seg_skip && seg_skip[q];
        if (skip || ns < 1 || nt < 1 || ns > WICP_CAP || nt > WICP_CAP) continue;
        __syncwarp();
        __syncwarp();
        if (lane < 9) smem[wid].Vw[lane] = (lane % 4 == 0) ? 1.0 : 0.0;     // svd warm-start slot: cold for a new pair
        __syncwarp();
        IcpResult r = warp_icp(src, src_idx, s0, ns, tgt, tgt_idx, t0, nt, T0 ? T0 + (size_t)q * 16 : nullptr, max_dist,
                               max_iter, rel_fit, rel_rmse, T + (size_t)q * 16, corr, smem[wid], lane, tie_eps);
        if (lane == 0) { fitness[q] = r.fitness; rmse[q] = r.rmse; iters[q] = r.iters; if (fragile) fragile[q] = (uint8_t)r.fragile; }
    }
}

extern "C" int f4l_patch_icp_ex(const float* src, const int32_t* src_idx, const int32_t* s_start,
                                const int32_t* s_count, const float* tgt, const int32_t* tgt_idx,
                                const int32_t* t_start, const int32_t* t_count, const uint8_t* seg_skip, int32_t Q,
                                const double* T0, double max_corr_dist, int32_t max_iter, double rel_fitness,
                                double rel_rmse, double* T, double* fitness, double* rmse, int32_t* iters,
                                int32_t* corr, uint8_t* fragile, double tie_eps, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (!(tie_eps > 0.0)) tie_eps = F4L_ICP_TIE_EPS;
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(src && tgt && s_start && t_start && T && fitness && rmse && iters, "null pointer");
    F4L_REQUIRE(max_corr_dist > 0.0, "max_correspondence_distance must be > 0 (Open3D raises too)");
    F4L_REQUIRE(max_iter >= 0, "max_iter < 0");
    const size_t smem = (size_t)ICP_SMEM_PTS * 3 * sizeof(float);
    static F4lPerDevice once;
    if (!once.done()) {
        if (!f4l_optin_smem(k_patch_icp_warp, ICPW_WARPS * sizeof(WarpIcpSmem), "k_patch_icp_warp") ||
            !f4l_optin_smem(k_patch_icp<false>, smem, "k_patch_icp") || !f4l_optin_smem(k_patch_icp<true>, smem, "k_patch_icp"))
            return F4L_E_CUDA;
        once.mark();
    }
    const int grid_w = f4l_div_up(Q, ICPW_WARPS) < 148 * 8 ? f4l_div_up(Q, ICPW_WARPS) : 148 * 8;
    f4l_mark("k_patch_icp_warp", (cudaStream_t)stream);
    k_patch_icp_warp<<<grid_w, ICPW_WARPS * 32, ICPW_WARPS * sizeof(WarpIcpSmem), (cudaStream_t)stream>>>(
        src, src_idx, s_start, s_count, tgt, tgt_idx, t_start, t_count, seg_skip, Q, T0, max_corr_dist, max_iter,
        rel_fitness, rel_rmse, T, fitness, rmse, iters, corr, fragile, tie_eps);
    const int grid = Q < 148 * 16 ? Q : 148 * 16;
    f4l_mark("k_patch_icp", (cudaStream_t)stream);
    if (fragile)
        k_patch_icp<true><<<grid, ICP_THREADS, smem, (cudaStream_t)stream>>>(
            src, src_idx, s_start, s_count, tgt, tgt_idx, t_start, t_count, seg_skip, Q, T0, max_corr_dist, max_iter,
            rel_fitness, rel_rmse, T, fitness, rmse, iters, corr, fragile, tie_eps);
    else
        k_patch_icp<false><<<grid, ICP_THREADS, smem, (cudaStream_t)stream>>>(
            src, src_idx, s_start, s_count, tgt, tgt_idx, t_start, t_count, seg_skip, Q, T0, max_corr_dist, max_iter,
            rel_fitness, rel_rmse, T, fitness, rmse, iters, corr, fragile, tie_eps);
    return f4l_finish("f4l_patch_icp", stream);
}

extern "C" int f4l_patch_icp(const float* src, const int32_t* src_idx, const int32_t* s_start,
                             const int32_t* s_count, const float* tgt, const int32_t* tgt_idx,
                             const int32_t* t_start, const int32_t* t_count, const uint8_t* seg_skip, int32_t Q,
                             const double* T0, double max_corr_dist, int32_t max_iter, double rel_fitness,
                             double rel_rmse, double* T, double* fitness, double* rmse, int32_t* iters,
                             int32_t* corr, void* stream) {
    return f4l_patch_icp_ex(src, src_idx, s_start, s_count, tgt, tgt_idx, t_start, t_count, seg_skip, Q, T0, max_corr_dist,
                            max_iter, rel_fitness, rel_rmse, T, fitness, rmse, iters, corr, nullptr, 0.0, stream);
}

// ------------------------------------------------------------------------------------------
// Segmented 1-NN (row A4): one CTA per segment pair, reference points staged in shared memory.
// The query is the f32-rounded T*p (what the reference hands to the KD-tree), distances in fp64
// as Open3D computes them, first minimal index wins.
#define SNN_THREADS 256
#define SNN_SMEM_PTS 8192

__global__ void __launch_bounds__(SNN_THREADS)
k_segmented_nn(const float* __restrict__ qpts, const int32_t* __restrict__ qidx,
               const int32_t* __restrict__ q_start, const int32_t* __restrict__ q_count,
               const float* __restrict__ rpts, const int32_t* __restrict__ ridx,
               const int32_t* __restrict__ r_start, const int32_t* __restrict__ r_count, int Q,
               const float* __restrict__ T, const float* __restrict__ thr, int32_t* __restrict__ nn,
               float* __restrict__ d2out) {
    extern __shared__ float sref[];
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        int s0, ns, t0, nt;
        seg_bounds(q_start, q_count, q, s0, ns);
        seg_bounds(r_start, r_count, q, t0, nt);
        const bool staged = nt <= SNN_SMEM_PTS;
        __syncthreads();
        if (staged) {
            for (int j = threadIdx.x; j < nt; j += SNN_THREADS) {
                float x, y, z;
                load_ptf(rpts, ridx, t0 + j, x, y, z);
                sref[3 * j] = x; sref[3 * j + 1] = y; sref[3 * j + 2] = z;
            }
        }
        __syncthreads();
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tv[3] = {0, 0, 0};
        if (T) {
            const float* Tq = T + (size_t)q * 16;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                R[i * 3] = Tq[i * 4]; R[i * 3 + 1] = Tq[i * 4 + 1]; R[i * 3 + 2] = Tq[i * 4 + 2];
                tv[i] = Tq[i * 4 + 3];
            }
        }
        const double th = thr ? (double)thr[q] : INFINITY;
        const double th2 = th * th;
        for (int i = threadIdx.x; i < ns; i += SNN_THREADS) {
            double x, y, z;
            load_pt(qpts, qidx, s0 + i, x, y, z);
            const double px = (double)(float)(R[0] * x + R[1] * y + R[2] * z + tv[0]);
            const double py = (double)(float)(R[3] * x + R[4] * y + R[5] * z + tv[1]);
            const double pz = (double)(float)(R[6] * x + R[7] * y + R[8] * z + tv[2]);
            double best = INFINITY;
            int bj = -1;
            for (int j = 0; j < nt; ++j) {
                float gx, gy, gz;
                if (staged) { gx = sref[3 * j]; gy = sref[3 * j + 1]; gz = sref[3 * j + 2]; }
                else load_ptf(rpts, ridx, t0 + j, gx, gy, gz);
                const double dx = px - (double)gx, dy = py - (double)gy, dz = pz - (double)gz;
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) { best = d2; bj = j; }
            }
            const bool ok = best < th2;   // base.py:82 strict
            nn[s0 + i] = ok ? bj : -1;
            if (d2out) d2out[s0 + i] = (float)best;
        }
    }
}

extern "C" int f4l_segmented_nn(const float* qpts, const int32_t* qidx, const int32_t* q_start,
                                const int32_t* q_count, const float* rpts, const int32_t* ridx,
                                const int32_t* r_start, const int32_t* r_count, int32_t Q, const float* T,
                                const float* thr, int32_t* nn, float* d2, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(qpts && rpts && q_start && r_start && nn, "null pointer");
    const size_t smem = (size_t)SNN_SMEM_PTS * 3 * sizeof(float);
    static F4lPerDevice once;
    if (!once.done()) {
        if (!f4l_optin_smem(k_segmented_nn, smem, "k_segmented_nn")) return F4L_E_CUDA;
        once.mark();
    }
    const int grid = Q < 148 * 8 ? Q : 148 * 8;
    f4l_mark("k_segmented_nn", (cudaStream_t)stream);
    k_segmented_nn<<<grid, SNN_THREADS, smem, (cudaStream_t)stream>>>(qpts, qidx, q_start, q_count, rpts, ridx,
                                                                     r_start, r_count, Q, T, thr, nn, d2);
    return f4l_finish("f4l_segmented_nn", stream);
}
