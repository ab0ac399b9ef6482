// K-d segmented weighted Kabsch / Procrustes, K-f transform apply, K-c rigidity check and
// segmented median.  One warp per segment for the reductions: a segment's points are read from
// HBM exactly once, moments are accumulated in fp64 about a per-segment pivot, the 3x3 SVD runs
// in fp64 on lane 0.
#include "common.cuh"
#include "rigid_device.cuh"

// ------------------------------------------------------------------------------------------
// K-d.  A warp owns a GROUP of 32 consecutive segments: the moments of each segment are accumulated
// cooperatively (coalesced, every point read from HBM once) and reduce-scattered over the warp with 16
// double shuffles; then every lane solves the 3x3 SVD of ITS segment, so the long fp64 dependency chain of
// the Jacobi sweeps runs 32-wide instead of once per warp.
#define KAB_WARPS 4

// butterfly reduce-scatter of the 16 moments: afterwards lane l holds the warp total of m[(l >> 1) & 15]
__device__ __forceinline__ double moments_reduce_scatter(const Moments& M, int lane) {
    double v8[8], v4[4], v2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double send = b4 ? M.m[i] : M.m[i + 8];
        const double keep = b4 ? M.m[i + 8] : M.m[i];
        v8[i] = keep + __shfl_xor_sync(F4L_FULL, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = b3 ? v8[i] : v8[i + 4];
        const double keep = b3 ? v8[i + 4] : v8[i];
        v4[i] = keep + __shfl_xor_sync(F4L_FULL, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b2 ? v4[i] : v4[i + 2];
        const double keep = b2 ? v4[i + 2] : v4[i];
        v2[i] = keep + __shfl_xor_sync(F4L_FULL, send, 4);
    }
    double v = (b1 ? v2[1] : v2[0]) + __shfl_xor_sync(F4L_FULL, b1 ? v2[0] : v2[1], 2);
    v += __shfl_xor_sync(F4L_FULL, v, 1);
    return v;
}

// pass 1: warp per segment, every point read from HBM exactly once, 16 raw moments to scratch.
// The kernel is a pure stream with ~25 fp64 ops per 24 bytes, so what limits it is bytes in flight, and
// with fp64 accumulators registers cap the occupancy: each warp therefore stages its segment through shared
// memory with cp.async (LDGSTS, 4-byte granules so that unaligned / gathered segments work) -- a whole
// 256-point chunk (6 KB) is in flight per warp without holding a single register.
#define KAB_CHUNK 256
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}

__global__ void __launch_bounds__(KAB_WARPS * 32)
k_kabsch_moments(const float* __restrict__ src, const float* __restrict__ tgt,
                 const int32_t* __restrict__ src_idx, const int32_t* __restrict__ tgt_idx,
                 const float* __restrict__ w, const int32_t* __restrict__ seg_start,
                 const int32_t* __restrict__ seg_count, int Q, float weight_thresh, int variant,
                 double* __restrict__ mom) {
    __shared__ float stage[KAB_WARPS][2][KAB_CHUNK * 3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int q = blockIdx.x * KAB_WARPS + wid;
    if (q >= Q) return;
    int s0, n;
    seg_bounds(seg_start, seg_count, q, s0, n);
    float* ss = stage[wid][0];
    float* st = stage[wid][1];
    Moments M;
    moments_zero(M);
    if (n > 0) {
        double ps[3], pt[3];
        load_pt(src, src_idx, s0, ps[0], ps[1], ps[2]);
        load_pt(tgt, tgt_idx, s0, pt[0], pt[1], pt[2]);
        for (int c0 = 0; c0 < n; c0 += KAB_CHUNK) {
            const int cnt = min(KAB_CHUNK, n - c0);
            if (!src_idx && !tgt_idx) {
                // packed pairs: the chunk is one contiguous run of 3*cnt floats per array
                const float* gs = src + (size_t)(s0 + c0) * 3;
                const float* gt = tgt + (size_t)(s0 + c0) * 3;
                for (int e = lane; e < 3 * cnt; e += 32) { cp_async4(ss + e, gs + e); cp_async4(st + e, gt + e); }
            } else {
                for (int j = lane; j < cnt; j += 32) {
                    const size_t a = src_idx ? (size_t)src_idx[s0 + c0 + j] : (size_t)(s0 + c0 + j);
                    const size_t b = tgt_idx ? (size_t)tgt_idx[s0 + c0 + j] : (size_t)(s0 + c0 + j);
#pragma unroll
                    for (int k = 0; k < 3; ++k) { cp_async4(ss + 3 * j + k, src + a * 3 + k); cp_async4(st + 3 * j + k, tgt + b * 3 + k); }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            for (int j = lane; j < cnt; j += 32) {
                double wi = 1.0;
                if (w) {
                    float wf = __ldg(w + s0 + c0 + j);
                    if (variant == 0 && wf < weight_thresh) wf = 0.f;
                    wi = (double)wf;
                }
                moments_add(M, wi, (double)ss[3 * j] - ps[0], (double)ss[3 * j + 1] - ps[1], (double)ss[3 * j + 2] - ps[2],
                            (double)st[3 * j] - pt[0], (double)st[3 * j + 1] - pt[1], (double)st[3 * j + 2] - pt[2]);
            }
            __syncwarp();
        }
    }
    const double v = moments_reduce_scatter(M, lane);
    if (!(lane & 1)) mom[(size_t)q * 16 + (lane >> 1)] = v;
}

// pass 2: THREAD per segment: reference formulas + 3x3 SVD, 32 segments per warp in flight
__global__ void __launch_bounds__(128)
k_kabsch_fit(const float* __restrict__ src, const float* __restrict__ tgt, const int32_t* __restrict__ src_idx,
             const int32_t* __restrict__ tgt_idx, const int32_t* __restrict__ seg_start,
             const int32_t* __restrict__ seg_count, int Q, double eps, int variant, double* __restrict__ mom,
             float* __restrict__ R, float* __restrict__ t, double* __restrict__ T64, uint8_t* __restrict__ flag) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    int s0, n;
    seg_bounds(seg_start, seg_count, q, s0, n);
    double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tv[3] = {0, 0, 0};
    bool bad = true;
    if (n > 0) {
        Moments M;
#pragma unroll
        for (int i = 0; i < 16; ++i) M.m[i] = mom[(size_t)q * 16 + i];
        double ps[3], pt[3];
        load_pt(src, src_idx, s0, ps[0], ps[1], ps[2]);
        load_pt(tgt, tgt_idx, s0, pt[0], pt[1], pt[2]);
        bad = fit_from_moments(M, ps, pt, eps, variant, Rm, tv);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[(size_t)q * 9 + i] = (float)Rm[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[(size_t)q * 3 + i] = (float)tv[i];
    if (T64) {
        double* T = T64 + (size_t)q * 16;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            T[i * 4 + 0] = Rm[i * 3 + 0]; T[i * 4 + 1] = Rm[i * 3 + 1]; T[i * 4 + 2] = Rm[i * 3 + 2];
            T[i * 4 + 3] = tv[i];
        }
        T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
    }
    if (flag) flag[q] = bad ? 1 : 0;
    // fp64 fit for the residual pass (overwrites this segment's moments)
#pragma unroll
    for (int i = 0; i < 9; ++i) mom[(size_t)q * 16 + i] = Rm[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) mom[(size_t)q * 16 + 9 + i] = tv[i];
}

// pass 3 (optional): residuals ||R s + t - tgt||, warp per segment
__global__ void __launch_bounds__(KAB_WARPS * 32)
k_kabsch_residuals(const float* __restrict__ src, const float* __restrict__ tgt, const int32_t* __restrict__ src_idx,
                   const int32_t* __restrict__ tgt_idx, const int32_t* __restrict__ seg_start,
                   const int32_t* __restrict__ seg_count, int Q, const double* __restrict__ fit, float* __restrict__ res) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * KAB_WARPS + (threadIdx.x >> 5);
    if (q >= Q) return;
    int s0, n;
    seg_bounds(seg_start, seg_count, q, s0, n);
    double Rs[9], ts[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rs[i] = fit[(size_t)q * 16 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) ts[i] = fit[(size_t)q * 16 + 9 + i];
    for (int i = lane; i < n; i += 32) {
        const int k = s0 + i;
        double sx, sy, sz, tx, ty, tz;
        load_pt(src, src_idx, k, sx, sy, sz);
        load_pt(tgt, tgt_idx, k, tx, ty, tz);
        const double rx = Rs[0] * sx + Rs[1] * sy + Rs[2] * sz + ts[0] - tx;
        const double ry = Rs[3] * sx + Rs[4] * sy + Rs[5] * sz + ts[1] - ty;
        const double rz = Rs[6] * sx + Rs[7] * sy + Rs[8] * sz + ts[2] - tz;
        res[k] = (float)sqrt(rx * rx + ry * ry + rz * rz);
    }
}

extern "C" int f4l_segmented_kabsch(const float* src, const float* tgt, const int32_t* src_idx,
                                    const int32_t* tgt_idx, const float* w, const int32_t* seg_start,
                                    const int32_t* seg_count, int32_t Q, float eps, float weight_thresh,
                                    int variant, float* R, float* t, double* T64, float* res,
                                    uint8_t* flag, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(src && tgt && seg_start && R && t, "null pointer");
    F4L_REQUIRE(variant == F4L_KABSCH_PROCRUSTES || variant == F4L_KABSCH_F2S3, "unknown variant");
    cudaStream_t st = (cudaStream_t)stream;
    // 128 B of scratch per segment (raw moments, then the fp64 fit), stream-ordered pool allocation
    double* mom = nullptr;
    static F4lPerDevice pool_ready;
    if (!pool_ready.done()) {
        // keep freed blocks in the default pool instead of returning them to the OS at every synchronisation
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready.mark();
    }
    if (cudaMallocAsync((void**)&mom, (size_t)Q * 16 * sizeof(double), st) != cudaSuccess) {
        f4l_set_error("f4l_segmented_kabsch: cudaMallocAsync of %zu bytes failed", (size_t)Q * 128);
        cudaGetLastError();
        return F4L_E_CUDA;
    }
    f4l_mark("k_kabsch_moments", st);
    k_kabsch_moments<<<f4l_div_up(Q, KAB_WARPS), KAB_WARPS * 32, 0, st>>>(src, tgt, src_idx, tgt_idx, w, seg_start, seg_count, Q,
                                                                         weight_thresh, variant, mom);
    f4l_mark("k_kabsch_fit", st);
    k_kabsch_fit<<<f4l_div_up(Q, 128), 128, 0, st>>>(src, tgt, src_idx, tgt_idx, seg_start, seg_count, Q, (double)eps, variant,
                                                    mom, R, t, T64, flag);
    if (res) {
        f4l_mark("k_kabsch_residuals", st);
        k_kabsch_residuals<<<f4l_div_up(Q, KAB_WARPS), KAB_WARPS * 32, 0, st>>>(src, tgt, src_idx, tgt_idx, seg_start, seg_count,
                                                                               Q, mom, res);
    }
    cudaFreeAsync(mom, st);
    return f4l_finish("f4l_segmented_kabsch", stream);
}

// ------------------------------------------------------------------------------------------
// K-f: one warp per segment, rows [p | R p + t] in f32 (the reference applies T in f32,
// base.py:3373); fp64 accumulate then round once.
__global__ void __launch_bounds__(256)
k_apply_transforms(const float* __restrict__ pts, const int32_t* __restrict__ idx,
                   const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_count,
                   const int32_t* __restrict__ out_start, const uint8_t* __restrict__ seg_skip, int Q,
                   const float* __restrict__ T, int inverse, float* __restrict__ dvf,
                   float* __restrict__ mag) {
    __shared__ __align__(16) float s_rows[8][192];
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    if (seg_skip && seg_skip[q]) return;
    int s0, n;
    seg_bounds(seg_start, seg_count, q, s0, n);
    const int o0 = out_start ? out_start[q] : s0;
    const float* Tq = T + (size_t)q * 16;
    double Rm[9], tv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        Rm[i * 3 + 0] = Tq[i * 4 + 0]; Rm[i * 3 + 1] = Tq[i * 4 + 1]; Rm[i * 3 + 2] = Tq[i * 4 + 2];
        tv[i] = Tq[i * 4 + 3];
    }
    // rows of a warp trip are contiguous in the output (32 x 24 B): stage them in shared memory and write them
    // back as three fully coalesced float2 stores instead of three 24-byte-strided ones
    float* buf = s_rows[threadIdx.x >> 5];
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const int cnt = min(32, n - i0);
        float ox = 0.f, oy = 0.f, oz = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
        if (i < n) {
            load_ptf(pts, idx, s0 + i, fx, fy, fz);
            const double x = fx, y = fy, z = fz;
            if (!inverse) {
                ox = (float)(Rm[0] * x + Rm[1] * y + Rm[2] * z + tv[0]);
                oy = (float)(Rm[3] * x + Rm[4] * y + Rm[5] * z + tv[1]);
                oz = (float)(Rm[6] * x + Rm[7] * y + Rm[8] * z + tv[2]);
                buf[lane * 6 + 0] = fx; buf[lane * 6 + 1] = fy; buf[lane * 6 + 2] = fz;
                buf[lane * 6 + 3] = ox; buf[lane * 6 + 4] = oy; buf[lane * 6 + 5] = oz;
            } else {
                // f32 subtraction first, as base.py:3389-3390 does: R^T (p - t)
                const double dx = (double)(fx - (float)tv[0]);
                const double dy = (double)(fy - (float)tv[1]);
                const double dz = (double)(fz - (float)tv[2]);
                ox = (float)(Rm[0] * dx + Rm[3] * dy + Rm[6] * dz);
                oy = (float)(Rm[1] * dx + Rm[4] * dy + Rm[7] * dz);
                oz = (float)(Rm[2] * dx + Rm[5] * dy + Rm[8] * dz);
                buf[lane * 6 + 0] = ox; buf[lane * 6 + 1] = oy; buf[lane * 6 + 2] = oz;
                buf[lane * 6 + 3] = fx; buf[lane * 6 + 4] = fy; buf[lane * 6 + 5] = fz;
            }
            if (mag) {
                const float ex = ox - fx, ey = oy - fy, ez = oz - fz;
                mag[o0 + i] = sqrtf(ex * ex + ey * ey + ez * ez);
            }
        }
        __syncwarp();
        float2* out2 = reinterpret_cast<float2*>(dvf + (size_t)(o0 + i0) * 6);
        const float2* b2 = reinterpret_cast<const float2*>(buf);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int e = lane + 32 * u;
            if (e < cnt * 3) out2[e] = b2[e];
        }
        __syncwarp();
    }
}

extern "C" int f4l_apply_transforms(const float* pts, const int32_t* idx, const int32_t* seg_start,
                                    const int32_t* seg_count, const int32_t* out_start,
                                    const uint8_t* seg_skip, int32_t Q, const float* T, int inverse,
                                    float* dvf, float* mag, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(pts && seg_start && T && dvf, "null pointer");
    const int warps = 8;
    f4l_mark("k_apply_transforms", (cudaStream_t)stream);
    k_apply_transforms<<<f4l_div_up(Q, warps), warps * 32, 0, (cudaStream_t)stream>>>(
        pts, idx, seg_start, seg_count, out_start, seg_skip, Q, T, inverse, dvf, mag);
    return f4l_finish("f4l_apply_transforms", stream);
}

// ------------------------------------------------------------------------------------------
// K-c rigidity check: one CTA per segment, points staged in shared memory (f32), the K(K-1)/2
// pair terms are spread over the CTA's threads.
#define RIG_MAX_SMEM_PTS 2048
__global__ void __launch_bounds__(256)
k_rigidity(const float* __restrict__ src, const float* __restrict__ tgt,
           const int32_t* __restrict__ src_idx, const int32_t* __restrict__ tgt_idx,
           const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_count, int Q,
           float thres, float* __restrict__ ratio_inlier, float* __restrict__ dist_mean) {
    extern __shared__ float sm[];
    __shared__ double red_sum[8];
    __shared__ unsigned long long red_cnt[8];
    const int q = blockIdx.x;
    int s0, n;
    seg_bounds(seg_start, seg_count, q, s0, n);
    double sum;
    unsigned long long cnt;
    block_rigidity(src, tgt, src_idx, tgt_idx, s0, n, thres, sm, red_sum, red_cnt, sum, cnt);
    if (threadIdx.x == 0) {
        double num_ele = 0.5 * (double)n * (double)(n - 1);
        // base.py:3315-3316: count over the full KxK matrix (2*pairs + K diagonal zeros) minus K
        dist_mean[q] = (float)(sum / num_ele);
        ratio_inlier[q] = (float)((double)(2ull * cnt) / (num_ele * 2.0));
    }
}

extern "C" int f4l_rigidity_check(const float* src, const float* tgt, const int32_t* src_idx,
                                  const int32_t* tgt_idx, const int32_t* seg_start,
                                  const int32_t* seg_count, int32_t Q, float thres_dist_diff,
                                  float* ratio_inlier, float* dist_mean, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(src && tgt && seg_start && ratio_inlier && dist_mean, "null pointer");
    size_t smem = (size_t)RIG_MAX_SMEM_PTS * 6 * sizeof(float);
    static F4lPerDevice once;
    if (!once.done()) {
        if (!f4l_optin_smem(k_rigidity, smem, "k_rigidity")) return F4L_E_CUDA;
        once.mark();
    }
    f4l_mark("k_rigidity", (cudaStream_t)stream);
    k_rigidity<<<Q, 256, smem, (cudaStream_t)stream>>>(src, tgt, src_idx, tgt_idx, seg_start, seg_count, Q,
                                                      thres_dist_diff, ratio_inlier, dist_mean);
    return f4l_finish("f4l_rigidity_check", stream);
}

// ------------------------------------------------------------------------------------------
// Segmented lower median (torch.median semantics): one CTA per segment.
__global__ void __launch_bounds__(256)
k_segmented_median(const float* __restrict__ x, const int32_t* __restrict__ seg_start,
                   const int32_t* __restrict__ seg_count, int Q, float* __restrict__ med) {
    __shared__ unsigned hist[256];
    __shared__ unsigned state[2];
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        int s0, n;
        seg_bounds(seg_start, seg_count, q, s0, n);
        __syncthreads();
        if (n <= 0) {
            if (threadIdx.x == 0) med[q] = nanf("");
            continue;
        }
        const float m = block_select_kth(x + s0, n, (n - 1) / 2, hist, state);
        if (threadIdx.x == 0) med[q] = m;
    }
}

extern "C" int f4l_segmented_median(const float* x, const int32_t* seg_start, const int32_t* seg_count, int32_t Q,
                                    float* med, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(x && seg_start && med, "null pointer");
    f4l_mark("k_segmented_median", (cudaStream_t)stream);
    k_segmented_median<<<Q < 148 * 8 ? Q : 148 * 8, 256, 0, (cudaStream_t)stream>>>(x, seg_start, seg_count, Q, med);
    return f4l_finish("f4l_segmented_median", stream);
}

// ------------------------------------------------------------------------------------------
// F2S3 pruning tail per supervoxel (src/models/outlier_classifier.py:71-105): one CTA per
// segment.  corr rows are [src | tgt] (stride 6).  res is both output and scratch.
__device__ __forceinline__ void load_row6(const float* __restrict__ corr, int k, double s[3], double t[3]) {
    const float2* r = reinterpret_cast<const float2*>(corr + (size_t)k * 6);
    const float2 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2);
    s[0] = a.x; s[1] = a.y; s[2] = b.x; t[0] = b.y; t[1] = c.x; t[2] = c.y;
}

__global__ void __launch_bounds__(128)
k_f2s3_prune_tail(const float* __restrict__ corr, const float* __restrict__ scores,
                  const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_count, int Q,
                  float coeff, float* __restrict__ R, float* __restrict__ t, uint8_t* __restrict__ robust,
                  float* __restrict__ res, float* __restrict__ median) {
    __shared__ unsigned hist[256];
    __shared__ unsigned state[2];
    __shared__ double sR[9], st[3];
    __shared__ int s_inl;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int q = blockIdx.x; q < Q; q += gridDim.x) {
        int s0, n;
        seg_bounds(seg_start, seg_count, q, s0, n);
        __syncthreads();
        bool is_robust = false;
        float med = nanf("");
        for (int round = 0; round < 2; ++round) {
            // weighted Kabsch (variant 1: functions.py:36-80); round 0: network scores, round 1: 0/1 inliers
            if (tid < 32) {
                double ps[3] = {0, 0, 0}, pt[3] = {0, 0, 0};
                if (n > 0) load_row6(corr, s0, ps, pt);
                Moments M;
                moments_zero(M);
                const float thr = coeff * med;
                for (int i = lane; i < n; i += 32) {
                    double s[3], g[3];
                    load_row6(corr, s0 + i, s, g);
                    double w = round == 0 ? (double)__ldg(scores + s0 + i) : ((res[s0 + i] < thr) ? 1.0 : 0.0);
                    moments_add(M, w, s[0] - ps[0], s[1] - ps[1], s[2] - ps[2], g[0] - pt[0], g[1] - pt[1], g[2] - pt[2]);
                }
                moments_warp_reduce(M);
                double Rm[9], tv[3];
                if (n > 0) fit_from_moments(M, ps, pt, 1e-7, 1, Rm, tv);
                else { Rm[0] = 1; Rm[1] = 0; Rm[2] = 0; Rm[3] = 0; Rm[4] = 1; Rm[5] = 0; Rm[6] = 0; Rm[7] = 0; Rm[8] = 1; tv[0] = tv[1] = tv[2] = 0; }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) sR[i] = Rm[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) st[i] = tv[i];
                }
            }
            __syncthreads();
            // residuals of this fit (functions.py:100-104)
            for (int i = tid; i < n; i += blockDim.x) {
                double s[3], g[3];
                load_row6(corr, s0 + i, s, g);
                const double rx = sR[0] * s[0] + sR[1] * s[1] + sR[2] * s[2] + st[0] - g[0];
                const double ry = sR[3] * s[0] + sR[4] * s[1] + sR[5] * s[2] + st[1] - g[1];
                const double rz = sR[6] * s[0] + sR[7] * s[1] + sR[8] * s[2] + st[2] - g[2];
                res[s0 + i] = (float)sqrt(rx * rx + ry * ry + rz * rz);
            }
            __syncthreads();
            if (round == 1 || n <= 0) break;
            med = block_select_kth(res + s0, n, (n - 1) / 2, hist, state);          // :80 torch.median
            if (tid == 0) s_inl = 0;
            __syncthreads();
            int c = 0;
            const float thr = coeff * med;
            for (int i = tid; i < n; i += blockDim.x) c += (res[s0 + i] < thr) ? 1 : 0;
            c = warp_sum(c);
            if (lane == 0 && c) atomicAdd(&s_inl, c);
            __syncthreads();
            is_robust = s_inl >= 5 && med < 0.5f;                                     // :91
            if (!is_robust) break;
        }
        if (tid < 9) R[(size_t)q * 9 + tid] = (float)sR[tid];
        if (tid < 3) t[(size_t)q * 3 + tid] = (float)st[tid];
        if (tid == 0) {
            robust[q] = is_robust ? 1 : 0;
            if (median) median[q] = med;
        }
    }
}

extern "C" int f4l_f2s3_prune_tail(const float* corr, const float* scores, const int32_t* seg_start,
                                   const int32_t* seg_count, int32_t Q, float coeff, float* R, float* t,
                                   uint8_t* robust, float* res, float* median, void* stream) {
    F4L_REQUIRE(Q >= 0, "Q < 0");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(corr && scores && seg_start && R && t && robust && res, "null pointer (res is required: it is scratch too)");
    f4l_mark("k_f2s3_prune_tail", (cudaStream_t)stream);
    k_f2s3_prune_tail<<<Q < 148 * 16 ? Q : 148 * 16, 128, 0, (cudaStream_t)stream>>>(corr, scores, seg_start, seg_count, Q,
                                                                                  coeff, R, t, robust, res, median);
    return f4l_finish("f4l_f2s3_prune_tail", stream);
}
