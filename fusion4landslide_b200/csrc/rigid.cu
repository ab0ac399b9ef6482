// K-d segmented weighted Kabsch / Procrustes, K-f transform apply, K-c rigidity check and
// segmented median.  A segment's points are read from HBM exactly once, moments are accumulated in fp64
// about a per-segment pivot, the 3x3 SVD runs in fp64, one segment per lane of a fitting warp.
#include "common.cuh"
#include "rigid_device.cuh"

// ------------------------------------------------------------------------------------------
// K-d segmented weighted Kabsch / Procrustes: k_kabsch_fused (moments + fit, below) and the optional residual pass.
#define KAB_WARPS 4
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
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

// ---- fused form: moments + fit in ONE persistent launch ---------------------------------------------------------
// CTAs are persistent (3 per SM) and walk over groups of 16 consecutive segments.  Warps 0-7 STREAM: warp w accumulates
// the moments of segments w and w + 8 of the group through a two-stage pipeline -- the next 128-point chunk is in flight
// while the current one is accumulated -- and leaves the 16 moments and the pivot of each segment in shared memory.
// Warp 8 FITS: its 32 lanes run the reference formulas and the 3x3 SVD of the 2 x 16 segments of the two previous groups
// while the streaming warps are already on the next ones (full / empty mbarriers over four moment tables), so the long
// fp64 dependency chain of the Jacobi sweeps (~1 800 dependent instructions per fit) is off the streaming path and costs
// one warp slot in nine; no second launch, no scratch round trip, no stream-ordered allocation.  16 segments per group
// keep the scheduling quantum at ~100 KB per CTA.
// Staging of packed pairs: ONE TMA bulk copy per array and chunk (cp.async.bulk global -> shared, completion on the
// stage's mbarrier), issued by lane 0: the 16-byte lines that cover the run, the run itself starting `mis` floats into
// the buffer (segments start at any multiple of 12 bytes).  A chunk costs ~30 instructions to stage instead of ~250 with
// per-lane 4-byte cp.async: the kernel is issue- / latency-bound, not DRAM-bound (ncu: 83 M warp instructions at 16 M
// pairs with per-lane staging, 51 M with bulk copies).
#define KF_WARPS 8                      // streaming warps; warp KF_WARPS fits
#ifndef KF_SEGS
#define KF_SEGS 1                       // segments per streaming warp and group
#endif
#define KF_GROUP (KF_WARPS * KF_SEGS)
#define KF_FIT_GROUPS (32 / KF_GROUP)   // groups the fitting warp takes at once (one segment per lane)
#define KF_CHUNK 128
#define KF_STAGE_FLOATS (KF_CHUNK * 3 + 8)
#define KF_MOM 23                       // 16 moments + 2 pivots, odd stride (conflict-free thread-per-segment reads)
#define KF_CTAS_PER_SM 3
#define KF_SLOTS (2 * KF_FIT_GROUPS)     // moment tables in flight: half being fitted, half being streamed
struct KfWarpStage {
    float s[KF_STAGE_FLOATS];
    float t[KF_STAGE_FLOATS];
    float w[KF_CHUNK + 8];
};
struct KfSmem {
    KfWarpStage st[KF_WARPS][2];
    double mom[KF_SLOTS][KF_GROUP][KF_MOM];
    unsigned long long bar[KF_WARPS][2];
    unsigned long long full[KF_SLOTS], empty[KF_SLOTS];
    int seg[KF_WARPS][KF_SEGS][2];      // (first item, count) of the warp's segments
};
static_assert(sizeof(KfWarpStage) % 16 == 0 && sizeof(KfWarpStage) >= 8 * 33 * 8, "stage: 16-byte multiples; hosts the reduction tile");
static_assert(offsetof(KfWarpStage, t) % 16 == 0 && offsetof(KfWarpStage, w) % 16 == 0, "bulk-copy destinations are 16-byte aligned");

__device__ __forceinline__ uint32_t kf_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kf_mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KF_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra KF_WAIT_%=;\n\t}" ::"r"(mbar), "r"(parity)
        : "memory");
}

// One bulk copy stages the run of F floats at g: the 16-byte lines that cover it, [g - mis, round-up of the end), land at
// `dst` (16-byte aligned) and the run starts at dst + mis floats (mis = g's offset inside its line).  Up to 12 bytes
// before / after the run ride along; they lie in lines that hold valid bytes of the same array, so they are mapped.
// Returns the bytes the stage's mbarrier has to expect.
__device__ __forceinline__ uint32_t kf_bulk_bytes(const float* g, int F) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(g) >> 2) & 3u);
    return ((mis + (uint32_t)F) * 4u + 15u) & ~15u;
}
__device__ __forceinline__ void kf_bulk(uint32_t dst, const float* g, uint32_t bytes, uint32_t mbar) {
    const float* g_al = reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g_al),
                 "r"(bytes), "r"(mbar)
                 : "memory");
}

__device__ __forceinline__ void moments_add_unit(Moments& M, double sx, double sy, double sz, double tx, double ty, double tz) {
    // moments_add with w = 1 (bit-identical: the products with 1.0 are exact)
    M.m[0] += 1.0;
    M.m[1] += sx; M.m[2] += sy; M.m[3] += sz;
    M.m[4] += tx; M.m[5] += ty; M.m[6] += tz;
    M.m[7] += sx * tx; M.m[8] += sx * ty; M.m[9] += sx * tz;
    M.m[10] += sy * tx; M.m[11] += sy * ty; M.m[12] += sy * tz;
    M.m[13] += sz * tx; M.m[14] += sz * ty; M.m[15] += sz * tz;
}

// warp total of the 16 moments -> row[0..15], through an 8 x 33 tile of doubles (two rounds of eight moments): lane l
// adds the eight partials of moment (l & 7) held by lanes 8 (l >> 3) .., two shuffles finish -- 64 instructions instead
// of the 150 of the butterfly reduce-scatter.  Deterministic order.
__device__ __forceinline__ void kf_reduce(const Moments& M, double* tile, double* row, int lane) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[i * 33 + lane] = M.m[8 * r + i];
        __syncwarp();
        const double* col = tile + (lane & 7) * 33 + (lane >> 3) * 8;
        double v = col[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) v += col[k];
        v += __shfl_xor_sync(F4L_FULL, v, 8);
        v += __shfl_xor_sync(F4L_FULL, v, 16);
        if (lane < 8) row[8 * r + lane] = v;
        __syncwarp();
    }
}

template <bool HAS_W>
__global__ void __launch_bounds__((KF_WARPS + 1) * 32, KF_CTAS_PER_SM)
k_kabsch_fused(const float* __restrict__ src, const float* __restrict__ tgt, const int32_t* __restrict__ src_idx,
               const int32_t* __restrict__ tgt_idx, const float* __restrict__ w, const int32_t* __restrict__ seg_start,
               const int32_t* __restrict__ seg_count, int Q, double eps, float weight_thresh, int variant,
               float* __restrict__ R, float* __restrict__ t, double* __restrict__ T64, uint8_t* __restrict__ flag,
               double* __restrict__ fit64) {
    extern __shared__ __align__(16) unsigned char kf_raw[];
    KfSmem& sm = *reinterpret_cast<KfSmem*>(kf_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_groups = (Q + KF_GROUP - 1) / KF_GROUP;
    const uint32_t sm_base = kf_u32(kf_raw);
    const uint32_t full_base = sm_base + (uint32_t)offsetof(KfSmem, full), empty_base = sm_base + (uint32_t)offsetof(KfSmem, empty);
    if (threadIdx.x < KF_SLOTS) {
        // every thread that wrote (read) a moment table arrives itself: its own release (acquire) orders its accesses
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full_base + threadIdx.x * 8u), "r"((uint32_t)KF_WARPS * 32u) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty_base + threadIdx.x * 8u), "r"((uint32_t)KF_GROUP) : "memory");
    }
    if (wid < KF_WARPS && lane < 2)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sm_base + (uint32_t)offsetof(KfSmem, bar) + (uint32_t)(wid * 2 + lane) * 8u),
                     "r"(1u)
                     : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (wid == KF_WARPS) {
        // ---- fitting warp: the 32 lanes fit the segments of KF_FIT_GROUPS consecutive groups of this CTA at once
        const int sub = lane / KF_GROUP, l = lane % KF_GROUP;
        for (int it = 0, g = blockIdx.x; g < n_groups; g += KF_FIT_GROUPS * gridDim.x, it += KF_FIT_GROUPS) {
#pragma unroll
            for (int u = 0; u < KF_FIT_GROUPS; ++u)
                if (g + u * (int)gridDim.x < n_groups)
                    kf_mbar_wait(full_base + (uint32_t)((it + u) & (KF_SLOTS - 1)) * 8u, (uint32_t)((it + u) / KF_SLOTS) & 1u);
            const int slot = (it + sub) & (KF_SLOTS - 1);
            const int gs = g + sub * (int)gridDim.x;
            const int q = gs * KF_GROUP + l;
            if (gs < n_groups && q < Q) {
                int s0, n;
                seg_bounds(seg_start, seg_count, q, s0, n);
                double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tv[3] = {0, 0, 0};
                bool bad = true;
                if (n > 0) {
                    const double* row = sm.mom[slot][l];
                    Moments Mq;
#pragma unroll
                    for (int i = 0; i < 16; ++i) Mq.m[i] = row[i];
                    double qs[3] = {row[16], row[17], row[18]}, qt[3] = {row[19], row[20], row[21]};
                    bad = fit_from_moments(Mq, qs, qt, eps, variant, Rm, tv);
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
                if (fit64) {                  // fp64 fit for the residual pass
#pragma unroll
                    for (int i = 0; i < 9; ++i) fit64[(size_t)q * 16 + i] = Rm[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) fit64[(size_t)q * 16 + 9 + i] = tv[i];
                }
            }
            if (gs < n_groups)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_base + (uint32_t)slot * 8u) : "memory");
        }
        return;
    }

    // ---- streaming warps ------------------------------------------------------------------------------------------
    const bool packed = !src_idx && !tgt_idx;
    const int (*seg)[2] = sm.seg[wid];
    const uint32_t bar_base = sm_base + (uint32_t)offsetof(KfSmem, bar) + (uint32_t)wid * 16u;
    const uint32_t stage_base = sm_base + (uint32_t)wid * 2u * (uint32_t)sizeof(KfWarpStage);
    // put the chunk (segment first item s0, chunk offset c0, n items) in flight into stage b
    auto issue = [&](int s0, int c0, int n, int b) {
        const int cnt = min(KF_CHUNK, n - c0);
        const uint32_t mbar = bar_base + (uint32_t)b * 8u;
        if (packed) {
            if (lane == 0) {
                const uint32_t dst = stage_base + (uint32_t)b * (uint32_t)sizeof(KfWarpStage);
                const float* gs = src + (size_t)(s0 + c0) * 3;
                const float* gt = tgt + (size_t)(s0 + c0) * 3;
                const uint32_t bs = kf_bulk_bytes(gs, 3 * cnt), bt = kf_bulk_bytes(gt, 3 * cnt);
                const uint32_t bw = HAS_W ? kf_bulk_bytes(w + s0 + c0, cnt) : 0u;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bs + bt + bw) : "memory");
                kf_bulk(dst, gs, bs, mbar);
                kf_bulk(dst + (uint32_t)offsetof(KfWarpStage, t), gt, bt, mbar);
                if (HAS_W) kf_bulk(dst + (uint32_t)offsetof(KfWarpStage, w), w + s0 + c0, bw, mbar);
            }
        } else {
            KfWarpStage& S = sm.st[wid][b];
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
            for (int j = lane; j < cnt; j += 32) {
                const size_t a = src_idx ? (size_t)src_idx[s0 + c0 + j] : (size_t)(s0 + c0 + j);
                const size_t bb = tgt_idx ? (size_t)tgt_idx[s0 + c0 + j] : (size_t)(s0 + c0 + j);
#pragma unroll
                for (int k = 0; k < 3; ++k) { cp_async4(S.s + 3 * j + k, src + a * 3 + k); cp_async4(S.t + 3 * j + k, tgt + bb * 3 + k); }
                if (HAS_W) cp_async4(S.w + j, w + s0 + c0 + j);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    };
    // offsets (in floats) of the arrays inside their 16-byte lines: the run of item k starts at (off + stride k) & 3
    const int off_s = packed ? (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u) : 0;
    const int off_t = packed ? (int)((reinterpret_cast<uintptr_t>(tgt) >> 2) & 3u) : 0;
    const int off_w = (packed && HAS_W) ? (int)((reinterpret_cast<uintptr_t>(w) >> 2) & 3u) : 0;
    auto advance = [&](int& s, int& c) {      // the item after (s, c); s == KF_SEGS: none
        c += KF_CHUNK;
        if (c >= seg[s][1]) {
            c = 0;
            do { ++s; } while (s < KF_SEGS && seg[s][1] == 0);
        }
    };
    int buf = 0;
    unsigned phase = 0;                       // bit b: parity the next wait on stage b expects
    int it = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
        const int mb = it & (KF_SLOTS - 1);
        const int q0 = g * KF_GROUP;
        if (lane < KF_SEGS) {
            const int q = q0 + wid + KF_WARPS * lane;
            int s0 = 0, n = 0;
            if (q < Q) seg_bounds(seg_start, seg_count, q, s0, n);
            sm.seg[wid][lane][0] = s0;
            sm.seg[wid][lane][1] = max(n, 0);
        }
        __syncwarp();
        int cs = 0, cc = 0;                   // item being computed
        if (seg[0][1] == 0) { cc = -KF_CHUNK; advance(cs, cc); }
        if (cs < KF_SEGS) issue(seg[cs][0], cc, seg[cs][1], buf);
        // this moment table is free once the fitting warp is done with the group that used it KF_SLOTS groups ago
        if (it >= KF_SLOTS) kf_mbar_wait(empty_base + (uint32_t)mb * 8u, (uint32_t)(it / KF_SLOTS - 1) & 1u);
        Moments M;
        moments_zero(M);
        double ps[3] = {0, 0, 0}, pt[3] = {0, 0, 0};
        while (cs < KF_SEGS) {
            const int n = seg[cs][1];
            const int k0 = packed ? seg[cs][0] + cc : 0;
            int is = cs, ic = cc;
            advance(is, ic);
            if (is < KF_SEGS) {
                issue(seg[is][0], ic, seg[is][1], buf ^ 1);
                if (!packed) asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else if (!packed) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            kf_mbar_wait(bar_base + (uint32_t)buf * 8u, (phase >> buf) & 1u);
            phase ^= 1u << buf;
            __syncwarp();
            const KfWarpStage& S = sm.st[wid][buf];
            const float* ss = S.s + ((off_s + 3 * k0) & 3);
            const float* st = S.t + ((off_t + 3 * k0) & 3);
            const float* sw = S.w + ((off_w + k0) & 3);
            if (cc == 0) {                    // pivot = the segment's first pair
                ps[0] = ss[0]; ps[1] = ss[1]; ps[2] = ss[2];
                pt[0] = st[0]; pt[1] = st[1]; pt[2] = st[2];
            }
            const int cnt = min(KF_CHUNK, n - cc);
#pragma unroll 2
            for (int j = lane; j < cnt; j += 32) {
                const double sx = (double)ss[3 * j] - ps[0], sy = (double)ss[3 * j + 1] - ps[1], sz = (double)ss[3 * j + 2] - ps[2];
                const double tx = (double)st[3 * j] - pt[0], ty = (double)st[3 * j + 1] - pt[1], tz = (double)st[3 * j + 2] - pt[2];
                if (HAS_W) {
                    float wf = sw[j];
                    if (variant == 0 && wf < weight_thresh) wf = 0.f;
                    moments_add(M, (double)wf, sx, sy, sz, tx, ty, tz);
                } else {
                    moments_add_unit(M, sx, sy, sz, tx, ty, tz);
                }
            }
            __syncwarp();
            if (cc + KF_CHUNK >= n) {         // segment complete: the stage just consumed hosts the reduction tile
                double* row = sm.mom[mb][wid + KF_WARPS * cs];
                kf_reduce(M, reinterpret_cast<double*>(&sm.st[wid][buf]), row, lane);
                if (lane == 0) {
                    row[16] = ps[0]; row[17] = ps[1]; row[18] = ps[2];
                    row[19] = pt[0]; row[20] = pt[1]; row[21] = pt[2];
                }
                moments_zero(M);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile's generic writes before a later bulk copy lands here
                __syncwarp();
            }
            cs = is; cc = ic;
            buf ^= 1;
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_base + (uint32_t)mb * 8u) : "memory");
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
    static F4lPerDevice optin;
    if (!optin.done()) {
        if (!f4l_optin_smem(k_kabsch_fused<false>, sizeof(KfSmem), "k_kabsch_fused") ||
            !f4l_optin_smem(k_kabsch_fused<true>, sizeof(KfSmem), "k_kabsch_fused")) return F4L_E_CUDA;
        // keep freed blocks in the default pool instead of returning them to the OS at every synchronisation
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        optin.mark();
    }
    // 128 B of scratch per segment (the fp64 fit, for the residual pass), stream-ordered pool allocation
    double* mom = nullptr;
    if (res && cudaMallocAsync((void**)&mom, (size_t)Q * 16 * sizeof(double), st) != cudaSuccess) {
        f4l_set_error("f4l_segmented_kabsch: cudaMallocAsync of %zu bytes failed", (size_t)Q * 128);
        cudaGetLastError();
        return F4L_E_CUDA;
    }
    const int kf_grid = f4l_div_up(Q, KF_GROUP) < 148 * KF_CTAS_PER_SM ? f4l_div_up(Q, KF_GROUP) : 148 * KF_CTAS_PER_SM;
    f4l_mark("k_kabsch_fused", st);
    if (w)
        k_kabsch_fused<true><<<kf_grid, (KF_WARPS + 1) * 32, sizeof(KfSmem), st>>>(
            src, tgt, src_idx, tgt_idx, w, seg_start, seg_count, Q, (double)eps, weight_thresh, variant, R, t, T64, flag, mom);
    else
        k_kabsch_fused<false><<<kf_grid, (KF_WARPS + 1) * 32, sizeof(KfSmem), st>>>(
            src, tgt, src_idx, tgt_idx, w, seg_start, seg_count, Q, (double)eps, weight_thresh, variant, R, t, T64, flag, mom);
    if (res) {
        f4l_mark("k_kabsch_residuals", st);
        k_kabsch_residuals<<<f4l_div_up(Q, KAB_WARPS), KAB_WARPS * 32, 0, st>>>(src, tgt, src_idx, tgt_idx, seg_start, seg_count,
                                                                               Q, mom, res);
    }
    if (mom) cudaFreeAsync(mom, st);
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
