// K-b: exact descriptor-space nearest neighbour (D in {32,64}) on the 5th-generation tensor cores.
//
// Replaces torch.cdist + min at base.py:2783-2815 (global_matches_from_3d, exact branches), the
// hnswlib query at src/f2s3.py:273-281 and the two dense S x T matrices + row/column argmin of the
// coarse matching at base.py:2966-2995.
//
// The distance matrix is a dense contraction: argmin_j ||a_i - b_j||^2 = argmin_j (1/2||b_j||^2 - a_i.b_j).
//
//   k_desc_absmax / k_desc_scale   power-of-two scale that puts max|x| of both matrices in [0.5,1)
//   k_desc_pack      f32 rows -> fp16 operand tiles stored in HBM in the exact shared-memory image
//                    tcgen05.mma reads (K-major, no-swizzle 8x16-byte core matrices; one 128-row tile =
//                    [Kp/8][128][8] fp16, contiguous), plus 16 extra K columns: 1/2||b||^2 split in
//                    hi/mid/lo fp16 on the b side against 1.0 on the a side, so the accumulator IS the score
//   k_desc_nn_tc     persistent CTA per SM, warp-specialised:
//                      warp 0    producer: cp.async.bulk (UBLKCP) of b tiles through a 4-stage mbarrier ring
//                      warp 1    tcgen05.mma issuer: M128 x N128 x K16, kind::f16, fp32 accumulators in
//                                TMEM, 2 query sub-blocks x 2 buffers = all 512 TMEM columns
//                      warps 2-9 epilogue: tcgen05.ld 32 columns at a time, FMNMX3 min tree, and only when
//                                the chunk minimum beats (running minimum + margin) a slow path that records
//                                (score, column) candidates in shared memory
//                    The margin is a proven bound of the fp16/accumulation error, so the exact nearest
//                    neighbour is always among the recorded candidates.
//   k_desc_rerank    fp64 distances of the candidates of each row -> argmin, lower index on ties
//                    (torch.min semantics); rows whose candidate set overflowed are listed ...
//   k_desc_exact     ... and resolved by this fp64 brute-force kernel, which is also the whole
//                    implementation for small problems and for the xyz-gated coarse matching.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

#define DT_TILE 128          // rows per operand tile = UMMA M = UMMA N
#define DT_RB 256            // query rows per CTA row block (2 tiles)
#define DT_STAGES 4
#define DT_CAND 8
#define DT_EPI_WARPS 16      // epilogue warps: (TMEM lane quarter) x (row sub-block) x (column half)
#define DT_HALVES 2          // column halves of a 128-column accumulator tile, one epilogue warp each
#define DT_THREADS (64 + DT_EPI_WARPS * 32)   // producer warp, MMA warp, epilogue warps
#define DT_PAD_BIAS 60000.0f // score of a padding column: never a minimum

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded wait: a broken pipeline traps after ~4 s instead of hanging the device.  The clock is only read once the
// first probe has failed (a CS2R per wait on the fast path cost the MMA-issuing warp more than its MMAs).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 8000000000ll) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (128 B
// contiguous); lbo = byte distance between the two K-chunks of one MMA, sbo = between 8-row groups
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;   // descriptor version of sm_100
    return d;
}
// kind::f16 instruction descriptor: D f32, A/B f16, both K-major, M = N = 128
__device__ __forceinline__ uint32_t instr_desc_f16_128x128() {
    return (1u << 4) | ((uint32_t)(DT_TILE >> 3) << 17) | ((uint32_t)(DT_TILE >> 4) << 24);
}

#define TMEM_LD32(r, taddr)                                                                                          \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
        "%29,%30,%31}, [%32];"                                                                                       \
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), \
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),      \
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),     \
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])                   \
        : "r"(taddr))
// wait for the outstanding tcgen05.ld; the registers are in/out operands so no consumer can be
// scheduled above the wait
#define TMEM_WAIT32(r)                                                                                                \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                     \
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]),    \
                   "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]),           \
                   "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]),         \
                   "+f"(r[22]), "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]),         \
                   "+f"(r[29]), "+f"(r[30]), "+f"(r[31])::"memory")

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---------------------------------------------------------------------------------------------
// scale: sc = 2^e with max|x| * sc in [0.5, 1)  (exact scaling; fp16 range is then never exceeded)
struct DescScalars {
    unsigned absmax_bits;   // max |x| over a and b (float bits, non-negative floats order as uints)
    float scale;
    unsigned bmax_bits;     // max ||b_j * scale|| (float bits)
    int n_overflow;
    unsigned bmin2_bits, bmax2_bits;   // min / max of ||b_j * scale||^2 over the real rows (float bits)
    int use_nobias;         // 1: all reference rows have (almost) the same norm -> the kernel without the bias K-columns runs
    float norm_spread;      // (bmax2 - bmin2) / 4: what ignoring 1/2||b||^2 can move a score by (about the mid value)
};

__global__ void k_desc_scalars_init(DescScalars* s) {
    s->absmax_bits = 0u; s->scale = 1.f; s->bmax_bits = 0u; s->n_overflow = 0;
    s->bmin2_bits = 0x7f800000u; s->bmax2_bits = 0u; s->use_nobias = 0; s->norm_spread = 0.f;
}

__global__ void __launch_bounds__(256) k_desc_absmax(const float* __restrict__ x, size_t n, DescScalars* s) {
    float m = 0.f;
    const size_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x4 + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(F4L_FULL, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&s->absmax_bits, __float_as_uint(m));
}

__global__ void k_desc_scale(DescScalars* s) {
    const float m = __uint_as_float(s->absmax_bits);
    float sc = 1.f;
    if (m > 0.f && isfinite(m)) {
        int e;
        frexpf(m, &e);          // m = f * 2^e, f in [0.5,1)
        sc = ldexpf(1.f, -e);
    }
    s->scale = sc;
}

// After the reference side is packed: if every reference row has (almost) the same norm -- L2-normalised descriptors,
// the reference's case (src/models/local_feature_descriptor.py:107-109) -- 1/2||b_j||^2 is a constant up to
// `norm_spread`, the score is -a.b alone, and the kernel variant WITHOUT the 16 bias K-columns runs (a fifth to a
// third fewer MMAs); the spread is added to the candidate margin, so the result stays exact.  force: 0 auto, 1 bias.
__global__ void k_desc_choose(DescScalars* s, int force_bias) {
    const float lo = __uint_as_float(s->bmin2_bits), hi = __uint_as_float(s->bmax2_bits);
    const bool ok = isfinite(lo) && isfinite(hi) && hi > 0.f && (hi - lo) <= 4e-4f * hi;
    s->use_nobias = (ok && !force_bias) ? 1 : 0;
    s->norm_spread = ok ? 0.25f * (hi - lo) : 0.f;
}

// one warp per row; tile t of 128 rows occupies tile_bytes = 128*Kp*2 contiguous bytes:
//   offset(r, k) = ((k/8)*128 + r)*16 + (k%8)*2
template <int D>
__global__ void __launch_bounds__(256)
k_desc_pack(const float* __restrict__ x, int n, int n_pad, int is_ref, DescScalars* __restrict__ sc_,
            __half* __restrict__ packed, float* __restrict__ norm2) {
    constexpr int KP = D + 16;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    const float sc = sc_->scale;
    const int tile = row / DT_TILE, r = row % DT_TILE;
    __half* base = packed + (size_t)tile * DT_TILE * KP;
    double nn = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < D; k0 += 32) {
        const int k = k0 + lane;
        float v = 0.f;
        if (row < n) v = __ldg(x + (size_t)row * D + k) * sc;
        nn += (double)v * (double)v;
        // the query side is stored negated: accumulator = 1/2||b||^2 - a.b
        base[((size_t)(k >> 3) * DT_TILE + r) * 8 + (k & 7)] = __float2half_rn(is_ref ? v : -v);
    }
    nn = warp_sum(nn);
    if (lane < 16) {
        float v = 0.f;
        if (is_ref) {
            if (row < n) {
                const float h = (float)(0.5 * nn);
                const float hi = __half2float(__float2half_rn(h));
                const float mid = __half2float(__float2half_rn(h - hi));
                const float lo = __half2float(__float2half_rn(h - hi - mid));
                v = lane == 0 ? hi : (lane == 1 ? mid : (lane == 2 ? lo : 0.f));
            } else {
                v = lane == 0 ? DT_PAD_BIAS : 0.f;
            }
        } else {
            v = lane < 3 ? 1.f : 0.f;
        }
        const int k = D + lane;
        base[((size_t)(k >> 3) * DT_TILE + r) * 8 + (k & 7)] = __float2half_rn(v);
    }
    if (lane == 0) {
        const float nf = (float)nn;
        if (norm2) norm2[row] = nf;
        if (is_ref && row < n) {
            atomicMax(&sc_->bmax_bits, __float_as_uint(sqrtf(nf) * 1.0000002f));
            atomicMin(&sc_->bmin2_bits, __float_as_uint(nf * 0.9999998f));
            atomicMax(&sc_->bmax2_bits, __float_as_uint(nf * 1.0000002f));
        }
    }
}

// ---------------------------------------------------------------------------------------------
struct DescTcShared {
    unsigned long long full[DT_STAGES], empty[DT_STAGES], tfull[2], tempty[2], afull, aempty;
    uint32_t tmem_base;
    uint32_t pad;
};

// Slow path of the epilogue (per lane, rare): record one (score, column) candidate of this lane's row.
// Per-row state lives in shared memory: cv/ci [DT_CAND][DT_RB] slots, s_min[DT_RB], s_cnt[DT_RB] (low 16
// bits = used slots, bit 16 = overflow).  Returns the row's new threshold.
__device__ __noinline__ float desc_cand_insert(float v, int j, float margin2, float thr, float* cv, int* ci,
                                               float* s_min, int* s_cnt) {
    int cnt = *s_cnt;
    if (v < *s_min) { *s_min = v; thr = v + margin2; }
    const int n = cnt & 0xffff;
    int slot = -1;
    for (int s = 0; s < n; ++s)
        if (cv[s * DT_RB] > thr) { slot = s; break; }       // stale: the running minimum has dropped since
    if (slot < 0 && n < DT_CAND) { slot = n; cnt += 1; }
    if (slot >= 0) {
        cv[slot * DT_RB] = v;
        ci[slot * DT_RB] = j;
    } else {
        cnt |= 0x10000;                                     // overflow: resolved by k_desc_exact
    }
    *s_cnt = cnt;
    return thr;
}

// min of 8 consecutive columns, then of the chunk; the slow path (a column within the margin of the row's running
// minimum: ~ln(M) times per row over the whole stream, i.e. in ~5 % of a warp's chunks) only scans the groups whose
// minimum passed -- its instruction count, not its frequency, was a third of this kernel's time at D = 32
#define DESC_G8(r, o) fminf(fmin3(fmin3(r[o], r[o + 1], r[o + 2]), fmin3(r[o + 3], r[o + 4], r[o + 5]), r[o + 6]), r[o + 7])
#define DESC_SCAN8(r, o, g, c)                                                                                   \
    if (g <= thr) {                                                                                              \
        _Pragma("unroll") for (int k = 0; k < 8; ++k)                                                            \
            if (r[o + k] <= thr)                                                                                 \
                thr = desc_cand_insert(r[o + k], jh + (c) * 32 + o + k, margin2, thr, my_cv, my_ci, my_min, my_cnt); \
    }
#define DESC_CHUNK(r, c)                                                                                         \
    {                                                                                                            \
        const float g0 = DESC_G8(r, 0), g1 = DESC_G8(r, 8), g2 = DESC_G8(r, 16), g3 = DESC_G8(r, 24);            \
        const float m = fminf(fmin3(g0, g1, g2), g3);                                                            \
        if (DBG == 3) { thr = fminf(thr, m + margin2); }                                                         \
        else if (m <= thr) {                                                                                     \
            DESC_SCAN8(r, 0, g0, c) DESC_SCAN8(r, 8, g1, c) DESC_SCAN8(r, 16, g2, c) DESC_SCAN8(r, 24, g3, c)    \
        }                                                                                                        \
    }

template <int D, int DBG, int BIAS>
__global__ void __launch_bounds__(DT_THREADS, 1)
k_desc_nn_tc(const __half* __restrict__ a_packed, const __half* __restrict__ b_packed, int N, int M, int n_rb, int n_bt,
             const float* __restrict__ a_norm2, const DescScalars* __restrict__ scal,
             int32_t* __restrict__ cand_idx, float* __restrict__ cand_val, int32_t* __restrict__ cand_cnt,
             float* __restrict__ cand_thr) {
    // two variants are launched back to back; the one k_desc_choose did not pick leaves at once
    if ((scal->use_nobias != 0) == (BIAS != 0)) return;
    constexpr int KP = BIAS ? D + 16 : D;                         // K columns this variant multiplies
    constexpr int KSTEPS = KP / 16;
    constexpr uint32_t TILE_BYTES = DT_TILE * KP * 2;            // bytes staged per tile (the leading K-chunks)
    constexpr uint32_t TILE_STRIDE = DT_TILE * (D + 16) * 2;     // bytes per packed tile in HBM (always with bias chunks)
    constexpr uint32_t LBO = DT_TILE * 16;   // K-chunk stride
    constexpr uint32_t SBO = 128;            // 8-row group stride
    extern __shared__ __align__(1024) unsigned char dsm[];
    unsigned char* a_sm = dsm;                                   // 2 tiles
    unsigned char* b_sm = dsm + 2 * TILE_BYTES;                  // DT_STAGES tiles
    float* cv = reinterpret_cast<float*>(b_sm + DT_STAGES * TILE_BYTES);     // [DT_HALVES][DT_CAND][DT_RB]
    int* ci = reinterpret_cast<int*>(cv + DT_HALVES * DT_CAND * DT_RB);        // [DT_HALVES][DT_CAND][DT_RB]
    float* s_min = reinterpret_cast<float*>(ci + DT_HALVES * DT_CAND * DT_RB);  // [DT_HALVES][DT_RB]
    int* s_cnt = reinterpret_cast<int*>(s_min + DT_HALVES * DT_RB);              // [DT_HALVES][DT_RB]
    DescTcShared* sh = reinterpret_cast<DescTcShared*>(s_cnt + DT_HALVES * DT_RB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DT_STAGES; ++s) { mbar_init(smem_u32(&sh->full[s]), 1); mbar_init(smem_u32(&sh->empty[s]), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&sh->tfull[b]), 1); mbar_init(smem_u32(&sh->tempty[b]), DT_EPI_WARPS); }
        mbar_init(smem_u32(&sh->afull), 1);
        mbar_init(smem_u32(&sh->aempty), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sh->tmem_base;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            uint32_t it = 0, nblk = 0;
            for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x, ++nblk) {
                mbar_wait(smem_u32(&sh->aempty), (nblk & 1u) ^ 1u);
                mbar_expect_tx(smem_u32(&sh->afull), 2 * TILE_BYTES);
                const unsigned char* asrc = reinterpret_cast<const unsigned char*>(a_packed) + (size_t)rb * 2 * TILE_STRIDE;
                bulk_g2s(smem_u32(a_sm), asrc, TILE_BYTES, smem_u32(&sh->afull));
                bulk_g2s(smem_u32(a_sm + TILE_BYTES), asrc + TILE_STRIDE, TILE_BYTES, smem_u32(&sh->afull));
                for (int t = 0; t < n_bt; ++t, ++it) {
                    const uint32_t s = it % DT_STAGES, ph = (it / DT_STAGES) & 1u;
                    mbar_wait(smem_u32(&sh->empty[s]), ph ^ 1u);
                    mbar_expect_tx(smem_u32(&sh->full[s]), TILE_BYTES);
                    bulk_g2s(smem_u32(b_sm + s * TILE_BYTES),
                             reinterpret_cast<const unsigned char*>(b_packed) + (size_t)t * TILE_STRIDE, TILE_BYTES,
                             smem_u32(&sh->full[s]));
                }
            }
            // the last tcgen05.commit arrives asynchronously: do not let the CTA retire before it landed
            if (nblk > 0) mbar_wait(smem_u32(&sh->aempty), (nblk & 1u) ^ 1u);
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (uniform control flow and address arithmetic stay in the
        // uniform datapath), one elected lane issues.  The shared-memory descriptors are base + constant: the 14-bit
        // address field (bytes >> 4) of a tile never carries into the LBO field (all tiles lie below 256 KB). =====
        const uint32_t idesc = instr_desc_f16_128x128();
        const uint64_t a_desc0 = smem_desc(smem_u32(a_sm), LBO, SBO);
        const uint64_t b_desc0 = smem_desc(smem_u32(b_sm), LBO, SBO);
        constexpr uint32_t TILE_U = TILE_BYTES >> 4, KSTEP_U = (2 * LBO) >> 4;
        uint32_t it = 0, nblk = 0, tc = 0;
        for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x, ++nblk) {
            mbar_wait(smem_u32(&sh->afull), nblk & 1u);
            tc_fence_after();
            for (int t = 0; t < n_bt; ++t, ++it, ++tc) {
                const uint32_t s = it % DT_STAGES, ph = (it / DT_STAGES) & 1u;
                const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
                mbar_wait(smem_u32(&sh->tempty[buf]), bph ^ 1u);
                mbar_wait(smem_u32(&sh->full[s]), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t b_desc = b_desc0 + (uint64_t)(s * TILE_U);
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb) {
                        const uint32_t d_addr = tmem + buf * 256u + mb * 128u;
#pragma unroll
                        for (int kk = 0; kk < KSTEPS; ++kk)
                            tc_mma_f16(d_addr, a_desc0 + (uint64_t)(mb * TILE_U + kk * KSTEP_U), b_desc + (uint64_t)(kk * KSTEP_U),
                                       idesc, kk > 0 ? 1u : 0u);
                    }
                    tc_commit(smem_u32(&sh->empty[s]));      // smem stage reusable once these MMAs retire
                    tc_commit(smem_u32(&sh->tfull[buf]));    // accumulators ready for the epilogue
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(smem_u32(&sh->aempty));            // a tiles reusable
            __syncwarp();
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32).  16 warps: every (row sub-block,
        // lane quarter) is served by two warps, one per 64-column half of the accumulator tile, each with its own
        // candidate list -- the per-warp dependency chain (TMEM load -> min tree) per tile is what bounds this kernel
        // at D = 32, so it is kept short and spread over many warps =====
        const int ew = warp - 2;
        const int quarter = warp & 3, mb = (ew >> 2) & 1, half = ew >> 3;
        const int rl = mb * 128 + quarter * 32 + lane;           // row inside the row block
        float* my_cv = cv + half * DT_CAND * DT_RB + rl;
        int* my_ci = ci + half * DT_CAND * DT_RB + rl;
        float* my_min = s_min + half * DT_RB + rl;
        int* my_cnt = s_cnt + half * DT_RB + rl;
        const float bmax = __uint_as_float(scal->bmax_bits);
        const float spread = BIAS ? 0.f : scal->norm_spread;
        uint32_t tc = 0;
        for (int rb = blockIdx.x; rb < n_rb; rb += gridDim.x) {
            const size_t row = (size_t)rb * DT_RB + rl;
            const float na = sqrtf(a_norm2[row]);
            // |score' - score| <= E  (fp16 rounding of both operands, fp32 accumulation of <= 80 terms,
            // fp16 subnormal flush, and without the bias columns the spread of 1/2||b||^2); the true minimum is
            // within 2E of the approximate minimum.  Padding rows record nothing after their first element.
            const float E = 9.785e-4f * na * bmax + 3.06e-5f * (na * bmax + 0.5f * bmax * bmax) + 8e-6f + spread;
            const float margin2 = row < (size_t)N ? 2.f * E : -INFINITY;
            float thr = INFINITY;
            *my_min = INFINITY;
            *my_cnt = 0;
            for (int t = 0; t < n_bt; ++t, ++tc) {
                const uint32_t buf = tc & 1u, bph = (tc >> 1) & 1u;
                mbar_wait(smem_u32(&sh->tfull[buf]), bph);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256u + mb * 128u + half * 64u;
                // this warp's 64 columns into registers, then hand the accumulator buffer back to the MMA warp
                // BEFORE the min trees: the tensor pipe never waits for the CUDA-core work
                float rA[32], rB[32];
                if (DBG != 2) {
                    TMEM_LD32(rA, taddr);
                    TMEM_LD32(rB, taddr + 32);
                    TMEM_WAIT32(rA);
                    TMEM_WAIT32(rB);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&sh->tempty[buf]));
                const int jh = t * DT_TILE + half * 64;
                if (!BIAS && t == n_bt - 1) {
                    // without the bias columns a padding column scores 0 and could pass for a minimum: mask it
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        if (jh + k >= M) rA[k] = INFINITY;
                        if (jh + 32 + k >= M) rB[k] = INFINITY;
                    }
                }
                if (DBG == 0 || DBG == 3) {
                    DESC_CHUNK(rA, 0)
                    DESC_CHUNK(rB, 1)
                } else if (DBG == 1) {          // experiment: TMEM reads only
                    thr = fminf(thr, rA[0] + rB[0]);
                }
            }
            // row block done: this thread owns the candidates of its (row, column half)
            const int cnt = *my_cnt;
            const int n = cnt & 0xffff;
            const size_t slot = row * DT_HALVES + half;
            for (int s = 0; s < DT_CAND; ++s) {
                cand_idx[slot * DT_CAND + s] = s < n ? my_ci[s * DT_RB] : -1;
                cand_val[slot * DT_CAND + s] = s < n ? my_cv[s * DT_RB] : INFINITY;
            }
            cand_cnt[slot] = cnt;
            cand_thr[slot] = thr;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}


// ---------------------------------------------------------------------------------------------
// fp64 re-rank of the candidates: warp per row
template <int D>
__global__ void __launch_bounds__(256)
k_desc_rerank(const float* __restrict__ a, const float* __restrict__ b, int N, int M,
              const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_val,
              const int32_t* __restrict__ cand_cnt, const float* __restrict__ cand_thr,
              int32_t* __restrict__ out_idx, float* __restrict__ out_d2, int32_t* __restrict__ ovf_list,
              DescScalars* __restrict__ scal, uint8_t* __restrict__ tie, double tie_eps) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= N) return;
    const int cnt0 = cand_cnt[(size_t)i * DT_HALVES], cnt1 = cand_cnt[(size_t)i * DT_HALVES + 1];
    if ((cnt0 | cnt1) & 0x10000) {
        if (lane == 0) ovf_list[atomicAdd(&scal->n_overflow, 1)] = i;
        return;
    }
    // the two column halves kept candidates within the margin of THEIR minimum: the row's threshold is the lower one
    const float thr = fminf(cand_thr[(size_t)i * DT_HALVES], cand_thr[(size_t)i * DT_HALVES + 1]);
    float av[D / 32];
#pragma unroll
    for (int u = 0; u < D / 32; ++u) av[u] = __ldg(a + (size_t)i * D + u * 32 + lane);
    double best = INFINITY, second = INFINITY;
    int bj = -1;
    for (int s = 0; s < DT_HALVES * DT_CAND; ++s) {
        const int h = s / DT_CAND;
        if ((s % DT_CAND) >= ((h ? cnt1 : cnt0) & 0xffff)) continue;
        const int j = cand_idx[(size_t)i * DT_HALVES * DT_CAND + s];
        if (j < 0 || j >= M || !(cand_val[(size_t)i * DT_HALVES * DT_CAND + s] <= thr)) continue;
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < D / 32; ++u) {
            const double d = (double)av[u] - (double)__ldg(b + (size_t)j * D + u * 32 + lane);
            acc += d * d;
        }
        acc = warp_sum(acc);
        if (acc < best || (acc == best && j < bj)) { second = best; best = acc; bj = j; }
        else if (acc < second) second = acc;
    }
    if (lane == 0) {
        out_idx[i] = bj;
        out_d2[i] = (float)best;
        // every row within tie_eps of the minimum is a candidate (the margin is >= 2e-3 in score units), so the runner-up
        // among the candidates decides the flag
        if (tie) tie[i] = (bj >= 0 && second - best <= tie_eps) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 brute force: CTA = 256 threads, EX_ROWS query rows staged in shared memory, each thread walks
// reference rows j = tid, tid+256, ... keeping the best (d2, j) of every staged query; first minimal
// index wins.  Optional xyz gate (base.py:2969: pairs farther apart than max_mag are excluded).
#define EX_ROWS 16
#define EX_THREADS 256

template <int D>
__global__ void __launch_bounds__(EX_THREADS)
k_desc_exact(const float* __restrict__ a, const float* __restrict__ b, int N, int M,
             const int32_t* __restrict__ rows, const int* __restrict__ n_rows_dev,
             const float* __restrict__ a_xyz, const float* __restrict__ b_xyz, double max_mag2,
             int32_t* __restrict__ out_idx, float* __restrict__ out_d2, uint8_t* __restrict__ tie, double tie_eps) {
    __shared__ float qa[EX_ROWS][D];
    __shared__ float qx[EX_ROWS][3];
    __shared__ int qrow[EX_ROWS];
    __shared__ double red_d[EX_THREADS / 32][EX_ROWS];
    __shared__ double red_s[EX_THREADS / 32][EX_ROWS];
    __shared__ int red_j[EX_THREADS / 32][EX_ROWS];
    const int n_rows = rows ? *n_rows_dev : N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool gate = a_xyz != nullptr;
    for (int r0 = blockIdx.x * EX_ROWS; r0 < n_rows; r0 += gridDim.x * EX_ROWS) {
        __syncthreads();
        if (tid < EX_ROWS) {
            const int rr = r0 + tid;
            const int row = rr < n_rows ? (rows ? rows[rr] : rr) : -1;
            qrow[tid] = row;
            if (gate && row >= 0) { qx[tid][0] = a_xyz[3 * (size_t)row]; qx[tid][1] = a_xyz[3 * (size_t)row + 1]; qx[tid][2] = a_xyz[3 * (size_t)row + 2]; }
        }
        __syncthreads();
        for (int e = tid; e < EX_ROWS * D; e += EX_THREADS) {
            const int q = e / D, k = e % D;
            qa[q][k] = qrow[q] >= 0 ? __ldg(a + (size_t)qrow[q] * D + k) : 0.f;
        }
        __syncthreads();
        double best[EX_ROWS], sec[EX_ROWS];                 // sec: runner-up distance (tie flag)
        int bj[EX_ROWS];
#pragma unroll
        for (int q = 0; q < EX_ROWS; ++q) { best[q] = INFINITY; sec[q] = INFINITY; bj[q] = -1; }
        for (int j = tid; j < M; j += EX_THREADS) {
            float bv[D];
            const float4* bp = reinterpret_cast<const float4*>(b + (size_t)j * D);
#pragma unroll
            for (int k = 0; k < D / 4; ++k) {
                const float4 v = __ldg(bp + k);
                bv[4 * k] = v.x; bv[4 * k + 1] = v.y; bv[4 * k + 2] = v.z; bv[4 * k + 3] = v.w;
            }
            float bx = 0.f, by = 0.f, bz = 0.f;
            if (gate) { bx = __ldg(b_xyz + 3 * (size_t)j); by = __ldg(b_xyz + 3 * (size_t)j + 1); bz = __ldg(b_xyz + 3 * (size_t)j + 2); }
#pragma unroll
            for (int q = 0; q < EX_ROWS; ++q) {
                if (gate) {
                    const double dx = (double)qx[q][0] - (double)bx, dy = (double)qx[q][1] - (double)by,
                                 dz = (double)qx[q][2] - (double)bz;
                    if (dx * dx + dy * dy + dz * dz > max_mag2) continue;
                }
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double d = (double)qa[q][k] - (double)bv[k];
                    acc = fma(d, d, acc);
                }
                if (acc < best[q]) { sec[q] = best[q]; best[q] = acc; bj[q] = j; }
                else if (acc < sec[q]) sec[q] = acc;
            }
        }
        // block argmin per query (lower index on equal distance; -1 = nothing passed the gate)
#pragma unroll
        for (int q = 0; q < EX_ROWS; ++q) {
            double d = best[q], sd = sec[q];
            int j = bj[q] < 0 ? 0x7fffffff : bj[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(F4L_FULL, d, o);
                const double os = __shfl_xor_sync(F4L_FULL, sd, o);
                const int oj = __shfl_xor_sync(F4L_FULL, j, o);
                sd = fmin(fmin(sd, os), fmax(d, od));       // runner-up of the union
                if (od < d || (od == d && oj < j)) { d = od; j = oj; }
            }
            if (lane == 0) { red_d[wid][q] = d; red_s[wid][q] = sd; red_j[wid][q] = j; }
        }
        __syncthreads();
        if (tid < EX_ROWS && qrow[tid] >= 0) {
            double d = red_d[0][tid], sd = red_s[0][tid];
            int j = red_j[0][tid];
            for (int w = 1; w < EX_THREADS / 32; ++w) {
                const double od = red_d[w][tid];
                const int oj = red_j[w][tid];
                sd = fmin(fmin(sd, red_s[w][tid]), fmax(d, od));
                if (od < d || (od == d && oj < j)) { d = od; j = oj; }
            }
            out_idx[qrow[tid]] = j == 0x7fffffff ? -1 : j;
            out_d2[qrow[tid]] = (float)d;
            if (tie) tie[qrow[tid]] = (j != 0x7fffffff && sd - d <= tie_eps) ? 1 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// (b)+(c) scatter of the global 3D matches, base.py:2872-2889:
//   keep[i] = ||src_sub[i] - tgt_sub[labels[i]]|| <= max_magnitude
//   corres[:,0] = arange, corres[:,1] = -1;  corres[voxel2pts_src[keep], 1] = voxel2pts_tgt[labels[keep]]
// Duplicate raw targets (several voxels -> same raw point, quirk q5) are last-writer-wins and
// non-deterministic in the reference on CUDA; here the LARGEST voxel index wins (atomicMax on the voxel
// id, then a resolve pass), which is what a sequential scatter gives.
__global__ void __launch_bounds__(256) k_scatter_init(int64_t* __restrict__ corres, int32_t* __restrict__ winner, int n_raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_raw) return;
    corres[2 * (size_t)i] = i;
    corres[2 * (size_t)i + 1] = -1;
    winner[i] = -1;
}
__global__ void __launch_bounds__(256)
k_scatter_vote(const int32_t* __restrict__ labels, const float* __restrict__ src_sub, const float* __restrict__ tgt_sub,
               int n_sub, const int64_t* __restrict__ v2p_src, float max_mag, int32_t* __restrict__ winner, int n_raw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sub) return;
    const int l = labels[i];
    if (l < 0) return;
    const float dx = src_sub[3 * (size_t)i] - tgt_sub[3 * (size_t)l], dy = src_sub[3 * (size_t)i + 1] - tgt_sub[3 * (size_t)l + 1],
                dz = src_sub[3 * (size_t)i + 2] - tgt_sub[3 * (size_t)l + 2];
    // torch.norm of the f32 difference (base.py:2875)
    const float mag = sqrtf(dx * dx + dy * dy + dz * dz);
    if (!(mag <= max_mag)) return;
    const long long p = v2p_src[i];
    if (p >= 0 && p < n_raw) atomicMax(&winner[p], i);
}
__global__ void __launch_bounds__(256)
k_scatter_resolve(const int32_t* __restrict__ labels, const int64_t* __restrict__ v2p_tgt, const int32_t* __restrict__ winner,
                  int64_t* __restrict__ corres, int n_raw) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_raw) return;
    const int i = winner[p];
    if (i >= 0) corres[2 * (size_t)p + 1] = v2p_tgt[labels[i]];
}

// ---------------------------------------------------------------------------------------------
static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct DescWs {
    __half *a_packed, *b_packed;
    float *a_norm2, *cand_val, *cand_thr;
    int32_t *cand_idx, *cand_cnt, *ovf_list;
    DescScalars* scal;
    size_t total;
};

static DescWs desc_layout(void* base, int N, int M, int D) {
    DescWs w;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p + off; off += al256(bytes); return q; };
    const size_t KP = (size_t)D + 16;
    const size_t n_pad = ((size_t)N + DT_RB - 1) / DT_RB * DT_RB;
    const size_t m_pad = ((size_t)M + DT_TILE - 1) / DT_TILE * DT_TILE;
    w.scal = (DescScalars*)take(sizeof(DescScalars));
    w.a_packed = (__half*)take(n_pad * KP * 2);
    w.b_packed = (__half*)take(m_pad * KP * 2);
    w.a_norm2 = (float*)take(n_pad * 4);
    w.cand_idx = (int32_t*)take(n_pad * DT_HALVES * DT_CAND * 4);
    w.cand_val = (float*)take(n_pad * DT_HALVES * DT_CAND * 4);
    w.cand_cnt = (int32_t*)take(n_pad * DT_HALVES * 4);
    w.cand_thr = (float*)take(n_pad * DT_HALVES * 4);
    w.ovf_list = (int32_t*)take((size_t)N * 4);
    w.total = off;
    return w;
}

extern "C" size_t f4l_desc_nn_workspace_bytes(int32_t N, int32_t M, int32_t D, int both_dirs) {
    if (N < 0 || M < 0) return 0;
    size_t a = desc_layout(nullptr, N, M, D).total;
    if (both_dirs) {
        const size_t b = desc_layout(nullptr, M, N, D).total;
        a = a > b ? a : b;
    }
    return a + 256;
}

template <int D>
static int desc_one_direction(const float* a, int N, const float* b, int M, const float* a_xyz, const float* b_xyz,
                              float max_mag, int algo, int32_t* idx, float* d2, void* ws_base, cudaStream_t st,
                              uint8_t* tie, double tie_eps) {
    const bool gate = a_xyz && b_xyz && max_mag > 0.f;
    const double mm2 = (double)max_mag * (double)max_mag;
    bool use_tc = algo == F4L_DESC_TENSOR;
    if (algo == F4L_DESC_AUTO) use_tc = !gate && (double)N * (double)M >= 134217728.0;   // 2^27 pairs
    if (gate && use_tc) {
        f4l_set_error("f4l_desc_nn: the xyz gate is only implemented by the exact fp64 kernel (algo 0 or 2)");
        return F4L_E_ARG;
    }
    if (!use_tc) {
        const int grid = f4l_div_up(N, EX_ROWS) < 148 * 8 ? f4l_div_up(N, EX_ROWS) : 148 * 8;
        f4l_mark("k_desc_exact", st);
        k_desc_exact<D><<<grid, EX_THREADS, 0, st>>>(a, b, N, M, nullptr, nullptr, gate ? a_xyz : nullptr,
                                                      gate ? b_xyz : nullptr, mm2, idx, d2, tie, tie_eps);
        return F4L_OK;
    }
    DescWs w = desc_layout(ws_base, N, M, D);
    const int n_pad = f4l_div_up(N, DT_RB) * DT_RB, m_pad = f4l_div_up(M, DT_TILE) * DT_TILE;
    f4l_mark("k_desc_prep", st);
    k_desc_scalars_init<<<1, 1, 0, st>>>(w.scal);
    k_desc_absmax<<<148 * 4, 256, 0, st>>>(a, (size_t)N * D, w.scal);
    k_desc_absmax<<<148 * 4, 256, 0, st>>>(b, (size_t)M * D, w.scal);
    k_desc_scale<<<1, 1, 0, st>>>(w.scal);
    f4l_count_launches(3);
    f4l_mark("k_desc_pack", st);
    k_desc_pack<D><<<f4l_div_up(n_pad, 8), 256, 0, st>>>(a, N, n_pad, 0, w.scal, w.a_packed, w.a_norm2);
    k_desc_pack<D><<<f4l_div_up(m_pad, 8), 256, 0, st>>>(b, M, m_pad, 1, w.scal, w.b_packed, nullptr);
    static int force_bias = -1;
    if (force_bias < 0) { const char* e = getenv("F4L_DESC_FORCE_BIAS"); force_bias = e ? atoi(e) : 0; }
    k_desc_choose<<<1, 1, 0, st>>>(w.scal, force_bias);
    f4l_count_launches(2);
    constexpr int KP = D + 16;
    const size_t smem = (size_t)(2 + DT_STAGES) * DT_TILE * KP * 2 + (size_t)DT_HALVES * (DT_CAND * DT_RB * 8 + DT_RB * 8) + sizeof(DescTcShared) + 16;
    const int n_rb = n_pad / DT_RB, n_bt = m_pad / DT_TILE;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_rb < sms ? n_rb : sms;
    // F4L_DESC_DBG=1|2 selects an epilogue-ablation build of the kernel (profiling experiments only:
    // the results are garbage); unset = the real kernel
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("F4L_DESC_DBG"); dbg = e ? atoi(e) : 0; }
    f4l_mark("k_desc_nn_tc", st);
#define F4L_LAUNCH_TC(DBGV)                                                                                          \
    {                                                                                                                \
        cudaFuncSetAttribute(k_desc_nn_tc<D, DBGV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        cudaFuncSetAttribute(k_desc_nn_tc<D, DBGV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        k_desc_nn_tc<D, DBGV, 0><<<grid, DT_THREADS, smem, st>>>(w.a_packed, w.b_packed, N, M, n_rb, n_bt, w.a_norm2, w.scal, \
                                                                w.cand_idx, w.cand_val, w.cand_cnt, w.cand_thr);     \
        k_desc_nn_tc<D, DBGV, 1><<<grid, DT_THREADS, smem, st>>>(w.a_packed, w.b_packed, N, M, n_rb, n_bt, w.a_norm2, w.scal, \
                                                                w.cand_idx, w.cand_val, w.cand_cnt, w.cand_thr);     \
        f4l_count_launches(1);                                                                                       \
    }
    if (dbg == 1) F4L_LAUNCH_TC(1)
    else if (dbg == 2) F4L_LAUNCH_TC(2)
    else if (dbg == 3) F4L_LAUNCH_TC(3)
    else F4L_LAUNCH_TC(0)
#undef F4L_LAUNCH_TC
    f4l_mark("k_desc_rerank", st);
    k_desc_rerank<D><<<f4l_div_up(N, 8), 256, 0, st>>>(a, b, N, M, w.cand_idx, w.cand_val, w.cand_cnt, w.cand_thr, idx, d2,
                                                      w.ovf_list, w.scal, tie, tie_eps);
    f4l_mark("k_desc_exact", st);
    k_desc_exact<D><<<148, EX_THREADS, 0, st>>>(a, b, N, M, w.ovf_list, &w.scal->n_overflow, nullptr, nullptr, -1.0, idx, d2,
                                                tie, tie_eps);
    return F4L_OK;
}

extern "C" int f4l_desc_nn_ex(const float* a, int32_t N, const float* b, int32_t M, int32_t D, const float* a_xyz,
                              const float* b_xyz, float max_mag, int both_dirs, int algo, int32_t* row_idx, float* row_d2,
                              int32_t* col_idx, float* col_d2, uint8_t* row_tie, uint8_t* col_tie, double tie_eps,
                              void* workspace, size_t workspace_bytes, void* stream);

extern "C" int f4l_desc_nn(const float* a, int32_t N, const float* b, int32_t M, int32_t D, const float* a_xyz,
                           const float* b_xyz, float max_mag, int both_dirs, int algo, int32_t* row_idx, float* row_d2,
                           int32_t* col_idx, float* col_d2, void* workspace, size_t workspace_bytes, void* stream) {
    return f4l_desc_nn_ex(a, N, b, M, D, a_xyz, b_xyz, max_mag, both_dirs, algo, row_idx, row_d2, col_idx, col_d2, nullptr,
                          nullptr, 0.0, workspace, workspace_bytes, stream);
}

extern "C" int f4l_desc_nn_ex(const float* a, int32_t N, const float* b, int32_t M, int32_t D, const float* a_xyz,
                              const float* b_xyz, float max_mag, int both_dirs, int algo, int32_t* row_idx, float* row_d2,
                              int32_t* col_idx, float* col_d2, uint8_t* row_tie, uint8_t* col_tie, double tie_eps,
                              void* workspace, size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(tie_eps >= 0.0, "tie_eps < 0");
    F4L_REQUIRE(D == 32 || D == 64, "D must be 32 or 64");
    F4L_REQUIRE(N >= 0 && M >= 0, "negative size");
    F4L_REQUIRE(algo >= 0 && algo <= 2, "unknown algo");
    F4L_REQUIRE(row_idx && row_d2, "null output");
    F4L_REQUIRE(!both_dirs || (col_idx && col_d2), "col outputs are null");
    F4L_REQUIRE((a_xyz == nullptr) == (b_xyz == nullptr), "a_xyz and b_xyz must be given together");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0 && (!both_dirs || M == 0)) return F4L_OK;
    F4L_REQUIRE(a && b, "null input");
    const size_t need = f4l_desc_nn_workspace_bytes(N, M, D, both_dirs);
    if (algo != F4L_DESC_EXACT && (!workspace || workspace_bytes < need)) {
        f4l_set_error("f4l_desc_nn: workspace too small (%zu < %zu)", workspace_bytes, need);
        return F4L_E_WORKSPACE;
    }
    void* ws = (void*)(((uintptr_t)workspace + 255) / 256 * 256);
    int rc = F4L_OK;
    if (M == 0) {
        cudaMemsetAsync(row_idx, 0xff, (size_t)N * 4, st);
        cudaMemsetAsync(row_d2, 0x7f, (size_t)N * 4, st);   // NaN pattern-free: 0x7f7f7f7f = 3.39e38
        if (row_tie) cudaMemsetAsync(row_tie, 0, (size_t)N, st);
    } else if (N > 0) {
        rc = D == 32 ? desc_one_direction<32>(a, N, b, M, a_xyz, b_xyz, max_mag, algo, row_idx, row_d2, ws, st, row_tie, tie_eps)
                     : desc_one_direction<64>(a, N, b, M, a_xyz, b_xyz, max_mag, algo, row_idx, row_d2, ws, st, row_tie, tie_eps);
        if (rc != F4L_OK) return rc;
    }
    if (both_dirs && M > 0) {
        if (N == 0) {
            cudaMemsetAsync(col_idx, 0xff, (size_t)M * 4, st);
            cudaMemsetAsync(col_d2, 0x7f, (size_t)M * 4, st);
            if (col_tie) cudaMemsetAsync(col_tie, 0, (size_t)M, st);
        } else {
            rc = D == 32 ? desc_one_direction<32>(b, M, a, N, b_xyz, a_xyz, max_mag, algo, col_idx, col_d2, ws, st, col_tie, tie_eps)
                         : desc_one_direction<64>(b, M, a, N, b_xyz, a_xyz, max_mag, algo, col_idx, col_d2, ws, st, col_tie, tie_eps);
            if (rc != F4L_OK) return rc;
        }
    }
    return f4l_finish("f4l_desc_nn", stream);
}

extern "C" size_t f4l_scatter_global_matches_workspace_bytes(int32_t n_raw) { return al256((size_t)(n_raw < 0 ? 0 : n_raw) * 4); }

extern "C" int f4l_scatter_global_matches(const int32_t* labels, const float* src_sub, const float* tgt_sub, int32_t n_sub,
                                          const int64_t* voxel2pts_src, const int64_t* voxel2pts_tgt, float max_magnitude,
                                          int64_t* corres, int32_t n_raw, void* workspace, size_t workspace_bytes,
                                          void* stream) {
    F4L_REQUIRE(n_sub >= 0 && n_raw >= 0, "negative size");
    F4L_REQUIRE(corres || n_raw == 0, "null output");
    if (n_raw == 0) return F4L_OK;
    F4L_REQUIRE(n_sub == 0 || (labels && src_sub && tgt_sub && voxel2pts_src && voxel2pts_tgt), "null input");
    if (!workspace || workspace_bytes < (size_t)n_raw * 4) {
        f4l_set_error("f4l_scatter_global_matches: workspace too small");
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* winner = (int32_t*)workspace;
    f4l_mark("k_scatter_matches", st);
    k_scatter_init<<<f4l_div_up(n_raw, 256), 256, 0, st>>>(corres, winner, n_raw);
    if (n_sub > 0) {
        k_scatter_vote<<<f4l_div_up(n_sub, 256), 256, 0, st>>>(labels, src_sub, tgt_sub, n_sub, voxel2pts_src, max_magnitude,
                                                              winner, n_raw);
        k_scatter_resolve<<<f4l_div_up(n_raw, 256), 256, 0, st>>>(labels, voxel2pts_tgt, winner, corres, n_raw);
        f4l_count_launches(2);
    }
    return f4l_finish("f4l_scatter_global_matches", stream);
}
