// the per-segment pieces of the two small networks that sit inside the reference's per-patch
// loops, for all segments (supervoxels / superpoints, CSR) of a tile in one launch each.
//   k_seg_scale_maxabs   src/f2s3.py:343          rows of a supervoxel divided by its max |value|
//   k_seg_norm2_relu     src/models/outlier_classifier.py:15-23   InstanceNorm2d(eps) -> BatchNorm2d(eps, batch
//                        statistics, single-sample batch) -> ReLU (-> + residual, :29-33) per segment and channel
//   k_seg_attention_pool src/feature_aggregation/cluster_feature_net_self_attention.py:18-33, :91   mean over the
//                        points of a superpoint of softmax(Q K^T * scale) V (the fc layer commutes with the mean)
//   k_seg_mean           :97   per-superpoint centroid of the voxel coordinates
// The dense per-point products in between (1x1 convolutions = (K,C)x(C,C)) are plain library GEMMs on the host side.
#include <stdlib.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(256)
k_seg_scale_maxabs(const TI* __restrict__ x, const int32_t* __restrict__ ptr, int C, float* __restrict__ out) {
    __shared__ TI red[8];
    const int q = blockIdx.x;
    const size_t s = (size_t)ptr[q] * C, e = (size_t)ptr[q + 1] * C;
    TI m = 0;
    for (size_t i = s + threadIdx.x; i < e; i += blockDim.x) m = fmax(m, fabs(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(F4L_FULL, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
    // torch.divide in the input precision, then .float() (src/f2s3.py:343,346)
    for (size_t i = s + threadIdx.x; i < e; i += blockDim.x) out[i] = (float)(x[i] / m);
}

// blockDim = (CX channels, RY row lanes); channel c of the block = blockIdx.y * CX + threadIdx.x
#define NORM_CX 32
#define NORM_RY 8
__global__ void __launch_bounds__(NORM_CX * NORM_RY)
k_seg_norm2_relu(const float* __restrict__ y, const int32_t* __restrict__ ptr, int C, float eps,
                 const float* __restrict__ residual, float* __restrict__ out) {
    __shared__ float red[NORM_RY][NORM_CX];
    const int q = blockIdx.x, c = blockIdx.y * NORM_CX + threadIdx.x, ty = threadIdx.y;
    const int s = ptr[q], n = ptr[q + 1] - s;
    if (n <= 0) return;
    const bool live = c < C;
    const float* base = y + (size_t)s * C + (live ? c : 0);
    float acc = 0.f;
    if (live)
        for (int r = ty; r < n; r += NORM_RY) acc += base[(size_t)r * C];
    red[ty][threadIdx.x] = acc;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_RY; ++k) mean += red[k][threadIdx.x];
    mean /= (float)n;
    __syncthreads();
    acc = 0.f;
    if (live)
        for (int r = ty; r < n; r += NORM_RY) { const float d = base[(size_t)r * C] - mean; acc += d * d; }
    red[ty][threadIdx.x] = acc;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_RY; ++k) var += red[k][threadIdx.x];
    var /= (float)n;                                          // biased variance (both norm layers)
    // InstanceNorm: z = (x - mean) / sqrt(var + eps).  BatchNorm over the same single sample: mean(z) = 0,
    // var(z) = var / (var + eps)  ->  out = z / sqrt(var / (var + eps) + eps)
    const float inv1 = rsqrtf(var + eps);
    const float var2 = var * inv1 * inv1;
    const float scale = inv1 * rsqrtf(var2 + eps);
    if (live) {
        float* o = out + (size_t)s * C + c;
        const float* rs = residual ? residual + (size_t)s * C + c : nullptr;
        for (int r = ty; r < n; r += NORM_RY) {
            float v = fmaxf((base[(size_t)r * C] - mean) * scale, 0.f);
            if (rs) v += rs[(size_t)r * C];
            o[(size_t)r * C] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// One CTA of ATT_T threads per segment; a thread owns one query of the current round of ATT_T queries (q and the
// running output in registers), keys / values stream through shared memory in chunks of ATT_CH rows (every lane
// reads the same row: shared-memory broadcast).  Online softmax per query; the CTA accumulates sum_i o_i / l_i.
#define ATT_T 64
#define ATT_CH 64
template <int HD>
__global__ void __launch_bounds__(ATT_T)
k_seg_attention_pool(const float* __restrict__ Qm, const float* __restrict__ Km, const float* __restrict__ Vm,
                     const int32_t* __restrict__ ptr, float scale, float* __restrict__ out) {
    __shared__ __align__(16) float ks[ATT_CH][HD];
    __shared__ __align__(16) float vs[ATT_CH][HD];
    __shared__ float pool[ATT_T / 32][HD];
    const int seg = blockIdx.x, tid = threadIdx.x;
    const int s = ptr[seg], n = ptr[seg + 1] - s;
    float accum[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) accum[d] = 0.f;
    for (int r0 = 0; r0 < n; r0 += ATT_T) {
        const int qi = r0 + tid;
        const bool live = qi < n;
        float qv[HD], o[HD];
        {
            const float4* qp = reinterpret_cast<const float4*>(Qm + (size_t)(s + (live ? qi : 0)) * HD);
#pragma unroll
            for (int d = 0; d < HD / 4; ++d) {
                const float4 t = qp[d];
                qv[4 * d] = t.x * scale; qv[4 * d + 1] = t.y * scale; qv[4 * d + 2] = t.z * scale; qv[4 * d + 3] = t.w * scale;
            }
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = 0.f;
        float mx = -INFINITY, l = 0.f;
        for (int k0 = 0; k0 < n; k0 += ATT_CH) {
            const int kc = min(ATT_CH, n - k0);
            __syncthreads();
            for (int i = tid; i < kc * (HD / 4); i += ATT_T) {
                const int row = i / (HD / 4), col = i % (HD / 4);
                reinterpret_cast<float4*>(&ks[row][0])[col] = reinterpret_cast<const float4*>(Km + (size_t)(s + k0 + row) * HD)[col];
                reinterpret_cast<float4*>(&vs[row][0])[col] = reinterpret_cast<const float4*>(Vm + (size_t)(s + k0 + row) * HD)[col];
            }
            __syncthreads();
            for (int j = 0; j < kc; ++j) {
                float sdot = 0.f;
#pragma unroll
                for (int d = 0; d < HD / 4; ++d) {
                    const float4 kk = reinterpret_cast<const float4*>(&ks[j][0])[d];
                    sdot = fmaf(qv[4 * d], kk.x, sdot); sdot = fmaf(qv[4 * d + 1], kk.y, sdot);
                    sdot = fmaf(qv[4 * d + 2], kk.z, sdot); sdot = fmaf(qv[4 * d + 3], kk.w, sdot);
                }
                if (sdot > mx) {                       // rescale the running sums (rare after the first keys)
                    const float f = __expf(mx - sdot);
                    l *= f;
#pragma unroll
                    for (int d = 0; d < HD; ++d) o[d] *= f;
                    mx = sdot;
                }
                const float p = __expf(sdot - mx);
                l += p;
#pragma unroll
                for (int d = 0; d < HD / 4; ++d) {
                    const float4 vv = reinterpret_cast<const float4*>(&vs[j][0])[d];
                    o[4 * d] = fmaf(p, vv.x, o[4 * d]); o[4 * d + 1] = fmaf(p, vv.y, o[4 * d + 1]);
                    o[4 * d + 2] = fmaf(p, vv.z, o[4 * d + 2]); o[4 * d + 3] = fmaf(p, vv.w, o[4 * d + 3]);
                }
            }
        }
        if (live) {
            const float inv = 1.f / l;
#pragma unroll
            for (int d = 0; d < HD; ++d) accum[d] = fmaf(o[d], inv, accum[d]);
        }
    }
    // sum over the queries of the CTA, divide by n
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        float v = accum[d];
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(F4L_FULL, v, o2);
        if ((tid & 31) == 0) pool[tid >> 5][d] = v;
    }
    __syncthreads();
    for (int d = tid; d < HD; d += ATT_T) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < ATT_T / 32; ++w) v += pool[w][d];
        out[(size_t)seg * HD + d] = n > 0 ? v / (float)n : __int_as_float(0x7fc00000);   // torch.mean of an empty set: nan
    }
}


// ------------------------------------------------------------------------------------------------------------
// Tensor-core form of the pooling.  Two observations: (1) mean_i sum_j a_ij V_j = sum_j c_j V_j with c_j = (1/n) sum_i a_ij,
// so the (n x n) x d product "attention x V" is never formed -- only the scores Q K^T, n^2 exponentials and one (1 x n) x d
// product; (2) the scores are a dense contraction: mma.sync m16n8k8 TF32 with the 3xTF32 split (x = hi + lo, both TF32;
// Qh Kh + Qh Kl + Ql Kh) keeps fp32-level accuracy (relative error ~2^-21), which the mutual-NN decisions downstream
// need.  The flash-style CUDA-core kernel above is bound by its shared-memory broadcast loads (one LDS.128 per four FMAs:
// 29 % of the FMA peak); here a K chunk is read as mma fragments (conflict-free LDS.32 of pre-split hi / lo planes).
// One CTA (4 warps) per segment; a warp owns 16 query rows of a round of 64; the scores of the round live in shared
// memory (the softmax needs a row's sum before its entries can be normalised and column-summed).
#define ATM_T 256
#define ATM_CH 64
#define ATM_SCORE_FLOATS 38400            // at most 150 KB of scores: 64 rows x 600, 32 x 1200, 16 x 2400
#define ATM_CS_CAP 2400

__device__ __forceinline__ uint32_t tf32_of(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 8 warps: warp w owns the 16 query rows (w & 3) of the round and the 32-key half (w >> 2) of every staged chunk.
// score_floats: size of the score buffer of this launch (the host sizes it for the longest segment, so short segments --
// the lower hierarchy levels -- leave room for several CTAs per SM).
template <int HD>
__global__ void __launch_bounds__(ATM_T)
k_seg_attention_pool_mma(const float* __restrict__ Qm, const float* __restrict__ Km, const float* __restrict__ Vm,
                         const int32_t* __restrict__ ptr, float scale, int score_floats, int cs_cap,
                         float* __restrict__ out) {
    constexpr int KS = HD / 8;                            // k-steps of 8
    constexpr int KPAD = HD + 4;                          // row stride of the K planes: (g, t) -> 32 distinct banks
    extern __shared__ __align__(16) float att_sm[];
    float* S = att_sm;                                    // [rows of the round][stride]
    uint32_t* khi = reinterpret_cast<uint32_t*>(S + score_floats);          // [ATM_CH][KPAD] tf32 bits
    uint32_t* klo = khi + ATM_CH * KPAD;
    float* rowmax = reinterpret_cast<float*>(klo + ATM_CH * KPAD);          // [2][64] per key half
    float* inv = rowmax + 128;                            // [64]  1 / (n * row sum)
    float* cs = inv + 64;                                 // [cs_cap] column sums
    __shared__ float red[ATM_T / HD][HD];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rt = warp & 3, kh = warp >> 2;              // row tile, key half
    const int s = ptr[seg], n = ptr[seg + 1] - s;
    if (n <= 0) {
        for (int d = tid; d < HD; d += ATM_T) out[(size_t)seg * HD + d] = __int_as_float(0x7fc00000);   // mean of nothing
        return;
    }
    (void)cs_cap;
    scale *= 1.4426950408889634f;                         // scores in units of log2: softmax = exp2(s - max) / sum
    const int stride = ((n + 23) / 32) * 32 + 8;          // >= n, == 8 (mod 32): conflict-free float2 stores of the C frags
    const int R = score_floats / stride >= 64 ? 64 : (score_floats / stride >= 32 ? 32 : 16);
    for (int j = tid; j < n; j += ATM_T) cs[j] = 0.f;
    const int wrow = rt * 16;                             // first row of this warp inside the round
    const bool wactive = wrow < R;
    for (int r0 = 0; r0 < n; r0 += R) {
        // ---- query fragments of the warp's 16 rows, scaled, split into hi / lo TF32 -------------------------------
        uint32_t ah[KS][4], al[KS][4];
        if (wactive) {
            const int ra = r0 + wrow + g, rb = ra + 8;
            const float* qa = Qm + (size_t)(s + (ra < n ? ra : 0)) * HD;
            const float* qb = Qm + (size_t)(s + (rb < n ? rb : 0)) * HD;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const float v[4] = {ra < n ? qa[kk * 8 + t] * scale : 0.f, rb < n ? qb[kk * 8 + t] * scale : 0.f,
                                    ra < n ? qa[kk * 8 + t + 4] * scale : 0.f, rb < n ? qb[kk * 8 + t + 4] * scale : 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ah[kk][e] = tf32_of(v[e]);
                    al[kk][e] = tf32_of(v[e] - __uint_as_float(ah[kk][e]));
                }
            }
        }
        float mxa = -INFINITY, mxb = -INFINITY;           // running maxima of rows g and g + 8 (this lane's columns)
        for (int k0 = 0; k0 < n; k0 += ATM_CH) {
            const int kc = min(ATM_CH, n - k0);
            __syncthreads();
            for (int i = tid; i < ATM_CH * HD; i += ATM_T) {
                const int row = i / HD, col = i % HD;
                const float v = row < kc ? Km[(size_t)(s + k0 + row) * HD + col] : 0.f;
                const uint32_t h = tf32_of(v);
                khi[row * KPAD + col] = h;
                klo[row * KPAD + col] = tf32_of(v - __uint_as_float(h));
            }
            __syncthreads();
            if (wactive && kh * 32 < kc) {
                float acc[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
                for (int kk = 0; kk < KS; ++kk) {
                    uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int o = (kh * 32 + nt * 8 + g) * KPAD + kk * 8 + t;
                        bh0[nt] = khi[o]; bh1[nt] = khi[o + 4];
                        bl0[nt] = klo[o]; bl1[nt] = klo[o + 4];
                    }
                    // the small terms first; independent accumulators between dependent MMAs
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], al[kk], bh0[nt], bh1[nt]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ah[kk], bl0[nt], bl1[nt]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ah[kk], bh0[nt], bh1[nt]);
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int col = k0 + kh * 32 + nt * 8 + 2 * t;
                    if (col < n) {                       // (the chunk's zero-padded keys beyond n are not part of the row)
                        *reinterpret_cast<float2*>(S + (wrow + g) * stride + col) = make_float2(acc[nt][0], acc[nt][1]);
                        *reinterpret_cast<float2*>(S + (wrow + g + 8) * stride + col) = make_float2(acc[nt][2], acc[nt][3]);
                        mxa = fmaxf(mxa, acc[nt][0]); mxb = fmaxf(mxb, acc[nt][2]);
                    }
                    if (col + 1 < n) { mxa = fmaxf(mxa, acc[nt][1]); mxb = fmaxf(mxb, acc[nt][3]); }
                }
            }
        }
        if (wactive) {
            mxa = fmaxf(mxa, __shfl_xor_sync(F4L_FULL, mxa, 1)); mxa = fmaxf(mxa, __shfl_xor_sync(F4L_FULL, mxa, 2));
            mxb = fmaxf(mxb, __shfl_xor_sync(F4L_FULL, mxb, 1)); mxb = fmaxf(mxb, __shfl_xor_sync(F4L_FULL, mxb, 2));
            if (t == 0) { rowmax[kh * 64 + wrow + g] = mxa; rowmax[kh * 64 + wrow + g + 8] = mxb; }
        }
        __syncthreads();
        if (wactive) {
            // exponentials and row sums: the two warps of a row tile take 8 rows each
            for (int rr = kh * 8; rr < kh * 8 + 8; ++rr) {
                const int row = wrow + rr;
                const bool live = r0 + row < n;
                const float m = fmaxf(rowmax[row], rowmax[64 + row]);
                float l = 0.f;
                if (live) {
                    float* sr = S + row * stride;
                    float l1 = 0.f;
                    int j = lane;
                    for (; j + 32 < n; j += 64) {                    // two independent chains per lane
                        const float e0 = ex2_approx(sr[j] - m), e1 = ex2_approx(sr[j + 32] - m);
                        sr[j] = e0; sr[j + 32] = e1;
                        l += e0; l1 += e1;
                    }
                    if (j < n) { const float e0 = ex2_approx(sr[j] - m); sr[j] = e0; l += e0; }
                    l += l1;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(F4L_FULL, l, o);
                if (lane == 0) inv[row] = live ? 1.f / (l * (float)n) : 0.f;
            }
        }
        __syncthreads();
        const int rows = min(R, n - r0);
        for (int j = tid; j < n; j += ATM_T) {
            float c0 = 0.f, c1 = 0.f;
            int rr = 0;
            for (; rr + 1 < rows; rr += 2) {
                c0 = fmaf(S[rr * stride + j], inv[rr], c0);
                c1 = fmaf(S[(rr + 1) * stride + j], inv[rr + 1], c1);
            }
            if (rr < rows) c0 = fmaf(S[rr * stride + j], inv[rr], c0);
            cs[j] += c0 + c1;
        }
    }
    __syncthreads();
    {
        const int d = tid % HD, part = tid / HD;
        constexpr int PARTS = ATM_T / HD;
        float acc = 0.f;
        for (int j = part; j < n; j += PARTS) acc = fmaf(cs[j], Vm[(size_t)(s + j) * HD + d], acc);
        red[part][d] = acc;
        __syncthreads();
        if (part == 0) {
            float v = 0.f;
#pragma unroll
            for (int p = 0; p < PARTS; ++p) v += red[p][d];
            out[(size_t)seg * HD + d] = v;
        }
    }
}

__global__ void __launch_bounds__(128)
k_seg_mean(const float* __restrict__ x, const int32_t* __restrict__ ptr, int P, int C, float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P) return;
    const int s = ptr[warp], n = ptr[warp + 1] - s;
    for (int c = 0; c < C; ++c) {
        double a = 0.0;
        for (int r = lane; r < n; r += 32) a += (double)x[(size_t)(s + r) * C + c];
        a = warp_sum(a);
        if (lane == 0) out[(size_t)warp * C + c] = n > 0 ? (float)(a / n) : __int_as_float(0x7fc00000);
    }
}

// ------------------------------------------------------------------------------------------------------------
extern "C" int f4l_segment_scale_maxabs(const void* x, int32_t x_is_f64, const int32_t* seg_ptr, int32_t Q, int32_t C, float* out,
                                        void* stream) {
    F4L_REQUIRE(Q >= 0 && C > 0, "bad size");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(x && seg_ptr && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    f4l_mark("k_seg_scale_maxabs", st);
    if (x_is_f64) k_seg_scale_maxabs<double><<<Q, 256, 0, st>>>((const double*)x, seg_ptr, C, out);
    else k_seg_scale_maxabs<float><<<Q, 256, 0, st>>>((const float*)x, seg_ptr, C, out);
    return f4l_finish("f4l_segment_scale_maxabs", stream);
}

extern "C" int f4l_segment_norm2_relu(const float* y, const int32_t* seg_ptr, int32_t Q, int32_t C, float eps,
                                      const float* residual, float* out, void* stream) {
    F4L_REQUIRE(Q >= 0 && C > 0 && eps > 0.f, "bad size");
    if (Q == 0) return F4L_OK;
    F4L_REQUIRE(y && seg_ptr && out, "null pointer");
    F4L_REQUIRE(f4l_div_up(C, NORM_CX) <= 65535, "too many channels");
    cudaStream_t st = (cudaStream_t)stream;
    f4l_mark("k_seg_norm2_relu", st);
    k_seg_norm2_relu<<<dim3(Q, f4l_div_up(C, NORM_CX)), dim3(NORM_CX, NORM_RY), 0, st>>>(y, seg_ptr, C, eps, residual, out);
    return f4l_finish("f4l_segment_norm2_relu", stream);
}

extern "C" int f4l_segment_attention_pool(const float* Qm, const float* Km, const float* Vm, const int32_t* seg_ptr,
                                          int32_t P, int32_t hidden, float scale, int32_t max_seg_rows, float* out,
                                          void* stream) {
    F4L_REQUIRE(P >= 0, "bad size");
    F4L_REQUIRE(hidden == 32 || hidden == 64, "hidden dimension must be 32 or 64");
    if (P == 0) return F4L_OK;
    F4L_REQUIRE(Qm && Km && Vm && seg_ptr && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // tensor-core kernel when the longest segment is known and fits its score buffer (<= 2392 rows); max_seg_rows <= 0
    // ("unknown") or longer segments: the flash-style CUDA-core kernel.  F4L_ATT_FLASH=1 forces the latter (experiments).
    static int force_flash = -1;
    if (force_flash < 0) { const char* e = getenv("F4L_ATT_FLASH"); force_flash = e ? atoi(e) : 0; }
    const int max_stride = ((max_seg_rows + 23) / 32) * 32 + 8;
    if (!force_flash && max_seg_rows > 0 && max_stride <= ATM_CS_CAP) {
        // score buffer for 64 rows of the longest segment, capped at 150 KB (longer segments: 32- or 16-row rounds)
        const int score_floats = 64 * max_stride < ATM_SCORE_FLOATS ? 64 * max_stride : ATM_SCORE_FLOATS;
        const int cs_cap = max_stride;
        const size_t smem = ((size_t)score_floats + 2 * ATM_CH * (hidden + 4) + 192 + cs_cap) * sizeof(float);
        const size_t smem_max = ((size_t)ATM_SCORE_FLOATS + 2 * ATM_CH * (64 + 4) + 192 + ATM_CS_CAP) * sizeof(float);
        static F4lPerDevice once;
        if (!once.done()) {
            if (!f4l_optin_smem(k_seg_attention_pool_mma<64>, smem_max, "k_seg_attention_pool_mma<64>") ||
                !f4l_optin_smem(k_seg_attention_pool_mma<32>, smem_max, "k_seg_attention_pool_mma<32>"))
                return F4L_E_CUDA;
            once.mark();
        }
        f4l_mark("k_seg_attention_pool_mma", st);
        if (hidden == 64) k_seg_attention_pool_mma<64><<<P, ATM_T, smem, st>>>(Qm, Km, Vm, seg_ptr, scale, score_floats, cs_cap, out);
        else k_seg_attention_pool_mma<32><<<P, ATM_T, smem, st>>>(Qm, Km, Vm, seg_ptr, scale, score_floats, cs_cap, out);
        return f4l_finish("f4l_segment_attention_pool", stream);
    }
    f4l_mark("k_seg_attention_pool", st);
    if (hidden == 64) k_seg_attention_pool<64><<<P, ATT_T, 0, st>>>(Qm, Km, Vm, seg_ptr, scale, out);
    else k_seg_attention_pool<32><<<P, ATT_T, 0, st>>>(Qm, Km, Vm, seg_ptr, scale, out);
    return f4l_finish("f4l_segment_attention_pool", stream);
}

extern "C" int f4l_segment_mean(const float* x, const int32_t* seg_ptr, int32_t P, int32_t C, float* out, void* stream) {
    F4L_REQUIRE(P >= 0 && C > 0, "bad size");
    if (P == 0) return F4L_OK;
    F4L_REQUIRE(x && seg_ptr && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    f4l_mark("k_seg_mean", st);
    k_seg_mean<<<f4l_div_up((long long)P * 32, 128), 128, 0, st>>>(x, seg_ptr, P, C, out);
    return f4l_finish("f4l_segment_mean", stream);
}
