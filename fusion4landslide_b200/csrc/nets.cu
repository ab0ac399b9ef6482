// the per-segment pieces of the two small networks that sit inside the reference's per-patch
// loops, for all segments (supervoxels / superpoints, CSR) of a tile in one launch each.
//   k_seg_scale_maxabs   src/f2s3.py:343          rows of a supervoxel divided by its max |value|
//   k_seg_norm2_relu     src/models/outlier_classifier.py:15-23   InstanceNorm2d(eps) -> BatchNorm2d(eps, batch
//                        statistics, single-sample batch) -> ReLU (-> + residual, :29-33) per segment and channel
//   k_seg_attention_pool src/feature_aggregation/cluster_feature_net_self_attention.py:18-33, :91   mean over the
//                        points of a superpoint of softmax(Q K^T * scale) V (the fc layer commutes with the mean)
//   k_seg_mean           :97   per-superpoint centroid of the voxel coordinates
// The dense per-point products in between (1x1 convolutions = (K,C)x(C,C)) are plain library GEMMs on the host side.
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
                                          int32_t P, int32_t hidden, float scale, float* out, void* stream) {
    F4L_REQUIRE(P >= 0, "bad size");
    F4L_REQUIRE(hidden == 32 || hidden == 64, "hidden dimension must be 32 or 64");
    if (P == 0) return F4L_OK;
    F4L_REQUIRE(Qm && Km && Vm && seg_ptr && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
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
