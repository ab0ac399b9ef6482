// B4 and F1: the integer / elementwise parts of the coarse stage.
//   k_vote_tgt_patch   2D-vote coarse matching (base.py:3016-3035): for every source patch the target-patch
//                      label that most of its points' 2D-lifted matches fall into (torch.unique + argsort)
//   k_magnitude_mask   displacement-magnitude gates (base.py:2875-2876, :1635-1636; f2s3.py:392,419-424)
#include "common.cuh"

#define VOTE_SLOTS 512
#define VOTE_WARPS 4

// warp per source patch; open-addressing label->count table in shared memory.  Ties in the count are
// resolved towards the smallest label and flagged (the reference's argsort order is unspecified there);
// flag 255 = more than VOTE_SLOTS distinct labels (not resolved).
__global__ void __launch_bounds__(VOTE_WARPS * 32)
k_vote_tgt_patch(const int64_t* __restrict__ corr2d, const int32_t* __restrict__ sp_idx, const int32_t* __restrict__ sp_ptr,
                 int P, const int32_t* __restrict__ label_tgt, int n_tgt, const int32_t* __restrict__ label_to_local,
                 int n_labels, int32_t* __restrict__ best, int32_t* __restrict__ best_count, uint8_t* __restrict__ flag) {
    __shared__ int keys[VOTE_WARPS][VOTE_SLOTS];
    __shared__ int cnts[VOTE_WARPS][VOTE_SLOTS];
    __shared__ int ovf[VOTE_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int* K = keys[wid];
    int* C = cnts[wid];
    for (int p = blockIdx.x * VOTE_WARPS + wid; p < P; p += gridDim.x * VOTE_WARPS) {
        for (int s = lane; s < VOTE_SLOTS; s += 32) { K[s] = -1; C[s] = 0; }
        if (lane == 0) ovf[wid] = 0;
        __syncwarp();
        const int s0 = sp_ptr[p], s1 = sp_ptr[p + 1];
        for (int i = s0 + lane; i < s1; i += 32) {
            const long long t = corr2d[2 * (size_t)sp_idx[i] + 1];
            if (t < 0 || t >= n_tgt) continue;                       // -1 = no 2D match (base.py:3020)
            const int lab = label_tgt[t];
            unsigned h = ((unsigned)lab * 2654435761u) % VOTE_SLOTS;
            bool done = false;
            for (int probe = 0; probe < VOTE_SLOTS && !done; ++probe) {
                const int old = atomicCAS(&K[h], -1, lab);
                if (old == -1 || old == lab) { atomicAdd(&C[h], 1); done = true; }
                else h = (h + 1) % VOTE_SLOTS;
            }
            if (!done) ovf[wid] = 1;
        }
        __syncwarp();
        int bc = 0, bl = 0x7fffffff, nb = 0;
        for (int s = lane; s < VOTE_SLOTS; s += 32) {
            const int c = C[s], l = K[s];
            if (c > bc || (c == bc && c > 0 && l < bl)) { bc = c; bl = l; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const int oc = __shfl_xor_sync(F4L_FULL, bc, o), ol = __shfl_xor_sync(F4L_FULL, bl, o);
            if (oc > bc || (oc == bc && oc > 0 && ol < bl)) { bc = oc; bl = ol; }
        }
        for (int s = lane; s < VOTE_SLOTS; s += 32) nb += (bc > 0 && C[s] == bc) ? 1 : 0;
        nb = warp_sum(nb);
        if (lane == 0) {
            int out = -1;
            if (bc > 0) {
                out = bl;
                if (label_to_local) out = (bl >= 0 && bl < n_labels) ? label_to_local[bl] : -1;   // base.py:3062-3064
            }
            best[p] = out;
            best_count[p] = bc;
            flag[p] = ovf[wid] ? 255 : (nb > 1 ? 1 : 0);
        }
        __syncwarp();
    }
}

extern "C" int f4l_vote_tgt_patch(const int64_t* corr2d, const int32_t* sp_idx, const int32_t* sp_ptr, int32_t P,
                                  const int32_t* label_tgt, int32_t n_tgt, const int32_t* label_to_local, int32_t n_labels,
                                  int32_t* best, int32_t* best_count, uint8_t* flag, void* stream) {
    F4L_REQUIRE(P >= 0 && n_tgt >= 0, "negative size");
    if (P == 0) return F4L_OK;
    F4L_REQUIRE(corr2d && sp_idx && sp_ptr && label_tgt && best && best_count && flag, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = f4l_div_up(P, VOTE_WARPS) < 148 * 8 ? f4l_div_up(P, VOTE_WARPS) : 148 * 8;
    f4l_mark("k_vote_tgt_patch", st);
    k_vote_tgt_patch<<<grid, VOTE_WARPS * 32, 0, st>>>(corr2d, sp_idx, sp_ptr, P, label_tgt, n_tgt, label_to_local, n_labels,
                                                       best, best_count, flag);
    return f4l_finish("f4l_vote_tgt_patch", stream);
}

// rows [src | tgt] f32: mag = ||src - tgt|| in f32 (torch.norm of the f32 difference), mask = mag <= max
// (strict == 0) or mag < max (strict != 0).  max_mag may come from a device scalar scaled by `factor`
// (the 30 x median gate, f2s3.py:427-441).
__global__ void __launch_bounds__(256)
k_magnitude_mask(const float* __restrict__ rows, int K, int stride, float max_mag, const float* __restrict__ d_max, float factor,
                 int strict, float* __restrict__ mag, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float* r = rows + (size_t)i * stride;
    const float dx = r[0] - r[3], dy = r[1] - r[4], dz = r[2] - r[5];
    const float m = sqrtf(dx * dx + dy * dy + dz * dz);
    const float lim = d_max ? d_max[0] * factor : max_mag;
    if (mag) mag[i] = m;
    mask[i] = (strict ? (m < lim) : (m <= lim)) ? 1 : 0;
}

extern "C" int f4l_magnitude_mask(const float* rows, int32_t K, int32_t stride, float max_mag, const float* d_max, float factor,
                                  int strict, float* mag, uint8_t* mask, void* stream) {
    F4L_REQUIRE(K >= 0 && stride >= 6, "bad size");
    if (K == 0) return F4L_OK;
    F4L_REQUIRE(rows && mask, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    f4l_mark("k_magnitude_mask", st);
    k_magnitude_mask<<<f4l_div_up(K, 256), 256, 0, st>>>(rows, K, stride, max_mag, d_max, factor, strict, mag, mask);
    return f4l_finish("f4l_magnitude_mask", stream);
}
