// A1: median point-cloud resolution = max over the two epochs of the median distance to the nearest OTHER
// point (base.py:2716-2754, src/f2s3.py:481-508: k=2 self kNN, np.median of the 2nd-neighbour distance).
//
// Both epochs go through the SAME nine stream operations (blockIdx.y = epoch), not 21 launches each:
//   memset(control + histograms + cell table) -> k_a1_bbox (last CTA of an epoch: grid parameters)
//   -> k_a1_count -> cub scan (ONE scan over both epochs' tables: epoch 1's ranges continue after epoch 0's)
//   -> k_a1_scatter (float4 rows in cell order; the table itself is advanced, no second memset, no cell-id array)
//   -> k_a1_search (2nd-neighbour squared distance in cell order + level-1 radix histogram in the same pass;
//      last CTA: first radix digit of both middle ranks)
//   -> k_a1_select x2 (levels 2 and 3 of the radix select; last CTA of the last level: median, max over epochs).
// "Last CTA" = the CTA that draws the final ticket of its epoch after a __threadfence(); there is no host
// round trip and no cooperative launch, so the sequence is capturable in a CUDA graph next to other streams.
//
// HBM traffic per point: bbox 12 B + count 12 B + scatter 12+16 B + search 16+4 B + select 2 x 4 B = 80 B.
#include <cub/device/device_scan.cuh>

#include "knn_grid.cuh"

#define A1_BINS 2048
#define A1_LEVELS 3

struct A1Sel {
    unsigned prefix[2];
    unsigned mask[2];
    int k[2];
};

struct A1Ctrl {                 // zeroed by the leading memset
    unsigned ticket[4][2];      // [stage][epoch]
    unsigned final_ticket;
    unsigned pad[7];
    float res[2];
    A1Sel sel[2];
};

struct A1Args {
    const float* p0;            // epoch 0 / 1 (scalars, not arrays: a dynamically indexed kernel parameter
    const float* p1;            // array is copied to local memory)
    int n0, n1;                 // epoch 1's rows follow epoch 0's in `sorted` / `d2` (offset n0)
    int mc;                     // cells reserved per epoch
    int bbox_blocks;
    float cell_factor;
    A1Ctrl* ctrl;
    int* hist;                  // [epoch][level][rank][A1_BINS]
    float* bbox_part;           // [epoch][bbox_blocks][6]
    GridParams* gp;             // [epoch]
    int* table;                 // raw table A: 2*mc + 2 ints; A[0] = 0; counts / running ends live at A + 1
    float4* sorted;
    float* d2;
    float* out;
};

// radix digits of a NON-NEGATIVE float's bit pattern (orders like the value): 11 + 11 + 9 bits
__constant__ int c_a1_shift[A1_LEVELS] = {20, 9, 0};
__constant__ int c_a1_bits[A1_LEVELS] = {11, 11, 9};

__device__ __forceinline__ bool a1_last_block(unsigned* ticket) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// One CTA: resolve the next radix digit of both ranks from the global histograms (warp 0: rank 0, warp 1: rank 1).
__device__ inline void a1_pick(const int* hist0, const int* hist1, int level, A1Sel* st) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= 2) return;
    const int* h = w ? hist1 : hist0;
    const int nb = 1 << c_a1_bits[level], per = nb >> 5;
    int sum = 0;
    for (int i = 0; i < per; ++i) sum += __ldcg(h + lane * per + i);
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(F4L_FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int before = incl - sum;
    const int k = st->k[w];
    __syncwarp();
    if (k >= before && k < before + sum) {
        int cum = before, b = 0;
        for (int i = 0; i < per; ++i) {
            const int c = __ldcg(h + lane * per + i);
            if (k < cum + c) { b = i; break; }
            cum += c;
        }
        st->prefix[w] |= (unsigned)(lane * per + b) << c_a1_shift[level];
        st->mask[w] |= (unsigned)(nb - 1) << c_a1_shift[level];
        st->k[w] = k - cum;
    }
}

// ---- stage 1: bounding boxes + grid parameters ------------------------------------------------
__global__ void __launch_bounds__(256) k_a1_bbox(A1Args a) {
    const int e = blockIdx.y;
    const float* __restrict__ p = e ? a.p1 : a.p0;
    const int n = e ? a.n1 : a.n0;
    __shared__ float smn[8][3], smx[8][3];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // 3 consecutive floats per point; a thread streams whole points, the warp's 384 bytes are contiguous
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(p + (size_t)i * 3 + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(F4L_FULL, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(F4L_FULL, mx[c], o));
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { smn[wid][c] = mn[c]; smx[wid][c] = mx[c]; }
    }
    __syncthreads();
    float* part = a.bbox_part + ((size_t)e * a.bbox_blocks + blockIdx.x) * 6;
    if (threadIdx.x < 3) {
        const int c = threadIdx.x;
        float lo = smn[0][c], hi = smx[0][c];
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, smn[w][c]); hi = fmaxf(hi, smx[w][c]); }
        part[c] = lo;
        part[3 + c] = hi;
    }
    if (!a1_last_block(&a.ctrl->ticket[0][e])) return;
    // last CTA of this epoch: fold the partial boxes, choose the grid, arm the rank select
    if (wid == 0) {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int b = lane; b < (int)gridDim.x; b += 32) {
            const float* q = a.bbox_part + ((size_t)e * a.bbox_blocks + b) * 6;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                lo[c] = fminf(lo[c], __ldcg(q + c));
                hi[c] = fmaxf(hi[c], __ldcg(q + 3 + c));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[c] = fminf(lo[c], __shfl_xor_sync(F4L_FULL, lo[c], o));
                hi[c] = fmaxf(hi[c], __shfl_xor_sync(F4L_FULL, hi[c], o));
            }
        }
        if (lane == 0) {
            grid_params_compute(lo, hi, n, 0.f, a.cell_factor, a.mc, a.gp + e);
            A1Sel* st = &a.ctrl->sel[e];
            st->prefix[0] = st->prefix[1] = 0u;
            st->mask[0] = st->mask[1] = 0u;
            st->k[0] = (n - 1) / 2;     // np.median: mean of the sorted elements (n-1)/2 and n/2
            st->k[1] = n / 2;
        }
    }
}

// ---- stage 2/3: counting sort into cell order ---------------------------------------------------
__global__ void __launch_bounds__(256) k_a1_count(A1Args a) {
    const int e = blockIdx.y;
    const float* __restrict__ p = e ? a.p1 : a.p0;
    const int n = e ? a.n1 : a.n0;
    const GridParams g = a.gp[e];
    int* __restrict__ T = a.table + 1 + (size_t)e * a.mc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = __ldg(p + (size_t)i * 3), y = __ldg(p + (size_t)i * 3 + 1), z = __ldg(p + (size_t)i * 3 + 2);
        int cx, cy, cz;
        cell_of(g, x, y, z, cx, cy, cz);
        atomicAdd(T + (cz * g.ny + cy) * g.nx + cx, 1);
    }
}

// T[c] walks from the start of cell c to its end, so afterwards A[c] (= T[c-1]) is the start of cell c and
// A[c+1] its end: the search reads the table through A, with A[0] = 0 from the leading memset.
__global__ void __launch_bounds__(256) k_a1_scatter(A1Args a) {
    const int e = blockIdx.y;
    const float* __restrict__ p = e ? a.p1 : a.p0;
    const int n = e ? a.n1 : a.n0;
    const GridParams g = a.gp[e];
    int* __restrict__ T = a.table + 1 + (size_t)e * a.mc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = __ldg(p + (size_t)i * 3), y = __ldg(p + (size_t)i * 3 + 1), z = __ldg(p + (size_t)i * 3 + 2);
        int cx, cy, cz;
        cell_of(g, x, y, z, cx, cy, cz);
        const int pos = atomicAdd(T + (cz * g.ny + cy) * g.nx + cx, 1);
        a.sorted[pos] = make_float4(x, y, z, __int_as_float(i));
    }
}

// ---- stage 4: 2nd-neighbour distance of every point + level-1 histogram ------------------------
#ifndef A1S_MIN_BLOCKS
#define A1S_MIN_BLOCKS 10     // 48 registers, 40 warps per SM: ~5 % on the search (long-scoreboard bound)
#endif
__global__ void __launch_bounds__(128, A1S_MIN_BLOCKS) k_a1_search(A1Args a) {
    __shared__ int sh[A1_BINS];
    const int e = blockIdx.y;
    const int n = e ? a.n1 : a.n0;
    for (int i = threadIdx.x; i < A1_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const GridParams g = a.gp[e];
    const int* __restrict__ cell_start = a.table + (size_t)e * a.mc;
    const float4* __restrict__ sorted = a.sorted;
    const float4* __restrict__ qs = a.sorted + (e ? a.n0 : 0);
    float* __restrict__ d2 = a.d2 + (e ? a.n0 : 0);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float4 q = __ldg(qs + t);
        int cx, cy, cz;
        cell_of(g, q.x, q.y, q.z, cx, cy, cz);
        TopD<2> tk;
        tk.init();
        grid_ring_search(sorted, cell_start, g, q.x, q.y, q.z, cx, cy, cz, 2, INFINITY, tk);
        const float dd = tk.kth(1);          // v[0] is the point itself (or a duplicate of it)
        d2[t] = dd;
        atomicAdd(&sh[__float_as_uint(dd) >> 20], 1);
    }
    __syncthreads();
    int* hist = a.hist + (size_t)e * A1_LEVELS * 2 * A1_BINS;     // level 0: both ranks share one histogram
    for (int i = threadIdx.x; i < A1_BINS; i += blockDim.x) {
        const int v = sh[i];
        if (v) atomicAdd(hist + i, v);
    }
    if (!a1_last_block(&a.ctrl->ticket[1][e])) return;
    a1_pick(hist, hist, 0, &a.ctrl->sel[e]);
}

// ---- stage 5: radix-select levels 2 and 3 ------------------------------------------------------
__global__ void __launch_bounds__(256) k_a1_select(A1Args a, int level) {
    __shared__ int sh[2][A1_BINS];
    const int e = blockIdx.y;
    const int n = e ? a.n1 : a.n0;
    const int nb = 1 << c_a1_bits[level], shift = c_a1_shift[level];
    for (int i = threadIdx.x; i < 2 * A1_BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    A1Sel* st = &a.ctrl->sel[e];
    const unsigned p0 = st->prefix[0], m0 = st->mask[0], p1 = st->prefix[1], m1 = st->mask[1];
    const float* __restrict__ d2 = a.d2 + (e ? a.n0 : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned u = __float_as_uint(__ldg(d2 + i));
        const unsigned b = (u >> shift) & (unsigned)(nb - 1);
        if ((u & m0) == p0) atomicAdd(&sh[0][b], 1);
        if ((u & m1) == p1) atomicAdd(&sh[1][b], 1);
    }
    __syncthreads();
    int* hist = a.hist + ((size_t)e * A1_LEVELS + level) * 2 * A1_BINS;
    for (int i = threadIdx.x; i < 2 * A1_BINS; i += blockDim.x) {
        const int v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
    if (!a1_last_block(&a.ctrl->ticket[1 + level][e])) return;
    a1_pick(hist, hist + A1_BINS, level, st);
    if (level + 1 < A1_LEVELS) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        // all 31 value bits are resolved: prefix IS the bit pattern of the rank's squared distance
        const double m = 0.5 * (sqrt((double)__uint_as_float(st->prefix[0])) + sqrt((double)__uint_as_float(st->prefix[1])));
        a.ctrl->res[e] = (float)m;
        __threadfence();
        if (atomicAdd(&a.ctrl->final_ticket, 1u) == 1u) {       // second epoch to finish: max over the epochs
            __threadfence();
            const volatile float* r = a.ctrl->res;
            a.out[0] = fmaxf(r[0], r[1]);
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------
struct A1Ws {
    A1Ctrl* ctrl;
    int* hist;
    int* table;
    size_t zero_bytes;     // ctrl .. table are contiguous: one memset
    float* bbox_part;
    GridParams* gp;
    float4* sorted;
    float* d2;
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
    int mc;
};

static const int kA1BboxBlocks = 148 * 2;

static A1Ws a1_layout(void* base, int n_src, int n_tgt) {
    A1Ws w;
    const int n = n_src > n_tgt ? n_src : n_tgt;
    w.mc = knn_max_cells(n);
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b + off; off += align_up(bytes); return (void*)p; };
    w.ctrl = (A1Ctrl*)take(sizeof(A1Ctrl));
    w.hist = (int*)take((size_t)2 * A1_LEVELS * 2 * A1_BINS * 4);
    w.table = (int*)take(((size_t)2 * w.mc + 2) * 4);
    w.zero_bytes = off;
    w.bbox_part = (float*)take((size_t)2 * kA1BboxBlocks * 6 * 4);
    w.gp = (GridParams*)take(2 * sizeof(GridParams));
    w.sorted = (float4*)take(((size_t)n_src + n_tgt) * 16);
    w.d2 = (float*)take(((size_t)n_src + n_tgt) * 4);
    size_t cb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, 2 * w.mc);
    w.cub_bytes = cb;
    w.cub_tmp = take(cb);
    w.total = off;
    return w;
}

extern "C" size_t f4l_median_resolution_workspace_bytes(int32_t n_src, int32_t n_tgt) {
    if (n_src < 0) n_src = 0;
    if (n_tgt < 0) n_tgt = 0;
    return a1_layout(nullptr, n_src, n_tgt).total;
}

extern "C" int f4l_median_resolution(const float* src, int32_t n_src, const float* tgt, int32_t n_tgt, float* out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(src && tgt && out && workspace, "null pointer");
    F4L_REQUIRE(n_src >= 2 && n_tgt >= 2, "need at least 2 points per epoch");
    F4L_REQUIRE((long long)n_src + n_tgt < (1LL << 31), "too many points for one call");
    const A1Ws w = a1_layout(workspace, n_src, n_tgt);
    if (workspace_bytes < w.total) {
        f4l_set_error("f4l_median_resolution: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n_src > n_tgt ? n_src : n_tgt;
    A1Args a;
    a.p0 = src; a.p1 = tgt;
    a.n0 = n_src; a.n1 = n_tgt;
    a.mc = w.mc;
    a.bbox_blocks = kA1BboxBlocks;
    a.cell_factor = f4l_knn_cell_factor();
    a.ctrl = w.ctrl; a.hist = w.hist; a.bbox_part = w.bbox_part; a.gp = w.gp; a.table = w.table;
    a.sorted = w.sorted; a.d2 = w.d2; a.out = out;

    f4l_mark("#memset_a1", st);
    cudaMemsetAsync(w.ctrl, 0, w.zero_bytes, st);
    const int bb = min(f4l_div_up(n, 256), kA1BboxBlocks);
    f4l_mark("k_a1_bbox", st);
    k_a1_bbox<<<dim3(bb, 2), 256, 0, st>>>(a);
    const int blocks = min(f4l_div_up(n, 256), 148 * 8);
    f4l_mark("k_a1_count", st);
    k_a1_count<<<dim3(blocks, 2), 256, 0, st>>>(a);
    size_t cb = w.cub_bytes;
    f4l_count_launches(1); f4l_mark("cub_exclusive_scan", st);   // cub: init + scan kernels
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.table + 1, w.table + 1, 2 * w.mc, st);
    f4l_mark("k_a1_scatter", st);
    k_a1_scatter<<<dim3(blocks, 2), 256, 0, st>>>(a);
    const int sblocks = min(f4l_div_up(n, 128), 148 * 16);
    f4l_mark("k_a1_search", st);
    k_a1_search<<<dim3(sblocks, 2), 128, 0, st>>>(a);
    const int hblocks = min(f4l_div_up(n, 256 * 8), 148 * 2);
    for (int level = 1; level < A1_LEVELS; ++level) {
        f4l_mark("k_a1_select", st);
        k_a1_select<<<dim3(hblocks, 2), 256, 0, st>>>(a, level);
    }
    return f4l_finish("f4l_median_resolution", stream);
}
