// K-a: exact xyz k-nearest neighbours on a uniform grid, and a radix select for medians.
//
// Layout in HBM: reference points are binned into a dense grid (x fastest) and rewritten as
// float4 {x,y,z,bits(original index)} in cell order, so every candidate load is one aligned
// 16-byte read and the 3 cells of a grid row that a query visits are ONE contiguous range.
// Queries are processed in the same cell order (self query: the sorted array itself; otherwise
// the queries are binned on the same grid), so the threads of a warp visit the same rows and the
// candidates are served from L1.  Each point crosses HBM O(1) times: bin pass (12 B read + 16 B
// write), search (16 B read) and the k results (8k B written).
//
// Exactness: rings of cells are searched until the k-th best distance is strictly below the
// distance to the border of the searched block (or the radius is covered); ties in squared
// distance are broken towards the lower original index, independent of the in-cell order.
#include <cub/device/device_scan.cuh>

#include "knn_grid.cuh"

// ---- bounding box ------------------------------------------------------------------------
__global__ void k_bbox_init(unsigned* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ p, int n, unsigned* bb) {
    __shared__ float smn[8][3], smx[8][3];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = __ldg(p + (size_t)i * 3 + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(F4L_FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(F4L_FULL, mx[a], o));
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { smn[wid][a] = mn[a]; smx[wid][a] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {                 // one atomic pair per axis per CTA
        const int a = threadIdx.x;
        float lo = smn[0][a], hi = smx[0][a];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, smn[w][a]); hi = fmaxf(hi, smx[w][a]); }
        atomicMin(bb + a, f2ord(lo));
        atomicMax(bb + 3 + a, f2ord(hi));
    }
}

__global__ void k_grid_params(const unsigned* bb, int m, float cell_in, float factor, int max_cells, GridParams* gp) {
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {
        mn[a] = ord2f(bb[a]);
        mx[a] = ord2f(bb[3 + a]);
    }
    grid_params_compute(mn, mx, m, cell_in, factor, max_cells, gp);
}

#include <stdlib.h>
float f4l_knn_cell_factor() {
    static const float f = [] {
        const char* s = getenv("F4L_KNN_CELL_FACTOR");
        float v = s ? (float)atof(s) : 0.f;
        return (v >= 0.5f && v <= 8.f) ? v : 1.5f;
    }();
    return f;
}

__global__ void __launch_bounds__(256)
k_bin_count(const float* __restrict__ p, int n, const GridParams* __restrict__ gp, int* __restrict__ cid,
            int* __restrict__ cell_count) {
    const GridParams g = *gp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float x = __ldg(p + (size_t)i * 3), y = __ldg(p + (size_t)i * 3 + 1), z = __ldg(p + (size_t)i * 3 + 2);
        int cx, cy, cz;
        cell_of(g, x, y, z, cx, cy, cz);
        int c = (cz * g.ny + cy) * g.nx + cx;
        cid[i] = c;
        atomicAdd(cell_count + c, 1);
    }
}

__global__ void __launch_bounds__(256)
k_bin_scatter(const float* __restrict__ p, int n, const int* __restrict__ cid,
              const int* __restrict__ cell_start, int* __restrict__ cell_fill, float4* __restrict__ sorted) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int c = cid[i];
        int pos = cell_start[c] + atomicAdd(cell_fill + c, 1);
        sorted[pos] = make_float4(__ldg(p + (size_t)i * 3), __ldg(p + (size_t)i * 3 + 1),
                                  __ldg(p + (size_t)i * 3 + 2), __int_as_float(i));
    }
}

// variant that advances the (scanned) table itself: table[c] walks from the cell's start to its end
__global__ void __launch_bounds__(256)
k_bin_scatter_advance(const float* __restrict__ p, int n, const int* __restrict__ cid,
                      int* __restrict__ table, float4* __restrict__ sorted) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int pos = atomicAdd(table + cid[i], 1);
        sorted[pos] = make_float4(__ldg(p + (size_t)i * 3), __ldg(p + (size_t)i * 3 + 1),
                                  __ldg(p + (size_t)i * 3 + 2), __int_as_float(i));
    }
}

// (idx, d2)[N,k] in the callers' query order.  (The median-resolution variant, which needs only the k-th squared
// distance of every query, lives in medres.cu.)
template <int K>
__global__ void __launch_bounds__(128)
k_grid_search(const float4* __restrict__ queries_sorted, int n, const float4* __restrict__ sorted,
              const int* __restrict__ cell_start, const GridParams* __restrict__ gp, int k, float max_r2,
              int* __restrict__ out_idx, float* __restrict__ out_d2) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GridParams g = *gp;
    const float4 q = __ldg(queries_sorted + t);
    const int qi = __float_as_int(q.w);
    int cx, cy, cz;
    cell_of(g, q.x, q.y, q.z, cx, cy, cz);
    TopK<K> tk;
    tk.init();
    const int kk = K < k ? K : k;
    grid_ring_search(sorted, cell_start, g, q.x, q.y, q.z, cx, cy, cz, kk, max_r2, tk);
    if (K == 2 && k == 2) {                 // the common pair: one 8-byte store per array
        float d0 = tk.d(0), d1 = tk.d(1);
        int i0 = tk.i(0), i1 = tk.i(1);
        if (!(d0 < max_r2)) { d0 = INFINITY; i0 = -1; }
        if (!(d1 < max_r2)) { d1 = INFINITY; i1 = -1; }
        reinterpret_cast<int2*>(out_idx)[qi] = make_int2(i0, i1);
        reinterpret_cast<float2*>(out_d2)[qi] = make_float2(d0, d1);
        return;
    }
#pragma unroll
    for (int a = 0; a < K; ++a) {
        if (a < k) {
            float dd = tk.d(a);
            int ii = tk.i(a);
            if (!(dd < max_r2)) { dd = INFINITY; ii = -1; }
            out_idx[(size_t)qi * k + a] = ii;
            out_d2[(size_t)qi * k + a] = dd;
        }
    }
    for (int a = K; a < k; ++a) { out_idx[(size_t)qi * k + a] = -1; out_d2[(size_t)qi * k + a] = INFINITY; }
}

// Tie flags: tie[i] = 1 when two adjacent distances among the first k + 1 neighbours of query i
// differ by no more than eps_rel * (the larger one) + 1e-12 -- the rows whose index order another exact search (the
// reference's kd-trees) may legitimately report differently.  Same search with one more slot.
template <int K1>
__global__ void __launch_bounds__(128)
k_grid_ties(const float4* __restrict__ queries_sorted, int n, const float4* __restrict__ sorted,
            const int* __restrict__ cell_start, const GridParams* __restrict__ gp, int k, float max_r2, float eps_rel,
            uint8_t* __restrict__ tie) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GridParams g = *gp;
    const float4 q = __ldg(queries_sorted + t);
    const int qi = __float_as_int(q.w);
    int cx, cy, cz;
    cell_of(g, q.x, q.y, q.z, cx, cy, cz);
    TopK<K1> tk;
    tk.init();
    grid_ring_search(sorted, cell_start, g, q.x, q.y, q.z, cx, cy, cz, k + 1 < K1 ? k + 1 : K1, INFINITY, tk);
    bool flag = false;
#pragma unroll
    for (int a = 0; a + 1 < K1; ++a) {
        if (a < k) {
            const float d0 = tk.d(a), d1 = tk.d(a + 1);
            if (d0 < max_r2 && d1 < INFINITY && d1 - d0 <= eps_rel * d1 + 1e-12f) flag = true;
        }
    }
    tie[qi] = flag ? 1 : 0;
}

__global__ void k_fill_none(int n, int k, int* idx, float* d2) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * k) { idx[t] = -1; d2[t] = INFINITY; }
}

// ---- workspace layout ------------------------------------------------------------------------
struct KnnWs {
    GridParams* gp;
    unsigned* bb;
    int* cell_count;   // max_cells+1 (counts, then exclusive scan in cell_start)
    int* cell_start;   // max_cells+1
    int* cid;          // max(N,M)
    float4* sorted_r;  // M
    float4* sorted_q;  // N
    void* cub_tmp;
    size_t cub_bytes;
    size_t total;
};

static KnnWs knn_layout(void* base, int N, int M) {
    KnnWs w;
    const int mc = knn_max_cells(M);
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b + off; off += align_up(bytes); return (void*)p; };
    w.gp = (GridParams*)take(sizeof(GridParams));
    w.bb = (unsigned*)take(6 * sizeof(unsigned));
    w.cell_count = (int*)take((size_t)(mc + 1) * 4);
    w.cell_start = (int*)take((size_t)(mc + 1) * 4);
    w.cid = (int*)take((size_t)(N > M ? N : M) * 4);
    w.sorted_r = (float4*)take((size_t)M * 16);
    w.sorted_q = (float4*)take((size_t)N * 16);
    size_t cb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, mc + 1);
    w.cub_bytes = cb;
    w.cub_tmp = take(cb);
    w.total = off;
    return w;
}

extern "C" size_t f4l_knn_grid_workspace_bytes(int32_t N, int32_t M) {
    if (N < 0) N = 0;
    if (M < 0) M = 0;
    return knn_layout(nullptr, N, M).total;
}

static int bin_points(const float* p, int n, const KnnWs& w, int mc, float4* sorted, cudaStream_t st) {
    f4l_mark("#memset_cells", st);
    cudaMemsetAsync(w.cell_count, 0, (size_t)(mc + 1) * 4, st);
    const int blocks = min(f4l_div_up(n, 256), 148 * 8);
    f4l_mark("k_bin_count", st);
    k_bin_count<<<blocks, 256, 0, st>>>(p, n, w.gp, w.cid, w.cell_count);
    size_t cb = w.cub_bytes;
    f4l_count_launches(1); f4l_mark("cub_exclusive_scan", st);   // cub: init + scan kernels
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.cell_count, w.cell_start, mc + 1, st);
    f4l_mark("#memset_cells", st);
    cudaMemsetAsync(w.cell_count, 0, (size_t)(mc + 1) * 4, st);
    f4l_mark("k_bin_scatter", st);
    k_bin_scatter<<<blocks, 256, 0, st>>>(p, n, w.cid, w.cell_start, w.cell_count, sorted);
    return f4l_check_launch("f4l_knn_grid/bin");
}

static int knn_grid_impl(const float* q, int32_t N, const float* r, int32_t M, int32_t k, float max_radius,
                         float cell, int32_t* idx, float* d2, void* workspace, size_t workspace_bytes,
                         void* stream, uint8_t* tie = nullptr, float eps_rel = 0.f) {
    cudaStream_t st = (cudaStream_t)stream;
    KnnWs w = knn_layout(workspace, N, M);
    if (workspace_bytes < w.total) {
        f4l_set_error("f4l_knn_grid: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    const int mc = knn_max_cells(M);
    f4l_mark("k_bbox_init", st);
    k_bbox_init<<<1, 32, 0, st>>>(w.bb);
    f4l_mark("k_bbox", st);
    k_bbox<<<min(f4l_div_up(M, 256), 148 * 4), 256, 0, st>>>(r, M, w.bb);
    f4l_mark("k_grid_params", st);
    k_grid_params<<<1, 1, 0, st>>>(w.bb, M, cell, f4l_knn_cell_factor(), mc, w.gp);
    int rc = bin_points(r, M, w, mc, w.sorted_r, st);
    if (rc) return rc;
    // keep the reference table: the query binning below needs its own counters
    const float4* qs = w.sorted_r;
    int* ref_start = w.cell_start;
    if (q != r || N != M) {
        // bin the queries on the same grid (queries outside the box clamp to border cells); their
        // table is scanned in place in cell_count, cell_start stays the reference table
        f4l_mark("#memset_cells", st);
        cudaMemsetAsync(w.cell_count, 0, (size_t)(mc + 1) * 4, st);
        const int blocks = min(f4l_div_up(N, 256), 148 * 8);
        f4l_mark("k_bin_count", st);
        k_bin_count<<<blocks, 256, 0, st>>>(q, N, w.gp, w.cid, w.cell_count);
        size_t cb = w.cub_bytes;
        f4l_count_launches(1); f4l_mark("cub_exclusive_scan", st);
        cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.cell_count, w.cell_count, mc + 1, st);  // in place
        f4l_mark("k_bin_scatter_advance", st);
        k_bin_scatter_advance<<<blocks, 256, 0, st>>>(q, N, w.cid, w.cell_count, w.sorted_q);
        qs = w.sorted_q;
    }
    const float max_r2 = max_radius > 0.f ? max_radius * max_radius : INFINITY;
    const int blocks = f4l_div_up(N, 128);
    if (tie) {
        f4l_mark("k_grid_ties", st);
        if (k == 1) k_grid_ties<2><<<blocks, 128, 0, st>>>(qs, N, w.sorted_r, ref_start, w.gp, k, max_r2, eps_rel, tie);
        else if (k <= 3) k_grid_ties<4><<<blocks, 128, 0, st>>>(qs, N, w.sorted_r, ref_start, w.gp, k, max_r2, eps_rel, tie);
        else k_grid_ties<8><<<blocks, 128, 0, st>>>(qs, N, w.sorted_r, ref_start, w.gp, k, max_r2, eps_rel, tie);
        return f4l_finish("f4l_knn_grid_ties", stream);
    }
    f4l_mark("k_grid_search", st);
#define KNN_LAUNCH(KK) k_grid_search<KK><<<blocks, 128, 0, st>>>(qs, N, w.sorted_r, ref_start, w.gp, k, max_r2, idx, d2)
    if (k == 1) KNN_LAUNCH(1);
    else if (k == 2) KNN_LAUNCH(2);
    else if (k <= 4) KNN_LAUNCH(4);
    else KNN_LAUNCH(8);
#undef KNN_LAUNCH
    return f4l_finish("f4l_knn_grid/search", stream);
}

extern "C" int f4l_knn_grid(const float* q, int32_t N, const float* r, int32_t M, int32_t k, float max_radius,
                            float cell, int32_t* idx, float* d2, void* workspace, size_t workspace_bytes,
                            void* stream) {
    F4L_REQUIRE(N >= 0 && M >= 0, "negative size");
    F4L_REQUIRE(k >= 1 && k <= 8, "k must be in [1,8]");
    if (N == 0) return F4L_OK;
    F4L_REQUIRE(q && idx && d2, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        f4l_mark("k_fill_none", st);
        k_fill_none<<<f4l_div_up((long long)N * k, 256), 256, 0, st>>>(N, k, idx, d2);
        return f4l_finish("f4l_knn_grid/fill", stream);
    }
    F4L_REQUIRE(r && workspace, "null pointer");
    return knn_grid_impl(q, N, r, M, k, max_radius, cell, idx, d2, workspace, workspace_bytes, stream);
}

extern "C" int f4l_knn_grid_ties(const float* q, int32_t N, const float* r, int32_t M, int32_t k, float max_radius,
                                 float cell, float eps_rel, uint8_t* tie, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    F4L_REQUIRE(N >= 0 && M >= 0, "negative size");
    F4L_REQUIRE(k >= 1 && k <= 7, "k must be in [1,7]");
    F4L_REQUIRE(eps_rel >= 0.f, "eps_rel < 0");
    if (N == 0) return F4L_OK;
    F4L_REQUIRE(q && tie, "null pointer");
    if (M == 0) {
        cudaMemsetAsync(tie, 0, (size_t)N, (cudaStream_t)stream);
        return F4L_OK;
    }
    F4L_REQUIRE(r && workspace, "null pointer");
    return knn_grid_impl(q, N, r, M, k, max_radius, cell, nullptr, nullptr, workspace, workspace_bytes, stream, tie, eps_rel);
}

// ---- radix select (k-th smallest) -------------------------------------------------------------
// Three passes of 11/11/10 bits over the order-preserving integer image of the floats; two ranks
// are resolved in the same passes (np.median of an even count needs both middle elements).
struct SelState {
    unsigned prefix[2];
    unsigned mask[2];
    int k[2];
};

__global__ void k_sel_init(SelState* st, int* hist, int k, int k2) {
    if (threadIdx.x == 0) {
        st->prefix[0] = st->prefix[1] = 0u;
        st->mask[0] = st->mask[1] = 0u;
        st->k[0] = k;
        st->k[1] = k2 >= 0 ? k2 : k;
    }
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0;
}

__global__ void __launch_bounds__(256)
k_sel_hist(const float* __restrict__ x, int n, int stride, int offset, const SelState* __restrict__ st,
           int shift, int bits, int* __restrict__ hist) {
    __shared__ int sh[2][2048];
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const unsigned p0 = st->prefix[0], m0 = st->mask[0], p1 = st->prefix[1], m1 = st->mask[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned u = f2ord(__ldg(x + (size_t)i * stride + offset));
        unsigned b = (u >> shift) & ((1u << bits) - 1u);
        if ((u & m0) == p0) atomicAdd(&sh[0][b], 1);
        if ((u & m1) == p1) atomicAdd(&sh[1][b], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) {
        int v = (&sh[0][0])[i];
        if (v) atomicAdd(hist + i, v);
    }
}

__global__ void __launch_bounds__(512) k_sel_pick(SelState* st, int* hist, int shift, int bits) {
    // 512 threads: threads [0,256) resolve rank 0, [256,512) rank 1; each owns 8 consecutive bins
    __shared__ int part[2][256];
    const int r = threadIdx.x >> 8, t = threadIdx.x & 255;
    const int nb = 1 << bits;
    int local[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int b = t * 8 + i;
        local[i] = b < nb ? hist[r * 2048 + b] : 0;
        sum += local[i];
    }
    part[r][t] = sum;
    __syncthreads();
    // exclusive prefix of the 256 partial sums (Hillis-Steele in shared memory)
    for (int o = 1; o < 256; o <<= 1) {
        const int v = t >= o ? part[r][t - o] : 0;
        __syncthreads();
        part[r][t] += v;
        __syncthreads();
    }
    const int k = st->k[r];
    const int before = part[r][t] - sum;     // elements in bins before this thread's first bin
    __syncthreads();
    if (k >= before && k < before + sum) {
        int cum = before, b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (k >= cum + local[i]) { cum += local[i]; b = i + 1; }
            else break;
        }
        st->prefix[r] |= (unsigned)(t * 8 + b) << shift;
        st->mask[r] |= (unsigned)(nb - 1) << shift;
        st->k[r] = k - cum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 2048; i += blockDim.x) hist[i] = 0;
}

__global__ void k_sel_out(const SelState* st, float* out, int two) {
    out[0] = ord2f(st->prefix[0]);
    if (two) out[1] = ord2f(st->prefix[1]);
}

extern "C" size_t f4l_select_kth_workspace_bytes(int32_t n) {
    (void)n;
    return align_up(sizeof(SelState)) + align_up(2 * 2048 * sizeof(int));
}

extern "C" int f4l_select_kth(const float* x, int32_t n, int32_t stride, int32_t offset, int32_t k, int32_t k2,
                              float* out, void* workspace, size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(x && out && workspace, "null pointer");
    F4L_REQUIRE(n >= 1 && stride >= 1 && offset >= 0 && k >= 0 && k < n && k2 < n, "bad rank / size");
    if (workspace_bytes < f4l_select_kth_workspace_bytes(n)) {
        f4l_set_error("f4l_select_kth: workspace too small");
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SelState* s = (SelState*)workspace;
    int* hist = (int*)((char*)workspace + align_up(sizeof(SelState)));
    f4l_mark("k_sel_init", st);
    k_sel_init<<<1, 256, 0, st>>>(s, hist, k, k2);
    const int blocks = min(f4l_div_up(n, 256 * 8), 148 * 4);
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
        f4l_mark("k_sel_hist", st);
        k_sel_hist<<<blocks, 256, 0, st>>>(x, n, stride, offset, s, shifts[p], bits[p], hist);
        f4l_mark("k_sel_pick", st);
        k_sel_pick<<<1, 512, 0, st>>>(s, hist, shifts[p], bits[p]);
    }
    f4l_mark("k_sel_out", st);
    k_sel_out<<<1, 1, 0, st>>>(s, out, k2 >= 0);
    return f4l_finish("f4l_select_kth", stream);
}

