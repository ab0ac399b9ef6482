// K-g: the reference's `Piecewise_ICP(cfg)` (src/piecewise_icp.py:89-216) as one launch
// sequence without host round trips: Open3D-octree leaf cells -> per-cell fp64 centroids -> 1-NN between
// the centroid sets -> mean+std stability threshold -> per-cell translation rows.
//
// The octree is restated as integer leaf codes (base-8 digits x+2y+4z, root digit most significant)
// produced by the same comparisons as Octree::InsertPointRecurse; a stable radix sort by code puts the
// points of a leaf in index order and the leaves in DFS (traversal) order.  The hard-coded 250-point
// early stop on internal nodes (piecewise_icp.py:52) and `number_points_min` on leaves (:55) become range
// counts over the sorted leaf table.  All arithmetic is fp64 like Open3D / numpy.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

struct PwFrame {
    unsigned long long bb[6];   // ordered-u64 min xyz, max xyz of both clouds
    double lo[3], hi[3];
    double origin[3], size;
    int depth;
    int n_leaf[2];              // occupied leaves per cloud
    int n_cell[2];              // selected cells per cloud (Cs, Ct)
    double thr;
    int n_stable, n_rows, n_stable_rows;
};

__device__ __forceinline__ unsigned long long d2ord(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2d(unsigned long long u) {
    return __longlong_as_double((long long)((u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u));
}

// Open3D AxisAlignedBoundingBox::GetBoxPoints order (piecewise_icp.py:101-105 appends these 8 points)
__device__ __forceinline__ void pw_corner(const PwFrame& f, int k, double& x, double& y, double& z) {
    const double ex = f.hi[0] - f.lo[0], ey = f.hi[1] - f.lo[1], ez = f.hi[2] - f.lo[2];
    x = f.lo[0]; y = f.lo[1]; z = f.lo[2];
    switch (k) {
        case 1: x = f.lo[0] + ex; break;
        case 2: y = f.lo[1] + ey; break;
        case 3: z = f.lo[2] + ez; break;
        case 4: x = f.hi[0]; y = f.hi[1]; z = f.hi[2]; break;
        case 5: y = f.lo[1] + ey; z = f.lo[2] + ez; break;
        case 6: x = f.lo[0] + ex; z = f.lo[2] + ez; break;
        case 7: x = f.lo[0] + ex; y = f.lo[1] + ey; break;
        default: break;
    }
}
__device__ __forceinline__ void pw_point(const double* __restrict__ p, int n, const PwFrame& f, int i, double& x, double& y, double& z) {
    if (i < n) { x = p[3 * (size_t)i]; y = p[3 * (size_t)i + 1]; z = p[3 * (size_t)i + 2]; }
    else pw_corner(f, i - n, x, y, z);
}

__global__ void k_pw_init(PwFrame* f) {
    if (threadIdx.x < 3) f->bb[threadIdx.x] = ~0ull;
    else if (threadIdx.x < 6) f->bb[threadIdx.x] = 0ull;
}

__global__ void __launch_bounds__(256) k_pw_bbox(const double* __restrict__ p, int n, PwFrame* f) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = p[3 * (size_t)i + a];
            mn[a] = fmin(mn[a], v);
            mx[a] = fmax(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fmin(mn[a], __shfl_xor_sync(F4L_FULL, mn[a], o));
            mx[a] = fmax(mx[a], __shfl_xor_sync(F4L_FULL, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&f->bb[a], d2ord(mn[a]));
            atomicMax(&f->bb[3 + a], d2ord(mx[a]));
        }
    }
}

// piecewise_icp.py:96-118: union bounding box, depth, octree frame.  Both clouds contain the 8 corners
// of the union box, so both octrees share the frame: center=(min+max)/2, h=max(center-min),
// origin=min(min, center-h), size=2h (Open3D ConvertFromPointCloud, size_expand=0).
__global__ void k_pw_frame(PwFrame* f, double smax) {
    double ext = 0.0;
    for (int a = 0; a < 3; ++a) {
        f->lo[a] = ord2d(f->bb[a]);
        f->hi[a] = ord2d(f->bb[3 + a]);
        ext = fmax(ext, f->hi[a] - f->lo[a]);
    }
    int depth = (int)ceil(log2(ext / smax));
    if (!(depth > 0)) depth = 0;
    if (depth > 20) depth = 20;
    f->depth = depth;
    double h = 0.0, c[3];
    for (int a = 0; a < 3; ++a) {
        c[a] = (f->lo[a] + f->hi[a]) / 2.0;
        h = fmax(h, c[a] - f->lo[a]);
    }
    for (int a = 0; a < 3; ++a) f->origin[a] = fmin(f->lo[a], c[a] - h);
    f->size = 2.0 * h;
}

// leaf code by the comparisons of Octree::InsertPointRecurse; out-of-bound points (max faces) get ~0
__global__ void __launch_bounds__(256)
k_pw_codes(const double* __restrict__ p, int n, const PwFrame* __restrict__ fr, unsigned long long* __restrict__ key,
           int32_t* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n + 8) return;
    const PwFrame& f = *fr;
    double q[3];
    pw_point(p, n, f, i, q[0], q[1], q[2]);
    bool inb = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) inb = inb && (f.origin[a] <= q[a]) && (q[a] < f.origin[a] + f.size);
    unsigned long long code = 0ull;
    double no[3] = {f.origin[0], f.origin[1], f.origin[2]};
    double s = f.size;
    for (int l = 0; l < f.depth; ++l) {
        s = s / 2.0;
        unsigned digit = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const bool bit = q[a] >= no[a] + s;
            if (bit) { no[a] = no[a] + s; digit |= 1u << a; }
        }
        code = code * 8ull + digit;
    }
    key[i] = inb ? code : ~0ull;
    val[i] = i;
}

__global__ void __launch_bounds__(256)
k_pw_heads(const unsigned long long* __restrict__ key, int n, int32_t* __restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    head[i] = (k != ~0ull && (i == 0 || key[i - 1] != k)) ? 1 : 0;
}

// leaf table from the head scan: leaf_code[l], leaf_start[l] (+ sentinel = number of in-bound points)
__global__ void __launch_bounds__(256)
k_pw_leaves(const unsigned long long* __restrict__ key, const int32_t* __restrict__ head, const int32_t* __restrict__ hscan,
            int n, unsigned long long* __restrict__ leaf_code, int32_t* __restrict__ leaf_start, PwFrame* f, int cloud) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (head[i]) { leaf_code[hscan[i]] = key[i]; leaf_start[hscan[i]] = i; }
    const bool last_inb = key[i] != ~0ull && (i == n - 1 || key[i + 1] == ~0ull);
    if (last_inb) {
        const int nl = hscan[i] + head[i];
        leaf_start[nl] = i + 1;
        f->n_leaf[cloud] = nl;
    }
    if (i == 0 && key[0] == ~0ull) { leaf_start[0] = 0; f->n_leaf[cloud] = 0; }
}

__device__ __forceinline__ int pw_lower_bound(const unsigned long long* a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// a leaf is visited iff every internal ancestor (levels 0..depth-1) holds >= internal_min points
// (piecewise_icp.py:52) and it is kept iff it holds >= number_points_min points (:55)
__global__ void __launch_bounds__(256)
k_pw_select(const unsigned long long* __restrict__ leaf_code, const int32_t* __restrict__ leaf_start, const PwFrame* __restrict__ f,
            int cloud, int internal_min, int min_pts, int32_t* __restrict__ sel) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = f->n_leaf[cloud];
    if (l >= nl) return;
    const unsigned long long code = leaf_code[l];
    const int depth = f->depth;
    bool ok = (leaf_start[l + 1] - leaf_start[l]) >= min_pts;
    for (int lvl = 0; lvl < depth && ok; ++lvl) {
        const int sh = 3 * (depth - lvl);
        const unsigned long long anc = code >> sh;
        const unsigned long long first = anc << sh;
        const int a = pw_lower_bound(leaf_code, nl, first);
        const int b = (sh >= 64 || (anc + 1ull) << sh == 0ull) ? nl : pw_lower_bound(leaf_code, nl, (anc + 1ull) << sh);
        ok = (leaf_start[b] - leaf_start[a]) >= internal_min;
    }
    sel[l] = ok ? 1 : 0;
}

// warp per selected cell: fp64 centroid (piecewise_icp.py:58-61)
__global__ void __launch_bounds__(256)
k_pw_centroids(const double* __restrict__ p, int n, const PwFrame* __restrict__ f, int cloud, const int32_t* __restrict__ sel,
               const int32_t* __restrict__ sscan, const int32_t* __restrict__ leaf_start, const int32_t* __restrict__ order,
               double* __restrict__ cent, int32_t* __restrict__ cell_leaf, PwFrame* fw) {
    const int lane = threadIdx.x & 31;
    const int l = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nl = f->n_leaf[cloud];
    if (l >= nl) return;
    if (l == nl - 1 && lane == 0) fw->n_cell[cloud] = sscan[l] + sel[l];
    if (!sel[l]) return;
    const int c = sscan[l];
    const int s0 = leaf_start[l], s1 = leaf_start[l + 1];
    double sx = 0, sy = 0, sz = 0;
    for (int i = s0 + lane; i < s1; i += 32) {
        double x, y, z;
        pw_point(p, n, *f, order[i], x, y, z);
        sx += x; sy += y; sz += z;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
    if (lane == 0) {
        const double inv = (double)(s1 - s0);
        cent[3 * (size_t)c] = sx / inv; cent[3 * (size_t)c + 1] = sy / inv; cent[3 * (size_t)c + 2] = sz / inv;
        cell_leaf[c] = l;
    }
}

// A5: 1-NN among the target centroids for every source centroid (piecewise_icp.py:134-149), fp64
__global__ void __launch_bounds__(128)
k_pw_centroid_nn(const double* __restrict__ cs, const double* __restrict__ ct, const PwFrame* __restrict__ f,
                 int32_t* __restrict__ nn, double* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int Cs = f->n_cell[0], Ct = f->n_cell[1];
    if (i >= Cs) return;
    const double x = cs[3 * (size_t)i], y = cs[3 * (size_t)i + 1], z = cs[3 * (size_t)i + 2];
    double best = INFINITY;
    int bj = -1;
    for (int j = 0; j < Ct; ++j) {
        const double dx = x - ct[3 * (size_t)j], dy = y - ct[3 * (size_t)j + 1], dz = z - ct[3 * (size_t)j + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { best = d2; bj = j; }
    }
    nn[i] = bj;
    dist[i] = sqrt(best);
}

// F5: thr = mean(d) + std(d) (population std), stable = d <= thr (piecewise_icp.py:152-161); single CTA
__global__ void __launch_bounds__(1024)
k_pw_threshold(const double* __restrict__ dist, PwFrame* f, int32_t* __restrict__ stable) {
    __shared__ double red[32];
    __shared__ double s_mean, s_thr;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int C = f->n_cell[0];
    double s = 0.0;
    for (int i = tid; i < C; i += 1024) s += dist[i];
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (wid == 0) {
        double v = red[lane];
        v = warp_sum(v);
        if (lane == 0) s_mean = C > 0 ? v / (double)C : 0.0;
    }
    __syncthreads();
    const double mean = s_mean;
    s = 0.0;
    for (int i = tid; i < C; i += 1024) { const double d = dist[i] - mean; s += d * d; }
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (wid == 0) {
        double v = red[lane];
        v = warp_sum(v);
        if (lane == 0) { s_thr = mean + (C > 0 ? sqrt(v / (double)C) : 0.0); f->thr = s_thr; }
    }
    __syncthreads();
    const double thr = s_thr;
    int ns = 0;
    for (int i = tid; i < C; i += 1024) {
        const int st = dist[i] <= thr ? 1 : 0;
        stable[i] = st;
        ns += st;
    }
    ns = warp_sum(ns);
    if (lane == 0 && ns) atomicAdd(&f->n_stable, ns);
}

// keys for the lexicographic sort of the stable centroids (np.unique(axis=0), :169): unstable cells sort last
__global__ void __launch_bounds__(256)
k_pw_lex_keys(const double* __restrict__ cent, const int32_t* __restrict__ stable, const int32_t* __restrict__ perm_in,
              const PwFrame* __restrict__ f, int axis, int n_all, double* __restrict__ key, int32_t* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_all) return;
    int c = -1;
    if (i < f->n_cell[0]) c = perm_in ? perm_in[i] : i;     // the sorts are stable: valid entries stay in front
    key[i] = (c >= 0 && stable[c]) ? cent[3 * (size_t)c + axis] : INFINITY;
    val[i] = c;
}

// output order: stable cells (lexicographic centroid order) then unstable cells (traversal order)
__global__ void __launch_bounds__(256)
k_pw_out_order(const int32_t* __restrict__ lex_perm, const int32_t* __restrict__ stable, const int32_t* __restrict__ uscan,
               const int32_t* __restrict__ cell_leaf, const int32_t* __restrict__ leaf_start, const PwFrame* __restrict__ f,
               int32_t* __restrict__ out_cell, int32_t* __restrict__ out_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int C = f->n_cell[0];
    if (i >= C) return;
    const int ns = f->n_stable;
    if (i < ns) {
        const int c = lex_perm[i];
        out_cell[i] = c;
        out_cnt[i] = leaf_start[cell_leaf[c] + 1] - leaf_start[cell_leaf[c]];
    }
    if (!stable[i]) {
        const int pos = ns + uscan[i];
        out_cell[pos] = i;
        out_cnt[pos] = leaf_start[cell_leaf[i] + 1] - leaf_start[cell_leaf[i]];
    }
}

__global__ void k_pw_unstable_flag(const int32_t* __restrict__ stable, const PwFrame* __restrict__ f, int32_t* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < f->n_cell[0]) flag[i] = stable[i] ? 0 : 1;
}

// rows [p | p + (c_t - c_s)] (zero shift for stable cells) and magnitudes (:166-202)
__global__ void __launch_bounds__(256)
k_pw_emit(const double* __restrict__ p, int n, const PwFrame* __restrict__ f, const int32_t* __restrict__ out_cell,
          const int32_t* __restrict__ out_cnt, const int32_t* __restrict__ out_off, const int32_t* __restrict__ cell_leaf,
          const int32_t* __restrict__ leaf_start, const int32_t* __restrict__ order, const int32_t* __restrict__ stable,
          const double* __restrict__ cs, const double* __restrict__ ct, const int32_t* __restrict__ nn,
          double* __restrict__ dvfs, double* __restrict__ mag, PwFrame* fw) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int C = f->n_cell[0];
    if (i >= C) return;
    const int c = out_cell[i];
    const int off = out_off[i], cnt = out_cnt[i];
    if (lane == 0) {
        if (i == C - 1) fw->n_rows = off + cnt;
        if (i == f->n_stable - 1) fw->n_stable_rows = off + cnt;
    }
    double sx = 0, sy = 0, sz = 0;
    const bool st = stable[c] != 0;
    if (!st) {
        const int j = nn[c];
        sx = ct[3 * (size_t)j] - cs[3 * (size_t)c]; sy = ct[3 * (size_t)j + 1] - cs[3 * (size_t)c + 1]; sz = ct[3 * (size_t)j + 2] - cs[3 * (size_t)c + 2];
    }
    const int s0 = leaf_start[cell_leaf[c]];
    for (int k = lane; k < cnt; k += 32) {
        double x, y, z;
        pw_point(p, n, *f, order[s0 + k], x, y, z);
        double* row = dvfs + (size_t)(off + k) * 6;
        const double tx = st ? x : x + sx, ty = st ? y : y + sy, tz = st ? z : z + sz;
        row[0] = x; row[1] = y; row[2] = z; row[3] = tx; row[4] = ty; row[5] = tz;
        if (mag) {
            const double dx = x - tx, dy = y - ty, dz = z - tz;
            mag[off + k] = sqrt(dx * dx + dy * dy + dz * dz);
        }
    }
}

__global__ void k_pw_counts(const PwFrame* __restrict__ f, int32_t* __restrict__ counts, double* __restrict__ thr_out) {
    counts[0] = f->n_cell[0] > 0 ? f->n_rows : 0;
    counts[1] = f->n_stable > 0 ? f->n_stable_rows : 0;
    counts[2] = f->n_cell[0];
    counts[3] = f->n_cell[1];
    counts[4] = f->depth;
    counts[5] = f->n_cell[0] - f->n_stable;
    if (thr_out) thr_out[0] = f->thr;
}

// ---------------------------------------------------------------------------------------------
static inline size_t pal(size_t x) { return (x + 255) / 256 * 256; }

struct PwCloudWs {
    unsigned long long *key, *key_sorted, *leaf_code;
    int32_t *val, *order, *head, *hscan, *leaf_start, *sel, *sscan, *cell_leaf;
    double* cent;
};
struct PwWs {
    PwFrame* frame;
    PwCloudWs c[2];
    int32_t *nn, *stable, *uflag, *uscan, *lex_a, *lex_b, *out_cell, *out_cnt, *out_off;
    double *dist, *lkey_a, *lkey_b;
    void* cub;
    size_t cub_bytes, total;
};

static size_t pw_cub_bytes(int n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int32_t*)nullptr,
                                    (int32_t*)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, n);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (double*)nullptr, (double*)nullptr, (int32_t*)nullptr, (int32_t*)nullptr, n);
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

static PwWs pw_layout(void* base, int n_src, int n_tgt) {
    PwWs w;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p + off; off += pal(bytes); return (void*)q; };
    w.frame = (PwFrame*)take(sizeof(PwFrame));
    const int ns[2] = {n_src + 8, n_tgt + 8};
    for (int k = 0; k < 2; ++k) {
        const size_t n = (size_t)ns[k];
        w.c[k].key = (unsigned long long*)take(n * 8);
        w.c[k].key_sorted = (unsigned long long*)take(n * 8);
        w.c[k].leaf_code = (unsigned long long*)take(n * 8);
        w.c[k].val = (int32_t*)take(n * 4);
        w.c[k].order = (int32_t*)take(n * 4);
        w.c[k].head = (int32_t*)take(n * 4);
        w.c[k].hscan = (int32_t*)take(n * 4);
        w.c[k].leaf_start = (int32_t*)take((n + 1) * 4);
        w.c[k].sel = (int32_t*)take(n * 4);
        w.c[k].sscan = (int32_t*)take(n * 4);
        w.c[k].cell_leaf = (int32_t*)take(n * 4);
        w.c[k].cent = (double*)take(n * 24);
    }
    const size_t n = (size_t)ns[0];
    w.nn = (int32_t*)take(n * 4);
    w.stable = (int32_t*)take(n * 4);
    w.uflag = (int32_t*)take(n * 4);
    w.uscan = (int32_t*)take(n * 4);
    w.lex_a = (int32_t*)take(n * 4);
    w.lex_b = (int32_t*)take(n * 4);
    w.out_cell = (int32_t*)take(n * 4);
    w.out_cnt = (int32_t*)take(n * 4);
    w.out_off = (int32_t*)take(n * 4);
    w.dist = (double*)take(n * 8);
    w.lkey_a = (double*)take(n * 8);
    w.lkey_b = (double*)take(n * 8);
    w.cub_bytes = pw_cub_bytes(ns[0] > ns[1] ? ns[0] : ns[1]);
    w.cub = take(w.cub_bytes);
    w.total = off;
    return w;
}

extern "C" size_t f4l_piecewise_icp_workspace_bytes(int32_t n_src, int32_t n_tgt) {
    if (n_src < 0 || n_tgt < 0) return 0;
    return pw_layout(nullptr, n_src, n_tgt).total + 256;
}

extern "C" int f4l_piecewise_icp(const double* src64, int32_t n_src, const double* tgt64, int32_t n_tgt, double smax,
                                 int32_t number_points_min, int32_t internal_min_points, double* dvfs, double* mag,
                                 int32_t* counts, double* thr_out, double* cent_src, double* cent_tgt, int32_t* nn_out,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(n_src > 0 && n_tgt > 0, "empty cloud");
    F4L_REQUIRE(src64 && tgt64 && dvfs && counts, "null pointer");
    F4L_REQUIRE(smax > 0.0, "smax must be > 0");
    const size_t need = f4l_piecewise_icp_workspace_bytes(n_src, n_tgt);
    if (!workspace || workspace_bytes < need) {
        f4l_set_error("f4l_piecewise_icp: workspace too small (%zu < %zu)", workspace_bytes, need);
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PwWs w = pw_layout((void*)(((uintptr_t)workspace + 255) / 256 * 256), n_src, n_tgt);
    const double* pts[2] = {src64, tgt64};
    const int nn_[2] = {n_src, n_tgt};
    f4l_mark("k_pw_frame", st);
    cudaMemsetAsync(w.frame, 0, sizeof(PwFrame), st);
    k_pw_init<<<1, 32, 0, st>>>(w.frame);
    k_pw_bbox<<<148 * 4, 256, 0, st>>>(src64, n_src, w.frame);
    k_pw_bbox<<<148 * 4, 256, 0, st>>>(tgt64, n_tgt, w.frame);
    k_pw_frame<<<1, 1, 0, st>>>(w.frame, smax);
    f4l_count_launches(4);
    for (int k = 0; k < 2; ++k) {
        const int n = nn_[k], na = n + 8;
        PwCloudWs& c = w.c[k];
        f4l_mark("k_pw_codes", st);
        k_pw_codes<<<f4l_div_up(na, 256), 256, 0, st>>>(pts[k], n, w.frame, c.key, c.val);
        f4l_mark("cub_radix_sort_cells", st);
        size_t cb = w.cub_bytes;
        cub::DeviceRadixSort::SortPairs(w.cub, cb, c.key, c.key_sorted, c.val, c.order, na, 0, 64, st);
        f4l_mark("k_pw_leaves", st);
        k_pw_heads<<<f4l_div_up(na, 256), 256, 0, st>>>(c.key_sorted, na, c.head);
        cb = w.cub_bytes;
        cub::DeviceScan::ExclusiveSum(w.cub, cb, c.head, c.hscan, na, st);
        k_pw_leaves<<<f4l_div_up(na, 256), 256, 0, st>>>(c.key_sorted, c.head, c.hscan, na, c.leaf_code, c.leaf_start, w.frame, k);
        k_pw_select<<<f4l_div_up(na, 256), 256, 0, st>>>(c.leaf_code, c.leaf_start, w.frame, k, internal_min_points,
                                                        number_points_min, c.sel);
        cb = w.cub_bytes;
        cub::DeviceScan::ExclusiveSum(w.cub, cb, c.sel, c.sscan, na, st);
        f4l_count_launches(4);
        f4l_mark("k_pw_centroids", st);
        k_pw_centroids<<<f4l_div_up(na, 8), 256, 0, st>>>(pts[k], n, w.frame, k, c.sel, c.sscan, c.leaf_start, c.order, c.cent,
                                                         c.cell_leaf, w.frame);
    }
    const int na = n_src + 8;
    f4l_mark("k_pw_centroid_nn", st);
    k_pw_centroid_nn<<<f4l_div_up(na, 128), 128, 0, st>>>(w.c[0].cent, w.c[1].cent, w.frame, w.nn, w.dist);
    f4l_mark("k_pw_threshold", st);
    k_pw_threshold<<<1, 1024, 0, st>>>(w.dist, w.frame, w.stable);
    // lexicographic order of the stable centroids: stable LSD sorts by z, y, x
    f4l_mark("k_pw_order", st);
    const int grid = f4l_div_up(na, 256);
    size_t cb = w.cub_bytes;
    k_pw_lex_keys<<<grid, 256, 0, st>>>(w.c[0].cent, w.stable, nullptr, w.frame, 2, na, w.lkey_a, w.lex_a);
    cub::DeviceRadixSort::SortPairs(w.cub, cb, w.lkey_a, w.lkey_b, w.lex_a, w.lex_b, na, 0, 64, st);
    k_pw_lex_keys<<<grid, 256, 0, st>>>(w.c[0].cent, w.stable, w.lex_b, w.frame, 1, na, w.lkey_a, w.lex_a);
    cb = w.cub_bytes;
    cub::DeviceRadixSort::SortPairs(w.cub, cb, w.lkey_a, w.lkey_b, w.lex_a, w.lex_b, na, 0, 64, st);
    k_pw_lex_keys<<<grid, 256, 0, st>>>(w.c[0].cent, w.stable, w.lex_b, w.frame, 0, na, w.lkey_a, w.lex_a);
    cb = w.cub_bytes;
    cub::DeviceRadixSort::SortPairs(w.cub, cb, w.lkey_a, w.lkey_b, w.lex_a, w.lex_b, na, 0, 64, st);
    k_pw_unstable_flag<<<grid, 256, 0, st>>>(w.stable, w.frame, w.uflag);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub, cb, w.uflag, w.uscan, na, st);
    k_pw_out_order<<<grid, 256, 0, st>>>(w.lex_b, w.stable, w.uscan, w.c[0].cell_leaf, w.c[0].leaf_start, w.frame, w.out_cell,
                                        w.out_cnt);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub, cb, w.out_cnt, w.out_off, na, st);
    f4l_count_launches(10);
    f4l_mark("k_pw_emit", st);
    k_pw_emit<<<f4l_div_up(na, 8), 256, 0, st>>>(src64, n_src, w.frame, w.out_cell, w.out_cnt, w.out_off, w.c[0].cell_leaf,
                                                w.c[0].leaf_start, w.c[0].order, w.stable, w.c[0].cent, w.c[1].cent, w.nn, dvfs,
                                                mag, w.frame);
    k_pw_counts<<<1, 1, 0, st>>>(w.frame, counts, thr_out);
    f4l_count_launches(1);
    if (cent_src) cudaMemcpyAsync(cent_src, w.c[0].cent, (size_t)na * 24, cudaMemcpyDeviceToDevice, st);
    if (cent_tgt) cudaMemcpyAsync(cent_tgt, w.c[1].cent, (size_t)(n_tgt + 8) * 24, cudaMemcpyDeviceToDevice, st);
    if (nn_out) cudaMemcpyAsync(nn_out, w.nn, (size_t)na * 4, cudaMemcpyDeviceToDevice, st);
    return f4l_finish("f4l_piecewise_icp", stream);
}
