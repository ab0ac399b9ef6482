// Uniform-grid pieces shared by the general kNN (knn.cu) and the fused A1 pipeline (medres.cu).
#pragma once
#include "common.cuh"

struct GridParams {
    float minx, miny, minz;
    float inv_cell, cell;
    float inv[3];        // per-axis 1/cell; 0 on a flattened axis (all points fall in layer 0)
    int nx, ny, nz;
    int ncells;
    int pad;
};

// one thread: choose the cell size (surface-like data: `factor` x the mean spacing on the largest face
// of the box) and clamp the table to max_cells.
__device__ inline void grid_params_compute(const float mn[3], const float mx[3], int m, float cell_in, float factor,
                                           int max_cells, GridParams* gp) {
    float ex[3];
    for (int a = 0; a < 3; ++a) ex[a] = fmaxf(mx[a] - mn[a], 0.f);
    float face = fmaxf(ex[0] * ex[1], fmaxf(ex[0] * ex[2], ex[1] * ex[2]));
    float c = cell_in > 0.f ? cell_in : factor * sqrtf(face / fmaxf((float)m, 1.f));
    float longest = fmaxf(ex[0], fmaxf(ex[1], ex[2]));
    if (!(c > 0.f)) c = fmaxf(longest, 1e-3f);
    c = fmaxf(c, longest * 1e-4f);   // never more than 10^4 cells per axis
    // surface-like clouds (terrain scans): the thin axis is not binned at all -- a 2D table is 10-50x
    // smaller (memset / scan / cell lookups) and a query visits 3 row ranges per ring instead of 9.
    // The ring test below only ever uses binned axes, so the search stays exact.
    int thin = 0;
    if (ex[1] < ex[thin]) thin = 1;
    if (ex[2] < ex[thin]) thin = 2;
    float mid = INFINITY;
    for (int a = 0; a < 3; ++a)
        if (a != thin) mid = fminf(mid, ex[a]);
    const bool flat = ex[thin] <= 0.25f * mid;
    int nn[3];
    for (int it = 0; it < 64; ++it) {
        for (int a = 0; a < 3; ++a) nn[a] = (flat && a == thin) ? 1 : (int)floorf(ex[a] / c) + 1;
        if ((long long)nn[0] * nn[1] * nn[2] <= (long long)max_cells) break;
        c *= 1.2599211f;
    }
    gp->minx = mn[0]; gp->miny = mn[1]; gp->minz = mn[2];
    gp->cell = c; gp->inv_cell = 1.0f / c;
    for (int a = 0; a < 3; ++a) gp->inv[a] = (flat && a == thin) ? 0.f : 1.0f / c;
    gp->nx = nn[0]; gp->ny = nn[1]; gp->nz = nn[2];
    gp->ncells = nn[0] * nn[1] * nn[2];
}

__device__ __forceinline__ void cell_of(const GridParams& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = min(max((int)floorf((x - g.minx) * g.inv[0]), 0), g.nx - 1);
    cy = min(max((int)floorf((y - g.miny) * g.inv[1]), 0), g.ny - 1);
    cz = min(max((int)floorf((z - g.minz) * g.inv[2]), 0), g.nz - 1);
}

// ---- search --------------------------------------------------------------------------------
// The k best candidates are kept as 64-bit keys (bits(d^2) << 32 | original index): squared distances are
// non-negative floats, whose bit patterns order like unsigned integers, so ONE unsigned 64-bit compare is the
// lexicographic (distance, index) order -- deterministic under any visiting order, ties towards the lower
// index -- and an insertion is a branch-free min/max chain.
typedef unsigned long long u64;
#define KNN_KEY_NONE 0x7f800000ffffffffull      // (inf, -1)

template <int K>
struct TopK {
    u64 v[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int a = 0; a < K; ++a) v[a] = KNN_KEY_NONE;
    }
    __device__ __forceinline__ void push(u64 key) {
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const u64 lo = key < v[a] ? key : v[a];
            key = key < v[a] ? v[a] : key;
            v[a] = lo;
        }
    }
    __device__ __forceinline__ void push(float dd, float w) { push(((u64)__float_as_uint(dd) << 32) | (u64)__float_as_uint(w)); }
    __device__ __forceinline__ float kth(int kk) const { return d(kk); }
    __device__ __forceinline__ float d(int a) const { return __uint_as_float((unsigned)(v[a] >> 32)); }
    __device__ __forceinline__ int i(int a) const { return (int)(unsigned)(v[a] & 0xffffffffull); }
};


// distance-only top-K (the consumer needs the k-th distance, not who it is): an insertion is K fmin/fmax pairs
template <int K>
struct TopD {
    float v[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int a = 0; a < K; ++a) v[a] = INFINITY;
    }
    __device__ __forceinline__ void push(float x) {
#pragma unroll
        for (int a = 0; a < K; ++a) {
            const float lo = fminf(v[a], x);
            if (a + 1 < K) x = fmaxf(v[a], x);
            v[a] = lo;
        }
    }
    __device__ __forceinline__ void push(float dd, float) { push(dd); }
    __device__ __forceinline__ float kth(int kk) const { return v[kk]; }
};

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int knn_max_cells(int M) {
    long long c = (long long)M / 2 + 4096;      // auto cell size: ~M/4 cells on surface-like clouds
    if (c > (1LL << 27)) c = 1LL << 27;
    return (int)c;
}


template <class Top>
__device__ __forceinline__ void scan_range(const float4* __restrict__ sorted, int b, int e, float qx, float qy,
                                           float qz, Top& tk) {
#pragma unroll 4
    for (int j = b; j < e; ++j) {
        const float4 c = __ldg(sorted + j);
        const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
        tk.push(fmaf(dz, dz, fmaf(dy, dy, dx * dx)), c.w);
    }
}

// squared distance from the query to the border of the block of cells within R of (cx,cy,cz), shrunk by a few
// ulps (cell borders are computed in f32); INFINITY when the block covers the grid (no reference point lies
// outside the grid's bounding box).
__device__ __forceinline__ float ring_margin2(const GridParams& g, float qx, float qy, float qz, int cx, int cy,
                                              int cz, int R) {
    float margin = INFINITY;
    if (cx - R > 0) margin = fminf(margin, qx - (g.minx + (float)(cx - R) * g.cell));
    if (cx + R < g.nx - 1) margin = fminf(margin, (g.minx + (float)(cx + R + 1) * g.cell) - qx);
    if (cy - R > 0) margin = fminf(margin, qy - (g.miny + (float)(cy - R) * g.cell));
    if (cy + R < g.ny - 1) margin = fminf(margin, (g.miny + (float)(cy + R + 1) * g.cell) - qy);
    if (cz - R > 0) margin = fminf(margin, qz - (g.minz + (float)(cz - R) * g.cell));
    if (cz + R < g.nz - 1) margin = fminf(margin, (g.minz + (float)(cz + R + 1) * g.cell) - qz);
    if (margin == INFINITY) return INFINITY;
    // f32 error budget: the border positions (min + c * cell) round with |coordinate|, the binning
    // floor((x - min) * inv) with the extent of the grid -- both terms, so that a reference point binned one cell
    // off its geometric cell can never lie beyond a ring that was declared final
    const float extent = ((float)g.nx + (float)g.ny + (float)g.nz) * g.cell;
    margin = fmaxf(margin - 4e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + extent + g.cell), 0.f);
    return margin * margin;
}

// Rings of cells around the query's cell (cx,cy,cz) until the kk-th best distance is provably final: strictly
// below the distance to the border of the searched block, or the block covers the radius / the whole grid.
// cell_start[c] .. cell_start[c+1] is the range of cell c in `sorted`; the cells of a grid row are contiguous,
// so ring 1 (the whole 3 x 3 (x 3) block) is one range per row: its (up to 9) range bounds are fetched up
// front -- 18 independent loads in flight instead of a load -> scan -> load chain per row.
template <class Top>
__device__ __forceinline__ void grid_ring_search(const float4* __restrict__ sorted, const int* __restrict__ cell_start,
                                                 const GridParams& g, float qx, float qy, float qz, int cx, int cy,
                                                 int cz, int kk, float max_r2, Top& tk) {
    {
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1) + 1;
        int rb[9], re[9];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int z = cz + dz - 1, y = cy + dy - 1;
                const bool ok = z >= 0 && z < g.nz && y >= 0 && y < g.ny;
                const int row = (z * g.ny + y) * g.nx;
                rb[dz * 3 + dy] = ok ? __ldg(cell_start + row + x0) : 0;
                re[dz * 3 + dy] = ok ? __ldg(cell_start + row + x1) : 0;
            }
        }
        scan_range(sorted, rb[4], re[4], qx, qy, qz, tk);       // the query's own row first
#pragma unroll
        for (int r = 0; r < 9; ++r)
            if (r != 4) scan_range(sorted, rb[r], re[r], qx, qy, qz, tk);
        const float m2 = ring_margin2(g, qx, qy, qz, cx, cy, cz, 1);
        if (tk.kth(kk - 1) < m2 || m2 >= max_r2) return;
    }
    const int maxR = max(g.nx, max(g.ny, g.nz));
    for (int R = 2; R <= maxR; ++R) {
        const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
        const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
        for (int z = z0; z <= z1; ++z) {
            const bool zshell = (z == cz - R) || (z == cz + R);
            for (int y = y0; y <= y1; ++y) {
                const bool shell = zshell || (y == cy - R) || (y == cy + R);
                const int row = (z * g.ny + y) * g.nx;
                if (shell) {
                    scan_range(sorted, __ldg(cell_start + row + x0), __ldg(cell_start + row + x1 + 1), qx, qy, qz, tk);
                } else {
                    if (cx - R >= 0)
                        scan_range(sorted, __ldg(cell_start + row + cx - R), __ldg(cell_start + row + cx - R + 1), qx, qy, qz, tk);
                    if (cx + R <= g.nx - 1)
                        scan_range(sorted, __ldg(cell_start + row + cx + R), __ldg(cell_start + row + cx + R + 1), qx, qy, qz, tk);
                }
            }
        }
        const float m2 = ring_margin2(g, qx, qy, qz, cx, cy, cz, R);
        if (tk.kth(kk - 1) < m2 || m2 >= max_r2) break;
    }
}

// cell edge = factor x mean point spacing (on the largest face of the bounding box) when the caller passes
// cell <= 0.  F4L_KNN_CELL_FACTOR overrides the tuned default (experiments only).
float f4l_knn_cell_factor();
