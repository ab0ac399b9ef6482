// A tiny uniform grid over the targets of ONE patch, resident in shared memory next to them.
//
// The 1-NN assignment (A4, base.py:48-97) and the ICP correspondence step (E1) scan every target of a
// patch for every source point: n_s * n_t distance evaluations, the bulk of the CUDA-core work of the
// path.  Patches are surface pieces, so their targets are binned on the two widest axes of their bounding
// box (<= 16 x 16 cells, ~3 targets per cell), stored cell by cell (x fastest: the three cells of a grid
// row a query visits are ONE contiguous range of float4 {x, y, z, bits(original index)}), and a query
// visits rings of cells until the caller's certainty test holds against
//     d2' = min(second best distance seen, squared distance to the border of the visited block)
// -- every unvisited target is at least that border distance away, so d2' is a valid lower bound for
// all targets other than the best one and the f32 error-margin argument of the callers is unchanged.
// When the rings cover the whole grid the scan has degenerated into the full scan.  Results are
// therefore identical to the brute-force scan, candidate order does not matter (exact f32 ties are
// never `certain` and go to the callers' fp64 path).
#pragma once
#include "common.cuh"

struct PatchGrid {
    float minu, minv, cell, inv;
    int ncu, ncv, au, av;
};

__device__ __forceinline__ float pg_pick(float x, float y, float z, int a) { return a == 0 ? x : (a == 1 ? y : z); }

// mn / mx: bounding box of the (pivot-local) targets; gmax: cells per axis at most (<= 16)
__device__ __forceinline__ PatchGrid make_patch_grid(const float mn[3], const float mx[3], int nt, int gmax) {
    PatchGrid g;
    const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
    int thin = 0;
    float et = ex;
    if (ey < et) { thin = 1; et = ey; }
    if (ez < et) { thin = 2; et = ez; }
    g.au = thin == 0 ? 1 : 0;
    g.av = thin == 2 ? 1 : 2;
    const float eu = g.au == 0 ? ex : ey, ev = g.av == 2 ? ez : ey;
    g.minu = g.au == 0 ? mn[0] : mn[1];
    g.minv = g.av == 2 ? mn[2] : mn[1];
    int G = (int)sqrtf((float)nt * (1.0f / 3.0f));
    G = min(max(G, 1), gmax);
    const float longest = fmaxf(eu, ev);
    if (!(longest > 0.f) || !(longest < INFINITY)) {          // a single point (or garbage): one cell
        g.cell = 1.f; g.inv = 0.f; g.ncu = 1; g.ncv = 1;
        return g;
    }
    g.cell = longest / (float)G;
    g.inv = (float)G / longest;
    g.ncu = min(max((int)ceilf(eu * g.inv), 1), G);
    g.ncv = min(max((int)ceilf(ev * g.inv), 1), G);
    return g;
}

__device__ __forceinline__ void pg_cell_uv(const PatchGrid& g, float u, float v, int& cu, int& cv) {
    cu = min(max((int)floorf((u - g.minu) * g.inv), 0), g.ncu - 1);
    cv = min(max((int)floorf((v - g.minv) * g.inv), 0), g.ncv - 1);
}

__device__ __forceinline__ int pg_cell(const PatchGrid& g, float x, float y, float z) {
    int cu, cv;
    pg_cell_uv(g, pg_pick(x, y, z, g.au), pg_pick(x, y, z, g.av), cu, cv);
    return cv * g.ncu + cu;
}

__device__ __forceinline__ void pg_scan(const float4* __restrict__ pts, int b, int e, float qx, float qy, float qz,
                                        float& d1, int& j1, float& d2) {
    for (int j = b; j < e; ++j) {
        const float4 c = pts[j];
        const float dx = qx - c.x, dy = qy - c.y, dz = qz - c.z;
        const float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const bool better = dd < d1;
        d2 = better ? d1 : fminf(d2, dd);
        j1 = better ? __float_as_int(c.w) : j1;
        d1 = better ? dd : d1;
    }
}

// Top-2 of one query over the binned targets.  start[c] .. start[c + 1]: the items of cell c (x fastest).
// certain(d1, d2') -> bool is the caller's f32 error-margin test.  Returns its last verdict; d2 is the
// bound it was evaluated with (the true second-best distance once the whole grid has been visited).
template <class StartT, class Certain>
__device__ __forceinline__ bool pg_top2(const PatchGrid& g, const float4* __restrict__ pts,
                                        const StartT* __restrict__ start, float qx, float qy, float qz,
                                        Certain certain, float& d1, int& j1, float& d2) {
    const float qu = pg_pick(qx, qy, qz, g.au), qv = pg_pick(qx, qy, qz, g.av);
    int cu, cv;
    pg_cell_uv(g, qu, qv, cu, cv);
    const float slack = 1e-5f + 2e-6f * (fabsf(qu) + fabsf(qv));
    d1 = INFINITY; d2 = INFINITY; j1 = 0;
    for (int R = 1;; ++R) {
        const int v0 = max(cv - R, 0), v1 = min(cv + R, g.ncv - 1);
        const int u0 = max(cu - R, 0), u1 = min(cu + R, g.ncu - 1);
        for (int v = v0; v <= v1; ++v) {
            const int row = v * g.ncu;
            if (R == 1 || v == cv - R || v == cv + R) {
                pg_scan(pts, (int)start[row + u0], (int)start[row + u1 + 1], qx, qy, qz, d1, j1, d2);
            } else {
                if (cu - R >= 0) pg_scan(pts, (int)start[row + cu - R], (int)start[row + cu - R + 1], qx, qy, qz, d1, j1, d2);
                if (cu + R <= g.ncu - 1) pg_scan(pts, (int)start[row + cu + R], (int)start[row + cu + R + 1], qx, qy, qz, d1, j1, d2);
            }
        }
        // distance from the query to the border of the visited block; a side on the border of the grid
        // is infinitely far (no target lies outside the bounding box)
        float margin = INFINITY;
        if (cu - R > 0) margin = fminf(margin, qu - (g.minu + (float)(cu - R) * g.cell));
        if (cu + R < g.ncu - 1) margin = fminf(margin, (g.minu + (float)(cu + R + 1) * g.cell) - qu);
        if (cv - R > 0) margin = fminf(margin, qv - (g.minv + (float)(cv - R) * g.cell));
        if (cv + R < g.ncv - 1) margin = fminf(margin, (g.minv + (float)(cv + R + 1) * g.cell) - qv);
        if (margin == INFINITY) return certain(d1, d2);
        margin = fmaxf(margin - slack, 0.f);
        const float d2e = fminf(d2, margin * margin);
        if (certain(d1, d2e)) {
            d2 = d2e;
            return true;
        }
    }
}
