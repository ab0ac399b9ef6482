// DIPs patch front-end: src/data_loader.py:16-109 Preprocess_Dataset.extract_patch, the
// per-point Python loop that feeds the descriptor network in src/f2s3.py:104-134 and base.py:1981-2034.
//
// For every query point: all reference points within `radius` (Open3D KDTreeFlann.search_radius_vector_3d:
// d^2 < r^2 in fp64, sorted by distance), the local reference frame of data_loader.py:46-80 (normal = eigenvector
// of the smallest eigenvalue of the neighbour covariance, sign by the neighbour centroid; x axis = weighted mean of
// the tangential parts), the neighbours expressed in that frame and divided by the radius, and `num_points` of
// them (zero rows when there are fewer) -> (n, 3, num_points) f32.
//
// Layout: reference points are binned once per cloud into a uniform grid (cell edge = radius/4, the thin axis of
// a surface-like cloud is not binned) as double4 {x, y, z, bits(original index)} rows in cell order, so the cells
// of one grid row that a query's ball overlaps are ONE contiguous range.  One warp owns a query: pass 1 walks the
// ranges with 32 lanes, keeps the hits as positions in a shared-memory list and accumulates the fp64 moments;
// pass 2 (x axis) and pass 3 (output) walk the list.  All arithmetic on coordinates is fp64 like the reference's
// numpy; d^2 is formed with separate multiplies and adds in nanoflann's order, so ball membership is bit-identical.
//
// Which rows are kept: the reference calls np.random.choice(n, num_points, replace=False) on the GLOBAL numpy
// generator inside DataLoader workers -- not reproducible even by the reference itself.  Two modes:
//   ranks == nullptr: slot t takes list entry pi(t), pi a keyed bijection of [0, max(n, num_points)) (4-round
//                     Feistel network with cycle walking): a uniform sample without replacement in random order;
//   ranks != nullptr: slot t takes the neighbour whose DISTANCE RANK is ranks[q][t] (what the reference does with
//                     its `inds`), ties by original index -- the parity tests pass numpy's own choice() output.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

#define DIPS_WARPS 4
#ifndef DIPS_CAP
#define DIPS_CAP 1408          // neighbours per query kept on chip (the reference's radius rule yields ~940)
#endif
#ifndef DIPS_MIN_BLOCKS
#define DIPS_MIN_BLOCKS 8      // 64 registers, 6.5 KB of shared memory per warp: 32 warps per SM (measured: 8.2 -> 6.5 ms per tile)
#endif
#define DIPS_MAXP 256
#define DIPS_CAP_LARGE 8192    // the dense-cloud variant (f4l_dips_patches_large): one CTA of 4 warps per SM, 33 KB per warp

struct DipsGrid {
    unsigned long long lo[3], hi[3];   // ordered-uint images of the bounding box (atomicMin / atomicMax)
    double org[3];
    double cell, inv[3];
    int n[3];
    int ncells;
    int pad;
};

__device__ __forceinline__ unsigned long long d2ord(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2d(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

__global__ void k_dips_init(DipsGrid* g) {
    if (threadIdx.x < 3) { g->lo[threadIdx.x] = ~0ull; g->hi[threadIdx.x] = 0ull; }
}

__global__ void __launch_bounds__(256) k_dips_bbox(const double* __restrict__ p, int n, DipsGrid* g) {
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = p[(size_t)i * 3 + c];
            mn[c] = fmin(mn[c], v);
            mx[c] = fmax(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(F4L_FULL, mn[c], o));
            mx[c] = fmax(mx[c], __shfl_xor_sync(F4L_FULL, mx[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (mn[c] <= mx[c]) {
                atomicMin(&g->lo[c], d2ord(mn[c]));
                atomicMax(&g->hi[c], d2ord(mx[c]));
            }
        }
    }
}

__global__ void k_dips_params(DipsGrid* g, double radius, int max_cells) {
    if (threadIdx.x != 0) return;
    double lo[3], ex[3];
    for (int c = 0; c < 3; ++c) {
        lo[c] = ord2d(g->lo[c]);
        ex[c] = fmax(ord2d(g->hi[c]) - lo[c], 0.0);
    }
    int thin = 0;
    if (ex[1] < ex[thin]) thin = 1;
    if (ex[2] < ex[thin]) thin = 2;
    double mid = INFINITY;
    for (int c = 0; c < 3; ++c)
        if (c != thin) mid = fmin(mid, ex[c]);
    const bool flat = ex[thin] <= 0.25 * mid;
    double cell = 0.25 * radius;
    int nn[3];
    for (int it = 0; it < 64; ++it) {
        for (int c = 0; c < 3; ++c) nn[c] = (flat && c == thin) ? 1 : (int)fmin(floor(ex[c] / cell), 2.0e9) + 1;
        if ((double)nn[0] * nn[1] * nn[2] <= (double)max_cells) break;
        cell *= 1.25;
    }
    g->cell = cell;
    for (int c = 0; c < 3; ++c) {
        g->org[c] = lo[c];
        g->inv[c] = (flat && c == thin) ? 0.0 : 1.0 / cell;
        g->n[c] = nn[c];
    }
    g->ncells = nn[0] * nn[1] * nn[2];
}

__device__ __forceinline__ int dips_cell1(const DipsGrid& g, int c, double v) {
    const double f = floor((v - g.org[c]) * g.inv[c]);
    return (int)fmin(fmax(f, 0.0), (double)(g.n[c] - 1));
}

// cell id of every point + the cell histogram
__global__ void __launch_bounds__(256) k_dips_count(const double* __restrict__ p, int n, const DipsGrid* __restrict__ gp,
                                                    int* __restrict__ table, int* __restrict__ cell_of_pt,
                                                    int* __restrict__ iota) {
    const DipsGrid g = *gp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int cx = dips_cell1(g, 0, p[(size_t)i * 3]), cy = dips_cell1(g, 1, p[(size_t)i * 3 + 1]),
                  cz = dips_cell1(g, 2, p[(size_t)i * 3 + 2]);
        const int c = (cz * g.n[1] + cy) * g.n[0] + cx;
        cell_of_pt[i] = c;
        iota[i] = i;
        atomicAdd(table + c, 1);
    }
}

// rows in cell order; inside a cell in ORIGINAL INDEX order (the pairs were sorted by a stable radix sort), so the
// layout -- and with it the hit lists and the random-sample mode -- does not depend on atomics order
__global__ void __launch_bounds__(256) k_dips_gather(const double* __restrict__ p, int n, const int* __restrict__ order,
                                                     double4* __restrict__ sorted) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = order[j];
        sorted[j] = make_double4(p[(size_t)i * 3], p[(size_t)i * 3 + 1], p[(size_t)i * 3 + 2], __longlong_as_double((long long)i));
    }
}

// ---- the per-query kernel ------------------------------------------------------------------------
#define DIPS_ROWCAP 128
template <int CAP>
struct DipsWarpSmem {
    unsigned list[CAP];             // positions of the hits in `sorted`
    int rows[2 * DIPS_ROWCAP];      // pass 1: [begin | end) of the candidate range of every grid row the ball touches
};
struct DipsRankSmem {               // ranked mode only
    double d2[DIPS_CAP];
    int idx[DIPS_CAP];
    unsigned short by_rank[DIPS_CAP];
};

__device__ __forceinline__ unsigned dips_mix(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// keyed bijection of [0, m): 3-round Feistel network on 2*hb bits (2^(2 hb) >= m > 2^(2 hb - 2)), cycle walking;
// round function = high bits of a keyed multiply (one IMAD + shift per round)
__device__ __forceinline__ unsigned dips_perm(unsigned t, unsigned m, int hb, unsigned key) {
    const unsigned mask = (1u << hb) - 1u;
    const unsigned k0 = key | 1u, k1 = dips_mix(key) | 1u, k2 = dips_mix(key ^ 0x9e3779b9u) | 1u;
    unsigned v = t;
    do {
        unsigned l = v >> hb, r = v & mask;
        l ^= (((r + 1u) * k0) >> 13) & mask;
        r ^= (((l + 1u) * k1) >> 13) & mask;
        l ^= (((r + 1u) * k2) >> 13) & mask;
        v = (l << hb) | r;
    } while (v >= m);
    return v;
}

__device__ __forceinline__ double dips_d2(double dx, double dy, double dz) {
    // nanoflann L2_Simple_Adaptor: result += diff * diff per dimension, no contraction
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// sqrt in fp64 from an f32 reciprocal square root and two Newton steps (relative error ~1e-15): the frame weights
// (radius - dist)^2 need fp64-level accuracy but not the IEEE-rounded root
__device__ __forceinline__ double dips_sqrt(double x) {
    if (!(x > 1e-280)) return 0.0;
    double y = (double)rsqrtf((float)x);
    y = y * (1.5 - 0.5 * x * y * y);
    y = y * (1.5 - 0.5 * x * y * y);
    return x * y;
}

// distance from v to the interval [a, b] (0 inside); first / last cells are open towards the outside
__device__ __forceinline__ double dips_slab(double v, double a, double b, bool first, bool last) {
    double d = 0.0;
    if (v < a && !first) d = a - v;
    if (v > b && !last) d = v - b;
    return d;
}

template <bool RANKED, int CAP, int MIN_BLOCKS>
__global__ void __launch_bounds__(DIPS_WARPS * 32, MIN_BLOCKS)
k_dips_patches(const double* __restrict__ query, int nq, const double4* __restrict__ sorted,
               const int* __restrict__ cell_start, const DipsGrid* __restrict__ gp, double radius, int num_points,
               const int32_t* __restrict__ ranks, unsigned long long seed, float* __restrict__ patches,
               double* __restrict__ lrf, int32_t* __restrict__ count) {
    extern __shared__ __align__(16) unsigned char dips_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    DipsWarpSmem<CAP>& sm = reinterpret_cast<DipsWarpSmem<CAP>*>(dips_raw)[wid];
    DipsRankSmem* rk = RANKED ? reinterpret_cast<DipsRankSmem*>(dips_raw + DIPS_WARPS * sizeof(DipsWarpSmem<CAP>)) + wid : nullptr;
    const DipsGrid g = *gp;
    const double r2 = radius * radius;
    const double inv_r = 1.0 / radius;
    for (int q = blockIdx.x * DIPS_WARPS + wid; q < nq; q += gridDim.x * DIPS_WARPS) {
        const double qx = query[(size_t)q * 3], qy = query[(size_t)q * 3 + 1], qz = query[(size_t)q * 3 + 2];
        // ---- pass 1: hits, moments, nearest --------------------------------------------------------
        int lo[3], hi[3];
        {
            const double qq[3] = {qx, qy, qz};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // a margin of one cell-size ulp keeps points that sit exactly on a cell border inside the block
                lo[c] = (int)fmin(fmax(floor((qq[c] - radius - g.org[c]) * g.inv[c] - 1e-9), 0.0), (double)(g.n[c] - 1));
                hi[c] = (int)fmin(fmax(floor((qq[c] + radius - g.org[c]) * g.inv[c] + 1e-9), 0.0), (double)(g.n[c] - 1));
            }
        }
        int n = 0;
        double sx = 0, sy = 0, sz = 0, cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0;
        double best_d2 = INFINITY;
        unsigned best_pos = 0xffffffffu;
        // one candidate: membership test, moments, nearest (ties between duplicates: any of them, they are equal),
        // ordered compaction of the hits into the list
        auto visit = [&](int j, bool valid, const double4& p) {
            bool hit = false;
            if (valid) {
                const double dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
                const double d2 = dips_d2(dx, dy, dz);
                hit = d2 < r2;                                           // RadiusResultSet::addPoint: dist < radius
                if (hit) {
                    sx += dx; sy += dy; sz += dz;
                    cxx += dx * dx; cxy += dx * dy; cxz += dx * dz; cyy += dy * dy; cyz += dy * dz; czz += dz * dz;
                    if (d2 < best_d2) { best_d2 = d2; best_pos = (unsigned)j; }
                }
            }
            const unsigned m = __ballot_sync(F4L_FULL, hit);
            if (hit) {
                const int slot = n + __popc(m & ((1u << lane) - 1u));
                if (slot < CAP) sm.list[slot] = (unsigned)j;
            }
            n += __popc(m);
        };
        const int ny = hi[1] - lo[1] + 1, nrows = ny * (hi[2] - lo[2] + 1);
        if (nrows <= DIPS_ROWCAP) {
            // the lanes trim the rows in parallel: a row's cells can only hold hits where the ball is at least as
            // wide as the row's distance from the query allows
            __syncwarp();
            for (int r = lane; r < nrows; r += 32) {
                const int z = lo[2] + r / ny, y = lo[1] + r % ny;
                double w2 = r2;
                if (g.inv[1] != 0.0) {
                    const double d = dips_slab(qy, g.org[1] + y * g.cell, g.org[1] + (y + 1) * g.cell, y == 0, y == g.n[1] - 1);
                    w2 -= d * d;
                }
                if (g.inv[2] != 0.0) {
                    const double d = dips_slab(qz, g.org[2] + z * g.cell, g.org[2] + (z + 1) * g.cell, z == 0, z == g.n[2] - 1);
                    w2 -= d * d;
                }
                int b = 0, e = 0;
                if (w2 > -1e-9 * r2) {
                    const double w = sqrt(fmax(w2, 0.0)) + 1e-9 * radius;
                    const int x0 = (int)fmin(fmax(floor((qx - w - g.org[0]) * g.inv[0] - 1e-9), 0.0), (double)(g.n[0] - 1));
                    const int x1 = (int)fmin(fmax(floor((qx + w - g.org[0]) * g.inv[0] + 1e-9), 0.0), (double)(g.n[0] - 1));
                    const size_t row = ((size_t)z * g.n[1] + y) * g.n[0];
                    b = __ldg(cell_start + row + x0);
                    e = __ldg(cell_start + row + x1 + 1);
                }
                sm.rows[r] = b;
                sm.rows[DIPS_ROWCAP + r] = e;
            }
            __syncwarp();
            for (int r = 0; r < nrows; ++r) {
                const int b = sm.rows[r], e = sm.rows[DIPS_ROWCAP + r];
                for (int j0 = b; j0 < e; j0 += 64) {              // two loads in flight per lane
                    const int j1 = j0 + lane, j2 = j1 + 32;
                    const bool v1 = j1 < e, v2 = j2 < e;
                    double4 p1 = make_double4(0, 0, 0, 0), p2 = p1;
                    if (v1) p1 = sorted[j1];
                    if (v2) p2 = sorted[j2];
                    visit(j1, v1, p1);
                    if (j0 + 32 < e) visit(j2, v2, p2);
                }
            }
            __syncwarp();
        } else {
            for (int z = lo[2]; z <= hi[2]; ++z) {
                for (int y = lo[1]; y <= hi[1]; ++y) {
                    const size_t row = ((size_t)z * g.n[1] + y) * g.n[0];
                    const int b = __ldg(cell_start + row + lo[0]), e = __ldg(cell_start + row + hi[0] + 1);
                    for (int j0 = b; j0 < e; j0 += 32) {
                        const int j1 = j0 + lane;
                        double4 p1 = make_double4(0, 0, 0, 0);
                        if (j1 < e) p1 = sorted[j1];
                        visit(j1, j1 < e, p1);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) count[q] = n;
        float* outq = patches + (size_t)q * 3 * num_points;
        if (n > CAP) {
            // more neighbours than this variant keeps on chip: the patch is left zero and count[q] reports the size;
            // the host mirror re-runs such queries through f4l_dips_patches_large (CAP 8192) and raises beyond that
            for (int t = lane; t < 3 * num_points; t += 32) outq[t] = 0.f;
            if (lrf && lane < 9) lrf[(size_t)q * 9 + lane] = 0.0;
            continue;
        }
        // nearest neighbour of the warp (the reference drops sorted entry 0 from the frame estimate)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(F4L_FULL, best_d2, o);
            const unsigned op = __shfl_xor_sync(F4L_FULL, best_pos, o);
            if (od < best_d2 || (od == best_d2 && op < best_pos)) { best_d2 = od; best_pos = op; }
        }
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};     // rows: xp, yp, zp (= lRg^T)
        const bool framed = n > 10;                    // data_loader.py:45  ptall.shape[1] > 10
        if (framed) {
            sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
            cxx = warp_sum(cxx); cxy = warp_sum(cxy); cxz = warp_sum(cxz);
            cyy = warp_sum(cyy); cyz = warp_sum(cyz); czz = warp_sum(czz);
            const double4 pn = sorted[best_pos];
            const double nx = pn.x - qx, ny = pn.y - qy, nz = pn.z - qz;
            sx -= nx; sy -= ny; sz -= nz;
            // data_loader.py:50  cov = 1/3 * vect_diff vect_diff^T  (ptnn.shape[0] is 3: the scale is irrelevant)
            const double k3 = 1.0 / 3.0;
            double H[9];
            H[0] = k3 * (cxx - nx * nx); H[1] = k3 * (cxy - nx * ny); H[2] = k3 * (cxz - nx * nz);
            H[4] = k3 * (cyy - ny * ny); H[5] = k3 * (cyz - ny * nz); H[8] = k3 * (czz - nz * nz);
            H[3] = H[1]; H[6] = H[2]; H[7] = H[5];
            double U[9], S[3], V[9];
            svd3x3(H, U, S, V);                        // symmetric PSD: singular vectors = eigenvectors
            double zx = V[2], zy = V[5], zz = V[8];    // smallest eigenvalue (data_loader.py:53-55)
            // data_loader.py:58  zp = np_hat if sum(np_hat . (-vect_diff)) > 0 else -np_hat
            if (!(-(zx * sx + zy * sy + zz * sz) > 0.0)) { zx = -zx; zy = -zy; zz = -zz; }
            // ---- pass 2: x axis = normalised sum of alpha beta v over the neighbours (data_loader.py:60-72) ------
            double ax = 0, ay = 0, az = 0;
            auto axis_term = [&](unsigned pos, const double4& p) {
                if (pos == best_pos) return;
                const double dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
                const double pz = dx * zx + dy * zy + dz * zz;
                const double a = radius - dips_sqrt(dips_d2(dx, dy, dz));
                const double w = (a * a) * (pz * pz);
                ax += w * (dx - pz * zx); ay += w * (dy - pz * zy); az += w * (dz - pz * zz);
            };
            for (int t = lane; t < n; t += 64) {                     // two gathers in flight per lane
                const unsigned pos1 = sm.list[t];
                const bool v2 = t + 32 < n;
                const unsigned pos2 = v2 ? sm.list[t + 32] : best_pos;
                const double4 p1 = sorted[pos1];
                const double4 p2 = sorted[pos2];
                axis_term(pos1, p1);
                axis_term(pos2, p2);
            }
            ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
            const double nrm = sqrt(ax * ax + ay * ay + az * az);
            const double sc = (nrm < 1e-6) ? 1.0 / (nrm + 1e-6) : 1.0 / nrm;           // data_loader.py:66-72
            const double xx = ax * sc, xy = ay * sc, xz = az * sc;
            // yp = cross(xp, zp)
            R[0] = xx; R[1] = xy; R[2] = xz;
            R[3] = xy * zz - xz * zy; R[4] = xz * zx - xx * zz; R[5] = xx * zy - xy * zx;
            R[6] = zx; R[7] = zy; R[8] = zz;
        }
        if (lrf && lane < 9) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < 9; ++a)
                if (lane == a) v = R[a];
            lrf[(size_t)q * 9 + lane] = framed ? v : 0.0;
        }
        // ---- ranked mode: distance rank of every hit (ties by original index) ---------------------------
        if (RANKED) {
            for (int t = lane; t < n; t += 32) {
                const double4 p = sorted[sm.list[t]];
                rk->d2[t] = dips_d2(p.x - qx, p.y - qy, p.z - qz);
                rk->idx[t] = (int)__double_as_longlong(p.w);
            }
            __syncwarp();
            for (int t = lane; t < n; t += 32) {
                const double d = rk->d2[t];
                const int id = rk->idx[t];
                int rank = 0;
                for (int u = 0; u < n; ++u) {
                    const double du = rk->d2[u];
                    rank += (du < d || (du == d && rk->idx[u] < id)) ? 1 : 0;
                }
                rk->by_rank[rank] = (unsigned short)t;
            }
            __syncwarp();
        }
        // ---- pass 3: the kept rows ----------------------------------------------------------------------
        const int m = n > num_points ? n : num_points;            // rows after zero padding (data_loader.py:99-100)
        int hb = 1;
        while ((1u << (2 * hb)) < (unsigned)m) ++hb;
        const unsigned key = dips_mix((unsigned)seed ^ dips_mix((unsigned)(seed >> 32) + 0x632be59bu * (unsigned)q));
        for (int t = lane; t < num_points; t += 32) {
            int e;                                                // index into the padded, distance-sorted list
            if (RANKED) {
                e = ranks[(size_t)q * num_points + t];
                e = (e >= 0 && e < n) ? (int)rk->by_rank[e] : -1;
            } else {
                e = (int)dips_perm((unsigned)t, (unsigned)m, hb, key);
                if (e >= n) e = -1;
            }
            float ox = 0.f, oy = 0.f, oz = 0.f;
            if (e >= 0) {
                const double4 p = sorted[sm.list[e]];
                if (framed) {
                    const double dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
                    ox = (float)((R[0] * dx + R[1] * dy + R[2] * dz) * inv_r);
                    oy = (float)((R[3] * dx + R[4] * dy + R[5] * dz) * inv_r);
                    oz = (float)((R[6] * dx + R[7] * dy + R[8] * dz) * inv_r);
                } else {
                    ox = (float)(p.x * inv_r); oy = (float)(p.y * inv_r); oz = (float)(p.z * inv_r);   // data_loader.py:91-94
                }
            }
            outq[t] = ox; outq[num_points + t] = oy; outq[2 * num_points + t] = oz;     // 32 consecutive slots per store
        }
        __syncwarp();
    }
}

// ---- host --------------------------------------------------------------------------------------
static inline size_t dips_align(size_t x) { return (x + 255) / 256 * 256; }
static int dips_max_cells(int m) {
    long long c = (long long)m * 2 + 4096;
    if (c > (1LL << 26)) c = 1LL << 26;
    return (int)c;
}

struct DipsWs {
    DipsGrid* grid;
    int* table;          // table[c] = start of cell c in `sorted`, table[c + 1] = its end (mc + 1 scanned words)
    double4* sorted;
    int *keys_in, *keys_out, *vals_in, *vals_out;
    void* cub_tmp;
    size_t cub_bytes, total;
    int mc, key_bits;
};

static DipsWs dips_layout(void* base, int n_ref) {
    DipsWs w;
    w.mc = dips_max_cells(n_ref);
    w.key_bits = 1;
    while ((1LL << w.key_bits) < (long long)w.mc) ++w.key_bits;
    const size_t n = (size_t)(n_ref > 0 ? n_ref : 1);
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b + off; off += dips_align(bytes); return (void*)p; };
    w.grid = (DipsGrid*)take(sizeof(DipsGrid));
    w.table = (int*)take(((size_t)w.mc + 2) * 4);
    w.sorted = (double4*)take(n * sizeof(double4));
    w.keys_in = (int*)take(n * 4); w.keys_out = (int*)take(n * 4);
    w.vals_in = (int*)take(n * 4); w.vals_out = (int*)take(n * 4);
    size_t cb = 0, cs = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, w.mc + 1);
    cub::DeviceRadixSort::SortPairs(nullptr, cs, (const int*)nullptr, (int*)nullptr, (const int*)nullptr, (int*)nullptr,
                                    (int)n, 0, w.key_bits);
    w.cub_bytes = cb > cs ? cb : cs;
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

extern "C" size_t f4l_dips_workspace_bytes(int32_t n_ref) { return dips_layout(nullptr, n_ref < 0 ? 0 : n_ref).total; }

extern "C" int f4l_dips_build(const double* ref64, int32_t n_ref, double radius, void* workspace, size_t workspace_bytes,
                              void* stream) {
    F4L_REQUIRE(ref64 && workspace, "null pointer");
    F4L_REQUIRE(n_ref >= 1, "empty reference cloud");
    F4L_REQUIRE(radius > 0.0, "radius must be positive");
    const DipsWs w = dips_layout(workspace, n_ref);
    if (workspace_bytes < w.total) {
        f4l_set_error("f4l_dips_build: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    f4l_mark("#memset_dips", st);
    cudaMemsetAsync(w.table, 0, ((size_t)w.mc + 2) * 4, st);
    f4l_mark("k_dips_init", st);
    k_dips_init<<<1, 32, 0, st>>>(w.grid);
    const int blocks = min(f4l_div_up(n_ref, 256), 148 * 8);
    f4l_mark("k_dips_bbox", st);
    k_dips_bbox<<<blocks, 256, 0, st>>>(ref64, n_ref, w.grid);
    f4l_mark("k_dips_params", st);
    k_dips_params<<<1, 32, 0, st>>>(w.grid, radius, w.mc);
    f4l_mark("k_dips_count", st);
    k_dips_count<<<blocks, 256, 0, st>>>(ref64, n_ref, w.grid, w.table, w.keys_in, w.vals_in);
    size_t cb = w.cub_bytes;
    f4l_count_launches(1); f4l_mark("cub_exclusive_scan", st);
    cub::DeviceScan::ExclusiveSum(w.cub_tmp, cb, w.table, w.table, w.mc + 1, st);
    cb = w.cub_bytes;
    f4l_count_launches(3); f4l_mark("cub_radix_sort", st);
    cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const int*)w.keys_in, w.keys_out, (const int*)w.vals_in, w.vals_out,
                                    n_ref, 0, w.key_bits, st);
    f4l_mark("k_dips_gather", st);
    k_dips_gather<<<blocks, 256, 0, st>>>(ref64, n_ref, w.vals_out, w.sorted);
    return f4l_finish("f4l_dips_build", stream);
}

static int dips_patches_launch(const double* query64, int32_t n_query, int32_t n_ref, double radius, int32_t num_points,
                               const int32_t* ranks, uint64_t seed, float* patches, double* lrf, int32_t* count,
                               void* workspace, size_t workspace_bytes, void* stream, bool large, const char* what) {
    F4L_REQUIRE(n_query >= 0, "n_query < 0");
    if (n_query == 0) return F4L_OK;
    F4L_REQUIRE(query64 && patches && count && workspace, "null pointer");
    F4L_REQUIRE(n_ref >= 1, "empty reference cloud");
    F4L_REQUIRE(num_points >= 1 && num_points <= DIPS_MAXP, "num_points must be in [1, 256]");
    F4L_REQUIRE(radius > 0.0, "radius must be positive");
    F4L_REQUIRE(!(large && ranks), "the large variant has no ranked mode");
    const DipsWs w = dips_layout(workspace, n_ref);
    if (workspace_bytes < w.total) {
        f4l_set_error("%s: workspace too small (%zu < %zu)", what, workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = min(f4l_div_up(n_query, DIPS_WARPS), 148 * 16);
    if (large) {
        const size_t smem = DIPS_WARPS * sizeof(DipsWarpSmem<DIPS_CAP_LARGE>);
        static F4lPerDevice once;
        if (!once.done()) {
            if (!f4l_optin_smem(k_dips_patches<false, DIPS_CAP_LARGE, 1>, smem, "k_dips_patches<large>")) return F4L_E_CUDA;
            once.mark();
        }
        f4l_mark("k_dips_patches_large", st);
        k_dips_patches<false, DIPS_CAP_LARGE, 1><<<min(grid, 148), DIPS_WARPS * 32, smem, st>>>(
            query64, n_query, w.sorted, w.table, w.grid, radius, num_points, nullptr, seed, patches, lrf, count);
    } else if (ranks) {
        const size_t smem = DIPS_WARPS * (sizeof(DipsWarpSmem<DIPS_CAP>) + sizeof(DipsRankSmem));
        static F4lPerDevice once;
        if (!once.done()) {
            if (!f4l_optin_smem(k_dips_patches<true, DIPS_CAP, DIPS_MIN_BLOCKS>, smem, "k_dips_patches<ranked>")) return F4L_E_CUDA;
            once.mark();
        }
        f4l_mark("k_dips_patches_ranked", st);
        k_dips_patches<true, DIPS_CAP, DIPS_MIN_BLOCKS><<<grid, DIPS_WARPS * 32, smem, st>>>(
            query64, n_query, w.sorted, w.table, w.grid, radius, num_points, ranks, seed, patches, lrf, count);
    } else {
        const size_t smem = DIPS_WARPS * sizeof(DipsWarpSmem<DIPS_CAP>);
        f4l_mark("k_dips_patches", st);
        k_dips_patches<false, DIPS_CAP, DIPS_MIN_BLOCKS><<<grid, DIPS_WARPS * 32, smem, st>>>(
            query64, n_query, w.sorted, w.table, w.grid, radius, num_points, nullptr, seed, patches, lrf, count);
    }
    return f4l_finish(what, stream);
}

extern "C" int f4l_dips_patches(const double* query64, int32_t n_query, int32_t n_ref, double radius, int32_t num_points,
                                const int32_t* ranks, uint64_t seed, float* patches, double* lrf, int32_t* count,
                                void* workspace, size_t workspace_bytes, void* stream) {
    return dips_patches_launch(query64, n_query, n_ref, radius, num_points, ranks, seed, patches, lrf, count, workspace,
                               workspace_bytes, stream, false, "f4l_dips_patches");
}

extern "C" int f4l_dips_patches_large(const double* query64, int32_t n_query, int32_t n_ref, double radius,
                                      int32_t num_points, uint64_t seed, float* patches, double* lrf, int32_t* count,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    return dips_patches_launch(query64, n_query, n_ref, radius, num_points, nullptr, seed, patches, lrf, count, workspace,
                               workspace_bytes, stream, true, "f4l_dips_patches_large");
}
