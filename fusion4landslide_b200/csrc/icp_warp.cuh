// K-e, small-patch path: one WARP runs the whole ICP loop of a patch pair (<= WICP_CAP points per
// side), so 16+ pairs are in flight per SM, no block barrier is ever taken and the 3x3 SVD is
// computed redundantly by all lanes.
//
// Exact-result work reduction (the reference semantics -- fp64 distances, first minimal index --
// are preserved, see oracle/icp.py):
//   * nearest-neighbour scans run in f32 on patch-local coordinates (targets are exactly
//     representable after subtracting the pivot), keep the two smallest distances, and fall back
//     to an exact fp64 scan whenever the two are closer than the f32 error margin
//   * triangle-inequality pruning: a point keeps its neighbour j* without a scan while
//     d(p, j*) < d2_lb - moved, where d2_lb is a lower bound of the distance to the second nearest
//     distinct target at the time of its last scan and `moved` bounds how far any point of the
//     patch has travelled since.  ICP converges, so after two or three iterations almost nothing
//     is rescanned.
// The accepted-pair moments, the inlier test and the error sums always use exact fp64 distances.
#pragma once
#include "icp_device.cuh"

#ifndef WICP_CAP
#define WICP_CAP 224
#endif

#ifdef F4L_DEBUG_SCANS
__device__ unsigned long long g_dbg[24];   // 0: point scans, 1: exact-path scans, 2: iterations, 3: points*iters, 4: sum moved (um), 5: keep checks
// phase timers (debug build): DBG_T0 starts, DBG_T(slot) adds the cycles since the last mark to g_dbg[slot]
// 8 rigidity, 9 procrustes, 10 icp stage, 11 phase1, 12 lane scans, 13 exact scans, 14 phase3, 15 update(svd), 16 total
#define DBG_T0 long long dbg_t = clock64();
#define DBG_T(slot) { const long long dbg_n = clock64(); if (lane == 0) atomicAdd(&g_dbg[slot], (unsigned long long)(dbg_n - dbg_t)); dbg_t = dbg_n; }
#else
#define DBG_T0
#define DBG_T(slot)
#endif

struct WarpIcpSmem {
    float Bx[WICP_CAP + 2], By[WICP_CAP + 2], Bz[WICP_CAP + 2];   // target, pivot-local, f32-rounded: ONLY for the f32 scans
                               // (SoA so that two consecutive targets are one 8-byte word of a packed operand; odd counts are
                               // padded with a point at 1e18)
    float Bg[WICP_CAP * 3];    // target, as given (f32 global): every exact fp64 computation
    float A[WICP_CAP * 3];     // source, as given (f32 global)
    float pscan[WICP_CAP * 3]; // pivot-local position of each source point at its last scan
    float d2lb[WICP_CAP];      // lower bound of the distance to the 2nd nearest distinct target
    unsigned short jstar[WICP_CAP];   // current neighbour of each source point
    unsigned short list[WICP_CAP];
    double Vw[9];              // SVD warm-start basis (common.cuh svd3x3), identical in all lanes
};

__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(F4L_FULL, v, o); }

// Gather 4 rounds of 32 points (items base + 32u + lane of the list k0.., u < 4) with every load of the batch in
// flight before the first use: indices first, then the 12 coordinate words -- the staging loops are pure
// latency (index -> point, both usually L2 misses) and a round-by-round loop serialises them.
struct Gather4 {
    float x[4], y[4], z[4];
};
__device__ __forceinline__ void gather4(const float* __restrict__ pts, const int32_t* __restrict__ idx, int k0, int base,
                                        int n, int lane, Gather4& gq) {
    size_t row[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = base + 32 * u + lane;
        const int k = k0 + (i < n ? i : 0);
        row[u] = idx ? (size_t)__ldg(idx + k) : (size_t)k;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float* p = pts + row[u] * 3;
        gq.x[u] = __ldg(p); gq.y[u] = __ldg(p + 1); gq.z[u] = __ldg(p + 2);
    }
}

// packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two lanes of fp32 per instruction)
typedef unsigned long long u64p;
__device__ __forceinline__ u64p pk2(float lo, float hi) { u64p r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64p v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64p sub2(u64p a, u64p b) { u64p r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64p mul2(u64p a, u64p b) { u64p r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64p fma2(u64p a, u64p b, u64p c) { u64p r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }


__device__ __forceinline__ void warp_argmin_f64(double& d, int& j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = shfl_xor_d(d, o);
        const int oj = __shfl_xor_sync(F4L_FULL, j, o);
        if (od < d || (od == d && oj < j)) { d = od; j = oj; }
    }
}

// top-2 merge across the warp in fp64; every lane ends with the global (d1, j1, d2); ties towards the lower index
__device__ __forceinline__ void warp_top2_f64(double& d1, int& j1, double& d2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od1 = shfl_xor_d(d1, o);
        const int oj1 = __shfl_xor_sync(F4L_FULL, j1, o);
        const double od2 = shfl_xor_d(d2, o);
        const bool other_wins = od1 < d1 || (od1 == d1 && oj1 < j1);
        const double loser = other_wins ? d1 : od1;
        d2 = fmin(fmin(d2, od2), loser);
        if (other_wins) { d1 = od1; j1 = oj1; }
    }
}

// Whole-warp EXACT scan for one query (global fp64 position pg): the nearest target in fp64 (first minimal index)
// and a lower bound of the distance to the second nearest distinct target.  Only queries whose f32 scan could
// not separate its two best candidates come here.
__device__ inline void warp_scan_point(const WarpIcpSmem& sm, int nt, const double pg[3], int lane, int& jbest,
                                       float& d2lb, double& d2_best, double& d2_second) {
#ifdef F4L_DEBUG_SCANS
    if (lane == 0) atomicAdd(&g_dbg[1], 1ull);
#endif
    // exact path: ONE fp64 pass keeps the two smallest distances (first index on ties).  A duplicate of the
    // winner has exactly the winner's distance, so e2 > e1 means e2 already is the nearest DISTINCT target;
    // only e2 == e1 (a duplicate, or an exact tie of different points) needs the pass that skips duplicates.
    double e1 = INFINITY, e2 = INFINITY;
    int k1 = 0x7fffffff;
    for (int j = lane; j < nt; j += 32) {
        const double dx = pg[0] - (double)sm.Bg[3 * j], dy = pg[1] - (double)sm.Bg[3 * j + 1], dz = pg[2] - (double)sm.Bg[3 * j + 2];
        const double dd = dx * dx + dy * dy + dz * dz;
        if (dd < e1) { e2 = e1; e1 = dd; k1 = j; }
        else if (dd < e2) e2 = dd;
    }
    warp_top2_f64(e1, k1, e2);
    if (e2 == e1 && e1 != INFINITY) {
        const float wx = sm.Bg[3 * k1], wy = sm.Bg[3 * k1 + 1], wz = sm.Bg[3 * k1 + 2];
        e2 = INFINITY;
        int k2 = 0;
        for (int j = lane; j < nt; j += 32) {
            const float bx = sm.Bg[3 * j], by = sm.Bg[3 * j + 1], bz = sm.Bg[3 * j + 2];
            if (bx == wx && by == wy && bz == wz) continue;          // duplicate of the winner
            const double dx = pg[0] - (double)bx, dy = pg[1] - (double)by, dz = pg[2] - (double)bz;
            const double dd = dx * dx + dy * dy + dz * dz;
            if (dd < e2) { e2 = dd; k2 = j; }
        }
        warp_argmin_f64(e2, k2);
    }
    jbest = k1;
    d2_best = e1;
    d2_second = e2;          // nearest target with other coordinates than the winner
    d2lb = (e2 == INFINITY) ? INFINITY : fmaxf((float)(sqrt(e2) * (1.0 - 1e-7)) - 1e-7f, 0.f);
}

// One warp, one pair.  Requires 1 <= ns, nt <= WICP_CAP.  T0: 16 doubles or nullptr.
__device__ inline IcpResult warp_icp(const float* __restrict__ src, const int32_t* __restrict__ sidx, int s0,
                                     int ns, const float* __restrict__ tgt, const int32_t* __restrict__ tidx,
                                     int t0, int nt, const double* T0, double max_dist, int max_iter,
                                     double rel_fit, double rel_rmse, double* Tout, int32_t* __restrict__ corr,
                                     WarpIcpSmem& sm, int lane, double tie_eps = F4L_ICP_TIE_EPS) {
    DBG_T0
    // ---- stage ------------------------------------------------------------------------------
    double cB[3];
    {
        float x, y, z;
        load_ptf(tgt, tidx, t0, x, y, z);
        cB[0] = x; cB[1] = y; cB[2] = z;
    }
    for (int base = 0; base < ns; base += 128) {
        Gather4 a;
        gather4(src, sidx, s0, base, ns, lane, a);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            if (i < ns) { sm.A[3 * i] = a.x[u]; sm.A[3 * i + 1] = a.y[u]; sm.A[3 * i + 2] = a.z[u]; }
        }
    }
    for (int base = 0; base < nt; base += 128) {
        Gather4 b;
        gather4(tgt, tidx, t0, base, nt, lane, b);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = base + 32 * u + lane;
            if (j < nt) {
                const float x = b.x[u], y = b.y[u], z = b.z[u];
                sm.Bg[3 * j] = x; sm.Bg[3 * j + 1] = y; sm.Bg[3 * j + 2] = z;
                sm.Bx[j] = (float)((double)x - cB[0]); sm.By[j] = (float)((double)y - cB[1]); sm.Bz[j] = (float)((double)z - cB[2]);
            }
        }
    }
    if (lane == 0 && (nt & 1)) { sm.Bx[nt] = 1e18f; sm.By[nt] = 1e18f; sm.Bz[nt] = 1e18f; }
    // Several source points are often matched to the SAME target point: a repeated target can never be the
    // first-index nearest neighbour (its first copy is as close) and must not count as the "second nearest distinct"
    // one, yet it would send every query near it through the exact scan (d2 == d1).  Repeats are moved out of the
    // f32 scans' way; the fp64 arrays keep them (correspondence indices and the exact path are unchanged).
    // (repeats are recognised by their index into the target cloud through a small open-addressing table that
    // borrows the scan-state arrays, which are not live yet; without an index list nothing is removed)
    if (tidx) {
        static_assert(offsetof(WarpIcpSmem, Vw) - offsetof(WarpIcpSmem, pscan) >= 512 * 4 + 512 * 2, "dedup table must fit");
        int* keys = reinterpret_cast<int*>(sm.pscan);
        unsigned short* first = reinterpret_cast<unsigned short*>(keys + 512);
        __syncwarp();
        for (int h = lane; h < 512; h += 32) { keys[h] = -1; first[h] = 0xffffu; }
        __syncwarp();
        for (int j = lane; j < nt; j += 32) {
            const int key = __ldg(tidx + t0 + j);
            unsigned h = ((unsigned)key * 2654435761u) >> 23;
            while (true) {
                const int prev = atomicCAS(&keys[h], -1, key);
                if (prev == -1 || prev == key) break;
                h = (h + 1u) & 511u;
            }
            // 16-bit min through the 32-bit word that holds it
            unsigned* w = reinterpret_cast<unsigned*>(first) + (h >> 1);
            const unsigned sh = (h & 1u) * 16u;
            unsigned old = *w;
            while (((old >> sh) & 0xffffu) > (unsigned)j) {
                const unsigned want = (old & ~(0xffffu << sh)) | ((unsigned)j << sh);
                const unsigned got = atomicCAS(w, old, want);
                if (got == old) break;
                old = got;
            }
        }
        __syncwarp();
        for (int j = lane; j < nt; j += 32) {
            const int key = __ldg(tidx + t0 + j);
            unsigned h = ((unsigned)key * 2654435761u) >> 23;
            while (keys[h] != key) h = (h + 1u) & 511u;
            if (first[h] != (unsigned short)j) { sm.Bx[j] = 1e18f; sm.By[j] = 1e18f; sm.Bz[j] = 1e18f; }
        }
        __syncwarp();
    }
    double T[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) T[a] = T0 ? T0[a] : ((a % 5 == 0) ? 1.0 : 0.0);
    __syncwarp();

    DBG_T(10)
    const double max_d2 = max_dist * max_dist;
    IcpResult out;
    out.fitness = 0; out.rmse = 0; out.iters = 0; out.fragile = 0;
    int frag = 0;            // lane-local until the end
    double prev_fit = 0.0, prev_rmse = 0.0;
    for (int it = 0;; ++it) {
        // ---- phase 1: which points need a scan -------------------------------------------
        // a point keeps its neighbour j* while d(p, j*) < d2lb - |p - p_at_last_scan|  (triangle inequality)
        int nres = 0;
        for (int base = 0; base < ns; base += 32) {
            const int i = base + lane;
            bool need = false;
            if (i < ns) {
                need = true;
                if (it > 0) {
                    const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
                    const double px = T[0] * fx + T[1] * fy + T[2] * fz + T[3];
                    const double py = T[4] * fx + T[5] * fy + T[6] * fz + T[7];
                    const double pz = T[8] * fx + T[9] * fy + T[10] * fz + T[11];
                    const int js = sm.jstar[i];
                    const double dx = px - (double)sm.Bg[3 * js], dy = py - (double)sm.Bg[3 * js + 1], dz = pz - (double)sm.Bg[3 * js + 2];
                    const float dnow = sqrtf((float)(dx * dx + dy * dy + dz * dz)) * 1.00001f + 1e-7f;   // upper bound: the f32 root's 6e-8 is inside the 1e-5 margin
                    const float mx = (float)(px - cB[0]) - sm.pscan[3 * i], my = (float)(py - cB[1]) - sm.pscan[3 * i + 1],
                                mz = (float)(pz - cB[2]) - sm.pscan[3 * i + 2];
                    const float moved = sqrtf(mx * mx + my * my + mz * mz) * 1.00001f + 1e-6f;
                    need = !(dnow < sm.d2lb[i] - moved);
                }
            }
            const unsigned m = __ballot_sync(F4L_FULL, need);
            if (need) sm.list[nres + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
            nres += __popc(m);
        }
        __syncwarp();
        DBG_T(11)
#ifdef F4L_DEBUG_SCANS
        if (lane == 0) { atomicAdd(&g_dbg[0], (unsigned long long)nres); atomicAdd(&g_dbg[2], 1ull); atomicAdd(&g_dbg[3], (unsigned long long)ns); }
#endif
        // ---- phase 2: scans ------------------------------------------------------------------
        // f32 scan of the listed points on pivot-local coordinates (targets are stored f32-rounded in that frame,
        // |coordinate| < 4 m -> <= 1.2e-7 m each, so two candidates are safely ordered when d2 >= 1.004 d1 + 2e-8).
        // G lanes share a point (G = 1 while >= 17 points are listed, up to 32 for a single one), each walks every
        // G-th PAIR of targets with packed arithmetic; a candidate is the 32-bit key (bits(d^2) & ~255) | index, so
        // the two smallest are three integer min/max and ties resolve to the lower index.  The dropped 8 mantissa
        // bits (2^-15 relative) are covered by testing the truncated values with the factor 1.00404.
        int nexact = 0;
        {
            const int lg = nres > 16 ? 0 : (nres > 8 ? 1 : (nres > 4 ? 2 : (nres > 2 ? 3 : (nres > 1 ? 4 : 5))));
            const int G = 1 << lg, per_round = 32 >> lg;
            const int g = lane & (G - 1), slot = lane >> lg;
            const int npairs = (nt + 1) >> 1;
            for (int r0 = 0; r0 < nres; r0 += per_round) {
                const int r = r0 + slot;
                const bool active = r < nres;
                const int i = active ? sm.list[r] : sm.list[0];
                const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
                const float qx = (float)(T[0] * fx + T[1] * fy + T[2] * fz + T[3] - cB[0]);
                const float qy = (float)(T[4] * fx + T[5] * fy + T[6] * fz + T[7] - cB[1]);
                const float qz = (float)(T[8] * fx + T[9] * fy + T[10] * fz + T[11] - cB[2]);
                const u64p qx2 = pk2(qx, qx), qy2 = pk2(qy, qy), qz2 = pk2(qz, qz);
                unsigned k1 = 0xffffffffu, k2 = 0xffffffffu;
#pragma unroll 4
                for (int p = g; p < npairs; p += G) {
                    const u64p bx = *reinterpret_cast<const u64p*>(&sm.Bx[2 * p]);
                    const u64p by = *reinterpret_cast<const u64p*>(&sm.By[2 * p]);
                    const u64p bz = *reinterpret_cast<const u64p*>(&sm.Bz[2 * p]);
                    const u64p dx = sub2(qx2, bx), dy = sub2(qy2, by), dz = sub2(qz2, bz);
                    const u64p dd = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float da, db;
                    upk2(dd, da, db);
                    const unsigned ka = (__float_as_uint(da) & 0xffffff00u) | (unsigned)(2 * p);
                    const unsigned kb = (__float_as_uint(db) & 0xffffff00u) | (unsigned)(2 * p + 1);
                    k2 = min(k2, max(k1, ka)); k1 = min(k1, ka);
                    k2 = min(k2, max(k1, kb)); k1 = min(k1, kb);
                }
                for (int o = 1; o < G; o <<= 1) {
                    const unsigned o1 = __shfl_xor_sync(F4L_FULL, k1, o), o2 = __shfl_xor_sync(F4L_FULL, k2, o);
                    k2 = min(min(k2, o2), max(k1, o1));
                    k1 = min(k1, o1);
                }
                const float v1 = __uint_as_float(k1 & 0xffffff00u), v2 = __uint_as_float(k2 & 0xffffff00u);
                const bool certain = v2 >= v1 * 1.00404f + 2e-8f;
                const bool writer = active && g == 0;
                if (writer && certain) {
                    sm.jstar[i] = (unsigned short)(k1 & 0xffu);
                    sm.d2lb[i] = fmaxf(sqrtf(v2) * 0.9999f - 2e-6f, 0.f);
                    sm.pscan[3 * i] = qx; sm.pscan[3 * i + 1] = qy; sm.pscan[3 * i + 2] = qz;
                }
                __syncwarp();
                const unsigned m = __ballot_sync(F4L_FULL, writer && !certain);
                if (writer && !certain) sm.list[nexact + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;   // in place: nexact <= r0
                nexact += __popc(m);
                __syncwarp();
            }
        }
        DBG_T(12)
        for (int r = 0; r < nexact; ++r) {
            const int i = sm.list[r];
            const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
            double pg[3];
            pg[0] = T[0] * fx + T[1] * fy + T[2] * fz + T[3];
            pg[1] = T[4] * fx + T[5] * fy + T[6] * fz + T[7];
            pg[2] = T[8] * fx + T[9] * fy + T[10] * fz + T[11];
            int jb;
            float lb;
            double e1, e2;
            warp_scan_point(sm, nt, pg, lane, jb, lb, e1, e2);
            // points whose f32 scan separated the two best candidates (factor 1.004) or that kept their neighbour by the
            // triangle inequality (margin 1e-5) are far from a tie: only the exact path can see one
            if (e1 < max_d2 && e2 - e1 <= tie_eps * e1) frag |= F4L_ICP_FRAGILE_NN;
            if (lane == 0) {
                sm.jstar[i] = (unsigned short)jb;
                sm.d2lb[i] = lb;
                sm.pscan[3 * i] = (float)(pg[0] - cB[0]); sm.pscan[3 * i + 1] = (float)(pg[1] - cB[1]); sm.pscan[3 * i + 2] = (float)(pg[2] - cB[2]);
            }
        }
        __syncwarp();
        DBG_T(13)
        // ---- phase 3: accept / accumulate with exact fp64 distances -------------------------
        Moments M;
        moments_zero(M);
        double err2 = 0.0, cnt = 0.0;
        for (int i = lane; i < ns; i += 32) {
            const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
            const double px = T[0] * fx + T[1] * fy + T[2] * fz + T[3];
            const double py = T[4] * fx + T[5] * fy + T[6] * fz + T[7];
            const double pz = T[8] * fx + T[9] * fy + T[10] * fz + T[11];
            const int j = sm.jstar[i];
            const double bx = sm.Bg[3 * j], by = sm.Bg[3 * j + 1], bz = sm.Bg[3 * j + 2];
            const double dx = px - bx, dy = py - by, dz = pz - bz;
            const double d2 = dx * dx + dy * dy + dz * dz;
            const bool ok = d2 < max_d2;
            if (corr) corr[s0 + i] = ok ? j : -1;
            if (fabs(d2 - max_d2) <= tie_eps * max_d2) frag |= F4L_ICP_FRAGILE_INLIER;
            if (ok) {
                err2 += d2;
                cnt += 1.0;
                moments_add(M, 1.0, px - cB[0], py - cB[1], pz - cB[2], bx - cB[0], by - cB[1], bz - cB[2]);
            }
        }
        moments_warp_reduce(M);
        err2 = warp_sum(err2);
        cnt = warp_sum(cnt);
        const double fit = cnt > 0 ? cnt / (double)ns : 0.0;
        const double rmse = cnt > 0 ? sqrt(err2 / cnt) : 0.0;
        out.fitness = fit;
        out.rmse = rmse;
        bool stop = false;
        if (it > 0 && fabs(prev_fit - fit) < rel_fit && fabs(prev_rmse - rmse) < rel_rmse) stop = true;
        if (it > 0 && it < max_iter) frag |= icp_stop_fragile(prev_fit, fit, prev_rmse, rmse, rel_fit, rel_rmse, tie_eps);
        if (it >= max_iter) stop = true;
        DBG_T(14)
        if (stop) {
            out.iters = it;
            break;
        }
        prev_fit = fit;
        prev_rmse = rmse;
        // ---- update: U = umeyama(P[corr], tgt[corr]);  T <- U T -------------------------------
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        if (cnt > 0) fit_from_moments(M, cB, cB, 0.0, 2, R, t, sm.Vw);
        double Tn[12];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                Tn[r * 4 + c] = R[r * 3 + 0] * T[c] + R[r * 3 + 1] * T[4 + c] + R[r * 3 + 2] * T[8 + c] + (c == 3 ? t[r] : 0.0);
#pragma unroll
        for (int a = 0; a < 12; ++a) T[a] = Tn[a];
        DBG_T(15)
    }
#pragma unroll
    for (int a = 0; a < 12; ++a)
        if (lane == a) Tout[a] = T[a];
    if (lane >= 12 && lane < 16) Tout[lane] = lane == 15 ? 1.0 : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) frag |= __shfl_xor_sync(F4L_FULL, frag, o);
    out.fragile = frag;
    return out;
}


// Rigidity statistic (base.py:3308-3317) of the staged pairs by one warp.
// Lane l owns the row pair (i, n-1-i): together they hold n-1 column terms, so the 32 lanes are
// balanced; distances use sqrt.approx.f32 (<= 1 ulp; the reference's own cdist carries ~1e-3 m of
// GEMM-formulation noise at these coordinates).
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // ftz: one MUFU, no denormal rescaling (d^2 < 1e-38 m^2 -> 0)
    return r;
}

// ---- packed rigidity ----------------------------------------------------------------------------
// Pair schedule: row i meets the h = (n-1)/2 points that FOLLOW it cyclically (j = i+1 .. i+h mod n), which
// covers every unordered pair exactly once for odd n and leaves the n/2 "diameters" (i, i+n/2) for even n --
// every row has the same trip count, so lanes are balanced without the triangular tail.  A lane owns two
// ADJACENT rows (i, i+1): at offset s they meet points (i+s, i+s+1), at s+1 points (i+s+1, i+s+2) -- a sliding
// window, one new shared-memory word per coordinate per step -- and the two pairs of a step go through the
// packed FADD2/FMUL2/FFMA2 pipe (two fp32 lanes per instruction).  The window's halves alternate roles
// (step A: lo = row i, hi = row i+1; step B: lo = row i+1, hi = row i) so no register moves are needed.
// The arena holds the six coordinate arrays twice over (cyclic access without a wrap test); it aliases the
// ICP staging area, which is filled after the check.
struct RigRows { u64p ax, ay, az, bx, by, bz; };      // the lane's two rows, packed (lo, hi)

// one step: lo pair = (rows.lo, window.lo), hi pair = (rows.hi, window.hi); (s_lo, c_lo) / (s_hi, c_hi) collect them
__device__ __forceinline__ void rig_step(const RigRows& r, float x0, float x1, float y0, float y1, float z0, float z1,
                                         float u0, float u1, float v0, float v1, float w0, float w1, float thres,
                                         float& s_lo, unsigned& c_lo, float& s_hi, unsigned& c_hi) {
    const u64p ex = sub2(r.ax, pk2(x0, x1)), ey = sub2(r.ay, pk2(y0, y1)), ez = sub2(r.az, pk2(z0, z1));
    const u64p fx = sub2(r.bx, pk2(u0, u1)), fy = sub2(r.by, pk2(v0, v1)), fz = sub2(r.bz, pk2(w0, w1));
    const u64p e2 = fma2(ez, ez, fma2(ey, ey, mul2(ex, ex)));
    const u64p f2 = fma2(fz, fz, fma2(fy, fy, mul2(fx, fx)));
    float e_lo, e_hi, f_lo, f_hi;
    upk2(e2, e_lo, e_hi);
    upk2(f2, f_lo, f_hi);
    const float d_lo = fabsf(sqrt_approx(e_lo) - sqrt_approx(f_lo));
    const float d_hi = fabsf(sqrt_approx(e_hi) - sqrt_approx(f_hi));
    s_lo += d_lo; c_lo += (d_lo <= thres) ? 1u : 0u;
    s_hi += d_hi; c_hi += (d_hi <= thres) ? 1u : 0u;
}

static_assert(6 * (2 * WICP_CAP + 4) * 4 <= (int)offsetof(WarpIcpSmem, Vw), "rigidity arena must fit the warp's slice");
// floats per coordinate array for n pairs.  Each array holds the extended cyclic sequence j = 0 .. 2n+3
// (point j mod n; 4 zero pad words) DE-INTERLEAVED: even j in the first half, odd j in the second, so that
// lanes owning rows 2l, 2l+1 read consecutive words (a plain layout would be a stride-2, two-way bank conflict
// on every load of the sliding window).
__device__ __forceinline__ int rig_arena_stride(int n) { return 2 * n + 4; }
__device__ __forceinline__ int rig_slot(int j, int n) { return (j & 1) * (n + 2) + (j >> 1); }

// Stage the matched pairs (global rows k0 .. k0+n of the correspondence lists) into the arena.
__device__ inline void warp_rigidity_stage(float* arena, const float* __restrict__ src_pts, const float* __restrict__ tgt_pts,
                                           const int32_t* __restrict__ cs, const int32_t* __restrict__ ct, int k0, int n,
                                           int lane) {
    const int L = rig_arena_stride(n);
    for (int base = 0; base < n; base += 128) {
        Gather4 a, b;
        gather4(src_pts, cs, k0, base, n, lane, a);
        gather4(tgt_pts, ct, k0, base, n, lane, b);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = base + 32 * u + lane;
            if (i < n) {
                const int s0 = rig_slot(i, n), s1 = rig_slot(i + n, n);
                arena[s0] = a.x[u]; arena[s1] = a.x[u];
                arena[L + s0] = a.y[u]; arena[L + s1] = a.y[u];
                arena[2 * L + s0] = a.z[u]; arena[2 * L + s1] = a.z[u];
                arena[3 * L + s0] = b.x[u]; arena[3 * L + s1] = b.x[u];
                arena[4 * L + s0] = b.y[u]; arena[4 * L + s1] = b.y[u];
                arena[5 * L + s0] = b.z[u]; arena[5 * L + s1] = b.z[u];
            }
        }
    }
    if (lane < 24) arena[(lane >> 2) * L + rig_slot(2 * n + (lane & 3), n)] = 0.f;     // the window may read 3 words past 2n
}

// Rigidity statistic (base.py:3308-3317) of the staged pairs by one warp: sum of |d_src(i,j) - d_tgt(i,j)| over
// i < j and the number of pairs with that difference <= thres.  Distances use sqrt.approx.f32 (<= 1 ulp; the
// reference's own cdist carries ~1e-3 m of GEMM-formulation noise at these coordinates).
__device__ inline void warp_rigidity(const float* arena, int n, float thres, int lane, double& out_sum, unsigned& out_cnt) {
    const int L = rig_arena_stride(n);
    const float* AX = arena; const float* AY = arena + L; const float* AZ = arena + 2 * L;
    const float* BX = arena + 3 * L; const float* BY = arena + 4 * L; const float* BZ = arena + 5 * L;
    const int h = (n - 1) >> 1;
    float sum = 0.f;
    unsigned cnt = 0;
    for (int base = 0; base < n; base += 64) {
        // a round covers 64 rows; when fewer are left, the offsets 1..h are split over f groups of lanes
        const int left = n - base;
        int f = 1;
        while (f < 16 && left * (2 * f) <= 64) f *= 2;
        const int per = 32 / f;                          // lanes (row pairs) per group
        const int grp = lane / per;
        const int i1 = base + 2 * (lane - grp * per);    // even
        const int chunk = (h + f - 1) / f;
        const int sb = 1 + grp * chunk;
        const int se = min(h, sb + chunk - 1);           // offsets [sb, se]
        if (i1 < n && sb <= se) {
            const bool two = i1 + 1 < n;
            RigRows ra, rb;
            {
                const int r1 = rig_slot(i1, n), r2 = rig_slot(i1 + 1, n);
                const float a1x = AX[r1], a1y = AY[r1], a1z = AZ[r1], b1x = BX[r1], b1y = BY[r1], b1z = BZ[r1];
                const float a2x = AX[r2], a2y = AY[r2], a2z = AZ[r2], b2x = BX[r2], b2y = BY[r2], b2z = BZ[r2];
                ra.ax = pk2(a1x, a2x); ra.ay = pk2(a1y, a2y); ra.az = pk2(a1z, a2z);
                ra.bx = pk2(b1x, b2x); ra.by = pk2(b1y, b2y); ra.bz = pk2(b1z, b2z);
                rb.ax = pk2(a2x, a1x); rb.ay = pk2(a2y, a1y); rb.az = pk2(a2z, a1z);
                rb.bx = pk2(b2x, b1x); rb.by = pk2(b2y, b1y); rb.bz = pk2(b2z, b1z);
            }
            float s1 = 0.f, s2 = 0.f;
            unsigned c1 = 0, c2 = 0;
            // window words: x0 walks the points j, j+2, ... (one parity), x1 the points j+1, j+3, ... (the other)
            const int j = i1 + sb;
            int o0 = rig_slot(j, n), o1 = rig_slot(j + 1, n);
            float x0 = AX[o0], y0 = AY[o0], z0 = AZ[o0], u0 = BX[o0], v0 = BY[o0], w0 = BZ[o0];
            float x1 = AX[o1], y1 = AY[o1], z1 = AZ[o1], u1 = BX[o1], v1 = BY[o1], w1 = BZ[o1];
            int s = sb;
            for (; s < se; s += 2) {
                rig_step(ra, x0, x1, y0, y1, z0, z1, u0, u1, v0, v1, w0, w1, thres, s1, c1, s2, c2);
                ++o0;
                x0 = AX[o0]; y0 = AY[o0]; z0 = AZ[o0]; u0 = BX[o0]; v0 = BY[o0]; w0 = BZ[o0];
                rig_step(rb, x0, x1, y0, y1, z0, z1, u0, u1, v0, v1, w0, w1, thres, s2, c2, s1, c1);
                ++o1;
                x1 = AX[o1]; y1 = AY[o1]; z1 = AZ[o1]; u1 = BX[o1]; v1 = BY[o1]; w1 = BZ[o1];
            }
            if (s == se) rig_step(ra, x0, x1, y0, y1, z0, z1, u0, u1, v0, v1, w0, w1, thres, s1, c1, s2, c2);
            sum += s1;
            cnt += c1;
            if (two) { sum += s2; cnt += c2; }
        }
    }
    if (!(n & 1)) {
        // even n: the diameters (i, i + n/2), i < n/2
        const int hn = n >> 1;
        for (int i = lane; i < hn; i += 32) {
            const int p = rig_slot(i, n), q = rig_slot(i + hn, n);
            const float ex = AX[p] - AX[q], ey = AY[p] - AY[q], ez = AZ[p] - AZ[q];
            const float fx = BX[p] - BX[q], fy = BY[p] - BY[q], fz = BZ[p] - BZ[q];
            const float ds = sqrt_approx(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
            const float dt = sqrt_approx(fmaf(fz, fz, fmaf(fy, fy, fx * fx)));
            const float diff = fabsf(ds - dt);
            sum += diff;
            cnt += (diff <= thres) ? 1u : 0u;
        }
    }
    out_sum = warp_sum((double)sum);
    out_cnt = (unsigned)warp_sum((int)cnt);
}

// D1/D2 on the pairs already staged for the rigidity check (unit weights): same values, same pivot, same
// accumulation order as warp_fit_segment, without gathering the points from global memory a second time.
__device__ inline bool warp_fit_arena(const float* arena, int n, double eps, int variant, int lane, double R[9],
                                      double t[3], double* warm) {
    const int L = rig_arena_stride(n);
    double ps[3], pt[3];
    {
        const int p0 = rig_slot(0, n);
        ps[0] = arena[p0]; ps[1] = arena[L + p0]; ps[2] = arena[2 * L + p0];
        pt[0] = arena[3 * L + p0]; pt[1] = arena[4 * L + p0]; pt[2] = arena[5 * L + p0];
    }
    Moments M;
    moments_zero(M);
    for (int i = lane; i < n; i += 32) {
        const int p = rig_slot(i, n);
        const double sx = arena[p], sy = arena[L + p], sz = arena[2 * L + p];
        const double tx = arena[3 * L + p], ty = arena[4 * L + p], tz = arena[5 * L + p];
        moments_add(M, 1.0, sx - ps[0], sy - ps[1], sz - ps[2], tx - pt[0], ty - pt[1], tz - pt[2]);
    }
    moments_warp_reduce(M);
    return fit_from_moments(M, ps, pt, eps, variant, R, t, warm);
}
