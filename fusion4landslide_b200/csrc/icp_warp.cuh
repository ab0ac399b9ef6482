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

#define WICP_CAP 224

#ifdef F4L_DEBUG_SCANS
__device__ unsigned long long g_dbg[8];   // 0: point scans, 1: exact-path scans, 2: iterations, 3: points*iters, 4: sum moved (um), 5: keep checks
#endif

struct WarpIcpSmem {
    float4 B[WICP_CAP];        // target, pivot-local, f32-rounded: ONLY for the f32 scans
    float Bg[WICP_CAP * 3];    // target, as given (f32 global): every exact fp64 computation
    float A[WICP_CAP * 3];     // source, as given (f32 global)
    float pscan[WICP_CAP * 3]; // pivot-local position of each source point at its last scan
    float d2lb[WICP_CAP];      // lower bound of the distance to the 2nd nearest distinct target
    unsigned short jstar[WICP_CAP];   // current neighbour of each source point
    unsigned short list[WICP_CAP];
};

__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(F4L_FULL, v, o); }

// top-2 merge across the warp; every lane ends with the global (d1, j1, d2)
__device__ __forceinline__ void warp_top2_f32(float& d1, int& j1, float& d2) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od1 = __shfl_xor_sync(F4L_FULL, d1, o);
        const int oj1 = __shfl_xor_sync(F4L_FULL, j1, o);
        const float od2 = __shfl_xor_sync(F4L_FULL, d2, o);
        const bool other_wins = od1 < d1 || (od1 == d1 && oj1 < j1);
        const float loser = other_wins ? d1 : od1;
        d2 = fminf(fminf(d2, od2), loser);
        if (other_wins) { d1 = od1; j1 = oj1; }
    }
}

__device__ __forceinline__ void warp_argmin_f64(double& d, int& j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = shfl_xor_d(d, o);
        const int oj = __shfl_xor_sync(F4L_FULL, j, o);
        if (od < d || (od == d && oj < j)) { d = od; j = oj; }
    }
}

// Whole-warp scan for ONE query (global fp64 position pg, pivot cB): returns the exact nearest target
// (first minimal index) and a lower bound of the distance to the second nearest distinct target.
// f32 error budget: query and targets are rounded to f32 in the pivot-local frame (|coordinate| < 4 m
// -> <= 1.2e-7 m each), so two candidates are safely ordered when d2 >= 1.004 d1 + 2e-8 (m^2).
__device__ inline void warp_scan_point(const WarpIcpSmem& sm, int nt, const double pg[3], const double cB[3],
                                       int lane, int& jbest, float& d2lb) {
    const float qx = (float)(pg[0] - cB[0]), qy = (float)(pg[1] - cB[1]), qz = (float)(pg[2] - cB[2]);
    float d1 = INFINITY, d2 = INFINITY;
    int j1 = 0x7fffffff;
    for (int j = lane; j < nt; j += 32) {
        const float4 b = sm.B[j];
        const float dx = qx - b.x, dy = qy - b.y, dz = qz - b.z;
        const float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (dd < d1) { d2 = d1; d1 = dd; j1 = j; }
        else if (dd < d2) d2 = dd;
    }
    warp_top2_f32(d1, j1, d2);
    if (d2 >= d1 * 1.004f + 2e-8f) {           // f32 argmin is provably the fp64 argmin
        jbest = j1;
        d2lb = fmaxf(sqrtf(d2) * 0.9999f - 2e-6f, 0.f);
        return;
    }
#ifdef F4L_DEBUG_SCANS
    if (lane == 0) atomicAdd(&g_dbg[1], 1ull);
#endif
    // exact path: fp64 argmin with first-index ties, then the nearest target with other coordinates
    double e1 = INFINITY;
    int k1 = 0x7fffffff;
    for (int j = lane; j < nt; j += 32) {
        const double dx = pg[0] - (double)sm.Bg[3 * j], dy = pg[1] - (double)sm.Bg[3 * j + 1], dz = pg[2] - (double)sm.Bg[3 * j + 2];
        const double dd = dx * dx + dy * dy + dz * dz;
        if (dd < e1) { e1 = dd; k1 = j; }
    }
    warp_argmin_f64(e1, k1);
    const float wx = sm.Bg[3 * k1], wy = sm.Bg[3 * k1 + 1], wz = sm.Bg[3 * k1 + 2];
    double e2 = INFINITY;
    int k2 = 0;
    for (int j = lane; j < nt; j += 32) {
        const float bx = sm.Bg[3 * j], by = sm.Bg[3 * j + 1], bz = sm.Bg[3 * j + 2];
        if (bx == wx && by == wy && bz == wz) continue;          // duplicate of the winner
        const double dx = pg[0] - (double)bx, dy = pg[1] - (double)by, dz = pg[2] - (double)bz;
        const double dd = dx * dx + dy * dy + dz * dz;
        if (dd < e2) { e2 = dd; k2 = j; }
    }
    warp_argmin_f64(e2, k2);
    jbest = k1;
    d2lb = (e2 == INFINITY) ? INFINITY : fmaxf((float)(sqrt(e2) * (1.0 - 1e-7)) - 1e-7f, 0.f);
}

// One warp, one pair.  Requires 1 <= ns, nt <= WICP_CAP.  T0: 16 doubles or nullptr.
__device__ inline IcpResult warp_icp(const float* __restrict__ src, const int32_t* __restrict__ sidx, int s0,
                                     int ns, const float* __restrict__ tgt, const int32_t* __restrict__ tidx,
                                     int t0, int nt, const double* T0, double max_dist, int max_iter,
                                     double rel_fit, double rel_rmse, double* Tout, int32_t* __restrict__ corr,
                                     WarpIcpSmem& sm, int lane) {
    // ---- stage ------------------------------------------------------------------------------
    double cB[3];
    {
        float x, y, z;
        load_ptf(tgt, tidx, t0, x, y, z);
        cB[0] = x; cB[1] = y; cB[2] = z;
    }
    for (int i = lane; i < ns; i += 32) {
        float x, y, z;
        load_ptf(src, sidx, s0 + i, x, y, z);
        sm.A[3 * i] = x; sm.A[3 * i + 1] = y; sm.A[3 * i + 2] = z;
    }
    for (int j = lane; j < nt; j += 32) {
        float x, y, z;
        load_ptf(tgt, tidx, t0 + j, x, y, z);
        sm.Bg[3 * j] = x; sm.Bg[3 * j + 1] = y; sm.Bg[3 * j + 2] = z;
        sm.B[j] = make_float4((float)((double)x - cB[0]), (float)((double)y - cB[1]), (float)((double)z - cB[2]), 0.f);
    }
    double T[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) T[a] = T0 ? T0[a] : ((a % 5 == 0) ? 1.0 : 0.0);
    __syncwarp();

    const double max_d2 = max_dist * max_dist;
    IcpResult out;
    out.fitness = 0; out.rmse = 0; out.iters = 0;
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
                    const float dnow = (float)sqrt(dx * dx + dy * dy + dz * dz) * 1.00001f + 1e-7f;
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
#ifdef F4L_DEBUG_SCANS
        if (lane == 0) { atomicAdd(&g_dbg[0], (unsigned long long)nres); atomicAdd(&g_dbg[2], 1ull); atomicAdd(&g_dbg[3], (unsigned long long)ns); }
#endif
        // ---- phase 2: scans ------------------------------------------------------------------
        int nexact = nres;                 // items [0, nexact) of the list go through the whole-warp exact-capable scan
        if (nres >= 8) {
            // many points: one LANE per point, every lane walks all targets (broadcast reads), f32 top-2;
            // points whose two best are within the f32 error margin are re-listed for the exact scan
            nexact = 0;
            for (int r0 = 0; r0 < nres; r0 += 32) {
                const int r = r0 + lane;
                const bool active = r < nres;
                const int i = active ? sm.list[r] : sm.list[0];
                const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
                const float qx = (float)(T[0] * fx + T[1] * fy + T[2] * fz + T[3] - cB[0]);
                const float qy = (float)(T[4] * fx + T[5] * fy + T[6] * fz + T[7] - cB[1]);
                const float qz = (float)(T[8] * fx + T[9] * fy + T[10] * fz + T[11] - cB[2]);
                float d1 = INFINITY, d2 = INFINITY;
                int j1 = 0;
#pragma unroll 4
                for (int j = 0; j < nt; ++j) {
                    const float4 b = sm.B[j];
                    const float dx = qx - b.x, dy = qy - b.y, dz = qz - b.z;
                    const float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const bool better = dd < d1;
                    d2 = better ? d1 : fminf(d2, dd);
                    j1 = better ? j : j1;
                    d1 = better ? dd : d1;
                }
                const bool certain = d2 >= d1 * 1.004f + 2e-8f;
                if (active && certain) {
                    sm.jstar[i] = (unsigned short)j1;
                    sm.d2lb[i] = fmaxf(sqrtf(d2) * 0.9999f - 2e-6f, 0.f);
                    sm.pscan[3 * i] = qx; sm.pscan[3 * i + 1] = qy; sm.pscan[3 * i + 2] = qz;
                }
                __syncwarp();
                const unsigned m = __ballot_sync(F4L_FULL, active && !certain);
                if (active && !certain) sm.list[nexact + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;   // in place: nexact <= r0
                nexact += __popc(m);
                __syncwarp();
            }
        }
        for (int r = 0; r < nexact; ++r) {
            const int i = sm.list[r];
            const double fx = sm.A[3 * i], fy = sm.A[3 * i + 1], fz = sm.A[3 * i + 2];
            double pg[3];
            pg[0] = T[0] * fx + T[1] * fy + T[2] * fz + T[3];
            pg[1] = T[4] * fx + T[5] * fy + T[6] * fz + T[7];
            pg[2] = T[8] * fx + T[9] * fy + T[10] * fz + T[11];
            int jb;
            float lb;
            warp_scan_point(sm, nt, pg, cB, lane, jb, lb);
            if (lane == 0) {
                sm.jstar[i] = (unsigned short)jb;
                sm.d2lb[i] = lb;
                sm.pscan[3 * i] = (float)(pg[0] - cB[0]); sm.pscan[3 * i + 1] = (float)(pg[1] - cB[1]); sm.pscan[3 * i + 2] = (float)(pg[2] - cB[2]);
            }
        }
        __syncwarp();
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
        if (it >= max_iter) stop = true;
        if (stop) {
            out.iters = it;
            break;
        }
        prev_fit = fit;
        prev_rmse = rmse;
        // ---- update: U = umeyama(P[corr], tgt[corr]);  T <- U T -------------------------------
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
        if (cnt > 0) fit_from_moments(M, cB, cB, 0.0, 2, R, t);
        double Tn[12];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                Tn[r * 4 + c] = R[r * 3 + 0] * T[c] + R[r * 3 + 1] * T[4 + c] + R[r * 3 + 2] * T[8 + c] + (c == 3 ? t[r] : 0.0);
#pragma unroll
        for (int a = 0; a < 12; ++a) T[a] = Tn[a];
    }
#pragma unroll
    for (int a = 0; a < 12; ++a)
        if (lane == a) Tout[a] = T[a];
    if (lane >= 12 && lane < 16) Tout[lane] = lane == 15 ? 1.0 : 0.0;
    return out;
}


// Rigidity statistic (base.py:3308-3317) of the staged pairs by one warp.
// Lane l owns the row pair (i, n-1-i): together they hold n-1 column terms, so the 32 lanes are
// balanced; distances use sqrt.approx.f32 (<= 1 ulp; the reference's own cdist carries ~1e-3 m of
// GEMM-formulation noise at these coordinates).
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void rigidity_row(const WarpIcpSmem& sm, int i, int n, float thres, float& sum, unsigned& cnt) {
    const float ax = sm.A[3 * i], ay = sm.A[3 * i + 1], az = sm.A[3 * i + 2];
    const float bx = sm.Bg[3 * i], by = sm.Bg[3 * i + 1], bz = sm.Bg[3 * i + 2];
    float rowsum = 0.f;
#pragma unroll 2
    for (int j = i + 1; j < n; ++j) {
        const float ex = ax - sm.A[3 * j], ey = ay - sm.A[3 * j + 1], ez = az - sm.A[3 * j + 2];
        const float fx = bx - sm.Bg[3 * j], fy = by - sm.Bg[3 * j + 1], fz = bz - sm.Bg[3 * j + 2];
        const float ds = sqrt_approx(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
        const float dt = sqrt_approx(fmaf(fz, fz, fmaf(fy, fy, fx * fx)));
        const float diff = fabsf(ds - dt);
        rowsum += diff;
        cnt += (diff <= thres) ? 1u : 0u;
    }
    sum += rowsum;
}

__device__ inline void warp_rigidity(const WarpIcpSmem& sm, int n, float thres, int lane, double& out_sum,
                                     unsigned& out_cnt) {
    float sum = 0.f;
    unsigned cnt = 0;
    const int half = n >> 1;                       // row pairs (i, n-1-i), i < half; odd n: middle row alone
    for (int i = lane; i < half; i += 32) {
        rigidity_row(sm, i, n, thres, sum, cnt);
        rigidity_row(sm, n - 1 - i, n, thres, sum, cnt);
    }
    if ((n & 1) && lane == 0) rigidity_row(sm, half, n, thres, sum, cnt);
    out_sum = warp_sum((double)sum);
    out_cnt = (unsigned)warp_sum((int)cnt);
}
