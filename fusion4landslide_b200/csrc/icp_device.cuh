// K-e building block: point-to-point ICP of one patch pair by one CTA.  The kNN <-> Kabsch
// iterate loop of Open3D's registration_icp (see oracle/icp.py for the restated algorithm) runs
// entirely on-chip: source and target points are staged once in shared memory (f32, as given),
// every iteration transforms the source on the fly in fp64, finds each point's nearest target
// by exhaustive search over the staged target points, reduces the Umeyama moments of the
// accepted pairs across the CTA and solves the 3x3 SVD on one thread.
#pragma once
#include "common.cuh"
#include "rigid_device.cuh"

#define ICP_THREADS 128
#define ICP_WARPS (ICP_THREADS / 32)
#define ICP_SMEM_PTS 4096   // staged points (source + target) per CTA: 48 KB

struct IcpShared {
    double T[12];                 // current transform, row-major 3x4
    double red[ICP_WARPS][19];    // per-warp partials: 16 moments, error2, count, spare
    double fitness, rmse;
    int done;
};

struct IcpResult {
    double fitness, rmse;
    int iters;
    int fragile;   // OR of F4L_ICP_FRAGILE_*: a decision of the loop was within tie_eps of flipping
};
// Fragility of an ICP run (the tie[Q] of the boundary schema): the loop's result is a chain of discrete decisions; when one
// of them is within a relative tie_eps of going the other way, an implementation with a different fp64 operation order
// (Open3D's KD-tree + Eigen, the oracle) may take the other branch and end on a different, equally valid path.
// F4L_ICP_FRAGILE_NN / _INLIER / _STOP: include/f4l_b200.h
#define F4L_ICP_TIE_EPS 1e-9

__device__ __forceinline__ int icp_stop_fragile(double prev_fit, double fit, double prev_rmse, double rmse, double rel_fit,
                                                double rel_rmse, double tie_eps) {
    const bool f = fabs(fabs(prev_fit - fit) - rel_fit) <= tie_eps;
    const bool r = fabs(fabs(prev_rmse - rmse) - rel_rmse) <= tie_eps * fmax(fmax(prev_rmse, rmse), rel_rmse);
    return (f || r) ? F4L_ICP_FRAGILE_STOP : 0;
}

// src/tgt: base arrays; sidx/tidx: optional gathers; items [s0,s0+ns) and [t0,t0+nt).
// T0: 16 doubles row-major (or nullptr = identity).  Tout: 16 doubles.  corr (ns ints at s0) or
// nullptr.  pts: dynamic shared memory of 3*ICP_SMEM_PTS floats.
// FRAG: also track the nearest target with other coordinates than the winner and set IcpResult::fragile (one more
// compare per candidate in the exhaustive loop: +45 % on this kernel, so only callers that asked for the flag pay it).
template <bool FRAG>
__device__ inline IcpResult block_icp(const float* __restrict__ src, const int32_t* __restrict__ sidx,
                                      int s0, int ns, const float* __restrict__ tgt,
                                      const int32_t* __restrict__ tidx, int t0, int nt,
                                      const double* T0, double max_dist, int max_iter, double rel_fit,
                                      double rel_rmse, double* Tout, int32_t* __restrict__ corr,
                                      float* pts, IcpShared& sh, double tie_eps = F4L_ICP_TIE_EPS) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const bool staged = (ns + nt) <= ICP_SMEM_PTS;
    float* ssrc = pts;
    float* stgt = pts + 3 * ns;
    if (staged) {
        for (int i = tid; i < ns; i += ICP_THREADS) {
            float x, y, z;
            load_ptf(src, sidx, s0 + i, x, y, z);
            ssrc[3 * i] = x; ssrc[3 * i + 1] = y; ssrc[3 * i + 2] = z;
        }
        for (int j = tid; j < nt; j += ICP_THREADS) {
            float x, y, z;
            load_ptf(tgt, tidx, t0 + j, x, y, z);
            stgt[3 * j] = x; stgt[3 * j + 1] = y; stgt[3 * j + 2] = z;
        }
    }
    if (tid < 12) {
        double v = (tid % 5 == 0) ? 1.0 : 0.0;   // identity 3x4: entries 0,5,10
        if (T0) v = T0[tid];                      // rows 0..2 of the 4x4 are its first 12 entries
        sh.T[tid] = v;
    }
    if (tid == 0) sh.done = 0;
    __syncthreads();

    IcpResult out;
    out.fitness = 0.0; out.rmse = 0.0; out.iters = 0; out.fragile = 0;
    int frag = 0;
    if (ns <= 0 || nt <= 0) {
        if (tid < 16) Tout[tid] = tid < 12 ? sh.T[tid] : (tid == 15 ? 1.0 : 0.0);
        for (int i = tid; corr && i < ns; i += ICP_THREADS) corr[s0 + i] = -1;
        return out;
    }
    const double max_d2 = max_dist * max_dist;
    // pivot for the moment accumulation: first target point (pairs are within max_dist of it
    // up to the patch extent)
    double pv[3];
    {
        float x, y, z;
        if (staged) { x = stgt[0]; y = stgt[1]; z = stgt[2]; }
        else load_ptf(tgt, tidx, t0, x, y, z);
        pv[0] = x; pv[1] = y; pv[2] = z;
    }
    double prev_fit = 0.0, prev_rmse = 0.0;
    for (int it = 0;; ++it) {
        // ---- match under the current transform --------------------------------------------
        double T[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) T[a] = sh.T[a];
        Moments M;
        moments_zero(M);
        double err2 = 0.0, cnt = 0.0;
        for (int i = tid; i < ns; i += ICP_THREADS) {
            float fx, fy, fz;
            if (staged) { fx = ssrc[3 * i]; fy = ssrc[3 * i + 1]; fz = ssrc[3 * i + 2]; }
            else load_ptf(src, sidx, s0 + i, fx, fy, fz);
            const double px = T[0] * fx + T[1] * fy + T[2] * fz + T[3];
            const double py = T[4] * fx + T[5] * fy + T[6] * fz + T[7];
            const double pz = T[8] * fx + T[9] * fy + T[10] * fz + T[11];
            double best = INFINITY, second = INFINITY;     // second: nearest target with other coordinates than the winner
            int bj = -1;
            double bx = 0, by = 0, bz = 0;
            for (int j = 0; j < nt; ++j) {
                float gx, gy, gz;
                if (staged) { gx = stgt[3 * j]; gy = stgt[3 * j + 1]; gz = stgt[3 * j + 2]; }
                else load_ptf(tgt, tidx, t0 + j, gx, gy, gz);
                const double dx = px - (double)gx, dy = py - (double)gy, dz = pz - (double)gz;
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (d2 < best) { if (FRAG) second = best; best = d2; bj = j; bx = gx; by = gy; bz = gz; }
                else if (FRAG && d2 < second && !(d2 == best && (double)gx == bx && (double)gy == by && (double)gz == bz)) second = d2;
            }
            const bool ok = best < max_d2;     // strict, as the hybrid search of Open3D
            if (corr) corr[s0 + i] = ok ? bj : -1;
            if (FRAG && fabs(best - max_d2) <= tie_eps * max_d2) frag |= F4L_ICP_FRAGILE_INLIER;
            if (FRAG && ok && second - best <= tie_eps * best) frag |= F4L_ICP_FRAGILE_NN;
            if (ok) {
                err2 += best;
                cnt += 1.0;
                moments_add(M, 1.0, px - pv[0], py - pv[1], pz - pv[2], bx - pv[0], by - pv[1], bz - pv[2]);
            }
        }
        // ---- CTA reduction ----------------------------------------------------------------
        moments_warp_reduce(M);
        err2 = warp_sum(err2);
        cnt = warp_sum(cnt);
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 16; ++a) sh.red[wid][a] = M.m[a];
            sh.red[wid][16] = err2;
            sh.red[wid][17] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            Moments S;
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                double v = 0;
                for (int w = 0; w < ICP_WARPS; ++w) v += sh.red[w][a];
                S.m[a] = v;
            }
            double e2 = 0, c = 0;
            for (int w = 0; w < ICP_WARPS; ++w) { e2 += sh.red[w][16]; c += sh.red[w][17]; }
            const double fit = c > 0 ? c / (double)ns : 0.0;
            const double rmse = c > 0 ? sqrt(e2 / c) : 0.0;
            sh.fitness = fit;
            sh.rmse = rmse;
            bool stop = false;
            if (it > 0 && fabs(prev_fit - fit) < rel_fit && fabs(prev_rmse - rmse) < rel_rmse) stop = true;
            if (FRAG && it > 0 && it < max_iter) frag |= icp_stop_fragile(prev_fit, fit, prev_rmse, rmse, rel_fit, rel_rmse, tie_eps);
            if (it >= max_iter) stop = true;
            if (!stop) {
                // U = umeyama(P[corr], tgt[corr]);  T <- U T
                double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
                if (c > 0) fit_from_moments(S, pv, pv, 0.0, 2, R, t);
                double Tn[12];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        Tn[r * 4 + cc] = R[r * 3 + 0] * sh.T[0 * 4 + cc] + R[r * 3 + 1] * sh.T[1 * 4 + cc] +
                                         R[r * 3 + 2] * sh.T[2 * 4 + cc] + (cc == 3 ? t[r] : 0.0);
                }
#pragma unroll
                for (int a = 0; a < 12; ++a) sh.T[a] = Tn[a];
            }
            sh.done = stop ? 1 : 0;
            prev_fit = fit;
            prev_rmse = rmse;
        }
        __syncthreads();
        if (sh.done) {
            out.iters = it;
            break;
        }
    }
    out.fitness = sh.fitness;
    out.rmse = sh.rmse;
    // __syncthreads_or is a logical OR of predicates: one round per flag bit
    if (FRAG)
        out.fragile = (__syncthreads_or(frag & 1) ? 1 : 0) | (__syncthreads_or(frag & 2) ? 2 : 0) | (__syncthreads_or(frag & 4) ? 4 : 0);
    if (tid < 16) Tout[tid] = tid < 12 ? sh.T[tid] : (tid == 15 ? 1.0 : 0.0);
    return out;
}
