// Device building blocks shared by the standalone rigid kernels and the fused fine-matching
// kernels: point loads, the warp-level weighted moment reduction + fit, the block-level
// rigidity statistic.
#pragma once
#include "common.cuh"

__device__ __forceinline__ void load_pt(const float* __restrict__ pts, const int32_t* __restrict__ idx,
                                        int k, double& x, double& y, double& z) {
    const size_t j = idx ? (size_t)idx[k] : (size_t)k;
    const float* p = pts + j * 3;
    x = (double)__ldg(p);
    y = (double)__ldg(p + 1);
    z = (double)__ldg(p + 2);
}
__device__ __forceinline__ void load_ptf(const float* __restrict__ pts, const int32_t* __restrict__ idx,
                                         int k, float& x, float& y, float& z) {
    const size_t j = idx ? (size_t)idx[k] : (size_t)k;
    const float* p = pts + j * 3;
    x = __ldg(p);
    y = __ldg(p + 1);
    z = __ldg(p + 2);
}

// Raw weighted moments about pivots (ps, pt): m[0]=sum w, m[1..3]=sum w s', m[4..6]=sum w t',
// m[7..15]=sum w s' t'^T (row-major).  Reduced over the warp; every lane gets the totals.
struct Moments {
    double m[16];
};

__device__ __forceinline__ void moments_zero(Moments& M) {
#pragma unroll
    for (int i = 0; i < 16; ++i) M.m[i] = 0.0;
}
__device__ __forceinline__ void moments_add(Moments& M, double w, double sx, double sy, double sz,
                                            double tx, double ty, double tz) {
    M.m[0] += w;
    const double wsx = w * sx, wsy = w * sy, wsz = w * sz;
    M.m[1] += wsx; M.m[2] += wsy; M.m[3] += wsz;
    M.m[4] += w * tx; M.m[5] += w * ty; M.m[6] += w * tz;
    M.m[7] += wsx * tx; M.m[8] += wsx * ty; M.m[9] += wsx * tz;
    M.m[10] += wsy * tx; M.m[11] += wsy * ty; M.m[12] += wsy * tz;
    M.m[13] += wsz * tx; M.m[14] += wsz * ty; M.m[15] += wsz * tz;
}
__device__ __forceinline__ void moments_warp_reduce(Moments& M) {
#pragma unroll
    for (int i = 0; i < 16; ++i) M.m[i] = warp_sum(M.m[i]);
}

// Turn raw moments (weights NOT yet normalised) into R, t following the reference formulas.
//   variant 0 (scripts/weighted_svd.py:99-117): wn = w/(W+eps); c = sum wn x; H = sum wn (s-cs)(t-ct)^T
//   variant 1 (src/functions.py:36-80):        wn = w/(W+eps); c = sum wn x / (sum wn + eps); same H form
//   variant 2 (Eigen::umeyama, no scaling):    wn = w/W (true means), S(2) = -1 iff det<0
// ps, pt: pivots the moments were taken about.  Returns true when the result is not finite.
__device__ inline bool fit_from_moments(const Moments& M, const double ps[3], const double pt[3],
                                        double eps, int variant, double R[9], double t[3], double* warm = nullptr) {
    const double W = M.m[0];
    const double inv = (variant == 2) ? 1.0 / W : 1.0 / (W + eps);
    const double f = W * inv;  // sum of normalised weights
    double ms[3], mt[3], Mst[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ms[i] = M.m[1 + i] * inv;
        mt[i] = M.m[4 + i] * inv;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) Mst[i] = M.m[7 + i] * inv;
    // centroid used by the reference, expressed in pivot-local coordinates: c' = c - pivot
    double a[3], b[3];
    if (variant == 1) {
        const double g = 1.0 / (f + eps);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            a[i] = (ms[i] - eps * ps[i]) * g;   // (f p + ms)/(f+eps) - p
            b[i] = (mt[i] - eps * pt[i]) * g;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            a[i] = ms[i] - (1.0 - f) * ps[i];   // f p + ms - p
            b[i] = mt[i] - (1.0 - f) * pt[i];
        }
    }
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            H[i * 3 + j] = Mst[i * 3 + j] - ms[i] * b[j] - a[i] * mt[j] + f * a[i] * b[j];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 9; ++i) bad |= !isfinite(H[i]);
    if (bad || !(W > 0.0) && variant == 2) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        t[0] = t[1] = t[2] = 0;
        return true;
    }
    double U[9], S[3], V[9];
    svd3x3(H, U, S, V, warm);
    rotation_from_svd(U, V, variant, R);
    double cs[3], ct[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cs[i] = ps[i] + a[i];
        ct[i] = pt[i] + b[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = ct[i] - (R[i * 3 + 0] * cs[0] + R[i * 3 + 1] * cs[1] + R[i * 3 + 2] * cs[2]);
    return false;
}

// One warp fits one segment: items [s0, s0+n).  All lanes return R, t.
__device__ inline bool warp_fit_segment(const float* __restrict__ src, const float* __restrict__ tgt,
                                        const int32_t* __restrict__ src_idx,
                                        const int32_t* __restrict__ tgt_idx, const float* __restrict__ w,
                                        int s0, int n, double eps, float weight_thresh, int variant,
                                        int lane, double R[9], double t[3], double* warm = nullptr) {
    if (n <= 0) {
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
        t[0] = t[1] = t[2] = 0;
        return true;
    }
    double ps[3], pt[3];
    load_pt(src, src_idx, s0, ps[0], ps[1], ps[2]);
    load_pt(tgt, tgt_idx, s0, pt[0], pt[1], pt[2]);
    Moments M;
    moments_zero(M);
    for (int i = lane; i < n; i += 32) {
        const int k = s0 + i;
        double sx, sy, sz, tx, ty, tz;
        load_pt(src, src_idx, k, sx, sy, sz);
        load_pt(tgt, tgt_idx, k, tx, ty, tz);
        double wi = 1.0;
        if (w) {
            float wf = __ldg(w + k);
            if (variant == 0 && wf < weight_thresh) wf = 0.f;
            wi = (double)wf;
        }
        moments_add(M, wi, sx - ps[0], sy - ps[1], sz - ps[2], tx - pt[0], ty - pt[1], tz - pt[2]);
    }
    moments_warp_reduce(M);
    return fit_from_moments(M, ps, pt, eps, variant, R, t, warm);
}

// Block-level rigidity statistic over items [s0, s0+n): sum over pairs i<j of |dS_ij - dT_ij|
// and the count of pairs with |.| <= thres.  f32 distances like torch.cdist (direct differences;
// torch uses the matmul formulation only above 25 rows -- rounding differs in the last ulp, far
// below the 0.5 m threshold scale), fp64 accumulation.  sm needs 6*min(n,cap) floats.
template <int CAP = 2048>
__device__ inline void block_rigidity(const float* __restrict__ src, const float* __restrict__ tgt,
                                      const int32_t* __restrict__ src_idx,
                                      const int32_t* __restrict__ tgt_idx, int s0, int n, float thres,
                                      float* sm, double* red_sum, unsigned long long* red_cnt,
                                      double& out_sum, unsigned long long& out_cnt) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool staged = n <= CAP;
    if (staged) {
        for (int i = tid; i < n; i += nt) {
            float x, y, z;
            load_ptf(src, src_idx, s0 + i, x, y, z);
            sm[i * 6 + 0] = x; sm[i * 6 + 1] = y; sm[i * 6 + 2] = z;
            load_ptf(tgt, tgt_idx, s0 + i, x, y, z);
            sm[i * 6 + 3] = x; sm[i * 6 + 4] = y; sm[i * 6 + 5] = z;
        }
    }
    __syncthreads();
    double sum = 0.0;
    unsigned long long cnt = 0;
    // row i handled by thread i (mod nt); inner loop over j > i
    for (int i = tid; i < n; i += nt) {
        float ax, ay, az, bx, by, bz;
        if (staged) {
            ax = sm[i * 6]; ay = sm[i * 6 + 1]; az = sm[i * 6 + 2];
            bx = sm[i * 6 + 3]; by = sm[i * 6 + 4]; bz = sm[i * 6 + 5];
        } else {
            load_ptf(src, src_idx, s0 + i, ax, ay, az);
            load_ptf(tgt, tgt_idx, s0 + i, bx, by, bz);
        }
        float rowsum = 0.f;
        unsigned rowcnt = 0;
        // pair (i, j) and (n-1-i, .) balance: iterate j over all != i, count each pair once by j > i
        for (int j = i + 1; j < n; ++j) {
            float cx, cy, cz, dx, dy, dz;
            if (staged) {
                cx = sm[j * 6]; cy = sm[j * 6 + 1]; cz = sm[j * 6 + 2];
                dx = sm[j * 6 + 3]; dy = sm[j * 6 + 4]; dz = sm[j * 6 + 5];
            } else {
                load_ptf(src, src_idx, s0 + j, cx, cy, cz);
                load_ptf(tgt, tgt_idx, s0 + j, dx, dy, dz);
            }
            float ex = ax - cx, ey = ay - cy, ez = az - cz;
            float fx = bx - dx, fy = by - dy, fz = bz - dz;
            float ds = sqrtf(ex * ex + ey * ey + ez * ez);
            float dt = sqrtf(fx * fx + fy * fy + fz * fz);
            float diff = fabsf(ds - dt);
            rowsum += diff;
            rowcnt += (diff <= thres) ? 1u : 0u;
        }
        sum += (double)rowsum;
        cnt += rowcnt;
    }
    sum = warp_sum(sum);
    int c32 = warp_sum((int)cnt);  // per-thread counts fit in int: n <= ~1e5 rows * cols / threads
    const int wid = tid >> 5, lane = tid & 31;
    if (lane == 0) {
        red_sum[wid] = sum;
        red_cnt[wid] = (unsigned long long)(unsigned)c32;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0;
        unsigned long long c = 0;
        for (int k = 0; k < (nt >> 5); ++k) {
            s += red_sum[k];
            c += red_cnt[k];
        }
        out_sum = s;
        out_cnt = c;
    }
    __syncthreads();
}
