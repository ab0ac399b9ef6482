// Shared device helpers for libf4l_b200 (sm_100a only).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/f4l_b200.h"

#define F4L_WARP 32
#define F4L_FULL 0xffffffffu

// ---- host-side error plumbing -------------------------------------------------------------
void f4l_set_error(const char* fmt, ...);
int f4l_check_launch(const char* what);
void f4l_count_launches(int n);   // bookkeeping for f4l_launch_count()
// Timeline mark on `st` before a kernel launch: counts the launch and, when profiling is enabled
// (f4l_profile_enable), records a CUDA event; the time between two consecutive marks of a stream
// is attributed to the kernel named by the first.  name == nullptr closes the current interval.
void f4l_mark(const char* name, cudaStream_t st);
int f4l_finish(const char* what, void* stream);   // closing mark + launch error check

#define F4L_REQUIRE(cond, msg)                       \
    do {                                             \
        if (!(cond)) {                               \
            f4l_set_error("%s: %s", __func__, msg);  \
            return F4L_E_ARG;                        \
        }                                            \
    } while (0)

static inline int f4l_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// Per-device one-time setup (cudaFuncSetAttribute and memory-pool attributes belong to the CURRENT device, so a
// process-wide flag would skip them on the second GPU a process touches).  Usage:
//     static F4lPerDevice once;  if (!once.done()) { if (f4l_optin_smem(k, bytes) ...) once.mark(); }
struct F4lPerDevice {
    unsigned long long mask[2] = {0ull, 0ull};      // 128 devices; benign race: the setup is idempotent
    static int dev() { int d = 0; cudaGetDevice(&d); return d & 127; }
    bool done() const { const int d = dev(); return (__atomic_load_n(&mask[d >> 6], __ATOMIC_ACQUIRE) >> (d & 63)) & 1ull; }
    void mark() { const int d = dev(); __atomic_fetch_or(&mask[d >> 6], 1ull << (d & 63), __ATOMIC_RELEASE); }
};
// opt-in to `bytes` of dynamic shared memory for kernel `fn` on the current device; false (+ error text) on failure
template <class F>
static inline bool f4l_optin_smem(F* fn, size_t bytes, const char* name) {
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        f4l_set_error("cudaFuncSetAttribute(%s, %zu B dynamic shared memory): %s", name, bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return false;
    }
    return true;
}

// ---- warp reductions ----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(F4L_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(F4L_FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(F4L_FULL, v, o);
    return v;
}

// order-preserving map float <-> uint32 (radix select, atomic min/max on floats)
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Block-wide k-th smallest (0-based rank) of vals[0..n) by 4 passes of an 8-bit radix select.
// hist: 256 unsigned in shared memory, state: 2 unsigned in shared memory.  All threads return it.
__device__ inline float block_select_kth(const float* vals, int n, int rank, unsigned* hist, unsigned* state) {
    if (threadIdx.x == 0) { state[0] = 0u; state[1] = (unsigned)rank; }
    unsigned mask = 0u;
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = state[0];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned u = f2ord(vals[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned r = state[1], cum = 0u;
            int b = 0;
            for (b = 0; b < 256; ++b) {
                if (r < cum + hist[b]) break;
                cum += hist[b];
            }
            if (b == 256) b = 255;
            state[0] = prefix | ((unsigned)b << shift);
            state[1] = r - cum;
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    return ord2f(state[0]);
}

// segment bounds: CSR when count == nullptr
__device__ __forceinline__ void seg_bounds(const int32_t* start, const int32_t* count, int q,
                                           int& s, int& n) {
    s = start[q];
    n = count ? count[q] : (start[q + 1] - s);
}

// ---- 3x3 SVD, fp64, one-sided Jacobi (Hestenes) ---------------------------------------------
// H (row-major) = U diag(S) V^T with S descending (torch.svd convention).  U, V row-major.
// Rank-deficient columns of U are completed to an orthonormal basis.
// warm (optional, 9 doubles, in/out): an orthonormal basis (v[c][r] at warm[c*3+r]) that nearly diagonalises
// H^T H -- the basis a previous, similar H converged to.  The iteration then starts from a = H warm instead of
// a = H, and one or two sweeps are enough (an ICP loop fits a slowly changing covariance up to 31 times).  The
// converged basis is written back.  Any orthonormal start gives the same decomposition up to rounding.
// Columns p, q count as orthogonal when cos^2 of their angle is below F4L_SVD_THR = 1e-30 (|cos| <= 1e-15).  A tighter
// bound (round 1: 1.44e-32) sits inside the rounding noise of the dot product, so converged columns kept "rotating by
// noise" for one or two extra sweeps -- ~180 dependent fp64 instructions each on chains that are pure latency (the
// tail of k_kabsch_fused, every ICP iteration of the fit kernels).  The decomposition moves by ~1e-15 relative.
#ifndef F4L_SVD_THR
#define F4L_SVD_THR 1e-30
#endif
__device__ inline void svd3x3(const double H[9], double U[9], double S[3], double V[9], double* warm = nullptr) {
    double a[3][3];  // a[c][r]: column c
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};  // v[c][r]
    if (warm) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) v[c][r] = warm[c * 3 + r];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) a[c][r] = H[r * 3 + 0] * v[c][0] + H[r * 3 + 1] * v[c][1] + H[r * 3 + 2] * v[c][2];
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) a[c][r] = H[r * 3 + c];
    }
    // Rotation angles without divisions or square roots on the dependency chain: with a = beta - alpha,
    // b = 2 gamma, r = sqrt(a^2 + b^2) the Hestenes rotation (smaller root of tan 2x = b/a) is
    //   cos 2x = |a|/r,  c = sqrt((1 + cos 2x)/2),  s = sign(a) b / (2 r c)
    // -- two rsqrt per rotation instead of three divisions and three square roots (fp64 div/sqrt are
    // multi-hundred-cycle sequences and this chain is pure latency for the warp that owns the patch).
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            double alpha = a[p][0] * a[p][0] + a[p][1] * a[p][1] + a[p][2] * a[p][2];
            double beta = a[q][0] * a[q][0] + a[q][1] * a[q][1] + a[q][2] * a[q][2];
            double gamma = a[p][0] * a[q][0] + a[p][1] * a[q][1] + a[p][2] * a[q][2];
            if (gamma == 0.0 || gamma * gamma <= F4L_SVD_THR * (alpha * beta)) continue;
            rotated = true;
            const double da = beta - alpha, db = 2.0 * gamma;
            const double inv_r = rsqrt(da * da + db * db);
            const double x = 0.5 + 0.5 * fabs(da) * inv_r;       // cos^2 of the rotation angle, in [0.5, 1]
            const double inv_c = rsqrt(x);
            const double c = x * inv_c;
            const double s = copysign(0.5, da) * db * inv_r * inv_c;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double ap = a[p][r], aq = a[q][r];
                a[p][r] = c * ap - s * aq;
                a[q][r] = s * ap + c * aq;
                double vp = v[p][r], vq = v[q][r];
                v[p][r] = c * vp - s * vq;
                v[q][r] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    if (warm) {
        __syncwarp();       // a warp may share one slot (all lanes hold the same values): no lane reads after a write
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r) warm[c * 3 + r] = v[c][r];
    }
    double sv[3], n2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        n2[c] = a[c][0] * a[c][0] + a[c][1] * a[c][1] + a[c][2] * a[c][2];
        sv[c] = n2[c] > 0.0 ? n2[c] * rsqrt(n2[c]) : 0.0;
    }
    // sort columns by singular value, descending: a 3-element network of conditional column swaps (selects on
    // registers -- indexing a[][] / v[][] with a run-time column number would push both arrays, and with them
    // every rotation of the sweeps above, into local memory)
#define F4L_SVD_CSWAP(i, j)                                                                  \
    {                                                                                        \
        const bool sw = sv[i] < sv[j];                                                       \
        double tmp;                                                                          \
        tmp = sv[i]; sv[i] = sw ? sv[j] : tmp; sv[j] = sw ? tmp : sv[j];                     \
        tmp = n2[i]; n2[i] = sw ? n2[j] : tmp; n2[j] = sw ? tmp : n2[j];                     \
        _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                      \
            tmp = a[i][r]; a[i][r] = sw ? a[j][r] : tmp; a[j][r] = sw ? tmp : a[j][r];       \
            tmp = v[i][r]; v[i][r] = sw ? v[j][r] : tmp; v[j][r] = sw ? tmp : v[j][r];       \
        }                                                                                    \
    }
    F4L_SVD_CSWAP(0, 1)
    F4L_SVD_CSWAP(1, 2)
    F4L_SVD_CSWAP(0, 1)
#undef F4L_SVD_CSWAP
    double u[3][3];
    const double tiny = 1e-300;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        S[c] = sv[c];
        double inv = sv[c] > tiny ? rsqrt(n2[c]) : 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            u[c][r] = a[c][r] * inv;
            V[r * 3 + c] = v[c][r];
        }
    }
    // complete U for (numerically) zero singular values
    const double rel = 1e-14 * S[0];
    if (!(S[0] > tiny)) {
        u[0][0] = 1; u[0][1] = 0; u[0][2] = 0;
        u[1][0] = 0; u[1][1] = 1; u[1][2] = 0;
        u[2][0] = 0; u[2][1] = 0; u[2][2] = 1;
    } else {
        if (!(S[1] > rel)) {
            // any unit vector orthogonal to u0
            int m = 0;
            double d = u[0][0];
            if (fabs(u[0][1]) < fabs(d)) { m = 1; d = u[0][1]; }
            if (fabs(u[0][2]) < fabs(d)) { m = 2; d = u[0][2]; }
            const double e[3] = {m == 0 ? 1.0 : 0.0, m == 1 ? 1.0 : 0.0, m == 2 ? 1.0 : 0.0};
            double w0 = e[0] - d * u[0][0], w1 = e[1] - d * u[0][1], w2 = e[2] - d * u[0][2];
            double nrm = rsqrt(w0 * w0 + w1 * w1 + w2 * w2);
            u[1][0] = w0 * nrm; u[1][1] = w1 * nrm; u[1][2] = w2 * nrm;
        }
        if (!(S[2] > rel)) {
            u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
            u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
            u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) U[r * 3 + c] = u[c][r];
}

__device__ __forceinline__ double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// R = V diag(1,1,d) U^T.  mode 0: d = sign(det(V U^T)) (weighted_svd.py:114); mode 1: raw det
// (functions.py:73-77); mode 2: Eigen::umeyama S(2) = -1 iff det(U) det(V) < 0 else +1.
__device__ inline void rotation_from_svd(const double U[9], const double V[9], int mode, double R[9]) {
    double dd = det3(U) * det3(V);
    double d;
    if (mode == 0) d = (dd > 0.0) ? 1.0 : ((dd < 0.0) ? -1.0 : 0.0);
    else if (mode == 1) d = dd;
    else d = (dd < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[i * 3 + j] = V[i * 3 + 0] * U[j * 3 + 0] + V[i * 3 + 1] * U[j * 3 + 1] + d * V[i * 3 + 2] * U[j * 3 + 2];
}
