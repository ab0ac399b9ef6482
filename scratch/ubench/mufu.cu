// micro-benchmark: MUFU.SQRT / MUFU.RSQ / FFMA2 / LDS throughput per SM with 16 warps/SM (4 per scheduler)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, int iters, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x + i;
    unsigned long long p[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i + 4]);
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += 128) sm[i] = i;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i & 3]));
            if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (MODE == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)(((threadIdx.x * 2 + i + it) & 4095) * 4))); a[i] += v; }
            if (MODE == 5) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)(((threadIdx.x + i + it) & 4095) * 4))); a[i] += v; }
            if (MODE == 6) asm volatile("sqrt.approx.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (float)(p[i] & 0xffff);
    out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* out) {
    const int iters = 4096, blocks = 148 * 4;
    k<MODE><<<blocks, 128>>>(out, 16, 1.f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128>>>(out, iters, 1.f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // warp-instructions per SM = 16 warps * iters * 8
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-28s %.3f ms  -> %.2f cycles per warp-instr per SM-quadrant (4 warps/quadrant)\n", name, ms, cyc / (4.0 * iters * 8));
}
extern "C" int ubench_main() {
    float* out; cudaMalloc(&out, 148 * 4 * 128 * 4);
    run<0>("MUFU.SQRT ftz", out);
    run<6>("sqrt.approx (no ftz)", out);
    run<1>("MUFU.RSQ ftz", out);
    run<2>("FFMA2", out);
    run<3>("FFMA", out);
    return 0;
}
