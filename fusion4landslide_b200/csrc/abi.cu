// Error plumbing and version of libf4l_b200.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void f4l_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int f4l_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        f4l_set_error("%s: %s", what, cudaGetErrorString(e));
        return F4L_E_CUDA;
    }
    return F4L_OK;
}

extern "C" int f4l_abi_version(void) { return F4L_ABI_VERSION; }
extern "C" const char* f4l_last_error(void) { return g_err; }
