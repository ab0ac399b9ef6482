// Peer-visible buffers for the fused displacement-field exchange (include/f4l_b200.h, f4l_peer_*).
// One process per GPU: the owner exports a CUDA IPC handle, every other rank maps it and passes the
// mapped pointer to f4l_fine_matching (f4l_fine_buffers.peer_dense), whose D5 kernel then stores each
// dense row into all peers over NVLink while it computes -- no separate collective for the field.
#include <string.h>

#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == F4L_PEER_HANDLE_BYTES, "CUDA IPC handle size");

#define PEER_CUDA(call, what)                                             \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) {                                          \
            f4l_set_error("%s: %s", what, cudaGetErrorString(e_));        \
            cudaGetLastError();                                           \
            return F4L_E_CUDA;                                            \
        }                                                                 \
    } while (0)

extern "C" int f4l_peer_alloc(size_t bytes, void** d_ptr, unsigned char* h_handle) {
    F4L_REQUIRE(d_ptr && h_handle && bytes > 0, "null pointer / empty buffer");
    void* p = nullptr;
    PEER_CUDA(cudaMalloc(&p, bytes), "f4l_peer_alloc/cudaMalloc");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        f4l_set_error("f4l_peer_alloc/cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(p);
        return F4L_E_CUDA;
    }
    memcpy(h_handle, &h, sizeof(h));
    *d_ptr = p;
    return F4L_OK;
}

extern "C" int f4l_peer_open(const unsigned char* h_handle, void** d_ptr) {
    F4L_REQUIRE(d_ptr && h_handle, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    void* p = nullptr;
    PEER_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "f4l_peer_open/cudaIpcOpenMemHandle");
    *d_ptr = p;
    return F4L_OK;
}

extern "C" int f4l_peer_close(void* d_ptr) {
    F4L_REQUIRE(d_ptr, "null pointer");
    PEER_CUDA(cudaIpcCloseMemHandle(d_ptr), "f4l_peer_close");
    return F4L_OK;
}

extern "C" int f4l_peer_free(void* d_ptr) {
    F4L_REQUIRE(d_ptr, "null pointer");
    PEER_CUDA(cudaFree(d_ptr), "f4l_peer_free");
    return F4L_OK;
}

extern "C" int f4l_peer_enable_access(int32_t peer_device) {
    int cur = -1;
    PEER_CUDA(cudaGetDevice(&cur), "f4l_peer_enable_access/cudaGetDevice");
    if (cur == peer_device) return F4L_OK;
    int can = 0;
    PEER_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device), "f4l_peer_enable_access/cudaDeviceCanAccessPeer");
    if (!can) {
        f4l_set_error("f4l_peer_enable_access: device %d cannot access device %d", cur, (int)peer_device);
        return F4L_E_CUDA;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return F4L_OK; }
    PEER_CUDA(e, "f4l_peer_enable_access/cudaDeviceEnablePeerAccess");
    return F4L_OK;
}
