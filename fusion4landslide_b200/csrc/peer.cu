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

// ---------------------------------------------------------------------------------------------------------------
// f4l_peer_push: the displacement-field all-gather as a copy kernel of its own, tile by tile, next to the compute.
// The fused variant (k_apply_assign storing into the peers) concentrates 7 x 18.7 MB of NVLink traffic into the
// 0.15 ms that kernel runs -- more than the links carry -- and stretches it (measured at 8 GPUs: 5.10 ms per step
// against 4.11 ms without any exchange).  Here a few CTAs per tile move the finished rows on a side stream while the
// NEXT tile's fits run: local HBM -> shared memory (one TMA bulk load) -> every peer's field (one TMA bulk store per
// peer, the shared-memory tile is read by the copy engine of the SM, the threads do not touch the data), double
// buffered, one elected thread per CTA.  The row count is a device scalar (f4l_fine_buffers.counts[0]): no host
// round trip.
#define PUSH_CHUNK 32768u
#define PUSH_STAGES 3

struct PushPeers { char* p[F4L_MAX_PEERS]; };

__device__ __forceinline__ uint32_t push_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k_peer_push(const char* __restrict__ src, const int32_t* __restrict__ d_rows,
                                                  int32_t row_bytes, long long max_bytes, int n_peers, PushPeers peers) {
    extern __shared__ __align__(128) unsigned char push_sm[];
    __shared__ unsigned long long bar[PUSH_STAGES];
    long long bytes = (long long)d_rows[0] * row_bytes;
    if (bytes > max_bytes) bytes = max_bytes;
    if (bytes <= 0 || n_peers <= 0) return;
    // 16-byte granularity of the bulk copies: an 8-byte head when the slot starts at 8 mod 16 (rows are 24 bytes and
    // every rank's field has the same layout, so source and destinations share the misalignment), a < 16-byte tail
    const unsigned head = (unsigned)((16u - ((uintptr_t)src & 15u)) & 15u);
    const long long hb = head < bytes ? head : bytes;
    const long long body = (bytes - hb) & ~15ll;
    const long long tail0 = hb + body;
    if (blockIdx.x == 0) {
        for (long long o = threadIdx.x; o < hb + (bytes - tail0); o += 32) {
            const long long off = o < hb ? o : tail0 + (o - hb);
            const char v = src[off];
            for (int p = 0; p < n_peers; ++p) peers.p[p][off] = v;
        }
    }
    if (threadIdx.x != 0 || body == 0) return;
    for (int s = 0; s < PUSH_STAGES; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(push_smem_u32(&bar[s])), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long n_chunks = (body + PUSH_CHUNK - 1) / PUSH_CHUNK;
    unsigned it = 0;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const unsigned s = it % PUSH_STAGES, ph = (it / PUSH_STAGES) & 1u;
        const long long off = hb + c * (long long)PUSH_CHUNK;
        const unsigned nb = (unsigned)((body - c * (long long)PUSH_CHUNK) < (long long)PUSH_CHUNK ? (body - c * (long long)PUSH_CHUNK) : PUSH_CHUNK);
        // the stores that read this stage two chunks ago must have read their source
        asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PUSH_STAGES - 1) : "memory");
        const uint32_t sm = push_smem_u32(push_sm + (size_t)s * PUSH_CHUNK), mb = push_smem_u32(&bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sm),
                     "l"(src + off), "r"(nb), "r"(mb)
                     : "memory");
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.b32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(mb), "r"(ph)
                : "memory");
        }
        for (int p = 0; p < n_peers; ++p)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(peers.p[p] + off), "r"(sm), "r"(nb)
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// fallback when a destination does not share the source's alignment mod 16 (no bulk copy possible): plain 4-byte stores
__global__ void __launch_bounds__(256) k_peer_push_plain(const unsigned* __restrict__ src, const int32_t* __restrict__ d_rows,
                                                         int32_t row_words, long long max_words, int n_peers, PushPeers peers) {
    long long words = (long long)d_rows[0] * row_words;
    if (words > max_words) words = max_words;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < words; i += (long long)gridDim.x * blockDim.x) {
        const unsigned v = __ldg(src + i);
        for (int p = 0; p < n_peers; ++p) reinterpret_cast<unsigned*>(peers.p[p])[i] = v;
    }
}

extern "C" int f4l_peer_push(const void* src, const int32_t* d_rows, int32_t row_bytes, int64_t max_rows,
                             void* const* peer_dst, int32_t n_peers, int32_t n_ctas, void* stream) {
    F4L_REQUIRE(n_peers >= 0 && n_peers <= F4L_MAX_PEERS, "n_peers out of range");
    F4L_REQUIRE(row_bytes > 0 && max_rows >= 0, "bad size");
    if (n_peers == 0 || max_rows == 0) return F4L_OK;
    F4L_REQUIRE(src && d_rows && peer_dst, "null pointer");
    PushPeers pp;
    bool aligned = true;
    for (int p = 0; p < F4L_MAX_PEERS; ++p) {
        pp.p[p] = p < n_peers ? (char*)peer_dst[p] : nullptr;
        F4L_REQUIRE(p >= n_peers || pp.p[p], "peer pointer is null");
        if (p < n_peers && (((uintptr_t)pp.p[p] ^ (uintptr_t)src) & 15u) != 0) aligned = false;
    }
    if (!aligned) {
        F4L_REQUIRE(row_bytes % 4 == 0 && ((uintptr_t)src & 3u) == 0, "rows must be 4-byte aligned");
        for (int p = 0; p < n_peers; ++p) F4L_REQUIRE(((uintptr_t)pp.p[p] & 3u) == 0, "rows must be 4-byte aligned");
        f4l_mark("k_peer_push_plain", (cudaStream_t)stream);
        k_peer_push_plain<<<148 * 2, 256, 0, (cudaStream_t)stream>>>((const unsigned*)src, d_rows, row_bytes / 4,
                                                                    (long long)max_rows * (row_bytes / 4), n_peers, pp);
        return f4l_finish("f4l_peer_push", stream);
    }
    const size_t smem = (size_t)PUSH_STAGES * PUSH_CHUNK;
    static F4lPerDevice once;
    if (!once.done()) {
        if (!f4l_optin_smem(k_peer_push, smem, "k_peer_push")) return F4L_E_CUDA;
        once.mark();
    }
    const long long max_bytes = (long long)max_rows * row_bytes;
    long long chunks = (max_bytes + PUSH_CHUNK - 1) / PUSH_CHUNK;
    int grid = n_ctas > 0 ? n_ctas : 32;
    if (grid > chunks) grid = (int)(chunks > 0 ? chunks : 1);
    f4l_mark("k_peer_push", (cudaStream_t)stream);
    k_peer_push<<<grid, 32, smem, (cudaStream_t)stream>>>((const char*)src, d_rows, row_bytes, max_bytes, n_peers, pp);
    return f4l_finish("f4l_peer_push", stream);
}
