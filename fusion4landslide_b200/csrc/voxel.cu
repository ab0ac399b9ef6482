// Voxel subsampling: Open3D PointCloud::VoxelDownSample as the reference calls it in
// base.py:1024-1025 (and :906-907, f2s3 tiles), the step right before A2 / B2.
//
// Open3D 0.19 (geometry/PointCloud.cpp, published algorithm; the wheel is absent here -> parity unpinned like the
// other Open3D pieces): voxel_min_bound = min_bound - voxel_size / 2; a point falls in voxel
// floor((p - voxel_min_bound) / voxel_size) per axis; every occupied voxel yields the MEAN of its points, summed in
// fp64 in point order.  Open3D emits the voxels in std::unordered_map iteration order (implementation-defined);
// here they come out in ascending (ix, iy, iz) order -- a permutation of the reference's rows.
//
// keys (3 x 21 bits) -> stable radix sort of (key, index) -> head flags + scan = voxel id per sorted row -> one
// thread per voxel sums its points in index order (bit-identical to a sequential accumulation) and divides.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

struct VoxFrame {
    unsigned long long lo[3], hi[3];   // ordered-uint images of the bounding box
    int overflow;                      // 1: more than 2^21 voxels along an axis
    int pad;
};

__device__ __forceinline__ unsigned long long vox_d2ord(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double vox_ord2d(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7fffffffffffffffull) : ~u));
}

__global__ void k_vox_init(VoxFrame* f) {
    if (threadIdx.x < 3) { f->lo[threadIdx.x] = ~0ull; f->hi[threadIdx.x] = 0ull; }
    if (threadIdx.x == 0) f->overflow = 0;
}

__global__ void __launch_bounds__(256) k_vox_bbox(const double* __restrict__ p, int n, VoxFrame* f) {
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
                atomicMin(&f->lo[c], vox_d2ord(mn[c]));
                atomicMax(&f->hi[c], vox_d2ord(mx[c]));
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_vox_keys(const double* __restrict__ p, int n, double voxel, VoxFrame* f,
                                                  unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    double vmb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) vmb[c] = vox_ord2d(f->lo[c]) - voxel * 0.5;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long key = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double r = floor((p[(size_t)i * 3 + c] - vmb[c]) / voxel);        // PointCloud.cpp: ref_coord, floor
            long long q = (long long)r;
            if (q < 0) q = 0;
            if (q >= (1ll << 21)) { q = (1ll << 21) - 1; f->overflow = 1; }
            key = (key << 21) | (unsigned long long)q;
        }
        keys[i] = key;
        vals[i] = i;
    }
}

__global__ void __launch_bounds__(256) k_vox_heads(const unsigned long long* __restrict__ keys, int n, int* __restrict__ head) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        head[j] = (j == 0 || keys[j] != keys[j - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_vox_mean(const double* __restrict__ p, int n, const unsigned long long* __restrict__ keys,
                                                  const int* __restrict__ order, const int* __restrict__ vid_incl,
                                                  const VoxFrame* __restrict__ f, double* __restrict__ centroids,
                                                  int32_t* __restrict__ voxel_of_point, int32_t* __restrict__ counts) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int id = vid_incl[j] - 1;
        if (voxel_of_point) voxel_of_point[order[j]] = id;
        if (j != 0 && keys[j] == keys[j - 1]) continue;          // not the first row of its voxel
        double sx = 0, sy = 0, sz = 0;
        int c = 0;
        for (int u = j; u < n && keys[u] == keys[j]; ++u) {      // index order (stable sort): a sequential sum
            const int i = order[u];
            sx += p[(size_t)i * 3]; sy += p[(size_t)i * 3 + 1]; sz += p[(size_t)i * 3 + 2];
            ++c;
        }
        centroids[(size_t)id * 3] = sx / (double)c;
        centroids[(size_t)id * 3 + 1] = sy / (double)c;
        centroids[(size_t)id * 3 + 2] = sz / (double)c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts[0] = f->overflow ? -1 : vid_incl[n - 1];
    }
}

static inline size_t vox_align(size_t x) { return (x + 255) / 256 * 256; }
struct VoxWs {
    VoxFrame* frame;
    unsigned long long *keys_in, *keys_out;
    int *vals_in, *vals_out, *head;
    void* cub_tmp;
    size_t cub_bytes, total;
};
static VoxWs vox_layout(void* base, int n) {
    VoxWs w;
    const size_t m = (size_t)(n > 0 ? n : 1);
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { char* p = b + off; off += vox_align(bytes); return (void*)p; };
    w.frame = (VoxFrame*)take(sizeof(VoxFrame));
    w.keys_in = (unsigned long long*)take(m * 8); w.keys_out = (unsigned long long*)take(m * 8);
    w.vals_in = (int*)take(m * 4); w.vals_out = (int*)take(m * 4);
    w.head = (int*)take(m * 4);
    size_t cs = 0, cb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cs, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int*)nullptr, (int*)nullptr, (int)m, 0, 63);
    cub::DeviceScan::InclusiveSum(nullptr, cb, (int*)nullptr, (int*)nullptr, (int)m);
    w.cub_bytes = cs > cb ? cs : cb;
    w.cub_tmp = take(w.cub_bytes);
    w.total = off;
    return w;
}

extern "C" size_t f4l_voxel_downsample_workspace_bytes(int32_t n) { return vox_layout(nullptr, n < 0 ? 0 : n).total; }

extern "C" int f4l_voxel_downsample(const double* pts64, int32_t n, double voxel_size, double* centroids,
                                    int32_t* voxel_of_point, int32_t* counts, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    F4L_REQUIRE(n >= 0, "n < 0");
    F4L_REQUIRE(counts, "null pointer");
    F4L_REQUIRE(voxel_size > 0.0, "voxel_size must be positive");       // PointCloud.cpp: "voxel_size <= 0" is an error
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(counts, 0, sizeof(int32_t), st);
        return f4l_finish("f4l_voxel_downsample", stream);
    }
    F4L_REQUIRE(pts64 && centroids && workspace, "null pointer");
    const VoxWs w = vox_layout(workspace, n);
    if (workspace_bytes < w.total) {
        f4l_set_error("f4l_voxel_downsample: workspace too small (%zu < %zu)", workspace_bytes, w.total);
        return F4L_E_WORKSPACE;
    }
    const int blocks = min(f4l_div_up(n, 256), 148 * 8);
    f4l_mark("k_vox_init", st);
    k_vox_init<<<1, 32, 0, st>>>(w.frame);
    f4l_mark("k_vox_bbox", st);
    k_vox_bbox<<<blocks, 256, 0, st>>>(pts64, n, w.frame);
    f4l_mark("k_vox_keys", st);
    k_vox_keys<<<blocks, 256, 0, st>>>(pts64, n, voxel_size, w.frame, w.keys_in, w.vals_in);
    size_t cb = w.cub_bytes;
    f4l_count_launches(8); f4l_mark("cub_radix_sort", st);
    cub::DeviceRadixSort::SortPairs(w.cub_tmp, cb, (const unsigned long long*)w.keys_in, w.keys_out, (const int*)w.vals_in,
                                    w.vals_out, n, 0, 63, st);
    f4l_mark("k_vox_heads", st);
    k_vox_heads<<<blocks, 256, 0, st>>>(w.keys_out, n, w.head);
    cb = w.cub_bytes;
    f4l_count_launches(1); f4l_mark("cub_inclusive_scan", st);
    cub::DeviceScan::InclusiveSum(w.cub_tmp, cb, w.head, w.head, n, st);
    f4l_mark("k_vox_mean", st);
    k_vox_mean<<<blocks, 256, 0, st>>>(pts64, n, w.keys_out, w.vals_out, w.head, w.frame, centroids, voxel_of_point, counts);
    return f4l_finish("f4l_voxel_downsample", stream);
}
