// Label -> CSR tables (the step right before the path).
//   f4l_labels_to_csr     prepare_pts2spt_dict (base.py:1301-1351): histogram of the patch labels, patches with
//                         count > min_pts kept in ascending label order, the points of a patch in ascending
//                         index order -- a stable radix sort of (label, index) instead of collections.Counter +
//                         one boolean mask per patch
//   f4l_gather_pairs_csr  spt_corres_src / spt_corres_tgt (base.py:3156-3157): the point lists of the matched
//                         patch pairs, concatenated in pair order
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

__global__ void __launch_bounds__(256)
k_csr_keys(const int64_t* __restrict__ labels, int n, unsigned long long* __restrict__ key, int32_t* __restrict__ val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (unsigned long long)labels[i] ^ 0x8000000000000000ull;    // order-preserving for signed labels
    val[i] = i;
}

__global__ void __launch_bounds__(256)
k_csr_heads(const unsigned long long* __restrict__ key, int n, int32_t* __restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0;
}

// run r (all labels, kept or not): start position in the sorted order
__global__ void __launch_bounds__(256)
k_csr_runs(const int32_t* __restrict__ head, const int32_t* __restrict__ hscan, int n, int32_t* __restrict__ run_start,
           int32_t* __restrict__ n_runs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (head[i]) run_start[hscan[i]] = i;
    if (i == n - 1) { n_runs[0] = hscan[i] + head[i]; run_start[hscan[i] + head[i]] = n; }
}

__global__ void __launch_bounds__(256)
k_csr_keep(const int32_t* __restrict__ run_start, const int32_t* __restrict__ n_runs, int min_pts, int n,
           int32_t* __restrict__ keep, int32_t* __restrict__ keep_cnt) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int k = 0, c = 0;
    if (r < n_runs[0]) {
        c = run_start[r + 1] - run_start[r];
        k = c > min_pts ? 1 : 0;                      // base.py:1314-1315: count > num_min_matches_for_small_patch
    }
    keep[r] = k;
    keep_cnt[r] = k ? c : 0;
}

// warp per run: kept runs copy their point indices (already in ascending index order: the sort is stable)
__global__ void __launch_bounds__(256)
k_csr_emit(const unsigned long long* __restrict__ key, const int32_t* __restrict__ order, const int32_t* __restrict__ run_start,
           const int32_t* __restrict__ n_runs, const int32_t* __restrict__ keep, const int32_t* __restrict__ kscan,
           const int32_t* __restrict__ cscan, int n, int64_t* __restrict__ patch_label, int32_t* __restrict__ ptr,
           int32_t* __restrict__ idx, int32_t* __restrict__ patch_of_point, int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nr = n_runs[0];
    if (r >= nr) return;
    const int s0 = run_start[r], s1 = run_start[r + 1];
    if (r == nr - 1 && lane == 0) {
        const int P = kscan[r] + keep[r], items = cscan[r] + (keep[r] ? (s1 - s0) : 0);
        counts[0] = P;
        counts[1] = items;
        ptr[P] = items;
    }
    if (!keep[r]) {
        for (int i = s0 + lane; i < s1; i += 32) patch_of_point[order[i]] = -1;
        return;
    }
    const int p = kscan[r], o0 = cscan[r];
    if (lane == 0) {
        patch_label[p] = (int64_t)(key[s0] ^ 0x8000000000000000ull);
        ptr[p] = o0;
    }
    for (int i = s0 + lane; i < s1; i += 32) {
        const int pt = order[i];
        idx[o0 + (i - s0)] = pt;
        patch_of_point[pt] = p;
    }
}

static inline size_t cal(size_t x) { return (x + 255) / 256 * 256; }

struct CsrWs {
    unsigned long long *key, *key_sorted;
    int32_t *val, *order, *head, *hscan, *run_start, *n_runs, *keep, *keep_cnt, *kscan, *cscan;
    void* cub;
    size_t cub_bytes, total;
};

static CsrWs csr_layout(void* base, int n) {
    CsrWs w;
    char* p = (char*)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* q = p + off; off += cal(bytes); return (void*)q; };
    const size_t N = (size_t)n + 1;
    w.key = (unsigned long long*)take(N * 8);
    w.key_sorted = (unsigned long long*)take(N * 8);
    w.val = (int32_t*)take(N * 4);
    w.order = (int32_t*)take(N * 4);
    w.head = (int32_t*)take(N * 4);
    w.hscan = (int32_t*)take(N * 4);
    w.run_start = (int32_t*)take((N + 1) * 4);
    w.n_runs = (int32_t*)take(256);
    w.keep = (int32_t*)take(N * 4);
    w.keep_cnt = (int32_t*)take(N * 4);
    w.kscan = (int32_t*)take(N * 4);
    w.cscan = (int32_t*)take(N * 4);
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int32_t*)nullptr,
                                    (int32_t*)nullptr, n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, n);
    w.cub_bytes = a > b ? a : b;
    w.cub = take(w.cub_bytes);
    w.total = off;
    return w;
}

extern "C" size_t f4l_labels_to_csr_workspace_bytes(int32_t n) { return n < 0 ? 0 : csr_layout(nullptr, n).total + 256; }

extern "C" int f4l_labels_to_csr(const int64_t* labels, int32_t n, int32_t min_pts, int64_t* patch_label, int32_t* ptr,
                                 int32_t* idx, int32_t* patch_of_point, int32_t* counts, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    F4L_REQUIRE(n >= 0, "negative size");
    F4L_REQUIRE(counts && ptr, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st);
        cudaMemsetAsync(ptr, 0, sizeof(int32_t), st);
        return F4L_OK;
    }
    F4L_REQUIRE(labels && patch_label && idx && patch_of_point, "null pointer");
    const size_t need = f4l_labels_to_csr_workspace_bytes(n);
    if (!workspace || workspace_bytes < need) {
        f4l_set_error("f4l_labels_to_csr: workspace too small (%zu < %zu)", workspace_bytes, need);
        return F4L_E_WORKSPACE;
    }
    CsrWs w = csr_layout((void*)(((uintptr_t)workspace + 255) / 256 * 256), n);
    const int g = f4l_div_up(n, 256);
    f4l_mark("k_csr_keys", st);
    k_csr_keys<<<g, 256, 0, st>>>(labels, n, w.key, w.val);
    f4l_mark("cub_radix_sort_labels", st);
    size_t cb = w.cub_bytes;
    cub::DeviceRadixSort::SortPairs(w.cub, cb, w.key, w.key_sorted, w.val, w.order, n, 0, 64, st);
    f4l_mark("k_csr_tables", st);
    k_csr_heads<<<g, 256, 0, st>>>(w.key_sorted, n, w.head);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub, cb, w.head, w.hscan, n, st);
    k_csr_runs<<<g, 256, 0, st>>>(w.head, w.hscan, n, w.run_start, w.n_runs);
    k_csr_keep<<<g, 256, 0, st>>>(w.run_start, w.n_runs, min_pts, n, w.keep, w.keep_cnt);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub, cb, w.keep, w.kscan, n, st);
    cb = w.cub_bytes;
    cub::DeviceScan::ExclusiveSum(w.cub, cb, w.keep_cnt, w.cscan, n, st);
    k_csr_emit<<<f4l_div_up(n, 8), 256, 0, st>>>(w.key_sorted, w.order, w.run_start, w.n_runs, w.keep, w.kscan, w.cscan, n,
                                                patch_label, ptr, idx, patch_of_point, counts);
    f4l_count_launches(6);
    return f4l_finish("f4l_labels_to_csr", stream);
}

// ---- pairs -> concatenated point lists ---------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pair_counts(const int32_t* __restrict__ ptr, const int32_t* __restrict__ sel, int Q, int32_t* __restrict__ cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) cnt[q] = ptr[sel[q] + 1] - ptr[sel[q]];
    if (q == Q) cnt[q] = 0;
}
__global__ void __launch_bounds__(256)
k_pair_gather(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ sel, int Q,
              const int32_t* __restrict__ out_ptr, int32_t* __restrict__ out_idx) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const int s0 = ptr[sel[q]], n = ptr[sel[q] + 1] - s0, o0 = out_ptr[q];
    for (int i = lane; i < n; i += 32) out_idx[o0 + i] = idx[s0 + i];
}

extern "C" size_t f4l_gather_pairs_csr_workspace_bytes(int32_t Q) {
    if (Q < 0) return 0;
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int32_t*)nullptr, (int32_t*)nullptr, Q + 1);
    return cal((size_t)(Q + 1) * 4) + cal(b) + 256;
}

// out_ptr (Q+1) int32; out_idx: capacity >= sum of the selected patch sizes (<= Q * largest patch; the caller
// reads out_ptr[Q] back or passes an upper bound)
extern "C" int f4l_gather_pairs_csr(const int32_t* ptr, const int32_t* idx, const int32_t* sel, int32_t Q, int32_t* out_ptr,
                                    int32_t* out_idx, int32_t out_capacity, void* workspace, size_t workspace_bytes,
                                    void* stream) {
    F4L_REQUIRE(Q >= 0, "negative size");
    F4L_REQUIRE(out_ptr, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    if (Q == 0) {
        cudaMemsetAsync(out_ptr, 0, sizeof(int32_t), st);
        return F4L_OK;
    }
    F4L_REQUIRE(ptr && idx && sel && workspace, "null pointer");
    if (workspace_bytes < f4l_gather_pairs_csr_workspace_bytes(Q)) {
        f4l_set_error("f4l_gather_pairs_csr: workspace too small");
        return F4L_E_WORKSPACE;
    }
    char* base = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
    int32_t* cnt = (int32_t*)base;
    void* cubtmp = base + cal((size_t)(Q + 1) * 4);
    size_t cb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cb, cnt, out_ptr, Q + 1);
    f4l_mark("k_pair_gather", st);
    k_pair_counts<<<f4l_div_up(Q + 1, 256), 256, 0, st>>>(ptr, sel, Q, cnt);
    cub::DeviceScan::ExclusiveSum(cubtmp, cb, cnt, out_ptr, Q + 1, st);
    if (out_idx && out_capacity > 0) k_pair_gather<<<f4l_div_up(Q, 8), 256, 0, st>>>(ptr, idx, sel, Q, out_ptr, out_idx);
    f4l_count_launches(2);
    return f4l_finish("f4l_gather_pairs_csr", stream);
}
