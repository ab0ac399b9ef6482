// Error plumbing and version of libf4l_b200.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

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

static std::atomic<long long> g_launches{0};
void f4l_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long f4l_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void f4l_launch_count_reset(void) { g_launches.store(0, std::memory_order_relaxed); }

// ---- optional in-stream kernel timing (bench.py roofline leg) --------------------------------
#include <map>
#include <mutex>
#include <string>
#include <vector>

struct ProfMark { const char* name; cudaEvent_t ev; cudaStream_t st; };
static bool g_prof_on = false;
static std::vector<ProfMark> g_marks;
static std::vector<cudaEvent_t> g_pool;
static std::map<std::string, std::pair<double, long long>> g_prof;   // name -> (total ms, launches)
static std::mutex g_prof_mu;

void f4l_mark(const char* name, cudaStream_t st) {
    if (name && name[0] != '#') g_launches.fetch_add(1, std::memory_order_relaxed);   // '#': not one of our kernels
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t ev;
    if (!g_pool.empty()) { ev = g_pool.back(); g_pool.pop_back(); }
    else cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    g_marks.push_back({name, ev, st});
}

int f4l_finish(const char* what, void* stream) {
    f4l_mark(nullptr, (cudaStream_t)stream);
    return f4l_check_launch(what);
}

extern "C" void f4l_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
}

// Synchronises the recorded events and folds them into the per-kernel table.
extern "C" int f4l_profile_collect(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (size_t i = 0; i + 1 < g_marks.size(); ++i) {
        const ProfMark& a = g_marks[i];
        const ProfMark& b = g_marks[i + 1];
        if (!a.name || a.st != b.st) continue;
        cudaEventSynchronize(b.ev);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, a.ev, b.ev) != cudaSuccess) continue;
        auto& e = g_prof[a.name];
        e.first += ms;
        e.second += 1;
    }
    for (auto& m : g_marks) g_pool.push_back(m.ev);
    g_marks.clear();
    cudaGetLastError();
    return (int)g_prof.size();
}

extern "C" int f4l_profile_get(int index, char* name, int name_cap, double* total_ms, long long* launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (index < 0 || index >= (int)g_prof.size()) return F4L_E_ARG;
    auto it = g_prof.begin();
    std::advance(it, index);
    snprintf(name, name_cap, "%s", it->first.c_str());
    *total_ms = it->second.first;
    *launches = it->second.second;
    return F4L_OK;
}

extern "C" void f4l_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.clear();
    for (auto& m : g_marks) g_pool.push_back(m.ev);
    g_marks.clear();
}

// ---- host-side helper of the host-buffer API --------------------------------------------------
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

extern "C" void f4l_host_pack_corr_targets(const int64_t* h_corr, int64_t n, int32_t* h_out, int32_t n_threads) {
    if (!h_corr || !h_out || n <= 0) return;
    int nt = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    if (n < 65536) nt = 1;
    auto work = [&](int t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t v = h_corr[2 * i + 1];
            h_out[i] = (v >= 0 && v <= 0x7fffffffLL) ? (int32_t)v : -1;
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
}

extern "C" long long f4l_host_expand_sparse(const float* h_once, const int32_t* h_pair_rows, int32_t Q, float* h_out,
                                            int32_t n_threads) {
    if (!h_once || !h_pair_rows || !h_out || Q <= 0) return 0;
    std::vector<long long> off((size_t)Q + 1, 0);
    for (int q = 0; q < Q; ++q) off[q + 1] = off[q] + (h_pair_rows[q] > 0 ? h_pair_rows[q] : 0);
    const long long total = off[Q];
    if (total == 0) return 0;
    int nt = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    if (total < 4096) nt = 1;
    auto work = [&](int t) {
        // pairs are dealt in contiguous blocks of roughly equal row counts
        const long long lo = total * t / nt, hi = total * (t + 1) / nt;
        int q = (int)(std::upper_bound(off.begin(), off.end(), lo) - off.begin()) - 1;
        for (; q < Q && off[q] < hi; ++q) {
            if (off[q] < lo) continue;                  // the pair belongs to the thread that owns its first row
            const size_t n = (size_t)(off[q + 1] - off[q]) * 6 * sizeof(float);
            const float* src = h_once + (size_t)off[q] * 6;
            float* dst = h_out + (size_t)off[q] * 12;
            std::memcpy(dst, src, n);
            std::memcpy(dst + (size_t)(off[q + 1] - off[q]) * 6, src, n);
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    return 2 * total;
}
