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
