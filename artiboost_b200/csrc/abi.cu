// Library-wide state of the C ABI: error string, launch counter, version, per-stage CUDA-event profiling.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ab {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- stage profiling: when enabled, every kernel launch of the library is bracketed by two events recorded on
// the launching stream; ab_profile_collect sums the elapsed times per stage.  Off by default (zero cost).
struct Span { cudaEvent_t a, b; int stage; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

bool profile_enabled() { return g_prof_on; }

int profile_begin(int stage, cudaStream_t st) {
    if (!g_prof_on) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    Span s{get_event(), get_event(), stage};
    cudaEventRecord(s.a, st);
    g_spans.push_back(s);
    return (int)g_spans.size() - 1;
}

void profile_end(int token, cudaStream_t st) {
    if (token < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_spans[token].b, st);
}
}  // namespace ab

extern "C" {
int ab_version(void) { return 100; }
const char* ab_last_error(void) { return ab::g_err; }
uint64_t ab_launch_count(void) { return ab::g_launches.load(std::memory_order_relaxed); }

int ab_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(ab::g_prof_mu);
    ab::g_prof_on = on != 0;
    return AB_OK;
}

int ab_profile_collect(double* ms_per_stage, int64_t* launches_per_stage, int n_stages) {
    AB_REQUIRE(ms_per_stage && launches_per_stage && n_stages > 0, "null output");
    std::lock_guard<std::mutex> lk(ab::g_prof_mu);
    for (int i = 0; i < n_stages; ++i) { ms_per_stage[i] = 0.0; launches_per_stage[i] = 0; }
    for (auto& s : ab::g_spans) {
        AB_CUDA(cudaEventSynchronize(s.b));
        float ms = 0.0f;
        AB_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
        if (s.stage >= 0 && s.stage < n_stages) { ms_per_stage[s.stage] += ms; launches_per_stage[s.stage] += 1; }
        ab::g_pool.push_back(s.a);
        ab::g_pool.push_back(s.b);
    }
    ab::g_spans.clear();
    return AB_OK;
}
}
