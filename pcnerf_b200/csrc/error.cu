// Library runtime: thread-local error channel, kernel-launch counter and the optional per-kernel-class CUDA-event
// profiler used by bench.py for the roofline figures (include/pcnerf_b200.h, "instrumentation").
#include "common.cuh"
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>

static thread_local char g_err[512] = "";

void pcn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* pcnerf_last_error(void) { return g_err; }
extern "C" int pcnerf_version(void) { return 104; }     // 104: pcnerf_affine_eval_alpha, pcnerf_affine_apply_rays

// ---------------------------------------------------------------------------------------------------------------
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfRec { cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs[PCN_K_COUNT];
static double g_work[PCN_K_COUNT];
static long long g_count[PCN_K_COUNT];

static const char* k_names[PCN_K_COUNT] = {"mlp_gemm_fwd", "mlp_gemm_dgrad", "mlp_gemm_wgrad", "mlp_small",
                                           "sample_encode", "composite_fwd", "composite_bwd", "aabb", "search", "affine",
                                           "affine_moments", "affine_algebra"};

PcnScope::PcnScope(int id_, cudaStream_t st_, double work, int nlaunch) : id(id_), st(st_), on(false) {
    g_launches.fetch_add(nlaunch, std::memory_order_relaxed);
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    on = true;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_work[id] += work;
    g_count[id] += nlaunch;
}

PcnScope::~PcnScope() {
    if (!on) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_recs[id].push_back({a, b});
}

extern "C" long long pcnerf_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

extern "C" void pcnerf_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < PCN_K_COUNT; ++i) {
        for (auto& r : g_recs[i]) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
        g_recs[i].clear();
        g_work[i] = 0;
        g_count[i] = 0;
    }
    g_prof_on.store(on ? 1 : 0);
}

extern "C" int pcnerf_prof_classes(void) { return PCN_K_COUNT; }
extern "C" const char* pcnerf_prof_name(int id) { return (id >= 0 && id < PCN_K_COUNT) ? k_names[id] : ""; }

extern "C" int pcnerf_prof_read(int id, double* out_ms, long long* out_launches, double* out_work) {
    PCN_CHECK_ARG(id >= 0 && id < PCN_K_COUNT, "prof_read: bad class id %d", id);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double ms = 0;
    for (auto& r : g_recs[id]) {
        PCN_CUDA(cudaEventSynchronize(r.b));
        float t = 0;
        PCN_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        ms += t;
    }
    if (out_ms) *out_ms = ms;
    if (out_launches) *out_launches = g_count[id];
    if (out_work) *out_work = g_work[id];
    return 0;
}
