// runtime.cu -- error / launch bookkeeping shared by every entry point of the C ABI.
#include "common.cuh"

#include <string.h>

#include <atomic>

namespace f3d {

std::atomic<long long> g_launches{0};  // process-wide: autograd issues the backward launches from its own thread
thread_local char g_err[256] = "";

int fail(int code, const char *what) {
    if (code > 0)
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(code)));
    else
        snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}

int check_launch(const char *what) {
    ++g_launches;
    const cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the sticky launch error so later calls report their own
        return fail(static_cast<int>(e), what);
    }
    return 0;
}

// ---- measurement aid: CUDA-event brackets around named launches, on the stream they are launched on ---------------------------
// Off by default (two predictable branches per launch).  Not usable while the stream is being captured into a CUDA graph:
// callers time eager launches.  `units` is the algorithmic work of the launch as the call site states it (bytes or flops,
// SURVEY.md section 8d), so that achieved = units / duration needs no second source.
namespace {
constexpr int kMaxTimed = 512;
struct TimedLaunch {
    const char *name;
    double units;
    cudaEvent_t a, b;
};
bool g_timer_on = false;
int g_timer_n = 0;
int g_timer_open = -1;
TimedLaunch g_timed[kMaxTimed];
}  // namespace

void ktimer_begin(const char *name, double units, cudaStream_t st) {
    g_timer_open = -1;
    if (!g_timer_on || g_timer_n >= kMaxTimed) return;
    TimedLaunch &r = g_timed[g_timer_n];
    if (!r.a && (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess)) return;
    r.name = name;
    r.units = units;
    if (cudaEventRecord(r.a, st) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    g_timer_open = g_timer_n++;
}
void ktimer_end(cudaStream_t st) {
    if (g_timer_open < 0) return;
    cudaEventRecord(g_timed[g_timer_open].b, st);
    g_timer_open = -1;
}

}  // namespace f3d

// enable != 0: start recording (clears earlier records); 0: stop.
F3D_API void f3d_debug_kernel_timer(int enable) {
    f3d::g_timer_on = enable != 0;
    if (enable) f3d::g_timer_n = 0;
    f3d::g_timer_open = -1;
}
// Waits for the recorded launches and returns their count (<= max): names (max x 48 chars, NUL-terminated), durations (ms) and the
// algorithmic units (bytes or flops) stated by each call site.
F3D_API int f3d_debug_kernel_timings(int max, char *names, float *ms, double *units) {
    int n = f3d::g_timer_n < max ? f3d::g_timer_n : max;
    for (int i = 0; i < n; ++i) {
        f3d::TimedLaunch &r = f3d::g_timed[i];
        float t = -1.0f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) {
            cudaGetLastError();
            t = -1.0f;
        }
        if (names) {
            strncpy(names + i * 48, r.name ? r.name : "", 47);
            names[i * 48 + 47] = 0;
        }
        if (ms) ms[i] = t;
        if (units) units[i] = r.units;
    }
    return n;
}

F3D_API int f3d_version(void) { return 100; }
F3D_API const char *f3d_last_error_string(void) { return f3d::g_err; }
F3D_API long long f3d_launch_count(void) { return f3d::g_launches.load(std::memory_order_relaxed); }
F3D_API void f3d_reset_launch_count(void) { f3d::g_launches.store(0, std::memory_order_relaxed); }
