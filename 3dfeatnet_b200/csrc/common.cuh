// common.cuh -- shared helpers of the sm_100a kernels (launch bookkeeping, error plumbing, warp utils).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/feat3dnet_b200.h"

#define F3D_API extern "C" __attribute__((visibility("default")))

namespace f3d {

// bookkeeping exported through f3d_launch_count() (process-wide) / f3d_last_error_string() (per thread)
extern std::atomic<long long> g_launches;
extern thread_local char g_err[256];

int fail(int code, const char *what);
int check_launch(const char *what);  // cudaPeekAtLastError -> code (0 ok), bumps the launch counter

// measurement aid (runtime.cu): CUDA-event bracket of one named launch on its own stream, recorded only while
// f3d_debug_kernel_timer(1) is in effect; `units` = the algorithmic bytes / flops of the launch
void ktimer_begin(const char *name, double units, cudaStream_t st);
void ktimer_end(cudaStream_t st);

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Squared distance with the association nvcc emits for the reference kernels on sm_100a
// (tf_sampling_g.cu:144, tf_grouping_g.cu:27):  FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Smallest float T with sqrt_rn(T) >= r  (r > 0).  Because IEEE sqrt is monotone,
//   fmaxf(sqrtf(s), 1e-20f) < r   <=>   s < T          for every s >= 0 when r > 1e-20f,
// which lets the scan compare squared distances and skip the square root (tf_grouping_g.cu:27-28).
__device__ __forceinline__ float ball_threshold(float r) {
    if (!(r > 1e-20f)) return 0.0f;           // no distance can be < r: the clamp keeps d >= 1e-20f
    if (r == __int_as_float(0x7f800000)) return r;  // every finite distance hits
    float t = __fmul_rn(r, r);
    if (t == __int_as_float(0x7f800000)) {
        t = __int_as_float(0x7f7fffff);
        if (__fsqrt_rn(t) < r) return __int_as_float(0x7f800000);
    }
    // walk down while the predecessor still satisfies sqrt >= r
    for (int it = 0; it < 8; ++it) {
        if (!(t > 0.0f)) break;
        const float p = __int_as_float(__float_as_int(t) - 1);
        if (__fsqrt_rn(p) >= r) t = p; else break;
    }
    // walk up while sqrt(t) < r
    for (int it = 0; it < 8; ++it) {
        if (__fsqrt_rn(t) < r) t = __int_as_float(__float_as_int(t) + 1); else break;
    }
    return t;
}

}  // namespace f3d
