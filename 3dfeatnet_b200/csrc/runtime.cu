// runtime.cu -- error / launch bookkeeping shared by every entry point of the C ABI.
#include "common.cuh"

#include <string.h>

namespace f3d {

thread_local long long g_launches = 0;
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

}  // namespace f3d

F3D_API int f3d_version(void) { return 100; }
F3D_API const char *f3d_last_error_string(void) { return f3d::g_err; }
F3D_API long long f3d_launch_count(void) { return f3d::g_launches; }
F3D_API void f3d_reset_launch_count(void) { f3d::g_launches = 0; }
