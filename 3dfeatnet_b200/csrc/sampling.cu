// sampling.cu -- farthest point sampling + gather_point for sm_100a.
//
// Replaces tf_ops/sampling/tf_sampling_g.cu:105-181 (farthestpointsamplingKernel, gatherpointKernel).
//
// FPS is a serial chain of m-1 block-wide arg-max rounds; the only lever is the latency of one round.
// The reference keeps the running distances in GLOBAL memory (temp) and spends 10 __syncthreads per
// round on a 9-level shared-memory tree.  Here, per cloud (one 1024-thread CTA):
//   * coordinates are staged once into shared memory as SoA (12 B/point, <= 196 KB at n = 16384),
//   * the running distances live in REGISTERS (PPT per thread) for the whole kernel,
//   * the arg-max is two redux.sync per warp + one 32-entry shared-memory hop: ONE __syncthreads per round,
//   * the winner's coordinates travel with the arg-max, so no dependent global load per round.
// The reference's tie rule is reproduced exactly: its thread t owns k = t (mod 512) and keeps its first
// strict maximum, the tree keeps the lower thread => the winner is the maximum distance with the lowest
// (k mod 512, k).  That pair is packed into one 32-bit tie key so it rides through redux.sync.min.
#include "common.cuh"

namespace f3d {

constexpr int kFpsThreads = 1024;

__device__ __forceinline__ unsigned fps_tie_key(int k) {  // orders by (k mod 512, k); bijective for k < 2^32
    return (static_cast<unsigned>(k & 511) << 23) | (static_cast<unsigned>(k) >> 9);
}
__device__ __forceinline__ int fps_tie_key_inv(unsigned t) { return static_cast<int>(((t & 0x7fffffu) << 9) | (t >> 23)); }

struct FpsSlots {  // double-buffered per-warp winners
    int d[2][32];
    unsigned key[2][32];
    float x[2][32], y[2][32], z[2][32];
};

// Block-wide arg-max of (best, besti) under the reference tie rule.  Returns the winner index and its
// coordinates to every thread.  One __syncthreads.  `par` alternates 0/1 between consecutive calls.
template <class CoordFn>
__device__ __forceinline__ int fps_block_argmax(FpsSlots &S, int par, float best, int besti, CoordFn coords,
                                                float &ox, float &oy, float &oz) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // distances are >= +0 (or the -1 sentinel), so their bit patterns order like signed ints
    const int bi = __float_as_int(best);
    const unsigned tk = fps_tie_key(besti);
    const int wmax = __reduce_max_sync(kFull, bi);
    const unsigned wmin = __reduce_min_sync(kFull, bi == wmax ? tk : 0xffffffffu);
    const unsigned who = __ballot_sync(kFull, bi == wmax && tk == wmin);
    if (lane == __ffs(who) - 1) {  // only the warp's winner fetches its coordinates
        float bx, by, bz;
        coords(besti, bx, by, bz);
        S.d[par][warp] = wmax;
        S.key[par][warp] = wmin;
        S.x[par][warp] = bx;
        S.y[par][warp] = by;
        S.z[par][warp] = bz;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    const int d2 = lane < nw ? S.d[par][lane] : static_cast<int>(0x80000000);
    const unsigned k2 = lane < nw ? S.key[par][lane] : 0xffffffffu;
    const int bmax = __reduce_max_sync(kFull, d2);
    const unsigned bmin = __reduce_min_sync(kFull, d2 == bmax ? k2 : 0xffffffffu);
    const int src = __ffs(__ballot_sync(kFull, d2 == bmax && k2 == bmin)) - 1;
    ox = S.x[par][src];
    oy = S.y[par][src];
    oz = S.z[par][src];
    return fps_tie_key_inv(bmin);
}

// ---- main path: n <= 1024*PPT, coordinates in shared memory, distances in registers ----------------------
// Thread t owns the points k = 4*(t + 1024*g) + q, g < PPT/4, q < 4 (float4 shared-memory loads).  Within a
// thread, k mod 512 grows with q and k grows with g, so walking (q major, g minor) with a strict '>' keeps
// the thread's lowest (k mod 512, k) among equal maxima.
template <int PPT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_smem_kernel(int n, int m, const float *__restrict__ inp, int *__restrict__ out) {
    constexpr int G = PPT / 4;
    constexpr int NP = kFpsThreads * PPT;
    extern __shared__ float4 fps_smem[];
    float *xs = reinterpret_cast<float *>(fps_smem);
    float *ys = xs + NP;
    float *zs = ys + NP;
    __shared__ FpsSlots slots;

    const int tid = threadIdx.x;
    const float *p = inp + static_cast<size_t>(blockIdx.x) * n * 3;
    int *o = out + static_cast<size_t>(blockIdx.x) * m;

    for (int i = tid; i < n * 3; i += kFpsThreads) {  // coalesced AoS read -> SoA
        const float v = __ldg(p + i);
        const int k = i / 3, c = i - 3 * k;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    for (int k = n + tid; k < NP; k += kFpsThreads) xs[k] = ys[k] = zs[k] = 0.0f;
    __syncthreads();

    float td[PPT];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int q = 0; q < 4; ++q) td[g * 4 + q] = (4 * (tid + kFpsThreads * g) + q) < n ? 1e38f : -1.0f;

    float ox = xs[0], oy = ys[0], oz = zs[0];
    if (tid == 0) o[0] = 0;

    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const float4 *ys4 = reinterpret_cast<const float4 *>(ys);
    const float4 *zs4 = reinterpret_cast<const float4 *>(zs);

    for (int j = 1; j < m; ++j) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const float4 X = xs4[tid + kFpsThreads * g];
            const float4 Y = ys4[tid + kFpsThreads * g];
            const float4 Z = zs4[tid + kFpsThreads * g];
            td[g * 4 + 0] = fminf(sqdist_ref(X.x - ox, Y.x - oy, Z.x - oz), td[g * 4 + 0]);
            td[g * 4 + 1] = fminf(sqdist_ref(X.y - ox, Y.y - oy, Z.y - oz), td[g * 4 + 1]);
            td[g * 4 + 2] = fminf(sqdist_ref(X.z - ox, Y.z - oy, Z.z - oz), td[g * 4 + 2]);
            td[g * 4 + 3] = fminf(sqdist_ref(X.w - ox, Y.w - oy, Z.w - oz), td[g * 4 + 3]);
        }
        float best = -1.0f;
        int besti = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float v = td[g * 4 + q];
                if (v > best) {
                    best = v;
                    besti = 4 * (tid + kFpsThreads * g) + q;
                }
            }
        const int old = fps_block_argmax(
            slots, j & 1, best, besti,
            [&](int k, float &x, float &y, float &z) { x = xs[k]; y = ys[k]; z = zs[k]; }, ox, oy, oz);
        if (tid == 0) o[j] = old;
    }
}

// ---- fallback for any n: running distances in a caller-provided (b,n) scratch, points from L2 ------------
// Thread t owns k = t, t+1024, ...: k mod 512 is constant per thread and k ascends, so the first strict
// maximum is again the thread's lowest (k mod 512, k).
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_global_kernel(int n, int m, const float *__restrict__ inp, float *__restrict__ temp, int *__restrict__ out) {
    __shared__ FpsSlots slots;
    const int tid = threadIdx.x;
    const float *p = inp + static_cast<size_t>(blockIdx.x) * n * 3;
    float *td = temp + static_cast<size_t>(blockIdx.x) * n;
    int *o = out + static_cast<size_t>(blockIdx.x) * m;
    for (int k = tid; k < n; k += kFpsThreads) td[k] = 1e38f;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (n > 0) {
        ox = __ldg(p + 0);
        oy = __ldg(p + 1);
        oz = __ldg(p + 2);
    }
    if (tid == 0) o[0] = 0;
    for (int j = 1; j < m; ++j) {
        float best = -1.0f, bx = 0.f, by = 0.f, bz = 0.f;
        int besti = 0;
        for (int k = tid; k < n; k += kFpsThreads) {
            const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
            const float d2 = fminf(sqdist_ref(x - ox, y - oy, z - oz), td[k]);
            td[k] = d2;
            if (d2 > best) {
                best = d2;
                besti = k;
                bx = x;
                by = y;
                bz = z;
            }
        }
        const int old = fps_block_argmax(
            slots, j & 1, best, besti, [&](int, float &x, float &y, float &z) { x = bx; y = by; z = bz; }, ox, oy, oz);
        if (tid == 0) o[j] = old;
    }
}

// gather_point: out[b,j,:] = inp[b,idx[b,j],:]  (tf_sampling_g.cu:172-181).  One thread per output float.
__global__ void gather_point_kernel(int n, int m, long long total, const float *__restrict__ inp,
                                    const int *__restrict__ idx, float *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long r = i / 3;
    const int c = static_cast<int>(i - r * 3);
    const long long bb = r / m;
    const int a = __ldg(idx + r);
    out[i] = __ldg(inp + (bb * n + a) * 3 + c);
}

template <int PPT>
static int launch_fps_smem(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(kFpsThreads) * PPT * 3 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "fps: cudaFuncSetAttribute");
    fps_smem_kernel<PPT><<<b, kFpsThreads, smem, st>>>(n, m, inp, out);
    return check_launch("fps_smem_kernel");
}

}  // namespace f3d

using namespace f3d;

F3D_API int f3d_farthest_point_sample(int b, int n, int m, const float *inp, float *temp, int *out, void *stream) {
    if (b < 0 || n <= 0 || m <= 0 || !inp || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample: bad arguments");
    if (b == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (n <= 1024 * 4) return launch_fps_smem<4>(b, n, m, inp, out, st);
    if (n <= 1024 * 8) return launch_fps_smem<8>(b, n, m, inp, out, st);
    if (n <= 1024 * 16) return launch_fps_smem<16>(b, n, m, inp, out, st);
    if (!temp) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "farthest_point_sample: n > 16384 needs temp of b*n floats");
    fps_global_kernel<<<b, kFpsThreads, 0, st>>>(n, m, inp, temp, out);
    return check_launch("fps_global_kernel");
}

F3D_API int f3d_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || !inp || !idx || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "gather_point: bad arguments");
    const long long total = 3LL * b * m;
    if (total == 0) return 0;
    gather_point_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(n, m, total, inp, idx, out);
    return check_launch("gather_point_kernel");
}
