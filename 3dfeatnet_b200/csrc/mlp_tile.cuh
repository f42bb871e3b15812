// mlp_tile.cuh -- the fp32 register-tile routine shared by the eval-mode forward (mlp_fp32.cu) and the training
// layers (train_layers.cu).
//
// 256 threads = 16 row groups (8 rows) x 16 column groups (CT columns); activations are kept TRANSPOSED in shared
// memory ([channel][row]) so a thread's 8 rows are two LDS.128 and the 16 row groups of a warp read 512 contiguous
// bytes; weights stream through the read-only path (two 32-byte segments per warp).
#pragma once
#include "common.cuh"

namespace f3d {

constexpr int kTileRows = 128;
constexpr int kMlpThreads = 256;

template <int CT>
__device__ __forceinline__ void load_w(const float *__restrict__ p, float (&w)[CT]) {
    if constexpr (CT == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else if constexpr (CT == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else if constexpr (CT == 2) {
        const float2 a = __ldg(reinterpret_cast<const float2 *>(p));
        w[0] = a.x; w[1] = a.y;
    } else {
        w[0] = __ldg(p);
    }
}

// acc[i][j] = sum_k in_t[k][rg*8+i] * W[k][col0 + cg*CT + j]   (k ascending, fp32 FMA)
template <int CT>
__device__ __forceinline__ void dense_tile(const float *in_t, int cin, const float *__restrict__ W, int ldw, int col0,
                                           float (&acc)[8][CT], int ld = kTileRows) {
    const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CT; ++j) acc[i][j] = 0.0f;
    const float *ap = in_t + rg * 8;
    const float *wp = W + col0 + cg * CT;
#pragma unroll 4
    for (int k = 0; k < cin; ++k) {
        const float4 a0 = *reinterpret_cast<const float4 *>(ap + k * ld);
        const float4 a1 = *reinterpret_cast<const float4 *>(ap + k * ld + 4);
        float w[CT];
        load_w<CT>(wp + static_cast<size_t>(k) * ldw, w);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
}

}  // namespace f3d
