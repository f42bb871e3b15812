// mlp_fp32.cu -- fused detector / descriptor forward, exact fp32 (CUDA-core FFMA) path, eval-mode BN folded.
//
// Replaces the TensorFlow graph of models/feat3dnet.py:90-187 + models/pointnet_common.py:32-135
// (group_point, tile, subtract, divide, [rotate], 3-4 x (1x1 conv + BN + ReLU), reduce_max, ... ) which the
// reference runs as one cuDNN/Eigen kernel per op with every (B,M,64,C) intermediate in HBM.  Here a CTA
// owns a tile of 128 grouped rows (two clusters at nsample = 64); the gather, the normalisation, every
// per-row layer and the max-pool happen in shared memory / registers, and only the pooled (B*M, C) vectors
// reach HBM.  The small per-cluster layers run in a second kernel over 128 clusters per CTA with the same
// tile routine.  This is the fp32 reference path (precision 0); the tensor-core path lives in mlp_tc.cu.
//
// Tile routine: 256 threads = 16 row groups (8 rows) x 16 column groups (CT columns); activations are kept
// TRANSPOSED in shared memory ([channel][row]) so a thread's 8 rows are two LDS.128 and the 16 row groups of
// a warp read 512 contiguous bytes; weights stream through the read-only path (two 32-byte segments per warp).
#include "common.cuh"
#include "mlp_tile.cuh"
#include "weights_layout.h"

namespace f3d {

// v = acc + bias (+ReLU); store transposed into out_t[col0 + cg*CT + j][row]
template <int CT, bool RELU>
__device__ __forceinline__ void bias_act_store(float (&acc)[8][CT], const float *__restrict__ bias, int col0, float *out_t) {
    const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
#pragma unroll
    for (int j = 0; j < CT; ++j) {
        const int col = col0 + cg * CT + j;
        const float bb = __ldg(bias + col);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = acc[i][j] + bb;
            if (RELU) v[i] = fmaxf(v[i], 0.0f);
            acc[i][j] = v[i];
        }
        float *o = out_t + col * kTileRows + rg * 8;
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// max over the rows of each cluster of acc (already biased/activated).  part: smem [16][16*CT].
// pooled_out(cl, col) receives the result for cl < cpt (clusters in this tile), S = rows per cluster (multiple of 8).
template <int CT, class Sink>
__device__ __forceinline__ void pool_rows(float (&acc)[8][CT], float *part, int S, int cpt, Sink sink) {
    const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
    constexpr int NC = 16 * CT;
    __syncthreads();  // part may still be read by a previous pool
#pragma unroll
    for (int j = 0; j < CT; ++j) {
        float mx = acc[0][j];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, acc[i][j]);
        part[rg * NC + cg * CT + j] = mx;
    }
    __syncthreads();
    const int gpc = S / 8;  // row groups per cluster
    for (int e = threadIdx.x; e < cpt * NC; e += kMlpThreads) {
        const int cl = e / NC, col = e - cl * NC;
        float mx = part[(cl * gpc) * NC + col];
        for (int g = 1; g < gpc; ++g) mx = fmaxf(mx, part[(cl * gpc + g) * NC + col]);
        sink(cl, col, mx);
    }
}

// rows of the tile <- (xyz[idx] - centre) / radius, optionally rotated about z by the cluster's orientation
// (pointnet_common.py:104-120).  in0_t: smem [3][128].
__device__ __forceinline__ void gather_rows(float *in0_t, long long tile, int S, long long num_clusters, int n, int m,
                                            float radius, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                                            const int *__restrict__ idx, const float *__restrict__ orientation) {
    const int r = threadIdx.x;
    if (r >= kTileRows) return;
    const int cpt = kTileRows / S;
    const long long cl = tile * cpt + r / S;
    float x = 0.f, y = 0.f, z = 0.f;
    if (cl < num_clusters) {
        const int s = r % S;
        int ii = __ldg(idx + cl * S + s);
        ii = min(max(ii, 0), n - 1);
        const long long bb = cl / m;
        const float *p = xyz + (bb * n + ii) * 3;
        const float *c = new_xyz + cl * 3;
        x = (__ldg(p) - __ldg(c)) / radius;
        y = (__ldg(p + 1) - __ldg(c + 1)) / radius;
        z = (__ldg(p + 2) - __ldg(c + 2)) / radius;
        if (orientation) {
            const float th = __ldg(orientation + cl);
            const float cs = cosf(th), sn = sinf(th);
            const float xr = x * cs - y * sn;
            const float yr = x * sn + y * cs;
            x = xr;
            y = yr;
        }
    }
    in0_t[r] = x;
    in0_t[kTileRows + r] = y;
    in0_t[2 * kTileRows + r] = z;
}

// ---------------------------------------------------------------------------------------------------
// detector, per-row part: 3 -> 64 -> 128 -> 256 (BN folded, ReLU) and max over the cluster's rows.
// pooled: (num_clusters, 256)
__global__ void __launch_bounds__(kMlpThreads, 2)
det_rows_fp32_kernel(long long num_clusters, int n, int m, int S, float radius, const float *__restrict__ xyz,
                     const float *__restrict__ new_xyz, const int *__restrict__ idx, const float *__restrict__ P,
                     WeightLayout L, float *__restrict__ pooled) {
    extern __shared__ float4 smem4[];
    float *in0_t = reinterpret_cast<float *>(smem4);  // [3][128]   (padded to 4 rows)
    float *a1_t = in0_t + 4 * kTileRows;              // [64][128]
    float *a2_t = a1_t + 64 * kTileRows;              // [128][128]
    float *part = a2_t + 128 * kTileRows;             // [16][128]
    const int cpt = kTileRows / S;
    const long long tile = blockIdx.x;

    gather_rows(in0_t, tile, S, num_clusters, n, m, radius, xyz, new_xyz, idx, nullptr);
    __syncthreads();
    {
        float acc[8][4];
        dense_tile<4>(in0_t, 3, P + L.off[W_DET0], 64, 0, acc);
        bias_act_store<4, true>(acc, P + L.off[B_DET0], 0, a1_t);
    }
    __syncthreads();
    {
        float acc[8][8];
        dense_tile<8>(a1_t, 64, P + L.off[W_DET1], 128, 0, acc);
        bias_act_store<8, true>(acc, P + L.off[B_DET1], 0, a2_t);
    }
    __syncthreads();
    for (int chunk = 0; chunk < 2; ++chunk) {
        float acc[8][8];
        dense_tile<8>(a2_t, 128, P + L.off[W_DET2], 256, chunk * 128, acc);
        const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
        (void)rg;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float bb = __ldg(P + L.off[B_DET2] + chunk * 128 + cg * 8 + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][j] = fmaxf(acc[i][j] + bb, 0.0f);
        }
        pool_rows<8>(acc, part, S, cpt, [&](int cl, int col, float v) {
            const long long c = tile * cpt + cl;
            if (c < num_clusters) pooled[c * 256 + chunk * 128 + col] = v;
        });
    }
}

// detector, per-cluster part: 256 -> 128 -> 64 (ReLU), attention = softplus(64->1), orientation =
// atan2 of the l2-normalised 64->2 output (feat3dnet.py:134-149).  One CTA = 128 clusters.
__global__ void __launch_bounds__(kMlpThreads, 1)
det_post_fp32_kernel(long long num_clusters, const float *__restrict__ pooled, const float *__restrict__ P,
                     WeightLayout L, float *__restrict__ attention, float *__restrict__ orientation) {
    extern __shared__ float4 smem4[];
    // the transposed input tile has a leading dimension of 132 floats: the coalesced-read / transposed-store loop below
    // then spreads a warp's 32 stores over 8 banks instead of 1, and float4 reads stay 16-byte aligned
    constexpr int kLdIn = kTileRows + 4;
    float *in_t = reinterpret_cast<float *>(smem4);  // [256][132]
    float *a1_t = in_t + 256 * kLdIn;                // [128][128]
    float *a2_t = in_t;                              // [64][128], reuses in_t after conv_post_0
    const long long c0 = static_cast<long long>(blockIdx.x) * kTileRows;
    for (int e = threadIdx.x; e < kTileRows * 256; e += kMlpThreads) {  // coalesced read, transposed store
        const int r = e >> 8, k = e & 255;
        in_t[k * kLdIn + r] = (c0 + r < num_clusters) ? __ldg(pooled + (c0 + r) * 256 + k) : 0.0f;
    }
    __syncthreads();
    {
        float acc[8][8];
        dense_tile<8>(in_t, 256, P + L.off[W_DETP0], 128, 0, acc, kLdIn);
        bias_act_store<8, true>(acc, P + L.off[B_DETP0], 0, a1_t);
    }
    __syncthreads();
    {
        float acc[8][4];
        dense_tile<4>(a1_t, 128, P + L.off[W_DETP1], 64, 0, acc);
        bias_act_store<4, true>(acc, P + L.off[B_DETP1], 0, a2_t);
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < kTileRows && c0 + r < num_clusters) {
        float att = __ldg(P + L.off[B_ATT]);
        float ox = __ldg(P + L.off[B_ORI]), oy = __ldg(P + L.off[B_ORI] + 1);
        for (int k = 0; k < 64; ++k) {
            const float a = a2_t[k * kTileRows + r];
            att = fmaf(a, __ldg(P + L.off[W_ATT] + k), att);
            ox = fmaf(a, __ldg(P + L.off[W_ORI] + 2 * k), ox);
            oy = fmaf(a, __ldg(P + L.off[W_ORI] + 2 * k + 1), oy);
        }
        attention[c0 + r] = att > 20.0f ? att : log1pf(expf(att));  // softplus
        const float inv = 1.0f / sqrtf(fmaxf(ox * ox + oy * oy, 1e-8f));  // tf.nn.l2_normalize(eps=1e-8)
        orientation[c0 + r] = atan2f(oy * inv, ox * inv);
    }
}

// ---------------------------------------------------------------------------------------------------
// descriptor, per-row part (feat3dnet.py:54-75): rotate, 3 -> 32 -> 64 (ReLU), max-pool, then
// conv_mid_0 on [h | tile(pool)] evaluated as h*W[:64] + (pool*W[64:] + b) (the second term once per
// cluster), no ReLU, max-pool.  pooled2: (num_clusters, MID)
__global__ void __launch_bounds__(kMlpThreads, 2)
desc_rows_fp32_kernel(long long num_clusters, int n, int m, int S, float radius, const float *__restrict__ xyz,
                      const float *__restrict__ new_xyz, const int *__restrict__ idx,
                      const float *__restrict__ orientation, const float *__restrict__ P, WeightLayout L,
                      float *__restrict__ pooled2) {
    extern __shared__ float4 smem4[];
    float *in0_t = reinterpret_cast<float *>(smem4);  // [4][128]
    float *a1_t = in0_t + 4 * kTileRows;              // [32][128]
    float *a2_t = a1_t + 32 * kTileRows;              // [64][128]
    float *part = a2_t + 64 * kTileRows;              // [16][128]
    float *pool1 = part + 16 * 128;                   // [16][64]   (cpt <= 16)
    float *cterm = pool1 + 16 * 64;                   // [16][256]
    const int cpt = kTileRows / S;
    const int MID = L.mid;
    const long long tile = blockIdx.x;

    gather_rows(in0_t, tile, S, num_clusters, n, m, radius, xyz, new_xyz, idx, orientation);
    __syncthreads();
    {
        float acc[8][2];
        dense_tile<2>(in0_t, 3, P + L.off[W_DESC0], 32, 0, acc);
        bias_act_store<2, true>(acc, P + L.off[B_DESC0], 0, a1_t);
    }
    __syncthreads();
    {
        float acc[8][4];
        dense_tile<4>(a1_t, 32, P + L.off[W_DESC1], 64, 0, acc);
        bias_act_store<4, true>(acc, P + L.off[B_DESC1], 0, a2_t);
        pool_rows<4>(acc, part, S, cpt, [&](int cl, int col, float v) { pool1[cl * 64 + col] = v; });
    }
    __syncthreads();
    for (int e = threadIdx.x; e < cpt * MID; e += kMlpThreads) {  // per-cluster constant of conv_mid_0
        const int cl = e / MID, col = e - cl * MID;
        float v = __ldg(P + L.off[B_MID] + col);
        const float *w = P + L.off[W_MID] + 64 * MID + col;
        for (int k = 0; k < 64; ++k) v = fmaf(pool1[cl * 64 + k], __ldg(w + k * MID), v);
        cterm[cl * MID + col] = v;
    }
    __syncthreads();
    for (int chunk = 0; chunk < MID / 128; ++chunk) {
        float acc[8][8];
        dense_tile<8>(a2_t, 64, P + L.off[W_MID], MID, chunk * 128, acc);
        const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
        const int cl = (rg * 8) / S;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float cc = cterm[cl * MID + chunk * 128 + cg * 8 + j];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][j] += cc;
        }
        pool_rows<8>(acc, part, S, cpt, [&](int cl2, int col, float v) {
            const long long c = tile * cpt + cl2;
            if (c < num_clusters) pooled2[c * MID + chunk * 128 + col] = v;
        });
    }
}

// descriptor, per-cluster part (feat3dnet.py:79-84,185): conv_post_0 MID -> F (BN folded, no ReLU), l2-normalise.
template <int CT>
__global__ void __launch_bounds__(kMlpThreads, 1)
desc_post_fp32_kernel(long long num_clusters, const float *__restrict__ pooled2, const float *__restrict__ P,
                      WeightLayout L, float *__restrict__ features) {
    extern __shared__ float4 smem4[];
    constexpr int F = 16 * CT;
    constexpr int kLdIn = kTileRows + 4;             // see det_post_fp32_kernel
    float *in_t = reinterpret_cast<float *>(smem4);  // [MID][132]
    float *o_t = in_t + 256 * kLdIn;                 // [F][128]
    const int MID = L.mid;
    const long long c0 = static_cast<long long>(blockIdx.x) * kTileRows;
    for (int e = threadIdx.x; e < kTileRows * MID; e += kMlpThreads) {
        const int r = e / MID, k = e - r * MID;
        in_t[k * kLdIn + r] = (c0 + r < num_clusters) ? __ldg(pooled2 + (c0 + r) * MID + k) : 0.0f;
    }
    __syncthreads();
    {
        float acc[8][CT];
        dense_tile<CT>(in_t, MID, P + L.off[W_POST], F, 0, acc, kLdIn);
        bias_act_store<CT, false>(acc, P + L.off[B_POST], 0, o_t);
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < kTileRows && c0 + r < num_clusters) {
        float ss = 0.0f;
        for (int k = 0; k < F; ++k) {
            const float v = o_t[k * kTileRows + r];
            ss = fmaf(v, v, ss);
        }
        const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-8f));
        for (int k = 0; k < F; ++k) features[(c0 + r) * F + k] = o_t[k * kTileRows + r] * inv;
    }
}

static bool supported_nsample(int S) { return S == 8 || S == 16 || S == 32 || S == 64 || S == 128; }

int detector_forward_fp32(int b, int n, int m, int S, float radius, const float *xyz, const float *new_xyz, const int *idx,
                          const float *packed, float *pooled_ws, float *attention, float *orientation, cudaStream_t st) {
    if (!supported_nsample(S)) return fail(F3D_ERR_UNSUPPORTED, "detector_forward: nsample must be 8,16,32,64 or 128");
    const long long nc = static_cast<long long>(b) * m;
    if (nc == 0) return 0;
    const WeightLayout L = make_weight_layout(32);
    const int cpt = kTileRows / S;
    const size_t smem_rows = sizeof(float) * (4 + 64 + 128 + 16) * kTileRows;
    const size_t smem_post = sizeof(float) * (256 * (kTileRows + 4) + 128 * kTileRows);
    cudaFuncSetAttribute(det_rows_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows));
    cudaFuncSetAttribute(det_post_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_post));
    det_rows_fp32_kernel<<<static_cast<unsigned>((nc + cpt - 1) / cpt), kMlpThreads, smem_rows, st>>>(
        nc, n, m, S, radius, xyz, new_xyz, idx, packed, L, pooled_ws);
    int rc = check_launch("det_rows_fp32_kernel");
    if (rc) return rc;
    det_post_fp32_kernel<<<static_cast<unsigned>((nc + kTileRows - 1) / kTileRows), kMlpThreads, smem_post, st>>>(
        nc, pooled_ws, packed, L, attention, orientation);
    return check_launch("det_post_fp32_kernel");
}

int detector_post_fp32(long long nc, const float *pooled, const float *packed, float *attention, float *orientation,
                       cudaStream_t st) {
    if (nc == 0) return 0;
    const WeightLayout L = make_weight_layout(32);
    const size_t smem_post = sizeof(float) * (256 * (kTileRows + 4) + 128 * kTileRows);
    cudaFuncSetAttribute(det_post_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_post));
    det_post_fp32_kernel<<<static_cast<unsigned>((nc + kTileRows - 1) / kTileRows), kMlpThreads, smem_post, st>>>(
        nc, pooled, packed, L, attention, orientation);
    return check_launch("det_post_fp32_kernel");
}

int descriptor_post_fp32(long long nc, const float *pooled2, const float *packed, int feature_dim, float *features,
                         cudaStream_t st) {
    if (nc == 0) return 0;
    const WeightLayout L = make_weight_layout(feature_dim);
    const size_t smem_post = sizeof(float) * (256 * (kTileRows + 4) + 128 * kTileRows);
    const unsigned gp = static_cast<unsigned>((nc + kTileRows - 1) / kTileRows);
#define F3D_POST2(CT)                                                                                                 \
    cudaFuncSetAttribute(desc_post_fp32_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                         static_cast<int>(smem_post));                                                                \
    desc_post_fp32_kernel<CT><<<gp, kMlpThreads, smem_post, st>>>(nc, pooled2, packed, L, features)
    switch (feature_dim) {
        case 16: F3D_POST2(1); break;
        case 32: F3D_POST2(2); break;
        case 64: F3D_POST2(4); break;
        default: F3D_POST2(8); break;
    }
#undef F3D_POST2
    return check_launch("desc_post_fp32_kernel");
}

int descriptor_forward_fp32(int b, int n, int m, int S, float radius, int feature_dim, const float *xyz,
                            const float *new_xyz, const int *idx, const float *orientation, const float *packed,
                            float *pooled_ws, float *features, cudaStream_t st) {
    if (!supported_nsample(S)) return fail(F3D_ERR_UNSUPPORTED, "descriptor_forward: nsample must be 8,16,32,64 or 128");
    if (feature_dim != 16 && feature_dim != 32 && feature_dim != 64 && feature_dim != 128)
        return fail(F3D_ERR_UNSUPPORTED, "descriptor_forward: feature_dim must be 16,32,64 or 128 (inference.py:41)");
    const long long nc = static_cast<long long>(b) * m;
    if (nc == 0) return 0;
    const WeightLayout L = make_weight_layout(feature_dim);
    const int cpt = kTileRows / S;
    const size_t smem_rows = sizeof(float) * ((4 + 32 + 64 + 16) * kTileRows + 16 * 64 + 16 * 256);
    const size_t smem_post = sizeof(float) * (256 * (kTileRows + 4) + 128 * kTileRows);
    cudaFuncSetAttribute(desc_rows_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_rows));
    desc_rows_fp32_kernel<<<static_cast<unsigned>((nc + cpt - 1) / cpt), kMlpThreads, smem_rows, st>>>(
        nc, n, m, S, radius, xyz, new_xyz, idx, orientation, packed, L, pooled_ws);
    int rc = check_launch("desc_rows_fp32_kernel");
    if (rc) return rc;
    const unsigned gp = static_cast<unsigned>((nc + kTileRows - 1) / kTileRows);
#define F3D_POST(CT)                                                                                                  \
    cudaFuncSetAttribute(desc_post_fp32_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,                      \
                         static_cast<int>(smem_post));                                                                \
    desc_post_fp32_kernel<CT><<<gp, kMlpThreads, smem_post, st>>>(nc, pooled_ws, packed, L, features)
    switch (feature_dim) {
        case 16: F3D_POST(1); break;
        case 32: F3D_POST(2); break;
        case 64: F3D_POST(4); break;
        default: F3D_POST(8); break;
    }
#undef F3D_POST
    return check_launch("desc_post_fp32_kernel");
}

}  // namespace f3d
