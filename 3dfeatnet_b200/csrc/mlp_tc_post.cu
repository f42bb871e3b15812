// mlp_tc_post.cu -- the per-cluster tails of the detector and the descriptor on the tensor cores ("bf16x3").
//
// Detector (models/feat3dnet.py:134-149): pooled (256) -> conv_post_0 (128, BN, ReLU) -> conv_post_1 (64, BN, ReLU) ->
//   attention = softplus(64 -> 1), orientation = atan2 of the l2-normalised (64 -> 2) output.  The two heads (3 outputs per cluster,
//   12 288 multiply-adds per 64-cluster tile) run on the CUDA cores in exact fp32 straight from conv_post_1's activations: as a padded
//   128-row MMA they cost 504 cycles of tensor pipe, an operand conversion, a proxy fence and a commit / wait round trip per tile.
// Descriptor (models/feat3dnet.py:71-84,185): the per-cluster half of conv_mid_0 -- rows 64..127 of its weight times the max-pooled conv1
//   vector, which the reference tiles over the 64 samples and concatenates (feat3dnet.py:66-70); by the split-weight identity it is one
//   column per cluster and is added here to the row kernel's max-pooled per-point half -- -> conv_post_0 (F, BN, no ReLU) -> l2-normalise.
//
// These layers are 1.6 % of the flops but ran at 140 + 55 us as fp32 FFMA kernels (one 128-cluster tile per SM, weights
// streamed through a 28 KB L1).  Here a tile is 64 clusters = the N axis of every MMA; the weights (bf16 hi/lo) are copied
// ONCE per CTA into tensor memory (tcgen05.cp) and every layer is D[channel x cluster] = W^T (TMEM) * X^T (shared memory),
// the same D^T = W^T X^T formulation as the row kernels: an epilogue thread owns one output channel of all 64 clusters,
// adds its bias, applies ReLU and writes the next operand cluster-contiguous (MN-major, 16-byte stores).  The per-cluster heads (3 rows of a padded 128-row
// weight) and the l2-normalisation need all channels of one cluster: the accumulator goes through a small shared-memory
// transpose and one thread per cluster finishes (softplus / atan2 / rsqrt).
// Phases of a tile are sequential (__syncthreads + one mbarrier for MMA completion): at 3.5 tiles per SM the kernel is
// latency- not throughput-bound, and the simple structure keeps it obviously correct.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "weights_layout.h"

#include <cuda_bf16.h>

namespace f3d {

using namespace tc;

namespace post {
constexpr int kTile = 64;                  // clusters per tile = MMA N
constexpr int kThreads = 256;
constexpr uint32_t kSbo = 128;
constexpr uint32_t kLboW = 128 * 16;       // weight images: 128 rows per K chunk
constexpr uint32_t kLboIn = kTile * 16;    // first operand: written 16 B per thread
// later operands are written by the epilogues (thread = channel, 32 clusters) CLUSTER-contiguous = MN-major: 16-byte stores
constexpr uint32_t kLboX = 128;              // K groups of 8 channels
constexpr uint32_t kSboX2 = 16 * 128;        // groups of 8 clusters, 128 channels (layer 2 operand)
constexpr uint32_t kSboX3 = 8 * 128;         // groups of 8 clusters, 64 channels (layer 3 operand)
constexpr uint32_t kX2Split = (kTile / 8) * kSboX2;  // 16 KB
constexpr uint32_t kX3Split = (kTile / 8) * kSboX3;  // 8 KB
constexpr uint32_t kStage = 64 * 1024;     // staging buffer: weight pieces at start-up, then the first operand of each tile
// shared memory map
constexpr uint32_t kOffStage = 0;
constexpr uint32_t kOffX2 = kOffStage + kStage;            // operand of layer 2: [split 2][cluster group 8][chunk 16][8 ch][8 clusters]
constexpr uint32_t kOffX3 = kOffX2 + 2 * kX2Split;         // operand of layer 3: [split 2][cluster group 8][chunk 8][8 ch][8 clusters]
constexpr uint32_t kOffOut = kOffX3 + 2 * kX3Split;        // fp32 [128][kTile + 1] accumulator transpose
constexpr uint32_t kOffBias = kOffOut + 128 * (kTile + 1) * 4;  // fp32 [512]
constexpr uint32_t kOffBars = kOffBias + 512 * 4;
constexpr uint32_t kSmemBytes = kOffBars + 64;
static_assert(kOffX2 % 128 == 0 && kOffX3 % 128 == 0 && kOffOut % 16 == 0 && kOffBars % 8 == 0, "alignment");
// tensor memory map (columns): detector W3 0..255 (2 splits x 128), W4 256..383 (2 x 64), D 448..511
//                               descriptor Wp 0..127 (2 x 64), Wb 128..191 (2 x 32), D 448..511
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemD = 448;
// global weight image (bytes): detector [W3 hi 64K][W3 lo 64K][W4 hi 32K][W4 lo 32K][fp32 512: biases 0..194, head weights 256..447]
constexpr uint32_t kDetImgBytes = 2 * 65536 + 2 * 32768 + 2048;
// descriptor [Wp hi 32K][Wp lo 32K][Wb hi 16K][Wb lo 16K][bias fp32 512]
constexpr uint32_t kDescImgBytes = 2 * 32768 + 2 * 16384 + 2048;
}  // namespace post

// One elected lane copies `nsteps` K-steps (16 bf16 = 8 TMEM columns each) of a staged weight piece into tensor memory.
__device__ __forceinline__ void post_cp_weights(uint32_t tmem_col, uint32_t smem_addr, int nsteps) {
    for (int k = 0; k < nsteps; ++k)
        tmem_cp_128x256b(tmem_col + k * 8, make_smem_desc(smem_addr + k * 2 * post::kLboW, post::kLboW, post::kSbo));
}

// fp32 rows (global, row stride `ld` floats) of 64 clusters -> bf16 hi/lo K-major operand image in the staging buffer, in two
// halves so that the loads of the NEXT tile can fly while the tensor pipe runs layer 1 of the current one
template <int K>
struct PostRows {
    static constexpr int kIter = K / 32;  // chunks of 8 channels per thread
    float4 va[kIter], vb[kIter];
    __device__ __forceinline__ void load(const float *__restrict__ src, int ld, long long c0, long long num_clusters) {
        const int r = threadIdx.x & 63, cq = threadIdx.x >> 6;
        const bool valid = c0 + r < num_clusters;
        const float *row = src + (c0 + r) * ld;
#pragma unroll
        for (int i = 0; i < kIter; ++i) {  // all loads in flight before the first conversion (L2 latency paid once)
            va[i] = vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                va[i] = __ldg(reinterpret_cast<const float4 *>(row + (cq + 4 * i) * 8));
                vb[i] = __ldg(reinterpret_cast<const float4 *>(row + (cq + 4 * i) * 8 + 4));
            }
        }
    }
    // rows of 2 x K floats (the two sample-half maxima of a cluster's conv1 channels, written by desc_rows_tc_kernel): their maximum
    __device__ __forceinline__ void load_max(const float *__restrict__ src, long long c0, long long num_clusters) {
        const int r = threadIdx.x & 63, cq = threadIdx.x >> 6;
        const bool valid = c0 + r < num_clusters;
        const float *row = src + (c0 + r) * (2 * K);
        float4 ua[kIter], ub[kIter];
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            va[i] = vb[i] = ua[i] = ub[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                va[i] = __ldg(reinterpret_cast<const float4 *>(row + (cq + 4 * i) * 8));
                vb[i] = __ldg(reinterpret_cast<const float4 *>(row + (cq + 4 * i) * 8 + 4));
                ua[i] = __ldg(reinterpret_cast<const float4 *>(row + K + (cq + 4 * i) * 8));
                ub[i] = __ldg(reinterpret_cast<const float4 *>(row + K + (cq + 4 * i) * 8 + 4));
            }
        }
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            va[i] = make_float4(fmaxf(va[i].x, ua[i].x), fmaxf(va[i].y, ua[i].y), fmaxf(va[i].z, ua[i].z), fmaxf(va[i].w, ua[i].w));
            vb[i] = make_float4(fmaxf(vb[i].x, ub[i].x), fmaxf(vb[i].y, ub[i].y), fmaxf(vb[i].z, ub[i].z), fmaxf(vb[i].w, ub[i].w));
        }
    }
    __device__ __forceinline__ void store(uint8_t *img, uint32_t split_bytes) const {
        const int r = threadIdx.x & 63, cq = threadIdx.x >> 6;
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
            const int c = cq + 4 * i;
            const float4 a = va[i], b = vb[i];
            const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            }
            uint8_t *dst = img + c * post::kLboIn + r * 16;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(dst + split_bytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
};

// 3-pass MMA group: D[128 x 64] = (Whi + Wlo) (TMEM) * (Xhi + Xlo)^T (shared), dropping lo*lo.  One elected lane.
__device__ __forceinline__ void post_mma(uint32_t d, uint32_t w_hi, uint32_t w_lo, uint32_t x_hi, uint32_t x_lo, uint32_t lbo, uint32_t sbo,
                                         int ksteps, uint32_t idesc) {
    uint32_t acc = 0;
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t wa = pass == 2 ? w_lo : w_hi;
        const uint32_t xb = pass == 1 ? x_lo : x_hi;
        for (int k = 0; k < ksteps; ++k) {
            umma_f16_ts(d, wa + k * 8, make_smem_desc(xb + k * 2 * lbo, lbo, sbo), idesc, acc);
            acc = 1;
        }
    }
}

// epilogue helper: this thread's channel of 32 clusters (columns col0..col0+31) from the accumulator
__device__ __forceinline__ void post_load_acc(uint32_t tmem_base, int q, int col0, uint32_t (&r)[32]) {
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + post::kTmemD + col0, r);
    tmem_ld_wait();
}

// bias (+ReLU) and hi/lo split of 32 accumulator values (this thread's channel, clusters col0..col0+31), stored
// cluster-contiguous (MN-major) as the next layer's operand: 8 consecutive clusters of a channel are 16 contiguous bytes
template <bool RELU>
__device__ __forceinline__ void post_store_operand(uint8_t *x, uint32_t split_bytes, uint32_t sbo, int ch, int col0, const uint32_t (&r)[32],
                                                   float bias) {
    uint8_t *base = x + (col0 >> 3) * sbo + (ch >> 3) * post::kLboX + (ch & 7) * 16;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float va = __uint_as_float(r[g * 8 + 2 * j]) + bias, vb = __uint_as_float(r[g * 8 + 2 * j + 1]) + bias;
            if (RELU) {
                va = fmaxf(va, 0.0f);
                vb = fmaxf(vb, 0.0f);
            }
            split_bf16x2(va, vb, hi[j], lo[j]);
        }
        *reinterpret_cast<uint4 *>(base + g * sbo) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(base + split_bytes + g * sbo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// MODE 0: detector tail (256 -> 128 -> 64 -> {attention, orientation}), `pooled` = the row kernel's max-pooled 256 channels.
// MODE 1: descriptor tail (pooled conv1 vector 64 -> + per-point half = conv_mid_0 128 -> F -> l2norm), `pooled` = the row kernel's
//         max-pooled per-point half of conv_mid_0 incl. bias (128 per cluster), `aux` = its two sample-half maxima of conv1 (2 x 64)
template <int MODE>
__global__ void __launch_bounds__(post::kThreads, 1)
post_tc_kernel(long long num_clusters, int feature_dim, const float *__restrict__ pooled, const float *__restrict__ aux,
               const uint8_t *__restrict__ wimg, float *__restrict__ out0, float *__restrict__ out1) {
    using namespace post;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + kOffBars);
    uint64_t *bar_m = bar_w + 1;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bar_w + 2);
    float *bias = reinterpret_cast<float *>(smem + kOffBias);
    float *outT = reinterpret_cast<float *>(smem + kOffOut);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;                 // TMEM lane quarter
    const int ch = q * 32 + lane;           // output channel owned in the epilogues
    const int col0 = (warp >> 2) * 32;      // warps 0-3: clusters 0..31, warps 4-7: clusters 32..63
    const uint32_t sbase = smem_u32(smem);

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_m, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_base_s, kTmemCols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    uint32_t wpar = 0, mpar = 0;

    // ---- weights -> tensor memory, piece by piece through the staging buffer (bulk TMA, then tcgen05.cp) -------------
    constexpr int kPieces = MODE == 0 ? 3 : 2;
    for (int piece = 0; piece < kPieces; ++piece) {
        // piece -> (global offset, bytes); detector: W3 hi | W3 lo | W4 hi+lo ; descriptor: Wp hi+lo | Wb hi+lo
        const uint32_t goff = piece * 65536u;
        const uint32_t bytes = MODE == 0 ? 65536u : (piece == 0 ? 65536u : 32768u);
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar_w, bytes);
            bulk_g2s(smem + kOffStage, wimg + goff, bytes, bar_w);  // one copy per piece: every cp.async.bulk costs its issuer ~450 cycles
        }
        mbar_wait(bar_w, wpar);
        wpar ^= 1;
        tcgen05_fence_after();
        if (warp == 0) {
            if (elect_one()) {
                if (MODE == 0) {
                    if (piece < 2) post_cp_weights(tmem_base + piece * 128, sbase + kOffStage, 16);       // W3 split `piece`, K = 256
                    else {
                        post_cp_weights(tmem_base + 256, sbase + kOffStage, 8);                             // W4 hi, K = 128
                        post_cp_weights(tmem_base + 320, sbase + kOffStage + 32768, 8);                     // W4 lo
                    }
                } else if (piece == 0) {
                    post_cp_weights(tmem_base + 0, sbase + kOffStage, 8);                                   // Wp hi, K = 128
                    post_cp_weights(tmem_base + 64, sbase + kOffStage + 32768, 8);                          // Wp lo
                } else {
                    post_cp_weights(tmem_base + 128, sbase + kOffStage, 4);                                 // Wb hi, K = 64
                    post_cp_weights(tmem_base + 160, sbase + kOffStage + 16384, 4);                         // Wb lo
                }
                umma_commit(bar_m);
            }
            __syncwarp();
        }
        mbar_wait(bar_m, mpar);  // the staging buffer may be overwritten once the copies have completed
        mpar ^= 1;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 512; i += kThreads)
        bias[i] = __ldg(reinterpret_cast<const float *>(wimg + (MODE == 0 ? kDetImgBytes : kDescImgBytes) - 2048) + i);
    __syncthreads();

    const uint32_t idesc = make_idesc(1, 128, kTile);
    const long long ntiles = (num_clusters + kTile - 1) / kTile;
    // ---- first operand from HBM/L2: pooled rows -> bf16 hi/lo, K-major.  The rows of tile t+1 are requested right after layer 1 of
    // tile t has been issued and converted into the staging buffer as soon as that MMA group has completed (its only reader).
    constexpr int K1 = MODE == 0 ? 256 : 64;
    constexpr uint32_t kSplit1 = (K1 / 8) * kLboIn;
    PostRows<K1> rows;
    auto load_rows = [&](long long c0) {
        if (MODE == 0) rows.load(pooled, K1, c0, num_clusters);
        else rows.load_max(aux, c0, num_clusters);
    };
    if (static_cast<long long>(blockIdx.x) < ntiles) {
        load_rows(static_cast<long long>(blockIdx.x) * kTile);
        rows.store(smem + kOffStage, kSplit1);
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long c0 = tile * kTile;
        fence_proxy_async_smem();
        __syncthreads();
        // ---- layer 1
        if (warp == 0) {
            tcgen05_fence_after();
            if (elect_one()) {
                post_mma(tmem_base + kTmemD, tmem_base + (MODE == 0 ? 0 : 128), tmem_base + (MODE == 0 ? 128 : 160), sbase + kOffStage,
                         sbase + kOffStage + kSplit1, kLboIn, kSbo, K1 / 16, idesc);
                umma_commit(bar_m);
            }
            __syncwarp();
        }
        const long long next = tile + gridDim.x;
        if (next < ntiles) load_rows(next * kTile);
        float pa[MODE == 0 ? 1 : 32];  // descriptor: the per-point half of conv_mid_0 of this thread's channel, 32 clusters (in flight during the MMA)
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pa[j] = c0 + col0 + j < num_clusters ? __ldg(pooled + (c0 + col0 + j) * 128 + ch) : 0.0f;
        }
        mbar_wait(bar_m, mpar);
        mpar ^= 1;
        tcgen05_fence_after();
        if (next < ntiles) rows.store(smem + kOffStage, kSplit1);  // layer 1 was the staging buffer's only reader
        uint32_t r[32];
        post_load_acc(tmem_base, q, col0, r);
        if (MODE == 0) {
            post_store_operand<true>(smem + kOffX2, kX2Split, kSboX2, ch, col0, r, bias[ch]);  // conv_post_0: 128 channels, ReLU
            tcgen05_fence_before();
            fence_proxy_async_smem();
            __syncthreads();
            // ---- layer 2: conv_post_1, 128 -> 64 (rows 64..127 of the weight are zero padding)
            if (warp == 0) {
                tcgen05_fence_after();
                if (elect_one()) {
                    post_mma(tmem_base + kTmemD, tmem_base + 256, tmem_base + 320, sbase + kOffX2, sbase + kOffX2 + kX2Split, kLboX, kSboX2, 8,
                             idesc | kIdescBMnMajor);
                    umma_commit(bar_m);
                }
                __syncwarp();
            }
            mbar_wait(bar_m, mpar);
            mpar ^= 1;
            tcgen05_fence_after();
            // conv_post_1's 64 activations (bias, ReLU) in fp32, channel-major with the clusters contiguous
            if (q < 2) {
                post_load_acc(tmem_base, q, col0, r);
                const float b2 = bias[128 + ch];
#pragma unroll
                for (int j = 0; j < 32; ++j) outT[ch * (kTile + 1) + col0 + j] = fmaxf(__uint_as_float(r[j]) + b2, 0.0f);
            }
            tcgen05_fence_before();
            __syncthreads();
            {  // heads on the CUDA cores: thread = (cluster, quarter of the 64 channels); partial sums into rows 64.. of outT
                const int c = threadIdx.x & 63, part = threadIdx.x >> 6;
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const int kc = part * 16 + k;
                    const float v = outT[kc * (kTile + 1) + c];
                    a0 = fmaf(v, bias[256 + kc], a0);          // attention (64 -> 1)
                    a1 = fmaf(v, bias[320 + 2 * kc], a1);      // orientation (64 -> 2)
                    a2 = fmaf(v, bias[320 + 2 * kc + 1], a2);
                }
                outT[(64 + part * 3 + 0) * (kTile + 1) + c] = a0;
                outT[(64 + part * 3 + 1) * (kTile + 1) + c] = a1;
                outT[(64 + part * 3 + 2) * (kTile + 1) + c] = a2;
            }
            __syncthreads();
            if (threadIdx.x < kTile && c0 + threadIdx.x < num_clusters) {
                const int c = threadIdx.x;
                float h[3];
#pragma unroll
                for (int o = 0; o < 3; ++o)
                    h[o] = ((outT[(64 + o) * (kTile + 1) + c] + outT[(67 + o) * (kTile + 1) + c]) + outT[(70 + o) * (kTile + 1) + c]) +
                           outT[(73 + o) * (kTile + 1) + c] + bias[192 + o];
                const float att = h[0], ox = h[1], oy = h[2];
                out0[c0 + c] = att > 20.0f ? att : log1pf(expf(att));                     // softplus
                const float inv = 1.0f / sqrtf(fmaxf(ox * ox + oy * oy, 1e-8f));          // tf.nn.l2_normalize(eps=1e-8)
                out1[c0 + c] = atan2f(oy * inv, ox * inv);
            }
            __syncthreads();  // outT and the operand buffers are reused by the next tile
        } else {
            // descriptor: conv_mid_0 = per-point half (max-pooled by the row kernel, bias included) + pooled half (this accumulator); no ReLU
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + pa[j]);
            post_store_operand<false>(smem + kOffX2, kX2Split, kSboX2, ch, col0, r, 0.0f);
            tcgen05_fence_before();
            fence_proxy_async_smem();
            __syncthreads();
            // ---- conv_post_0, 128 -> F (rows F..127 of the weight are zero padding)
            if (warp == 0) {
                tcgen05_fence_after();
                if (elect_one()) {
                    post_mma(tmem_base + kTmemD, tmem_base + 0, tmem_base + 64, sbase + kOffX2, sbase + kOffX2 + kX2Split, kLboX, kSboX2, 8,
                             idesc | kIdescBMnMajor);
                    umma_commit(bar_m);
                }
                __syncwarp();
            }
            mbar_wait(bar_m, mpar);
            mpar ^= 1;
            tcgen05_fence_after();
            post_load_acc(tmem_base, q, col0, r);
            // conv_post_0 (no ReLU) -> l2-normalise over the F channels of each cluster
            if (ch < feature_dim) {
#pragma unroll
                for (int j = 0; j < 32; ++j) outT[ch * (kTile + 1) + col0 + j] = __uint_as_float(r[j]) + bias[ch];
            }
            tcgen05_fence_before();
            __syncthreads();
            {  // four threads per cluster, a contiguous quarter of the channels each: 4 x 32 contiguous bytes per cluster row at F = 32
                const int c = threadIdx.x >> 2, qd = threadIdx.x & 3;
                const int kb = (feature_dim + 3) >> 2, k0 = qd * kb, k1 = min(feature_dim, k0 + kb);
                float ss = 0.0f;
                for (int k = k0; k < k1; ++k) {
                    const float v = outT[k * (kTile + 1) + c];
                    ss = fmaf(v, v, ss);
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-8f));
                if (c0 + c < num_clusters) {
                    float *dst = out0 + (c0 + c) * feature_dim;
                    if ((kb & 3) == 0 && (feature_dim & 3) == 0 && (reinterpret_cast<uintptr_t>(out0) & 15) == 0) {
                        for (int k = k0; k < k1; k += 4)
                            *reinterpret_cast<float4 *>(dst + k) =
                                make_float4(outT[k * (kTile + 1) + c] * inv, outT[(k + 1) * (kTile + 1) + c] * inv,
                                            outT[(k + 2) * (kTile + 1) + c] * inv, outT[(k + 3) * (kTile + 1) + c] * inv);
                    } else {
                        for (int k = k0; k < k1; ++k) dst[k] = outT[k * (kTile + 1) + c] * inv;
                    }
                }
            }
            __syncthreads();
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// weight images from the packed fp32 (BN-folded) weights: element (r,k) of A = W^T at (k/8)*kLboW + r*16 + (k%8)*2
__global__ void post_prep_kernel(const float *__restrict__ P, WeightLayout L, int mode, uint8_t *__restrict__ img) {
    using namespace post;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto put = [&](uint32_t hi_off, uint32_t lo_off, int r, int k, float w) {
        const uint32_t o = (k >> 3) * kLboW + r * 16 + (k & 7) * 2;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        *reinterpret_cast<__nv_bfloat16 *>(img + hi_off + o) = h;
        *reinterpret_cast<__nv_bfloat16 *>(img + lo_off + o) = __float2bfloat16_rn(w - __bfloat162float(h));
    };
    if (mode == 0) {
        if (i < 128 * 256) {  // conv_post_0: W (256,128) -> A[r][k] = W[k][r]
            const int r = i & 127, k = i >> 7;
            put(0, 65536, r, k, P[L.off[W_DETP0] + k * 128 + r]);
        } else if (i < 128 * 256 + 128 * 128) {  // conv_post_1: W (128,64), rows 64..127 zero
            const int e = i - 128 * 256;
            const int r = e & 127, k = e >> 7;
            put(131072, 131072 + 32768, r, k, r < 64 ? P[L.off[W_DETP1] + k * 64 + r] : 0.0f);
        } else if (i < 128 * 256 + 128 * 128 + 512) {  // biases, then the heads' fp32 weights (attention 64, orientation 64 x 2)
            const int e = i - 128 * 256 - 128 * 128;
            float b = 0.0f;
            if (e < 128) b = P[L.off[B_DETP0] + e];
            else if (e < 192) b = P[L.off[B_DETP1] + e - 128];
            else if (e == 192) b = P[L.off[B_ATT]];
            else if (e < 195) b = P[L.off[B_ORI] + e - 193];
            else if (e >= 256 && e < 320) b = P[L.off[W_ATT] + e - 256];
            else if (e >= 320 && e < 448) b = P[L.off[W_ORI] + e - 320];
            reinterpret_cast<float *>(img + kDetImgBytes - 2048)[e] = b;
        }
    } else {
        const int F = L.feature_dim;
        if (i < 128 * 128) {  // conv_post_0 of the descriptor: W (128,F), rows F..127 zero
            const int r = i & 127, k = i >> 7;
            put(0, 32768, r, k, r < F ? P[L.off[W_POST] + k * F + r] : 0.0f);
        } else if (i < 128 * 128 + 128 * 64) {  // Wb^T = W_mid[64:128, :]^T: the half of conv_mid_0 that sees the pooled conv1 vector
            const int e = i - 128 * 128;
            const int r = e & 127, k = e >> 7;
            put(65536, 65536 + 16384, r, k, P[L.off[W_MID] + (64 + k) * 128 + r]);
        } else if (i < 128 * 128 + 128 * 64 + 512) {
            const int e = i - 128 * 128 - 128 * 64;
            reinterpret_cast<float *>(img + kDescImgBytes - 2048)[e] = e < F ? P[L.off[B_POST] + e] : 0.0f;
        }
    }
}

static int post_num_sms() {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    return num_sms;
}

size_t post_tc_weight_bytes() { return post::kDetImgBytes; }  // >= kDescImgBytes

int detector_post_tc(long long nc, const float *pooled, const float *packed, uint8_t *wimg, float *attention, float *orientation,
                     bool build_image, int max_ctas, cudaStream_t st) {
    if (nc == 0) return 0;
    int rc = 0;
    if (build_image) {
        const int total = 128 * 256 + 128 * 128 + 512;
        post_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(packed, make_weight_layout(32), 0, wimg);
        rc = check_launch("post_prep_kernel");
        if (rc) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(post_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(post::kSmemBytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "post_tc: cudaFuncSetAttribute");
    const long long ntiles = (nc + post::kTile - 1) / post::kTile;
    const int ctas = (max_ctas > 0 && max_ctas < post_num_sms()) ? max_ctas : post_num_sms();
    const unsigned grid = static_cast<unsigned>(ntiles < ctas ? ntiles : ctas);
    ktimer_begin("post_tc_kernel<detector>", 2.0 * (256.0 * 128 + 128 * 64 + 64 * 3) * static_cast<double>(nc), st);
    post_tc_kernel<0><<<grid, post::kThreads, post::kSmemBytes, st>>>(nc, 0, pooled, nullptr, wimg, attention, orientation);
    ktimer_end(st);
    return check_launch("post_tc_kernel<detector>");
}

int descriptor_post_tc(long long nc, int feature_dim, const float *pooledA, const float *pmaxh, const float *packed, uint8_t *wimg,
                       float *features, bool build_image, int max_ctas, cudaStream_t st) {
    if (nc == 0) return 0;
    int rc = 0;
    if (build_image) {
        const int total = 128 * 128 + 128 * 64 + 512;
        post_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(packed, make_weight_layout(feature_dim), 1, wimg);
        rc = check_launch("post_prep_kernel");
        if (rc) return rc;
    }
    cudaError_t e = cudaFuncSetAttribute(post_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(post::kSmemBytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "post_tc: cudaFuncSetAttribute");
    const long long ntiles = (nc + post::kTile - 1) / post::kTile;
    const int ctas = (max_ctas > 0 && max_ctas < post_num_sms()) ? max_ctas : post_num_sms();
    const unsigned grid = static_cast<unsigned>(ntiles < ctas ? ntiles : ctas);
    ktimer_begin("post_tc_kernel<descriptor>", 2.0 * (64.0 * 128.0 + 128.0 * feature_dim) * static_cast<double>(nc), st);
    post_tc_kernel<1><<<grid, post::kThreads, post::kSmemBytes, st>>>(nc, feature_dim, pooledA, pmaxh, wimg, features, nullptr);
    ktimer_end(st);
    return check_launch("post_tc_kernel<descriptor>");
}

}  // namespace f3d
