// mlp_tc.cu -- fused detector forward on the Blackwell tensor cores (tcgen05.mma, TMEM accumulators, bulk-TMA
// weight staging), precision "bf16x3".
//
// Replaces the same TensorFlow graph as mlp_fp32.cu (models/feat3dnet.py:112-130: group, translate, /radius,
// conv 3->64->128->256 + BN + ReLU, reduce_max).  The two dense contractions (64->128 and 128->256, 97 % of the
// detector's flops) run on tcgen05; the K=3 first layer, bias/ReLU, the max-pool and the operand repacking stay on
// CUDA cores in the same kernel.  Nothing but the pooled (B*M,256) vectors reaches HBM.
//
// Numerics: every fp32 operand x is split as x = hi + lo with hi = bf16(x), lo = bf16(x - hi); a product is
// evaluated as hi*hi + hi*lo + lo*hi on kind::f16 MMAs with fp32 accumulation in TMEM.  The dropped terms are
// O(2^-16) relative, i.e. ~1e-5 -- two orders tighter than single-pass TF32 at 1.5x its tensor time and with the
// same shared-memory footprint (2 x 2 bytes per element), which is what lets all weights stay SM-resident.
//
// Formulation: D^T = W^T X^T -- the M (TMEM lane) axis is the OUTPUT CHANNEL, the N (TMEM column) axis is the
// sample.  One thread of the epilogue therefore owns one channel of every sample of a cluster: the max-pool over
// the 64 samples is a register reduction, the bias is a per-thread scalar, and the next layer's operand is written
// K-major with one 2-byte store per value.
//
// Persistent kernel, one CTA per SM, one cluster (64 samples) per tile, 13 warps:
//   warp 0      : TMEM allocation, bulk-TMA of the packed weights, then ONE lane issues every tcgen05.mma
//   warps 1-4   : producers -- gather + normalise + layer 0 (3->64, FFMA) -> operand X1 (bf16 hi/lo, K-major)
//   warps 5-8   : epilogue of even tiles, warps 9-12: epilogue of odd tiles
//                 E1: D1 (TMEM) -> +bias, ReLU -> operand X2 ;  E2: D2 (TMEM) -> max over samples, +bias, ReLU -> HBM
// conv2 (97 % of the MMA work) is issued per PAIR of tiles with N = 128: a 128x64x16 instruction holds the tensor pipe for 41.9
// cycles for 32 of math, a 128x128x16 one for 66.5 for 64 (profiles/r02_a_umma_instruction_shape.md).  The two tiles of a pair sit
// side by side in one X2 slot (sample-contiguous operand, 16 groups of 8 samples), the slots form a ring of two pairs, and the
// accumulator of each 128-channel block holds both tiles (128 columns).  Tensor memory: D1 64 + W1 64 + D2 2x128 + W2(hi) 128 = 512
// columns; the lo split of W2 (third product term) stays in shared memory and feeds its MMAs as an SS-mode A operand.
// Tensor-pipe order in steady state:  ... MMA2(p-1) | MMA1 of the next tiles | MMA2(p) ...  so E1 of pair p+1 runs under MMA2(p)
// and E2(p-1) under MMA2(p).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "weights_layout.h"

#include <cuda_bf16.h>

namespace f3d {

using namespace tc;

// ------------------------------------------------------------------------------------------------ self test
// One CTA: D[128 x N] = A[128 x K] * B[N x K]^T with operands given as canonical K-major no-swizzle bf16 images.
// Validates the descriptor encodings (LBO/SBO semantics, idesc, TMEM addressing) in isolation on the device.
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint8_t *__restrict__ a_img, const uint8_t *__restrict__ b_img, float *__restrict__ D, int N, int K,
                     uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, uint32_t a_bytes, uint32_t b_bytes, int a_in_tmem) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sa = smem;
    uint8_t *sb = smem + ((a_bytes + 127) & ~127u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_base_s, 512);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_load, a_bytes + b_bytes);
        bulk_g2s(sa, a_img, a_bytes, &bar_load);
        bulk_g2s(sb, b_img, b_bytes, &bar_load);
        mbar_wait(&bar_load, 0);
        tcgen05_fence_after();
        // a_in_tmem bit 1: the B image is MN-major (canonical no-swizzle: 8 K-rows x 16 B of 8 consecutive N)
        const uint32_t idesc = make_idesc(1, 128, static_cast<uint32_t>(N)) | ((a_in_tmem & 2) ? kIdescBMnMajor : 0u);
        a_in_tmem &= 1;
        if (a_in_tmem)  // stage A into TMEM columns 256.. (8 columns = 16 bf16 per K step), then run the MMAs from there
            for (int k = 0; k < K / 16; ++k)
                tmem_cp_128x256b(tmem_base + 256 + 8 * k, make_smem_desc(smem_u32(sa) + k * 2 * lbo_a, lbo_a, sbo_a));
        for (int k = 0; k < K / 16; ++k) {
            const uint64_t da = make_smem_desc(smem_u32(sa) + k * 2 * lbo_a, lbo_a, sbo_a);
            const uint64_t db = make_smem_desc(smem_u32(sb) + k * 2 * lbo_b, lbo_b, sbo_b);
            if (a_in_tmem)
                umma_f16_ts(tmem_base, tmem_base + 256 + 8 * k, db, idesc, k > 0 ? 1u : 0u);
            else
                umma_f16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tcgen05_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) D[static_cast<size_t>(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ detector rows
namespace det {
constexpr int kSamples = 64;                 // samples per cluster = MMA N
constexpr int kThreads = 14 * 32;            // warp 0 MMA2 issue, 1-4 producers, 5-12 epilogues, 13 MMA1 issue
constexpr uint32_t kSbo = 128;               // 8 rows x 16 B
constexpr uint32_t kLboW = 128 * 16;         // weights: 128 rows per K chunk
constexpr uint32_t kLboX1 = kSamples * 16;   // X1: written 16 B per thread, no padding needed
// X2 is stored SAMPLE-contiguous (MN-major B operand): a core matrix is 8 channels x 16 B (8 consecutive samples), so an
// epilogue thread (one channel, 64 samples) writes 8 x 16 B per split instead of 64 x 2 B
constexpr uint32_t kLboX2 = 128;                 // K groups of 8 channels
constexpr uint32_t kSboX2 = 16 * 128;            // groups of 8 samples (128 channels)
// Weight image in global memory (det_tc_prep_kernel), bytes:  [W2 lo | W1 | W2 hi | small], bulk-copied to the shared-memory
// offsets below.  W2 lo stays where it lands (A operand of the SS-mode MMAs); W1 and W2 hi are copied on into tensor memory
// (tcgen05.cp) and their staging area -- plus 32 KB behind it -- is recycled as the ring of X2 slots.
constexpr uint32_t kW1Split = 128 * 64 * 2;             // 16 KB per split
constexpr uint32_t kW2Blk = 128 * 128 * 2;              // 32 KB per (split, M block)
constexpr uint32_t kOffW2lo = 0;                        // [mblk 2][chunk 16][row 128][8] bf16                     64 KB
constexpr uint32_t kOffW1 = kOffW2lo + 2 * kW2Blk;      // [split 2][chunk 8][row 128][8] bf16 (staging)           32 KB
constexpr uint32_t kOffW2hi = kOffW1 + 2 * kW1Split;    // [mblk 2][chunk 16][row 128][8] bf16 (staging)           64 KB
constexpr uint32_t kImgBulk = kOffW2hi + 2 * kW2Blk;    // 160 KB copied to shared-memory offset 0
constexpr uint32_t kImgSmall = kImgBulk;                // image offset of the fp32 block: W0 float4[64], b1[128], b2[256]
constexpr uint32_t kSmallBytes = 64 * 16 + 128 * 4 + 256 * 4;
constexpr uint32_t kWeightBytes = kImgBulk + kSmallBytes;  // 166 400
// X2 ring: slot = pair & 1; one slot = [split 2][16 groups of 8 samples][128 channels x 16 B]: tile (pair*2 + g) owns sample groups
// g*8 .. g*8+7 of both splits, so the N = 128 operand of a pair is one uniform-stride block
constexpr uint32_t kX2Split = 16 * kSboX2;              // 32 KB
constexpr uint32_t kX2Slot = 2 * kX2Split;              // 64 KB
constexpr uint32_t kOffX2 = kOffW1;                     // aliases the W1 / W2 hi staging area (+ 32 KB)
constexpr uint32_t kOffW0 = kOffX2 + 2 * kX2Slot;       // 192 KB: float4 [64]: (w_x, w_y, w_z, bias) of layer 0 per channel
constexpr uint32_t kOffB1 = kOffW0 + 64 * 16;           // fp32 [128]
constexpr uint32_t kOffB2 = kOffB1 + 128 * 4;           // fp32 [256]
constexpr uint32_t kX1Split = 8 * kLboX1;               // 8 KB
constexpr uint32_t kOffX1 = kOffB2 + 256 * 4;           // [split 2][chunk 8][row 64][8] bf16
constexpr uint32_t kOffBars = kOffX1 + 2 * kX1Split;
constexpr uint32_t kSmemBytes = kOffBars + 20 * 8 + 16;
static_assert(kWeightBytes % 16 == 0 && kOffX1 % 128 == 0 && kOffX2 % 128 == 0 && kOffBars % 8 == 0 && kOffW0 % 16 == 0, "alignment");
static_assert(kOffX2 + 2 * kX2Slot >= kImgBulk, "the X2 ring covers the staging area");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
// TMEM columns: D1 at 0 (64), W1 at 64 + split*32 (64 K = 32 columns), D2 of M block mb at 128 + mb*128 (two tiles x 64 samples),
// W2 hi at 384 + mb*64 (128 K = 64 columns): all 512 columns in use.
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemW1 = 64;
constexpr uint32_t kTmemD2 = 128;
constexpr uint32_t kTmemW2 = 384;

// D2 (256 channels = two 128-lane M blocks A / B, each holding the two tiles of a pair) is single-buffered: its FULL / FREE barriers
// are per M block, so E2 drains block A while the tensor pipe computes block B, and MMA2(p+1) block A starts as soon as block B of
// pair p has been issued.  X2_FULL is per tile parity (the epilogue warpgroup that wrote it), X2_FREE per ring slot.
enum Bar { W_FULL = 0, W2_TMEM, X1_FULL, X1_FREE, X2_FULL0, X2_FULL1, X2_FREE0, X2_FREE1, D1_FULL0, D1_FULL1, D1_FREE,
           D2A_FULL, D2B_FULL, D2A_FREE, D2B_FREE, kNumBars };
static_assert(kNumBars <= 20, "barrier area");
// NOTE on mbarrier parity: a waiter may lag a barrier by at most ONE phase (try_wait.parity(p) is true as soon as the
// barrier is in the phase after p).  The two epilogue warpgroups alternate tiles, i.e. each sees every SECOND completion of
// a per-tile event, so the per-TILE barriers they wait on are per-warpgroup (D1_FULL[2]); per-PAIR events (D2A/B_FULL, X2_FREE of a
// slot) are seen by both warpgroups at every completion and use one barrier each, like the ones the MMA warps wait on (D2A/B_FREE,
// X1_FULL, X2_FULL[2]).
}  // namespace det

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

__global__ void __launch_bounds__(det::kThreads, 1)
det_rows_tc_kernel(long long num_clusters, int n, int m, float radius, const float *__restrict__ xyz,
                   const float *__restrict__ new_xyz, const int *__restrict__ idx, const uint8_t *__restrict__ wimg,
                   float *__restrict__ pooled, long long *__restrict__ dbg) {
    using namespace det;
    // optional timeline of CTA 0 (bring-up / profiling): dbg[tile*16 + slot] = clock64() at named points
    auto stamp = [&](int t, int slot) {
        if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) dbg[t * 16 + slot] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBars);
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(smem + kOffBars + kNumBars * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[W_FULL], 1);
        mbar_init(&bars[W2_TMEM], 1);
        mbar_init(&bars[X1_FULL], 128);
        mbar_init(&bars[X1_FREE], 1);
        mbar_init(&bars[D2A_FULL], 1);
        mbar_init(&bars[D2B_FULL], 1);
        mbar_init(&bars[D2A_FREE], 256);  // both epilogue warpgroups (one tile of the pair each)
        mbar_init(&bars[D2B_FREE], 256);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars[X2_FULL0 + b], 128);
            mbar_init(&bars[X2_FREE0 + b], 1);
            mbar_init(&bars[D1_FULL0 + b], 1);
        }
        mbar_init(&bars[D1_FREE], 128);
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

    // tiles of this CTA: cluster c = blockIdx.x + t * gridDim.x
    const long long first = blockIdx.x;
    const int T = first < num_clusters ? static_cast<int>((num_clusters - first + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 0) {
        {
            // ---- weights: one bulk-TMA burst, resident for the whole kernel ---------------------------------
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[W_FULL], kWeightBytes);
                // few large copies: the issue of one cp.async.bulk costs this thread ~450 cycles (tools/wgrad_tc_trace.py)
                for (uint32_t off = 0; off < kImgBulk; off += 65536) bulk_g2s(smem + off, wimg + off, min(65536u, kImgBulk - off), &bars[W_FULL]);
                bulk_g2s(smem + kOffW0, wimg + kImgSmall, kSmallBytes, &bars[W_FULL]);
            }
            __syncwarp();
            mbar_wait(&bars[W_FULL], 0);
            // ---- MMA issue loop: the whole warp walks it (uniform control flow), one elected lane issues --------
            const uint32_t idesc = make_idesc(1, 128, kSamples);
            const uint32_t sbase = smem_u32(smem);
            // ---- W2 hi (both M blocks) and W1 -> tensor memory, once; their staging area then becomes the X2 ring
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        tmem_cp_128x256b(tmem_base + kTmemW2 + mb * 64 + k * 8,
                                         make_smem_desc(sbase + kOffW2hi + mb * kW2Blk + k * 2 * kLboW, kLboW, kSbo));
#pragma unroll
                for (int sp = 0; sp < 2; ++sp)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tmem_cp_128x256b(tmem_base + kTmemW1 + sp * 32 + k * 8,
                                         make_smem_desc(sbase + kOffW1 + sp * kW1Split + k * 2 * kLboW, kLboW, kSbo));
                umma_commit(&bars[W2_TMEM]);
            }
            __syncwarp();
            const uint32_t idesc128 = make_idesc(1, 128, 2 * kSamples) | kIdescBMnMajor;
            const uint32_t idesc64 = idesc | kIdescBMnMajor;
            auto mma2 = [&](int pr) {  // pair pr = tiles 2pr, 2pr+1: B operand = X2 slot pr & 1 (both tiles side by side)
                const bool two = 2 * pr + 1 < T;
                mbar_wait(&bars[X2_FULL0], pr & 1);
                if (two) mbar_wait(&bars[X2_FULL1], pr & 1);
                tcgen05_fence_after();
                stamp(2 * pr, 1);
                const uint32_t id = two ? idesc128 : idesc64;
                const uint32_t xb0 = sbase + kOffX2 + (pr & 1) * kX2Slot;
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    mbar_wait(&bars[D2A_FREE + mb], (pr & 1) ^ 1);  // E2(pr-1) has moved this M block into registers
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem_base + kTmemD2 + mb * 128;
                        const uint32_t wa = tmem_base + kTmemW2 + mb * 64;             // W2 hi, tensor memory
                        const uint32_t wl = sbase + kOffW2lo + mb * kW2Blk;            // W2 lo, shared memory
                        uint32_t acc = 0;
#pragma unroll
                        for (int pass = 0; pass < 2; ++pass) {  // (Whi,Xhi) (Whi,Xlo)
                            const uint32_t xb = xb0 + (pass == 1 ? kX2Split : 0);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                umma_f16_ts(d, wa + k * 8, make_smem_desc(xb + k * 2 * kLboX2, kLboX2, kSboX2), id, acc);
                                acc = 1;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // (Wlo,Xhi): both operands from shared memory
                            umma_f16(d, make_smem_desc(wl + k * 2 * kLboW, kLboW, kSbo), make_smem_desc(xb0 + k * 2 * kLboX2, kLboX2, kSboX2), id, 1u);
                        if (mb == 1) umma_commit(&bars[X2_FREE0 + (pr & 1)]);
                        umma_commit(&bars[D2A_FULL + mb]);
                    }
                    __syncwarp();
                }
                stamp(2 * pr, 2);
            };
            // MMA1 (conv1) is issued by warp 13: two issuing warps hide each other's mbarrier-wait latency (~90 cycles per
            // try_wait, 4 per tile), which a single issuer left as bubbles in the tensor pipe.  Data dependencies are carried by
            // the barriers alone, so the interleaving of the two instruction streams in the pipe is free.
            for (int pr = 0; 2 * pr < T; ++pr) mma2(pr);
        }
    } else if (warp <= 4) {
        // ---- producers: gather + normalise + layer 0 -> X1 ------------------------------------------------------
        // Software-pipelined: the index of tile t+2 and the coordinates of tile t+1 are in flight while tile t is computed.
        mbar_wait(&bars[W_FULL], 0);  // W0 / b0 live in the weight image
        const float4 *W0 = reinterpret_cast<const float4 *>(smem + kOffW0);  // per channel: (w_x, w_y, w_z, bias)
        const int pt = threadIdx.x - 32;      // 0..127
        const int s = pt & 63, h = pt >> 6;   // sample, channel half
        uint8_t *x1 = smem + kOffX1 + s * 16;
        const unsigned stride = gridDim.x;
        // grouped_xyz /= radius (pointnet_common.py:47).  When radius is a power of two (the model's 2.0) multiplying by its
        // reciprocal is bit-identical to the division and avoids the IEEE-division slow path (taken whenever a lane's
        // numerator is 0, i.e. for the centre point of every cluster); otherwise the true division is kept.
        const float inv_r = 1.0f / radius;
        const bool pow2 = (__float_as_uint(radius) & 0x007fffffu) == 0u && radius > 1e-30f && radius < 1e30f;
        auto load_idx = [&](int t) -> int {
            if (t >= T) return 0;
            const unsigned cl = static_cast<unsigned>(first) + static_cast<unsigned>(t) * stride;
            return __ldg(idx + static_cast<size_t>(cl) * kSamples + s);
        };
        float px = 0.f, py = 0.f, pz = 0.f, qx = 0.f, qy = 0.f, qz = 0.f;
        auto load_xyz = [&](int t, int ii) {
            if (t >= T) return;
            const unsigned cl = static_cast<unsigned>(first) + static_cast<unsigned>(t) * stride;
            ii = min(max(ii, 0), n - 1);
            const float *p = xyz + (static_cast<size_t>(cl / static_cast<unsigned>(m)) * n + ii) * 3;
            const float *c = new_xyz + static_cast<size_t>(cl) * 3;
            px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
            qx = __ldg(c); qy = __ldg(c + 1); qz = __ldg(c + 2);
        };
        int i1 = load_idx(0);
        load_xyz(0, i1);
        i1 = load_idx(1);
        for (int t = 0; t < T; ++t) {
            if (warp == 1) stamp(t, 4);
            const float gx = pow2 ? (px - qx) * inv_r : (px - qx) / radius;
            const float gy = pow2 ? (py - qy) * inv_r : (py - qy) / radius;
            const float gz = pow2 ? (pz - qz) * inv_r : (pz - qz) / radius;
            const int i2 = load_idx(t + 2);
            load_xyz(t + 1, i1);
            i1 = i2;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float4 w = W0[h * 32 + j * 2 + e];
                    float a = w.w;
                    a = fmaf(gx, w.x, a);
                    a = fmaf(gy, w.y, a);
                    a = fmaf(gz, w.z, a);
                    v[e] = fmaxf(a, 0.0f);
                }
                split_bf16x2(v[0], v[1], hi[j], lo[j]);
            }
            if (warp == 1) stamp(t, 5);
            mbar_wait(&bars[X1_FREE], (t & 1) ^ 1);
            if (warp == 1) stamp(t, 6);
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // K chunk h*4+q holds channels h*32 + q*8 .. +7
                *reinterpret_cast<uint4 *>(x1 + (h * 4 + q) * kLboX1) = make_uint4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
                *reinterpret_cast<uint4 *>(x1 + kX1Split + (h * 4 + q) * kLboX1) =
                    make_uint4(lo[q * 4], lo[q * 4 + 1], lo[q * 4 + 2], lo[q * 4 + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[X1_FULL]);
        }
    } else if (warp == 13) {
        // ---- second MMA issuer: conv1 (X1 -> D1), A operand W1 from tensor memory ---------------------------------------
        const uint32_t idesc = make_idesc(1, 128, kSamples);
        const uint32_t sbase = smem_u32(smem);
        mbar_wait(&bars[W2_TMEM], 0);  // W1 / W2 have been copied into tensor memory
        auto mma1 = [&](int t) {
            mbar_wait(&bars[X1_FULL], t & 1);
            mbar_wait(&bars[D1_FREE], (t & 1) ^ 1);  // E1(t-1) has moved the single D1 accumulator into registers
            tcgen05_fence_after();
            stamp(t, 0);
            if (elect_one()) {
            const uint32_t d = tmem_base;
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {  // (Whi,Xhi) (Whi,Xlo) (Wlo,Xhi)
                const uint32_t wa = tmem_base + kTmemW1 + (pass == 2 ? 32 : 0);
                const uint32_t xb = sbase + kOffX1 + (pass == 1 ? kX1Split : 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16_ts(d, wa + k * 8, make_smem_desc(xb + k * 2 * kLboX1, kLboX1, kSbo), idesc, acc);
                    acc = 1;
                }
            }
            umma_commit(&bars[X1_FREE]);
            umma_commit(&bars[D1_FULL0 + (t & 1)]);
            }
            __syncwarp();
        };
        for (int t = 0; t < T; ++t) mma1(t);
    } else {
        // ---- epilogue warpgroups: g = 0 (warps 5-8) even tiles, g = 1 (warps 9-12) odd tiles ----------------------
        mbar_wait(&bars[W_FULL], 0);  // biases live in the weight image
        const int g = (warp - 5) >> 2;
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int ch = q * 32 + lane;        // output channel (TMEM lane) owned by this thread
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const float b1 = reinterpret_cast<const float *>(smem + kOffB1)[ch];
        const float *B2 = reinterpret_cast<const float *>(smem + kOffB2);
        // this warpgroup's tiles own sample groups g*8 .. g*8+7 of their pair's slot
        uint8_t *x2g = smem + kOffX2 + g * 8 * kSboX2 + (ch >> 3) * kLboX2 + (ch & 7) * 16;
        mbar_wait(&bars[W2_TMEM], 0);  // the X2 buffers alias the W2 staging area: wait until W2 sits in tensor memory
        uint32_t r0[32], r1[32];
        // E1(t): D1 -> +bias, ReLU, split -> this tile's half of X2 slot (pair & 1)
        auto e1 = [&](int t) {
            const int b = t & 1;
            const int pr = t >> 1;               // pair
            const uint32_t ph = pr & 1;          // parity of per-pair events = ring slot of the pair
            uint8_t *x2 = x2g + ph * kX2Slot;
            mbar_wait(&bars[D1_FULL0 + b], ph);
            tcgen05_fence_after();
            if (q == 1) stamp(t, 8);
            tmem_ld32(tmem_base + lane_addr, r0);
            tmem_ld32(tmem_base + lane_addr + 32, r1);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&bars[D1_FREE]);
#pragma unroll
            for (int sidx = 0; sidx < 64; sidx += 2) {  // r[sidx] becomes the hi pair (samples sidx, sidx+1), r[sidx+1] the lo pair
                uint32_t &ra = sidx < 32 ? r0[sidx & 31] : r1[sidx & 31];
                uint32_t &rb = sidx < 32 ? r0[(sidx + 1) & 31] : r1[(sidx + 1) & 31];
                const float va = fmaxf(__uint_as_float(ra) + b1, 0.0f), vb = fmaxf(__uint_as_float(rb) + b1, 0.0f);
                split_bf16x2(va, vb, ra, rb);
            }
            if (q == 1) stamp(t, 9);
            mbar_wait(&bars[X2_FREE0 + ph], ((pr >> 1) & 1) ^ 1);  // MMA2(pr-2) has finished reading this slot
            if (q == 1) stamp(t, 10);
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // samples 8j .. 8j+7 of this channel: 16 contiguous bytes per split
                const uint32_t *r = j < 4 ? &r0[(j & 3) * 8] : &r1[(j & 3) * 8];
                *reinterpret_cast<uint4 *>(x2 + j * kSboX2) = make_uint4(r[0], r[2], r[4], r[6]);
                *reinterpret_cast<uint4 *>(x2 + kX2Split + j * kSboX2) = make_uint4(r[1], r[3], r[5], r[7]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[X2_FULL0 + b]);
            if (q == 1) stamp(t, 11);
        };
        // E2(t): D2 -> max over the 64 samples, +bias, ReLU -> pooled (ReLU and +bias commute with max); M block A is released to
        // MMA2(pair+1) before block B has even been computed
        auto e2 = [&](int t) {
            const uint32_t ph = (t >> 1) & 1;
            const long long cl = first + static_cast<long long>(t) * gridDim.x;
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                mbar_wait(&bars[D2A_FULL + mb], ph);
                tcgen05_fence_after();
                if (q == 1 && mb == 0) stamp(t, 12);
                tmem_ld32(tmem_base + lane_addr + kTmemD2 + mb * 128 + g * 64, r0);   // this tile's 64 of the pair's 128 columns
                tmem_ld32(tmem_base + lane_addr + kTmemD2 + mb * 128 + g * 64 + 32, r1);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(&bars[D2A_FREE + mb]);
                float mv = __uint_as_float(r0[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(r0[j]));
#pragma unroll
                for (int j = 0; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(r1[j]));
                pooled[cl * 256 + mb * 128 + ch] = fmaxf(mv + B2[mb * 128 + ch], 0.0f);
            }
            if (q == 1) stamp(t, 13);
        };
        // Both warpgroups work on the SAME pair (one tile each), so the operand of the NEXT pair has to be written before this
        // pair's accumulator is drained: E1(t+2) runs under MMA2(pair of t), then E2(t) waits for that MMA2 to finish.
        if (g < T) e1(g);
        for (int t = g; t < T; t += 2) {
            if (t + 2 < T) e1(t + 2);
            e2(t);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// Builds the kernel's shared-memory weight image from the packed fp32 (BN-folded) weights: bf16 hi/lo splits of
// W1^T [128 x 64] and W2^T [256 x 128] in the canonical K-major core-matrix order, then W0, b0, b1, b2 in fp32.
__global__ void det_tc_prep_kernel(const float *__restrict__ P, WeightLayout L, uint8_t *__restrict__ wimg) {
    using namespace det;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 128 * 64) {  // W1: A[r][k] = W[k][r]
        const int r = i & 127, k = i >> 7;
        const float w = P[L.off[W_DET1] + k * 128 + r];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const uint32_t o = (k >> 3) * kLboW + r * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(wimg + kOffW1 + o) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(wimg + kOffW1 + kW1Split + o) = lo;
    } else if (i < 128 * 64 + 256 * 128) {  // W2: A[r][k] = W[k][r], r = mb*128 + rr
        const int e = i - 128 * 64;
        const int r = e & 255, k = e >> 8;
        const float w = P[L.off[W_DET2] + k * 256 + r];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        const int mb = r >> 7, rr = r & 127;
        const uint32_t o = (k >> 3) * kLboW + rr * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(wimg + kOffW2hi + mb * kW2Blk + o) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(wimg + kOffW2lo + mb * kW2Blk + o) = lo;
    } else {
        const int e = i - 128 * 64 - 256 * 128;
        float *f = reinterpret_cast<float *>(wimg + kImgSmall);
        if (e < 256) {  // per channel k: (W0[0][k], W0[1][k], W0[2][k], b0[k])
            const int k = e >> 2, c = e & 3;
            f[e] = c < 3 ? P[L.off[W_DET0] + c * 64 + k] : P[L.off[B_DET0] + k];
        } else if (e < 384) f[e] = P[L.off[B_DET1] + e - 256];
        else if (e < 640) f[e] = P[L.off[B_DET2] + e - 384];
    }
}

static long long *g_det_dbg = nullptr;  // set through f3d_debug_set_timeline (bring-up only)

int detector_rows_tc(long long num_clusters, int n, int m, float radius, const float *xyz, const float *new_xyz,
                     const int *idx, const float *packed, uint8_t *wimg, float *pooled, bool build_image, int max_ctas, cudaStream_t st) {
    if (num_clusters == 0) return 0;
    if (build_image) {
        const int total = 128 * 64 + 256 * 128 + 640;
        det_tc_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(packed, make_weight_layout(32), wimg);
        const int rc = check_launch("det_tc_prep_kernel");
        if (rc) return rc;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    cudaError_t e = cudaFuncSetAttribute(det_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(det::kSmemBytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "det_rows_tc: cudaFuncSetAttribute");
    const int ctas = (max_ctas > 0 && max_ctas < num_sms) ? max_ctas : num_sms;  // persistent: one CTA per SM the caller grants
    const unsigned grid = static_cast<unsigned>(num_clusters < ctas ? num_clusters : ctas);
    // algorithmic flops: 2 * rows * (3*64 + 64*128 + 128*256), rows = 64 samples per cluster
    ktimer_begin("det_rows_tc_kernel", 2.0 * 41152.0 * 64.0 * static_cast<double>(num_clusters), st);
    det_rows_tc_kernel<<<grid, det::kThreads, det::kSmemBytes, st>>>(num_clusters, n, m, radius, xyz, new_xyz, idx, wimg, pooled,
                                                                     g_det_dbg);
    ktimer_end(st);
    return check_launch("det_rows_tc_kernel");
}

}  // namespace f3d

using namespace f3d;

// Debug / bring-up entry point (not part of the reference surface): runs the single-CTA UMMA self test.
F3D_API int f3d_debug_umma_selftest(const void *a_img, const void *b_img, float *D, int N, int K, int lbo_a, int sbo_a, int lbo_b,
                                    int sbo_b, int a_bytes, int b_bytes, int a_in_tmem, void *stream) {
    if (!a_img || !b_img || !D || N < 8 || N > 256 || N % 8 || K < 16 || K % 16) return fail(F3D_ERR_INVALID_ARGUMENT, "umma_selftest: bad arguments");
    const size_t smem = ((static_cast<size_t>(a_bytes) + 127) & ~static_cast<size_t>(127)) + b_bytes + 128;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(static_cast<const uint8_t *>(a_img), static_cast<const uint8_t *>(b_img), D, N, K,
                                                             lbo_a, sbo_a, lbo_b, sbo_b, a_bytes, b_bytes, a_in_tmem);
    return check_launch("umma_selftest_kernel");
}

// Bring-up: device buffer of (tiles per CTA) x 16 int64 that receives CTA 0's clock64() timeline (NULL disables).
namespace f3d { extern long long *g_desc_dbg; }
F3D_API void f3d_debug_set_timeline(void *buf) { f3d::g_det_dbg = static_cast<long long *>(buf); }
F3D_API void f3d_debug_set_timeline_desc(void *buf) { f3d::g_desc_dbg = static_cast<long long *>(buf); }

F3D_API size_t f3d_detector_tc_weight_bytes(void) { return det::kWeightBytes; }
