// mlp_tc_desc.cu -- fused descriptor forward (per-row part) on the Blackwell tensor cores, precision "bf16x3".
//
// Replaces the TensorFlow graph of pointnet_sa_module (models/feat3dnet.py:54-75) on top of sample_and_group
// (models/pointnet_common.py:104-120): group, translate, /radius, rotate by the detector's orientation,
// conv 3->32->64 (+BN+ReLU), reduce_max, tile+concat, conv_mid_0 128->128 (+BN, no ReLU), reduce_max.
// Same machinery and the same D^T = W^T X^T formulation as mlp_tc.cu (TMEM lane = output channel, column = sample).
//
//   MMA1 : D1[128 x 64] = W1^T[128 x 32]  X1^T                   (conv1, 32 -> 64, per tile; rows 64..127 of the M axis REPEAT
//                                                                   rows 0..63, so all four epilogue warps of a warpgroup find
//                                                                   the 64 channels in their own 32 TMEM lanes)
//   E1   : relu(D1 + b1) -> X2 (operand of conv_mid_0, stored SAMPLE-contiguous = MN-major: one 16-byte store per 8 samples)
//          and, per channel, its max over the 64 samples -> P; warps 0/1 take samples 0..31, warps 2/3 samples 32..63
//   MMA2 : D2[128 x 128] = Wa^T[128 x 64]  X2^T                   (conv_mid_0, rows 0..63 of its weight: the per-point part) --
//                                                                   issued per PAIR of tiles: a 128x128x16 MMA holds the tensor
//                                                                   pipe 66.5 cycles, two 128x64x16 ones 2 x 41.9
//                                                                   (profiles/r02_a_umma_instruction_shape.md)
//   MMA3 : D3[128 x 8]   = Wb^T[128 x 64]  P^T                    (rows 64..127: the tiled max-pool part, ONE column per
//                                                                   cluster instead of 64 -- the split-weight identity; the two
//                                                                   tiles of a pair are rows 0 / 1 of the same 8-row operand)
//   E2   : pooled2 = max_s D2[:, s] + D3[:, tile] + b_mid         (no ReLU: final_relu=False, feat3dnet.py:71)
// D1 and D2 are double-buffered; D3 is single (its reader, E2 of the previous pair, is awaited before MMA3).
//
// 18 warps: 0 = MMA2/3 issue, 1-4 and 14-17 = two producer warpgroups (even / odd tiles, one X1 buffer each: with the pair MMAs the
// tensor pipe needs ~860 cycles per tile and one producer warpgroup ~1060), 5-12 = two epilogue warpgroups (even / odd tiles), 13 = MMA1
// issue.  The epilogue warpgroups work on the same pair, so each writes the operand of its NEXT tile before draining the current
// accumulator (see mlp_tc.cu).
// conv_post_0 + l2-normalise run afterwards over 128 clusters per CTA (desc_post_fp32_kernel).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "weights_layout.h"

#include <cuda_bf16.h>

namespace f3d {

using namespace tc;

namespace dsc {
constexpr int kSamples = 64;
constexpr int kThreads = 18 * 32;  // warp 0 MMA2/3 issue, 1-4 / 14-17 producers (even / odd tiles), 5-12 epilogues, 13 MMA1 issue
constexpr uint32_t kSbo = 128;
constexpr uint32_t kLboW = 128 * 16;
constexpr uint32_t kLboX1 = kSamples * 16;
constexpr uint32_t kLboX2 = 128;                        // X2 is MN-major: K groups of 8 channels 128 B apart,
constexpr uint32_t kSboX2 = 8 * 128;                    // groups of 8 samples 1 KB apart (64 channels)
constexpr uint32_t kLboP = 8 * 16;                      // P: one 8-row group per K chunk
constexpr uint32_t kW1Split = 128 * 32 * 2;             // 8 KB
constexpr uint32_t kWmSplit = 128 * 64 * 2;             // 16 KB
constexpr uint32_t kOffW1 = 0;                           // [split 2][chunk 4][row 128][8]
constexpr uint32_t kOffWa = kOffW1 + 2 * kW1Split;       // [split 2][chunk 8][row 128][8]
constexpr uint32_t kOffWb = kOffWa + 2 * kWmSplit;
constexpr uint32_t kOffW0 = kOffWb + 2 * kWmSplit;       // float4 [32]: (w_x, w_y, w_z, bias) of layer 0 per channel
constexpr uint32_t kOffB0 = kOffW0 + 3 * 32 * 4;         // (tail of the float4 table)
constexpr uint32_t kOffB1 = kOffB0 + 32 * 4;             // fp32 [64]
constexpr uint32_t kOffBm = kOffB1 + 64 * 4;             // fp32 [128]
constexpr uint32_t kWeightBytes = kOffBm + 128 * 4;      // 83 200
constexpr uint32_t kX1Split = 4 * kLboX1;                // 4 KB
constexpr uint32_t kX1Buf = 2 * kX1Split;                // one X1 operand (hi + lo); buffer = tile & 1 (one per producer warpgroup)
constexpr uint32_t kOffX1 = kWeightBytes;
// X2 ring: slot = pair & 1; one slot = [split 2][16 groups of 8 samples][64 channels x 16 B]; tile (pair*2 + g) owns sample groups
// g*8 .. g*8+7, so the N = 128 operand of a pair is one uniform-stride block
constexpr uint32_t kX2Split = 16 * kSboX2;               // 16 KB
constexpr uint32_t kX2Slot = 2 * kX2Split;               // 32 KB
constexpr uint32_t kOffX2 = kOffX1 + 2 * kX1Buf;
constexpr uint32_t kPSplit = 8 * kLboP;                  // 1 KB
constexpr uint32_t kPBuf = 2 * kPSplit;                  // per pair slot; row g of the 8-row group = tile g of the pair
constexpr uint32_t kOffP = kOffX2 + 2 * kX2Slot;         // [slot 2][split 2][chunk 8][row 8][8]
constexpr uint32_t kOffPm = kOffP + 2 * kPBuf;            // fp32 [warpgroup 2][128]: partial channel maxima of the sample halves
constexpr uint32_t kOffBars = kOffPm + 2 * 128 * 4;
constexpr uint32_t kSmemBytes = kOffBars + 20 * 8 + 16;
static_assert(kWeightBytes % 16 == 0 && kOffX1 % 128 == 0 && kOffX2 % 128 == 0 && kOffP % 128 == 0 && kOffBars % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
// TMEM columns: D1[tile & 1] at 0 / 64, D2[slot] at 128 + slot*128 (two tiles x 64 samples), D3 at 384 (8 used), then the weights
// that feed MMAs from tensor memory (copied once with tcgen05.cp): W1 at 416 + split*16 (32 K = 16 columns), Wa at 448 + split*32
// (64 K = 32 columns).  Wb (the N = 8 MMAs of the pooled term) stays in shared memory: an SS-mode 128x8x16 instruction reads its
// 4 KB A operand in ~33 cycles, the time such an instruction holds the pipe anyway (34 cycles), and the 64 columns it would take
// are what lets BOTH D1 and D2 be double-buffered.
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemD2 = 128, kTmemD3 = 384;
constexpr uint32_t kTmemW1 = 416, kTmemWa = 448;
// per-tile events the alternating epilogue warpgroups wait on are per-warpgroup barriers; per-pair events are per ring slot
enum Bar { W_FULL = 0, W_TMEM, X1_FULL0, X1_FULL1, X1_FREE0, X1_FREE1, X2_FULL0, X2_FULL1, X2_FREE0, X2_FREE1, D1_FULL0, D1_FULL1, D1_FREE0,
           D1_FREE1, D2_FULL0, D2_FULL1, D2_FREE0, D2_FREE1, kNumBars };
static_assert(kNumBars <= 20, "barrier area");
}  // namespace dsc

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

__global__ void __launch_bounds__(dsc::kThreads, 1)
desc_rows_tc_kernel(long long num_clusters, int n, int m, float radius, const float *__restrict__ xyz,
                    const float *__restrict__ new_xyz, const int *__restrict__ idx, const float *__restrict__ orientation,
                    const uint8_t *__restrict__ wimg, float *__restrict__ pooled2, long long *__restrict__ dbg) {
    using namespace dsc;
    auto stamp = [&](int t, int slot) {  // optional clock64() timeline of CTA 0 (bring-up)
        if (dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) dbg[t * 16 + slot] = clock64();
    };
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kOffBars);
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(smem + kOffBars + kNumBars * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(&bars[W_FULL], 1);
        mbar_init(&bars[W_TMEM], 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars[D1_FREE0 + b], 128);
            mbar_init(&bars[X1_FULL0 + b], 128);
            mbar_init(&bars[X1_FREE0 + b], 1);
            mbar_init(&bars[X2_FULL0 + b], 128);
            mbar_init(&bars[X2_FREE0 + b], 1);
            mbar_init(&bars[D1_FULL0 + b], 1);
            mbar_init(&bars[D2_FULL0 + b], 1);
            mbar_init(&bars[D2_FREE0 + b], 256);  // both epilogue warpgroups (one tile of the pair each)
        }
        fence_barrier_init();
    }
    // rows 2..7 of the pooled operand P are never written: zero the operand once (their D3 columns are never read either)
    for (uint32_t i = threadIdx.x; i < 2 * kPBuf / 4; i += kThreads) reinterpret_cast<uint32_t *>(smem + kOffP)[i] = 0;
    fence_proxy_async_smem();
    if (warp == 0) {
        tmem_alloc(tmem_base_s, kTmemCols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    const long long first = blockIdx.x;
    const int T = first < num_clusters ? static_cast<int>((num_clusters - first + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 0) {
        {
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[W_FULL], kWeightBytes);
                for (uint32_t off = 0; off < kWeightBytes; off += 65536) {  // few large copies (each issue costs this thread ~450 cycles)
                    const uint32_t sz = min(65536u, kWeightBytes - off);
                    bulk_g2s(smem + off, wimg + off, sz, &bars[W_FULL]);
                }
            }
            __syncwarp();
            mbar_wait(&bars[W_FULL], 0);
            const uint32_t idesc64 = make_idesc(1, 128, kSamples) | kIdescBMnMajor;
            const uint32_t idesc128 = make_idesc(1, 128, 2 * kSamples) | kIdescBMnMajor;
            const uint32_t idesc8 = make_idesc(1, 128, 8);
            const uint32_t sbase = smem_u32(smem);
            // ---- W1 and Wa (hi and lo splits) -> tensor memory, once; Wb feeds its MMAs from shared memory
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int sp = 0; sp < 2; ++sp) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tmem_cp_128x256b(tmem_base + kTmemW1 + sp * 16 + k * 8,
                                         make_smem_desc(sbase + kOffW1 + sp * kW1Split + k * 2 * kLboW, kLboW, kSbo));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tmem_cp_128x256b(tmem_base + kTmemWa + sp * 32 + k * 8,
                                         make_smem_desc(sbase + kOffWa + sp * kWmSplit + k * 2 * kLboW, kLboW, kSbo));
                }
                umma_commit(&bars[W_TMEM]);
            }
            __syncwarp();
            mbar_wait(&bars[W_TMEM], 0);
            auto mma23 = [&](int pr) {  // pair pr = tiles 2pr, 2pr+1: X2 / P slot pr & 1, accumulator D2[pr & 1], the single D3
                const int sl = pr & 1;
                const bool two = 2 * pr + 1 < T;
                mbar_wait(&bars[X2_FULL0], pr & 1);
                if (two) mbar_wait(&bars[X2_FULL1], pr & 1);
                mbar_wait(&bars[D2_FREE0 + sl], ((pr >> 1) & 1) ^ 1);  // E2(pr-2) has drained this accumulator
                tcgen05_fence_after();
                stamp(2 * pr, 1);
                if (elect_one()) {
                    const uint32_t d2 = tmem_base + kTmemD2 + sl * 128;
                    const uint32_t id = two ? idesc128 : idesc64;
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t wa = tmem_base + kTmemWa + (pass == 2 ? 32 : 0);
                        const uint32_t xb = sbase + kOffX2 + sl * kX2Slot + (pass == 1 ? kX2Split : 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16_ts(d2, wa + k * 8, make_smem_desc(xb + k * 2 * kLboX2, kLboX2, kSboX2), id, acc);
                            acc = 1;
                        }
                    }
                }
                __syncwarp();
                // D3 is single-buffered: E2(pr-1), which reads it, arrives on the OTHER slot's D2_FREE
                if (pr > 0) mbar_wait(&bars[D2_FREE0 + (sl ^ 1)], ((pr - 1) >> 1) & 1);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t d3 = tmem_base + kTmemD3;
                    uint32_t acc = 0;
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t wb = sbase + kOffWb + (pass == 2 ? kWmSplit : 0);
                        const uint32_t xb = sbase + kOffP + sl * kPBuf + (pass == 1 ? kPSplit : 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16(d3, make_smem_desc(wb + k * 2 * kLboW, kLboW, kSbo), make_smem_desc(xb + k * 2 * kLboP, kLboP, kSbo), idesc8, acc);
                            acc = 1;
                        }
                    }
                    umma_commit(&bars[X2_FREE0 + sl]);
                    umma_commit(&bars[D2_FULL0 + sl]);
                }
                __syncwarp();
                stamp(2 * pr, 2);
            };
            // MMA1 (conv1) is issued by warp 13: two issuing warps hide each other's mbarrier waits (see mlp_tc.cu)
            for (int pr = 0; 2 * pr < T; ++pr) mma23(pr);
        }
    } else if (warp == 13) {
        // ---- second MMA issuer: conv1 (X1[t & 1] -> D1), A operand W1 from tensor memory ---------------------------------
        const uint32_t idesc64 = make_idesc(1, 128, kSamples);
        const uint32_t sbase = smem_u32(smem);
        mbar_wait(&bars[W_TMEM], 0);  // the weights have been copied into tensor memory
        tcgen05_fence_after();
        auto mma1 = [&](int t) {
            mbar_wait(&bars[X1_FULL0 + (t & 1)], (t >> 1) & 1);
            mbar_wait(&bars[D1_FREE0 + (t & 1)], ((t >> 1) & 1) ^ 1);  // E1(t-2) has moved this D1 buffer into registers
            tcgen05_fence_after();
            stamp(t, 0);
            if (elect_one()) {
                const uint32_t d = tmem_base + (t & 1) * 64;
                uint32_t acc = 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t wa = tmem_base + kTmemW1 + (pass == 2 ? 16 : 0);
                    const uint32_t xb = sbase + kOffX1 + (t & 1) * kX1Buf + (pass == 1 ? kX1Split : 0);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        umma_f16_ts(d, wa + k * 8, make_smem_desc(xb + k * 2 * kLboX1, kLboX1, kSbo), idesc64, acc);
                        acc = 1;
                    }
                }
                umma_commit(&bars[X1_FREE0 + (t & 1)]);
                umma_commit(&bars[D1_FULL0 + (t & 1)]);
            }
            __syncwarp();
        };
        for (int t = 0; t < T; ++t) mma1(t);
    } else if (warp <= 4 || warp >= 14) {
        // ---- producers: gather + normalise + rotate + layer 0 (3 -> 32) -> X1 ------------------------------------
        // Two warpgroups: pg = 0 (warps 1-4) makes the even tiles, pg = 1 (warps 14-17) the odd ones, each into its own X1 buffer.
        // Software-pipelined over the warpgroup's own tile sequence: index of item i+2D and coordinates / orientation of item i+D
        // in flight while item i is computed.
        mbar_wait(&bars[W_FULL], 0);
        const float4 *W0 = reinterpret_cast<const float4 *>(smem + kOffW0);  // per channel: (w_x, w_y, w_z, bias)
        const int pg = warp >= 14 ? 1 : 0;
        const int pt = pg ? threadIdx.x - 14 * 32 : threadIdx.x - 32;
        const int s = pt & 63, h = pt >> 6;  // sample, channel half (16 channels = 2 K chunks)
        uint8_t *x1 = smem + kOffX1 + pg * kX1Buf + s * 16;
        const unsigned stride = gridDim.x;
        const int Tg = T > pg ? (T - pg + 1) / 2 : 0;  // items (tiles) of this warpgroup: tile = pg + 2 i
        const float inv_r = 1.0f / radius;  // exact replacement of the division when radius is a power of two (see mlp_tc.cu)
        const bool pow2 = (__float_as_uint(radius) & 0x007fffffu) == 0u && radius > 1e-30f && radius < 1e30f;
        auto cluster_of = [&](int i) -> unsigned { return static_cast<unsigned>(first) + static_cast<unsigned>(pg + 2 * i) * stride; };
        auto load_idx = [&](int i) -> int {
            if (i >= Tg) return 0;
            return __ldg(idx + static_cast<size_t>(cluster_of(i)) * kSamples + s);
        };
        // With depth 1 the kernel was bound by this dependent L2 gather (clock64 timeline: ~2100 cycles per producer iteration
        // against ~1000 of MMA + epilogue work); the slots are static (loop unrolled by D).
        constexpr int D = 3;
        struct Grp { float px, py, pz, qx, qy, qz, th; };
        Grp gq[D];
        int iq[D];
        auto load_xyz = [&](Grp &g, int i, int ii) {
            g.px = g.py = g.pz = g.qx = g.qy = g.qz = g.th = 0.f;
            if (i >= Tg) return;
            const unsigned cl = cluster_of(i);
            ii = min(max(ii, 0), n - 1);
            const float *p = xyz + (static_cast<size_t>(cl / static_cast<unsigned>(m)) * n + ii) * 3;
            const float *c = new_xyz + static_cast<size_t>(cl) * 3;
            g.px = __ldg(p); g.py = __ldg(p + 1); g.pz = __ldg(p + 2);
            g.qx = __ldg(c); g.qy = __ldg(c + 1); g.qz = __ldg(c + 2);
            if (orientation) g.th = __ldg(orientation + cl);
        };
#pragma unroll
        for (int d = 0; d < D; ++d) iq[d] = load_idx(d);
#pragma unroll
        for (int d = 0; d < D; ++d) load_xyz(gq[d], d, iq[d]);
#pragma unroll
        for (int d = 0; d < D; ++d) iq[d] = load_idx(D + d);
        for (int i0 = 0; i0 < Tg; i0 += D) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            const int i = i0 + d;
            if (i >= Tg) break;
            const int t = pg + 2 * i;
            if ((warp & 3) == 1) stamp(t, 4);
            const float px = gq[d].px, py = gq[d].py, pz = gq[d].pz, qx = gq[d].qx, qy = gq[d].qy, qz = gq[d].qz, th = gq[d].th;
            float gx = pow2 ? (px - qx) * inv_r : (px - qx) / radius;
            float gy = pow2 ? (py - qy) * inv_r : (py - qy) / radius;
            const float gz = pow2 ? (pz - qz) * inv_r : (pz - qz) / radius;
            if (orientation) {  // pointnet_common.py:110-120: x' = x c - y s ; y' = x s + y c
                float cs, sn;
                sincosf(th, &sn, &cs);
                const float xr = gx * cs - gy * sn;
                const float yr = gx * sn + gy * cs;
                gx = xr;
                gy = yr;
            }
            load_xyz(gq[d], i + D, iq[d]);
            iq[d] = load_idx(i + 2 * D);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float4 w = W0[h * 16 + j * 2 + e];
                    float a = w.w;
                    a = fmaf(gx, w.x, a);
                    a = fmaf(gy, w.y, a);
                    a = fmaf(gz, w.z, a);
                    v[e] = fmaxf(a, 0.0f);
                }
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[0], v[1]);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[0] - __low2float(h2), v[1] - __high2float(h2));
                hi[j] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[j] = *reinterpret_cast<const uint32_t *>(&l2);
            }
            if ((warp & 3) == 1) stamp(t, 5);
            mbar_wait(&bars[X1_FREE0 + pg], (i & 1) ^ 1);  // MMA1 of this warpgroup's previous tile has read the buffer
            if ((warp & 3) == 1) stamp(t, 6);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                *reinterpret_cast<uint4 *>(x1 + (h * 2 + q) * kLboX1) = make_uint4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
                *reinterpret_cast<uint4 *>(x1 + kX1Split + (h * 2 + q) * kLboX1) =
                    make_uint4(lo[q * 4], lo[q * 4 + 1], lo[q * 4 + 2], lo[q * 4 + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[X1_FULL0 + pg]);
          }
        }
    } else {
        // ---- epilogue warpgroups: g = 0 (warps 5-8) even tiles, g = 1 (warps 9-12) odd tiles -------------------------
        mbar_wait(&bars[W_FULL], 0);
        const int g = (warp - 5) >> 2;
        const int q = warp & 3;
        const int ch = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int hsel = q >> 1;                 // E1: which half of the 64 samples this warp handles
        const int c1 = (q & 1) * 32 + lane;      // E1: conv1 channel (TMEM lane q*32+lane holds channel (q*32+lane) mod 64)
        const float b1 = reinterpret_cast<const float *>(smem + kOffB1)[c1];
        const float bm = reinterpret_cast<const float *>(smem + kOffBm)[ch];
        // this warpgroup's tiles own sample groups g*8 .. g*8+7 of their pair's slot; 8 consecutive samples of one channel are 16
        // contiguous bytes; its pooled vector is row g of the pair's P operand
        uint8_t *x2g = smem + kOffX2 + (g * 8 + hsel * 4) * kSboX2 + (c1 >> 3) * kLboX2 + (c1 & 7) * 16;
        uint8_t *ppg = smem + kOffP + (c1 >> 3) * kLboP + g * 16 + (c1 & 7) * 2;
        float *pm = reinterpret_cast<float *>(smem + kOffPm) + g * 128;
        uint32_t r0[32], r1[32];
        auto e1 = [&](int t) {
            const int b = t & 1;
            const int pr = t >> 1;
            const uint32_t sl = pr & 1;
            mbar_wait(&bars[D1_FULL0 + b], pr & 1);
            tcgen05_fence_after();
            if (q == 1) stamp(t, 8);
            tmem_ld32(tmem_base + lane_addr + b * 64 + hsel * 32, r0);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&bars[D1_FREE0 + b]);
            float pmax = 0.0f;  // values are post-ReLU (>= 0)
            uint32_t hi[16], lo[16];
            // bias, ReLU, max-pool and the hi/lo split BEFORE waiting for the operand slot
#pragma unroll
            for (int sidx = 0; sidx < 32; sidx += 2) {
                const float va = fmaxf(__uint_as_float(r0[sidx]) + b1, 0.0f), vb = fmaxf(__uint_as_float(r0[sidx + 1]) + b1, 0.0f);
                pmax = fmaxf(pmax, fmaxf(va, vb));
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(va, vb);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(va - __low2float(h2), vb - __high2float(h2));
                hi[sidx >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[sidx >> 1] = *reinterpret_cast<const uint32_t *>(&l2);
            }
            pm[q * 32 + lane] = pmax;
            if (q == 1) stamp(t, 9);
            mbar_wait(&bars[X2_FREE0 + sl], ((pr >> 1) & 1) ^ 1);  // MMA2/3(pr-2) have finished reading this slot
            if (q == 1) stamp(t, 10);
            uint8_t *x2 = x2g + sl * kX2Slot;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4 *>(x2 + j * kSboX2) = make_uint4(hi[j * 4], hi[j * 4 + 1], hi[j * 4 + 2], hi[j * 4 + 3]);
                *reinterpret_cast<uint4 *>(x2 + kX2Split + j * kSboX2) = make_uint4(lo[j * 4], lo[j * 4 + 1], lo[j * 4 + 2], lo[j * 4 + 3]);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");  // both sample halves of every channel maximum are in pm
            if (q < 2) {
                uint8_t *pp = ppg + sl * kPBuf;
                const float full = fmaxf(pmax, pm[(q + 2) * 32 + lane]);
                const __nv_bfloat16 hp = __float2bfloat16_rn(full);
                *reinterpret_cast<__nv_bfloat16 *>(pp) = hp;
                *reinterpret_cast<__nv_bfloat16 *>(pp + kPSplit) = __float2bfloat16_rn(full - __bfloat162float(hp));
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[X2_FULL0 + b]);
            if (q == 1) stamp(t, 11);
        };
        auto e2 = [&](int t) {
            const int pr = t >> 1;
            const uint32_t sl = pr & 1;
            mbar_wait(&bars[D2_FULL0 + sl], (pr >> 1) & 1);
            tcgen05_fence_after();
            if (q == 1) stamp(t, 12);
            tmem_ld32(tmem_base + lane_addr + kTmemD2 + sl * 128 + g * 64, r0);   // this tile's 64 of the pair's 128 columns
            tmem_ld32(tmem_base + lane_addr + kTmemD2 + sl * 128 + g * 64 + 32, r1);
            tmem_ld_wait();
            float mv = __uint_as_float(r0[0]);
#pragma unroll
            for (int j = 1; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(r0[j]));
#pragma unroll
            for (int j = 0; j < 32; ++j) mv = fmaxf(mv, __uint_as_float(r1[j]));
            tmem_ld32(tmem_base + lane_addr + kTmemD3, r0);  // column g: the pooled row of this tile
            tmem_ld_wait();
            const float cterm = __uint_as_float(g ? r0[1] : r0[0]);
            tcgen05_fence_before();
            mbar_arrive(&bars[D2_FREE0 + sl]);
            const long long cl = first + static_cast<long long>(t) * gridDim.x;
            pooled2[cl * 128 + ch] = mv + cterm + bm;
            if (q == 1) stamp(t, 13);
        };
        // both warpgroups work on the same pair: the next tile's operand is written before the current accumulator is drained
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

// shared-memory weight image of the descriptor kernel from the packed fp32 (BN-folded) weights (MID = 128)
__global__ void desc_tc_prep_kernel(const float *__restrict__ P, WeightLayout L, uint8_t *__restrict__ wimg) {
    using namespace dsc;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    auto put = [&](uint32_t base, uint32_t split, uint32_t o, float w) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        *reinterpret_cast<__nv_bfloat16 *>(wimg + base + o) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(wimg + base + split + o) = __float2bfloat16_rn(w - __bfloat162float(hi));
    };
    if (i < 128 * 32) {  // W1^T; rows 64..127 repeat rows 0..63
        const int r = i & 127, k = i >> 7;
        put(kOffW1, kW1Split, (k >> 3) * kLboW + r * 16 + (k & 7) * 2, P[L.off[W_DESC1] + k * 64 + (r & 63)]);
    } else if (i < 128 * 32 + 128 * 64) {  // Wa^T = W_mid[0:64, :]^T
        const int e = i - 128 * 32;
        const int r = e & 127, k = e >> 7;
        put(kOffWa, kWmSplit, (k >> 3) * kLboW + r * 16 + (k & 7) * 2, P[L.off[W_MID] + k * 128 + r]);
    } else if (i < 128 * 32 + 2 * 128 * 64) {  // Wb^T = W_mid[64:128, :]^T
        const int e = i - 128 * 32 - 128 * 64;
        const int r = e & 127, k = e >> 7;
        put(kOffWb, kWmSplit, (k >> 3) * kLboW + r * 16 + (k & 7) * 2, P[L.off[W_MID] + (64 + k) * 128 + r]);
    } else {
        const int e = i - 128 * 32 - 2 * 128 * 64;
        float *f = reinterpret_cast<float *>(wimg + kOffW0);
        if (e < 128) {  // per channel k: (W0[0][k], W0[1][k], W0[2][k], b0[k])
            const int k = e >> 2, c = e & 3;
            f[e] = c < 3 ? P[L.off[W_DESC0] + c * 32 + k] : P[L.off[B_DESC0] + k];
        } else if (e < 192) f[e] = P[L.off[B_DESC1] + e - 128];
        else if (e < 320) f[e] = P[L.off[B_MID] + e - 192];
    }
}

long long *g_desc_dbg = nullptr;  // bring-up timeline buffer (f3d_debug_set_timeline with which = 1)

int descriptor_rows_tc(long long num_clusters, int n, int m, float radius, int feature_dim, const float *xyz,
                       const float *new_xyz, const int *idx, const float *orientation, const float *packed, uint8_t *wimg,
                       float *pooled2, bool build_image, int max_ctas, cudaStream_t st) {
    if (num_clusters == 0) return 0;
    if (build_image) {
        const int total = 128 * 32 + 2 * 128 * 64 + 320;
        desc_tc_prep_kernel<<<(total + 255) / 256, 256, 0, st>>>(packed, make_weight_layout(feature_dim), wimg);
        const int rc = check_launch("desc_tc_prep_kernel");
        if (rc) return rc;
    }
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    cudaError_t e = cudaFuncSetAttribute(desc_rows_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(dsc::kSmemBytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "desc_rows_tc: cudaFuncSetAttribute");
    const int ctas = (max_ctas > 0 && max_ctas < num_sms) ? max_ctas : num_sms;
    const unsigned grid = static_cast<unsigned>(num_clusters < ctas ? num_clusters : ctas);
    // algorithmic flops (split-weight form): 2 * (rows * 10336 + clusters * 8192)
    ktimer_begin("desc_rows_tc_kernel", 2.0 * (10336.0 * 64.0 + 8192.0) * static_cast<double>(num_clusters), st);
    desc_rows_tc_kernel<<<grid, dsc::kThreads, dsc::kSmemBytes, st>>>(num_clusters, n, m, radius, xyz, new_xyz, idx, orientation, wimg,
                                                                      pooled2, g_desc_dbg);
    ktimer_end(st);
    return check_launch("desc_rows_tc_kernel");
}

size_t descriptor_tc_weight_bytes() { return dsc::kWeightBytes; }

}  // namespace f3d
