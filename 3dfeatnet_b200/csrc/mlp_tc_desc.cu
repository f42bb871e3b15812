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
//   E2   : pooledA = max_s D2[:, s] + b_mid                         (no ReLU: final_relu=False, feat3dnet.py:71)
// The OTHER half of conv_mid_0 -- rows 64..127 of its weight times the tiled max-pool vector P, one column per cluster by the
// split-weight identity -- rides the tail kernel (post_tc_kernel<1>, one N = 64 MMA group per 64 clusters): as an N = 8 MMA per
// pair of tiles inside this kernel it held the tensor pipe 413 of a pair's 1714 cycles for 48 cycles of math (an instruction
// never takes less than ~34 cycles, profiles/r02_a_umma_instruction_shape.md), and its single accumulator serialised the pairs.
// This kernel writes pooledA and the two sample-half maxima of every conv1 channel (pmaxh); the tail forms
// pooled2 = pooledA + Wb^T max(pmaxh[0], pmaxh[1]).
// D1 and D2 are double-buffered.
//
// 22 warps: 0 = MMA2/3 issue, 1-4 and 14-17 = two producer warpgroups (even / odd tiles, one X1 buffer each: with the pair MMAs the
// tensor pipe needs ~860 cycles per tile and one producer warpgroup ~1060), 5-12 = two E1 warpgroups (even / odd tiles), 13 = MMA1
// issue, 18-21 = the E2 warpgroup (both tiles of every pair).  E1 (~1400 cycles per tile: accumulator -> bias / ReLU / hi-lo split ->
// operand stores -> proxy fence) was the longest stage when the same warpgroup also drained D2 (clock64 timeline
// profiles/r02_e_desc_tc_timeline.txt: an epilogue warpgroup was busy 2130 of the 2130 cycles of a pair period); with E2 on warps of its
// own the pair period is E1's alone.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "weights_layout.h"

#include <cuda_bf16.h>

#include <type_traits>

namespace f3d {

using namespace tc;

namespace dsc {
constexpr int kSamples = 64;
constexpr int kThreads = 22 * 32;  // warp 0 MMA2 issue, 1-4 / 14-17 producers (even / odd tiles), 5-12 E1, 13 MMA1 issue, 18-21 E2
constexpr uint32_t kSbo = 128;
constexpr uint32_t kLboW = 128 * 16;
constexpr uint32_t kLboX1 = kSamples * 16;
constexpr uint32_t kLboX2 = 128;                        // X2 is MN-major: K groups of 8 channels 128 B apart,
constexpr uint32_t kSboX2 = 8 * 128;                    // groups of 8 samples 1 KB apart (64 channels)
constexpr uint32_t kW1Split = 128 * 32 * 2;             // 8 KB
constexpr uint32_t kWmSplit = 128 * 64 * 2;             // 16 KB
constexpr uint32_t kOffW1 = 0;                           // [split 2][chunk 4][row 128][8]
constexpr uint32_t kOffWa = kOffW1 + 2 * kW1Split;       // [split 2][chunk 8][row 128][8]
constexpr uint32_t kOffW0 = kOffWa + 2 * kWmSplit;       // fp32 [16][8]: (w_x, w_y, w_z, bias) of layer 0, interleaved per channel pair
constexpr uint32_t kOffB0 = kOffW0 + 3 * 32 * 4;         // (tail of the float4 table)
constexpr uint32_t kOffB1 = kOffB0 + 32 * 4;             // fp32 [64]
constexpr uint32_t kOffBm = kOffB1 + 64 * 4;             // fp32 [128]
constexpr uint32_t kWeightBytes = kOffBm + 128 * 4;      // 50 432
constexpr uint32_t kX1Split = 4 * kLboX1;                // 4 KB
constexpr uint32_t kX1Buf = 2 * kX1Split;                // one X1 operand (hi + lo); buffer = tile & 1 (one per producer warpgroup)
constexpr uint32_t kOffX1 = kWeightBytes;
// X2 ring: slot = pair & 1; one slot = [split 2][16 groups of 8 samples][64 channels x 16 B]; tile (pair*2 + g) owns sample groups
// g*8 .. g*8+7, so the N = 128 operand of a pair is one uniform-stride block
constexpr uint32_t kX2Split = 16 * kSboX2;               // 16 KB
constexpr uint32_t kX2Slot = 2 * kX2Split;               // 32 KB
constexpr uint32_t kOffX2 = kOffX1 + 2 * kX1Buf;
constexpr uint32_t kOffBars = kOffX2 + 2 * kX2Slot;
constexpr uint32_t kSmemBytes = kOffBars + 20 * 8 + 16;
static_assert(kWeightBytes % 16 == 0 && kOffX1 % 128 == 0 && kOffX2 % 128 == 0 && kOffBars % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
// TMEM columns: D1[tile & 1] at 0 / 64, D2[slot] at 128 + slot*128 (two tiles x 64 samples), then the weights, which feed their
// MMAs from tensor memory (copied once with tcgen05.cp): W1 at 416 + split*16 (32 K = 16 columns), Wa at 448 + split*32 (64 K = 32).
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemD2 = 128;
constexpr uint32_t kTmemW1 = 416, kTmemWa = 448;
// per-tile events the alternating epilogue warpgroups wait on are per-warpgroup barriers; per-pair events are per ring slot
enum Bar { W_FULL = 0, W_TMEM, X1_FULL0, X1_FULL1, X1_FREE0, X1_FREE1, X2_FULL0, X2_FULL1, X2_FREE0, X2_FREE1, D1_FULL0, D1_FULL1, D1_FREE0,
           D1_FREE1, D2_FULL0, D2_FULL1, D2_FREE0, D2_FREE1, kNumBars };
static_assert(kNumBars <= 20, "barrier area");
}  // namespace dsc

__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

template <bool kTrace>  // kTrace: clock64() stamps of CTA 0 (tools/tc_timeline.py); the production instantiation carries none
__global__ void __launch_bounds__(dsc::kThreads, 1)  // 80 registers: 6 of the 22 warps share a scheduler's 16 384
desc_rows_tc_kernel(long long num_clusters, int n, int m, float radius, const float *__restrict__ xyz,
                    const float *__restrict__ new_xyz, const int *__restrict__ idx, const float *__restrict__ orientation,
                    const uint8_t *__restrict__ wimg, float *__restrict__ pooledA, float *__restrict__ pmaxh, long long *__restrict__ dbg) {
    using namespace dsc;
    auto stamp = [&](int t, int slot) {  // optional clock64() timeline of CTA 0 (bring-up)
        if constexpr (kTrace) {
            if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) dbg[t * 16 + slot] = clock64();
        }
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
            mbar_init(&bars[D2_FREE0 + b], 128);  // the E2 warpgroup (both tiles of the pair)
        }
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
            const uint32_t sbase = smem_u32(smem);
            // ---- W1 and Wa (hi and lo splits) -> tensor memory, once
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
            auto mma2 = [&](int pr) {  // pair pr = tiles 2pr, 2pr+1: X2 slot pr & 1, accumulator D2[pr & 1]
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
                    umma_commit(&bars[X2_FREE0 + sl]);
                    umma_commit(&bars[D2_FULL0 + sl]);
                }
                __syncwarp();
                stamp(2 * pr, 2);
            };
            // MMA1 (conv1) is issued by warp 13: two issuing warps hide each other's mbarrier waits (see mlp_tc.cu)
            for (int pr = 0; 2 * pr < T; ++pr) mma2(pr);
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
    } else if (warp <= 4 || (warp >= 14 && warp < 18)) {
        // ---- producers: gather + normalise + rotate + layer 0 (3 -> 32) -> X1 ------------------------------------
        // Two warpgroups: pg = 0 (warps 1-4) makes the even tiles, pg = 1 (warps 14-17) the odd ones, each into its own X1 buffer.
        // Software-pipelined over the warpgroup's own tile sequence: index of item i+2D and coordinates of item i+D in flight
        // while item i is computed.  The kernel's pace is the instruction count of its warps (six warps per scheduler, each a
        // serial chain: profiles/r02_ba_desc_tc_timeline.txt), so the per-item bookkeeping is kept off the issue slots: both
        // prefetch streams advance running pointers (no cluster / m division per item), the rotation's (cos, sin) of 32 items are
        // computed at once -- lane l of a warp takes item 32 k + l -- and handed out by shuffle, layer 0 and the hi/lo split run on
        // packed fp32 pairs (FFMA2 / FADD2), and the exact division by a radius that is not a power of two is a separate
        // instantiation of the loop instead of a select per coordinate.
        mbar_wait(&bars[W_FULL], 0);
        // per channel PAIR (a, b): {(wx_a, wx_b), (wy_a, wy_b)}, {(wz_a, wz_b), (bias_a, bias_b)}
        const ulonglong2 *W0 = reinterpret_cast<const ulonglong2 *>(smem + kOffW0);
        const int pg = warp >= 14 ? 1 : 0;
        const int pt = pg ? threadIdx.x - 14 * 32 : threadIdx.x - 32;
        const int s = pt & 63, h = pt >> 6;  // sample, channel half (16 channels = 2 K chunks)
        uint8_t *x1 = smem + kOffX1 + pg * kX1Buf + s * 16;
        const unsigned stride = gridDim.x;
        const int Tg = T > pg ? (T - pg + 1) / 2 : 0;  // items (tiles) of this warpgroup: tile = pg + 2 i
        const float inv_r = 1.0f / radius;  // exact replacement of the division when radius is a power of two (see mlp_tc.cu)
        const bool pow2 = (__float_as_uint(radius) & 0x007fffffu) == 0u && radius > 1e-30f && radius < 1e30f;
        const unsigned cl0 = static_cast<unsigned>(first) + static_cast<unsigned>(pg) * stride, step = 2u * stride;  // cluster of item i = cl0 + i step
        // stream 1: ball-query index of this thread's sample, items 0, 1, 2, ...
        const int *ip = idx + static_cast<size_t>(cl0) * kSamples + s;
        const size_t ip_step = static_cast<size_t>(step) * kSamples;
        int i_idx = 0;
        auto next_idx = [&]() -> int {
            int v = 0;
            if (i_idx < Tg) v = __ldg(ip);
            ip += ip_step;
            ++i_idx;
            return v;
        };
        // stream 2: the sample's point and the cluster centre; cloud of the cluster = cluster / m, advanced without dividing
        struct Grp { float px, py, pz, qx, qy, qz; };
        const unsigned um = static_cast<unsigned>(m), qstep = step / um, rstep = step % um;
        unsigned cl_x = cl0, bat_x = cl0 / um, rem_x = cl0 % um;
        int i_x = 0;
        auto next_xyz = [&](Grp &g, int ii) {
            g.px = g.py = g.pz = g.qx = g.qy = g.qz = 0.f;
            if (i_x < Tg) {
                ii = min(max(ii, 0), n - 1);
                const float *p = xyz + (static_cast<size_t>(bat_x) * n + ii) * 3;
                const float *c = new_xyz + static_cast<size_t>(cl_x) * 3;
                g.px = __ldg(p); g.py = __ldg(p + 1); g.pz = __ldg(p + 2);
                g.qx = __ldg(c); g.qy = __ldg(c + 1); g.qz = __ldg(c + 2);
            }
            cl_x += step; bat_x += qstep; rem_x += rstep;
            if (rem_x >= um) { rem_x -= um; ++bat_x; }
            ++i_x;
        };
        // stream 3: orientation of item 32 k + lane, one block of 32 items ahead
        auto load_th = [&](int i) -> float {
            return (orientation && i < Tg) ? __ldg(orientation + (cl0 + static_cast<unsigned>(i) * step)) : 0.0f;
        };
        float th_pref = load_th(lane), cs_l = 1.0f, sn_l = 0.0f;
        // With depth 1 the kernel was bound by the dependent L2 gather (clock64 timeline: ~2100 cycles per producer iteration
        // against ~1000 of MMA + epilogue work); the slots are static (loop unrolled by D).
        constexpr int D = 3;
        Grp gq[D];
        int iq[D];
#pragma unroll
        for (int d = 0; d < D; ++d) iq[d] = next_idx();
#pragma unroll
        for (int d = 0; d < D; ++d) next_xyz(gq[d], iq[d]);
#pragma unroll
        for (int d = 0; d < D; ++d) iq[d] = next_idx();
        auto produce = [&](auto pow2_c) {
          constexpr bool kPow2 = decltype(pow2_c)::value;
          for (int i0 = 0; i0 < Tg; i0 += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const int i = i0 + d;
              if (i >= Tg) break;
              const int t = pg + 2 * i;
              if ((warp & 3) == 1) stamp(t, 4);
              if (orientation && (i & 31) == 0) {  // (warp-uniform) the next 32 items' rotations
                  sincosf(th_pref, &sn_l, &cs_l);
                  th_pref = load_th(i + 32 + lane);
              }
              const float px = gq[d].px, py = gq[d].py, pz = gq[d].pz, qx = gq[d].qx, qy = gq[d].qy, qz = gq[d].qz;
              float gx = kPow2 ? (px - qx) * inv_r : (px - qx) / radius;
              float gy = kPow2 ? (py - qy) * inv_r : (py - qy) / radius;
              const float gz = kPow2 ? (pz - qz) * inv_r : (pz - qz) / radius;
              if (orientation) {  // pointnet_common.py:110-120: x' = x c - y s ; y' = x s + y c
                  const float cs = __shfl_sync(0xffffffffu, cs_l, i & 31), sn = __shfl_sync(0xffffffffu, sn_l, i & 31);
                  const float xr = gx * cs - gy * sn;
                  const float yr = gx * sn + gy * cs;
                  gx = xr;
                  gy = yr;
              }
              next_xyz(gq[d], iq[d]);
              iq[d] = next_idx();
              const uint64_t gx2 = pk2(gx, gx), gy2 = pk2(gy, gy), gz2 = pk2(gz, gz);
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                  const ulonglong2 wa = W0[(h * 8 + j) * 2], wb = W0[(h * 8 + j) * 2 + 1];
                  uint64_t a = fma2(gx2, wa.x, wb.y);
                  a = fma2(gy2, wa.y, a);
                  a = fma2(gz2, wb.x, a);
                  float v0, v1;
                  upk2(a, v0, v1);
                  split_bf16x2(fmaxf(v0, 0.0f), fmaxf(v1, 0.0f), hi[j], lo[j]);
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
        };
        if (pow2) produce(std::true_type{});
        else produce(std::false_type{});
    } else if (warp >= 18) {
        // ---- E2 warpgroup: pooledA = max over the samples of D2 + b_mid, both tiles of every pair --------------------
        mbar_wait(&bars[W_FULL], 0);
        const int q = warp & 3;
        const int ch = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const float bm = reinterpret_cast<const float *>(smem + kOffBm)[ch];
        uint32_t r0[32];
        for (int pr = 0; 2 * pr < T; ++pr) {
            const uint32_t sl = pr & 1;
            const bool two = 2 * pr + 1 < T;
            mbar_wait(&bars[D2_FULL0 + sl], (pr >> 1) & 1);
            tcgen05_fence_after();
            if (q == 1) stamp(2 * pr, 12);
            float mv[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (g == 1 && !two) break;  // (the N = 64 MMA of a last single tile leaves columns 64..127 unwritten)
                tmem_ld32(tmem_base + lane_addr + kTmemD2 + sl * 128 + g * 64, r0);
                tmem_ld_wait();
                float a = __uint_as_float(r0[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) a = fmaxf(a, __uint_as_float(r0[j]));
                tmem_ld32(tmem_base + lane_addr + kTmemD2 + sl * 128 + g * 64 + 32, r0);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) a = fmaxf(a, __uint_as_float(r0[j]));
                mv[g] = a;
            }
            tcgen05_fence_before();
            mbar_arrive(&bars[D2_FREE0 + sl]);
            const long long cl = first + static_cast<long long>(2 * pr) * gridDim.x;
            pooledA[cl * 128 + ch] = mv[0] + bm;
            if (two) pooledA[(cl + gridDim.x) * 128 + ch] = mv[1] + bm;
            if (q == 1) stamp(2 * pr, 13);
        }
    } else {
        // ---- E1 warpgroups: g = 0 (warps 5-8) even tiles, g = 1 (warps 9-12) odd tiles --------------------------------
        mbar_wait(&bars[W_FULL], 0);
        const int g = (warp - 5) >> 2;
        const int q = warp & 3;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int hsel = q >> 1;                 // E1: which half of the 64 samples this warp handles
        const int c1 = (q & 1) * 32 + lane;      // E1: conv1 channel (TMEM lane q*32+lane holds channel (q*32+lane) mod 64)
        const float b1 = reinterpret_cast<const float *>(smem + kOffB1)[c1];
        const uint64_t b1p = pk2(b1, b1);
        // this warpgroup's tiles own sample groups g*8 .. g*8+7 of their pair's slot; 8 consecutive samples of one channel are 16
        // contiguous bytes
        uint8_t *x2g = smem + kOffX2 + (g * 8 + hsel * 4) * kSboX2 + (c1 >> 3) * kLboX2 + (c1 & 7) * 16;
        float *pmx = pmaxh + hsel * 64 + c1;  // this thread's channel maximum over its half of the samples, per cluster
        uint32_t r0[32];
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
            for (int sidx = 0; sidx < 32; sidx += 2) {  // packed fp32 pairs: one FADD2 for the two bias additions, one for the two residuals
                float va, vb;
                upk2(add2(pk2(__uint_as_float(r0[sidx]), __uint_as_float(r0[sidx + 1])), b1p), va, vb);
                va = fmaxf(va, 0.0f);
                vb = fmaxf(vb, 0.0f);
                pmax = fmaxf(pmax, fmaxf(va, vb));
                split_bf16x2(va, vb, hi[sidx >> 1], lo[sidx >> 1]);
            }
            pmx[(first + static_cast<long long>(t) * gridDim.x) * 128] = pmax;
            if (q == 1) stamp(t, 9);
            mbar_wait(&bars[X2_FREE0 + sl], ((pr >> 1) & 1) ^ 1);  // MMA2/3(pr-2) have finished reading this slot
            if (q == 1) stamp(t, 10);
            uint8_t *x2 = x2g + sl * kX2Slot;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4 *>(x2 + j * kSboX2) = make_uint4(hi[j * 4], hi[j * 4 + 1], hi[j * 4 + 2], hi[j * 4 + 3]);
                *reinterpret_cast<uint4 *>(x2 + kX2Split + j * kSboX2) = make_uint4(lo[j * 4], lo[j * 4 + 1], lo[j * 4 + 2], lo[j * 4 + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(&bars[X2_FULL0 + b]);
            if (q == 1) stamp(t, 11);
        };
        for (int t = g; t < T; t += 2) e1(t);
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
    } else {
        const int e = i - 128 * 32 - 128 * 64;
        float *f = reinterpret_cast<float *>(wimg + kOffW0);
        if (e < 128) {  // per channel pair (2p, 2p+1): (wx, wx', wy, wy', wz, wz', b, b') -- the producers' packed fp32 pairs
            const int pr = e >> 3, c = (e >> 1) & 3, k = 2 * pr + (e & 1);
            f[e] = c < 3 ? P[L.off[W_DESC0] + c * 32 + k] : P[L.off[B_DESC0] + k];
        } else if (e < 192) f[e] = P[L.off[B_DESC1] + e - 128];
        else if (e < 320) f[e] = P[L.off[B_MID] + e - 192];
    }
}

long long *g_desc_dbg = nullptr;  // bring-up timeline buffer (f3d_debug_set_timeline with which = 1)

int descriptor_rows_tc(long long num_clusters, int n, int m, float radius, int feature_dim, const float *xyz,
                       const float *new_xyz, const int *idx, const float *orientation, const float *packed, uint8_t *wimg,
                       float *pooledA, float *pmaxh, bool build_image, int max_ctas, cudaStream_t st) {
    if (num_clusters == 0) return 0;
    if (build_image) {
        const int total = 128 * 32 + 128 * 64 + 320;
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
    auto kernel = g_desc_dbg ? desc_rows_tc_kernel<true> : desc_rows_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dsc::kSmemBytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "desc_rows_tc: cudaFuncSetAttribute");
    const int ctas = (max_ctas > 0 && max_ctas < num_sms) ? max_ctas : num_sms;
    const unsigned grid = static_cast<unsigned>(num_clusters < ctas ? num_clusters : ctas);
    // algorithmic flops (split-weight form): 2 * (rows * 10336 + clusters * 8192); the per-cluster term runs in the tail
    ktimer_begin("desc_rows_tc_kernel", 2.0 * (10336.0 * 64.0) * static_cast<double>(num_clusters), st);
    kernel<<<grid, dsc::kThreads, dsc::kSmemBytes, st>>>(num_clusters, n, m, radius, xyz, new_xyz, idx, orientation, wimg, pooledA, pmaxh,
                                                         g_desc_dbg);
    ktimer_end(st);
    return check_launch("desc_rows_tc_kernel");
}

size_t descriptor_tc_weight_bytes() { return dsc::kWeightBytes; }

}  // namespace f3d
