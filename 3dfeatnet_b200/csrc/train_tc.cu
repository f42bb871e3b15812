// train_tc.cu -- the contractions of the training-mode layers on the tensor cores (tcgen05 + TMEM), "bf16x3" precision.
//
// The three products of a 1x1-conv layer over R rows (reference models/layers.py:11-46 and the gradient graph TF derives):
//     forward  z  = x  W   (R x cin)(cin x cout)      lin_tc_kernel, A = W^T
//     dgrad    dx = dz W^T (R x cout)(cout x cin)     lin_tc_kernel, A = W
//     wgrad    dW = x^T dz (cin x R)(R x cout)        wgrad_tc_kernel
// are HBM-bound streams once the math is on the 5th-gen tensor cores (the fp32 FFMA versions in train_layers.cu ran at
// ~27 TFLOP/s and were 60 % of the training step).  fp32 operands are split on the fly into bf16 hi + lo and three MMAs
// (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) reproduce fp32 products to ~1e-5 relative, like the eval kernels.
//
// lin_tc_kernel:  D^T[channel x row] = A (TMEM, 128 output channels x K) * X^T (shared memory, 64 rows x K, K-major).
//   One CTA owns one block of 128 output channels (grid.y) whose hi/lo weight images are copied once into tensor memory
//   (bulk TMA -> tcgen05.cp) and walks 64-row tiles: all threads build the operand image, one elected lane issues the
//   3 x K/16 MMAs, the 8 warps read the accumulator back (lane = channel => a warp writes 128 contiguous bytes of a row,
//   and the BN batch statistics -- sum z, sum z^2 per channel -- are plain per-thread accumulations).  Phases of a tile are
//   sequential; two to four CTAs share an SM (TMEM columns K + 64 <= 256) and overlap each other.
// wgrad_tc_kernel: D[cin(128) x cout] += X^T (shared, K = rows) * dZ^T (shared, K = rows) over the CTA's row range, 32 rows
//   per stage (TMA ring, double-buffered operand images; see the kernel).  Per-CTA partials are reduced in a fixed order
//   afterwards (deterministic).
#include "common.cuh"
#include "dz_source.cuh"
#include "tc_ptx.cuh"

#include <cuda_bf16.h>

namespace f3d {

using namespace tc;

namespace ttc {
constexpr int kTile = 64;                 // rows per tile = MMA N of lin_tc
constexpr int kThreads = 256;
constexpr uint32_t kSbo = 128;
constexpr uint32_t kLboW = 128 * 16;      // weight image: 128 rows per 8-wide K chunk
// operand image: 64 rows per K chunk, padded so that the chunks a (quarter-)warp writes fall into distinct banks; the
// 3-split image pads by 16 B only (2-way conflicts on a few stores) so that two CTAs with a 2-deep ring still share an SM
// at K = 128
__host__ __device__ constexpr uint32_t lbo_x(int nsplit) { return kTile * 16 + (nsplit == 3 ? 16 : 32); }
}  // namespace ttc

__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        split_bf16x2(v[2 * j], v[2 * j + 1], h[j], l[j]);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// weight images: A[m][k] = src[m*sm + k*sk] (0 outside m_real x k_real), per 128-row block nsplit bf16 images
// [hi | lo] or [hi | mid | lo] (w = hi + mid + lo to 24 bits), element (r,k) at (k/8)*kLboW + r*16 + (k%8)*2
__global__ void lin_prep_kernel(const float *__restrict__ src, long long sm, long long sk, int m_real, int k_real, int kp, int mblocks,
                                int nsplit, uint8_t *__restrict__ img) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long per = 128LL * kp;
    if (i >= per * mblocks) return;
    const int mb = static_cast<int>(i / per);
    const int e = static_cast<int>(i - mb * per);
    const int r = e & 127, k = e >> 7;
    const int m = mb * 128 + r;
    float w = (m < m_real && k < k_real) ? src[m * sm + k * sk] : 0.0f;
    uint8_t *base = img + static_cast<size_t>(mb) * kp * 256 * nsplit;
    const uint32_t o = (k >> 3) * ttc::kLboW + r * 16 + (k & 7) * 2;
    for (int sp = 0; sp < nsplit; ++sp) {
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        *reinterpret_cast<__nv_bfloat16 *>(base + static_cast<size_t>(sp) * kp * 256 + o) = h;
        w -= __bfloat162float(h);
    }
}

// out[rows x nout] (+bias) = x[rows x k_real] * A^T; optional per-CTA column sums of out and out^2:
// part[((cta*2 + half)*2 + {0,1})*nout + channel].  grid = (channel blocks, row CTAs): the CTAs that share a row tile are
// neighbours in launch order, so the second one finds the tile in L2.
// The fp32 rows of a tile are one contiguous block of global memory: when k_real % 8 == 0 they are fetched by the TMA
// engine (bulk copies of up to 16 KB into a ring, kRing tiles ahead, mbarrier completion) so the loads of the next tiles
// are in flight during the conversion / MMA / epilogue of the current one; the conversion reads the ring with
// conflict-free LDS.128 and writes the bf16 hi/lo operand image (chunk stride padded by 32 B against bank conflicts).
// Measurement aid (f3d_debug_set_lin_tc_phases): phases of the lin_tc kernels that are SKIPPED -- bit 0 the operand conversion, bit 1 the
// MMAs, bit 2 the epilogue's global stores, bit 3 the TMA fetches.  Results are garbage with any bit set; the pipeline still runs.
__device__ int g_lin_dbg = 0;

// Epilogue of the lin_tc kernels: the N accumulator values of this thread's channel (rows r .. r + N - 1 of the tile) + bias -> global
// memory (row stride nout), optionally with the BN statistics.  Instruction count matters here (the epilogue warps are issue-bound):
// one pointer bump per row instead of a 64-bit row * nout product, no per-row bounds test on full tiles.
// POOL: the thread also tracks {largest key, its multiplicity, second largest distinct key} of its N rows (PoolEpilogue, dz_source.cuh);
// sgn = 0 or 0x80000000 flips z into key space.
template <int N, bool STATS, bool POOL>
__device__ __forceinline__ void store_channel_rows(float *__restrict__ o, int nout, int nvalid, const uint32_t (&r)[N], float bb, float &s1, float &s2,
                                                   uint32_t sgn, float &m1, float &c1, float &m2) {
    if (nvalid >= N) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float v = __uint_as_float(r[j]) + bb;
            *o = v;
            o += nout;
            if (STATS) { s1 += v; s2 = fmaf(v, v, s2); }
            if (POOL) {
                const float key = __uint_as_float(__float_as_uint(v) ^ sgn);
                const bool gt = key > m1, eq = key == m1;
                m2 = gt ? m1 : (!eq && key > m2) ? key : m2;
                c1 = gt ? 1.0f : eq ? c1 + 1.0f : c1;
                m1 = gt ? key : m1;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (j < nvalid) {
                const float v = __uint_as_float(r[j]) + bb;
                *o = v;
                o += nout;
                if (STATS) { s1 += v; s2 = fmaf(v, v, s2); }
            }
        }
    }
}

// contiguous global block -> shared memory through the TMA engine in as few bulk copies as possible: the issue of one cp.async.bulk costs
// the issuing thread ~450 cycles (clock64 trace, tools/wgrad_tc_trace.py), and that thread is the MMA issuer
constexpr uint32_t kBulkMax = 65536;
__device__ __forceinline__ void bulk_g2s_block(uint8_t *dst, const uint8_t *src, uint32_t bytes, uint64_t *bar) {
    for (uint32_t off = 0; off < bytes; off += kBulkMax) bulk_g2s(dst + off, src + off, bytes - off < kBulkMax ? bytes - off : kBulkMax, bar);
}

// Measurement aid (f3d_debug_lin_tc_trace): clock64() stamps of CTA (0, 0) of lin_tc_kernel, 16 slots per tile for threads 0 and 255
__device__ long long *g_lin_trace = nullptr;
constexpr int kTraceTiles = 64;

constexpr int kRingMax = 2;  // ring depth (1 when shared memory cannot hold two stages next to a 3-split operand image)

template <int nsplit>
__global__ void __launch_bounds__(ttc::kThreads, 4)
lin_tc_kernel(long long rows, int k_real, int kp, int nout, int kRing, uint32_t tmem_cols, const float *__restrict__ x, const uint8_t *__restrict__ wimg,
              const float *__restrict__ bias, const float *__restrict__ gbias, int gs, float *__restrict__ out, float *__restrict__ part,
              XSource X, PoolEpilogue PE) {
    using namespace ttc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const bool ring = (k_real & 7) == 0;
    const uint32_t wbytes = 256u * kp;                                  // one weight split: (kp/8) chunks x kLboW
    constexpr uint32_t kLboXp = lbo_x(nsplit);
    const uint32_t split = static_cast<uint32_t>(kp / 8) * kLboXp;      // operand image: nsplit splits x (kp/8) chunks x kLboXp
    const uint32_t opbytes = static_cast<uint32_t>(nsplit) * split;     // >= wbytes
    const uint32_t stage_bytes = ring ? kTile * static_cast<uint32_t>(k_real) * 4 : 0;
    uint8_t *ringbuf = smem + opbytes;
    uint64_t *bar_w = reinterpret_cast<uint64_t *>(smem + opbytes + kRing * stage_bytes);
    uint64_t *bar_m = bar_w + 1;
    uint64_t *bar_full = bar_w + 2;  // [kRing]
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bar_w + 2 + kRingMax);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2;
    const int ch = q * 32 + lane;
    const int col0 = half * 32;
    const uint32_t sbase = smem_u32(smem);
    const int mb = blockIdx.x;
    const long long cta = blockIdx.y, ncta = gridDim.y;
    const int dbg = g_lin_dbg;
    long long *tr = nullptr;
    if (g_lin_trace && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x == 0 || threadIdx.x == 255)) tr = g_lin_trace + (threadIdx.x ? kTraceTiles * 16 : 0);
#define F3D_LT(i) if (tr && it < kTraceTiles) tr[it * 16 + (i)] = clock64();

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_m, 1);
        for (int s = 0; s < kRing; ++s) mbar_init(bar_full + s, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_base_s, tmem_cols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    uint32_t wpar = 0, mpar = 0;

    // weights -> tensor memory: split sp -> columns [sp*kp/2, (sp+1)*kp/2)
    const uint8_t *wsrc = wimg + static_cast<size_t>(mb) * kp * 256 * nsplit;
    for (int piece = 0; piece < nsplit; ++piece) {
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(bar_w, wbytes);
            for (uint32_t off = 0; off < wbytes; off += 16384) {
                const uint32_t n = wbytes - off < 16384u ? wbytes - off : 16384u;
                bulk_g2s(smem + off, wsrc + static_cast<size_t>(piece) * wbytes + off, n, bar_w);
            }
        }
        mbar_wait(bar_w, wpar);
        wpar ^= 1;
        tcgen05_fence_after();
        if (warp == 0) {
            if (elect_one()) {
                for (int k = 0; k < kp / 16; ++k)
                    tmem_cp_128x256b(tmem_base + piece * (kp / 2) + k * 8, make_smem_desc(sbase + k * 2 * kLboW, kLboW, kSbo));
                umma_commit(bar_m);
            }
            __syncwarp();
        }
        mbar_wait(bar_m, mpar);
        mpar ^= 1;
        __syncthreads();
    }

    const long long ntiles = (rows + kTile - 1) / kTile;
    // warp 0: fetch the fp32 rows of `tile` into ring stage s
    auto fetch = [&](long long tile, int s) {
        const long long r0 = tile * kTile;
        const uint32_t valid = static_cast<uint32_t>(rows - r0 < kTile ? rows - r0 : kTile);
        const uint32_t bytes = valid * static_cast<uint32_t>(k_real) * 4;  // the tile's rows are one contiguous block
        if (lane == 0 && (dbg & 8)) {
            mbar_arrive_expect_tx(bar_full + s, 0);
        } else if (lane == 0) {
            mbar_arrive_expect_tx(bar_full + s, bytes);
            const uint8_t *src = reinterpret_cast<const uint8_t *>(x + r0 * k_real);
            bulk_g2s_block(ringbuf + s * stage_bytes, src, bytes, bar_full + s);
        }
        __syncwarp();
    };
    if (ring && warp == 0)
        for (int s = 0; s < kRing; ++s)
            if (cta + s * ncta < ntiles) fetch(cta + s * ncta, s);

    const uint32_t idesc = make_idesc(1, 128, kTile);
    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(nsplit) * (kp / 2);
    const int gch = mb * 128 + ch;
    const bool ch_ok = gch < nout;
    const float bb = (bias && ch_ok) ? __ldg(bias + gch) : 0.0f;
    float s1 = 0.0f, s2 = 0.0f;
    long long it = 0;
    for (long long tile = cta; tile < ntiles; tile += ncta, ++it) {
        const long long r0 = tile * kTile;
        const int stage = static_cast<int>(it % kRing);
        F3D_LT(0)
        if (ring) {
            mbar_wait(bar_full + stage, static_cast<uint32_t>((it / kRing) & 1));
            F3D_LT(1)
            // lane -> (row within a group of 4, chunk c mod 4, 16-byte half of the chunk): a quarter-warp reads 128 contiguous
            // bytes of one row, a half-warp writes 16 distinct 8-byte slots of the operand image (no bank conflicts)
            const int rsub = lane >> 3, c4 = (lane >> 1) & 3, h = lane & 1;
            const uint32_t row_bytes = static_cast<uint32_t>(k_real) * 4;
            for (int c = c4; c < ((dbg & 1) ? 0 : kp / 8); c += 4) {
                // X.coef: the rows in the ring are the previous layer's z; its BN + ReLU is applied here (XSource, dz_source.cuh)
                float4 xsc = make_float4(1.f, 1.f, 1.f, 1.f), xsh = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool chok = c * 8 + h * 4 < k_real;
                if (X.coef && chok) {
                    xsc = __ldg(reinterpret_cast<const float4 *>(X.coef + c * 8 + h * 4));
                    xsh = __ldg(reinterpret_cast<const float4 *>(X.coef + k_real + c * 8 + h * 4));
                }
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const int r = p * 32 + warp * 4 + rsub;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r0 + r < rows && chok) {
                        a = *reinterpret_cast<const float4 *>(ringbuf + stage * stage_bytes + r * row_bytes + h * 16 + c * 32);
                        if (X.coef) a = x_value(a, xsc, xsh, X.relu);
                    }
                    uint8_t *dst = smem + c * kLboXp + r * 16 + h * 8;
#pragma unroll
                    for (int sp = 0; sp < nsplit; ++sp) {  // a = hi (+ mid) + lo, each a bf16
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
                        *reinterpret_cast<uint2 *>(dst + sp * split) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                        bf16x2_residual(a.x, a.y, *reinterpret_cast<const uint32_t *>(&h0), a.x, a.y);
                            bf16x2_residual(a.z, a.w, *reinterpret_cast<const uint32_t *>(&h1), a.z, a.w);
                    }
                }
            }
        } else {  // few input channels (the xyz layers): straight from global memory
            const int r = threadIdx.x & 63, cq = threadIdx.x >> 6;
            const bool valid = r0 + r < rows;
            const float *row = x + (r0 + r) * k_real;
            for (int c = cq; c < kp / 8; c += 4) {
                float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (c * 8 + j < k_real) v[j] = __ldg(row + c * 8 + j);
                }
                uint8_t *dst = smem + c * kLboXp + r * 16;
#pragma unroll
                for (int sp = 0; sp < nsplit; ++sp) {
                    uint32_t hw[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        hw[j] = *reinterpret_cast<const uint32_t *>(&h2);
                        bf16x2_residual(v[2 * j], v[2 * j + 1], hw[j], v[2 * j], v[2 * j + 1]);
                    }
                    *reinterpret_cast<uint4 *>(dst + sp * split) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                }
            }
        }
        F3D_LT(2)
        fence_proxy_async_smem();
        __syncthreads();  // operand image complete; ring stage `stage` fully read
        F3D_LT(3)
        if (warp == 0) {
            tcgen05_fence_after();
            if (elect_one()) {
                // product terms (weight split, operand split): 2 splits -> hh, hl, lh (error ~2^-17 per product);
                // 3 splits -> + h*l2, l2*h, m*m (every term down to 2^-24: fp32-grade, used for the forward so that the
                // ReLU / max-pool routing decisions match an fp32 evaluation)
                constexpr int nterms = nsplit == 3 ? 6 : 3;
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < ((dbg & 2) ? 0 : nterms); ++term) {
                    const int ws = term == 2 ? 1 : term == 4 ? 2 : term == 5 ? 1 : 0;
                    const int xs = term == 1 ? 1 : term == 3 ? 2 : term == 5 ? 1 : 0;
                    const uint32_t wa = tmem_base + ws * (kp / 2);
                    const uint32_t xb = sbase + xs * split;
                    for (int k = 0; k < kp / 16; ++k) {
                        umma_f16_ts(tmem_d, wa + k * 8, make_smem_desc(xb + k * 2 * kLboXp, kLboXp, kSbo), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar_m);
            }
            __syncwarp();
            if (ring && tile + kRing * ncta < ntiles) fetch(tile + kRing * ncta, stage);  // refill the stage just consumed
        }
        F3D_LT(4)
        mbar_wait(bar_m, mpar);
        mpar ^= 1;
        tcgen05_fence_after();
        F3D_LT(5)
        uint32_t r[32];
        tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + col0, r);
        tmem_ld_wait();
        F3D_LT(6)
        if (ch_ok) {
            if (gbias) {
                // optional per-group additive term (rows of a group of gs consecutive rows share it): the pooled half of a
                // concat([x, tile(pooled)]) input contributes pooled * W_bottom once per group instead of once per row
                const long long g0 = (r0 + col0) / gs;
                const int rem0 = static_cast<int>((r0 + col0) - g0 * gs);
                const long long gmax = (rows - 1) / gs;
                if (rem0 + 31 < gs) {
                    const float gb = __ldg(gbias + g0 * nout + gch);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + gb);
                } else {
                    for (int j = 0; j < 32; ++j) {
                        long long grp = g0 + (rem0 + j) / gs;
                        grp = grp < gmax ? grp : gmax;
                        r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(gbias + grp * nout + gch));
                    }
                }
            }
            float *o = out + (r0 + col0) * nout + gch;
            const long long left = rows - (r0 + col0);
            const int nvalid = (dbg & 4) ? 0 : left < 32 ? static_cast<int>(left) : 32;
            float m1 = -INFINITY, c1 = 0.f, m2 = -INFINITY;
            if (PE.zext) {  // forward of a pool-only layer (full tiles only: rows % 64 == 0)
                const uint32_t sgn = __ldg(PE.gamma + gch) >= 0.f ? 0u : 0x80000000u;
                store_channel_rows<32, true, true>(o, nout, nvalid, r, bb, s1, s2, sgn, m1, c1, m2);
                float *ze = PE.zext + static_cast<size_t>((r0 + col0) >> 5) * 3 * nout + gch;
                ze[0] = m1;
                ze[nout] = c1;
                ze[2 * nout] = m2;
            } else if (part) {
                store_channel_rows<32, true, false>(o, nout, nvalid, r, bb, s1, s2, 0u, m1, c1, m2);
            } else {
                store_channel_rows<32, false, false>(o, nout, nvalid, r, bb, s1, s2, 0u, m1, c1, m2);
            }
        }
        F3D_LT(7)
        tcgen05_fence_before();
        __syncthreads();  // the operand image and the accumulator are reused by the next tile
        F3D_LT(8)
    }
#undef F3D_LT
    if (part && ch_ok) {
        float *p = part + (static_cast<size_t>(cta) * 2 + half) * 2 * nout;
        p[gch] = s1;
        p[nout + gch] = s2;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// lin_tc_pipe_kernel: the same contraction as lin_tc_kernel with the three phases of a tile running CONCURRENTLY on
// different warps (k_real % 8 == 0): warp 0 issues the TMA fetches (ring, up to 3 tiles ahead) and the MMAs, warps 1-8
// convert ring slot -> operand image (double-buffered), warps 9-16 drain the accumulator (double-buffered in TMEM) to
// global memory and keep the BN statistics.  mbarriers: ring_full (TMA bytes), img_full (the converter warps), mma_done
// (tcgen05.commit: frees the image AND publishes the accumulator), d_free (the epilogue warps).  Converters and epilogue are
// latency / issue-bound per warp (measured with tools/lin_tc_phases.py), hence eight warps each.
namespace lp {
constexpr int kConvWarps = 8;  // converter warps (rows of a tile split among them)
constexpr int kEpiWarps = 8;   // epilogue warps: two per TMEM lane quarter, each draining half of the tile's rows
constexpr int kThreads = 32 * (1 + kConvWarps + kEpiWarps);
constexpr int kMaxRing = 3;
// operand-image stride between 8-wide K chunks.  The fused dz converter writes a ROW of the tile per warp (32 channel quads -> 16 chunks
// x 2 halves): +16 puts consecutive chunks 4 banks apart, so its 8-byte stores take the minimum of two wavefronts.
__host__ __device__ constexpr uint32_t lbo(int nsplit, int nt, bool fused = false) {
    return static_cast<uint32_t>(nt) * 16 + ((nsplit == 3 || fused) ? 16 : 32);
}
}  // namespace lp

// FUSED: the input rows are not read from x but formed on the fly as the BN-backward gradient dz of a pool-only layer (DzSource,
// dz_source.cuh): the ring fetches the z tile plus the three pooled rows of the tile's group (S.gs % NT == 0), the converters evaluate
// dz_value() before the bf16 split.  dgb != NULL (S.gs == NT): the per-group column sums of dz -- the gradient of the layer's per-group
// additive term -- are reduced by the converter warps and written once per tile.
template <int nsplit, int NT, bool FUSED>
__global__ void __launch_bounds__(lp::kThreads, 1)
lin_tc_pipe_kernel(long long rows, int k_real, int kp, int nout, int nring, uint32_t tmem_cols, const float *__restrict__ x,
                   const uint8_t *__restrict__ wimg, const float *__restrict__ bias, const float *__restrict__ gbias, int gs,
                   float *__restrict__ out, float *__restrict__ part, DzSource S, float *__restrict__ dgb, XSource X, PoolEpilogue PE) {
    using namespace ttc;
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t kLbo = lp::lbo(nsplit, NT, FUSED);
    const uint32_t wbytes = 256u * kp;
    const uint32_t split = static_cast<uint32_t>(kp / 8) * kLbo;
    const uint32_t img_bytes = static_cast<uint32_t>(nsplit) * split;
    const uint32_t pool_bytes = FUSED ? 3u * static_cast<uint32_t>(k_real) * 4 : 0u;  // pooled max | pooled gradient | 1 / ties of the tile's group
    const uint32_t slot_bytes = NT * static_cast<uint32_t>(k_real) * 4 + pool_bytes;
    uint8_t *ringbuf = smem + 2 * img_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * img_bytes + nring * slot_bytes);
    float *coef_s = reinterpret_cast<float *>(smem + 2 * img_bytes + nring * slot_bytes + 128);  // FUSED: [7][k_real], then gred [4][k_real]
    float *gred = coef_s + kDzCoefs * k_real;
    if (FUSED) {
        for (int i = threadIdx.x; i < kDzCoefs * k_real; i += lp::kThreads) coef_s[i] = __ldg(S.coef + i);
    }
    uint64_t *bar_w = bars, *bar_wm = bars + 1, *ring_full = bars + 2, *img_full = bars + 2 + lp::kMaxRing, *mma_done = img_full + 2,
             *d_free = mma_done + 2;
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(d_free + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const int mb = blockIdx.x;
    const long long cta = blockIdx.y, ncta = gridDim.y;
    const int dbg = g_lin_dbg;

    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_wm, 1);
        for (int i = 0; i < lp::kMaxRing; ++i) mbar_init(ring_full + i, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(img_full + i, lp::kConvWarps);
            mbar_init(mma_done + i, 1);
            mbar_init(d_free + i, lp::kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_base_s, tmem_cols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    // weights -> tensor memory: split sp -> columns [sp*kp/2, (sp+1)*kp/2); staged through the (still unused) image area
    {
        uint32_t wpar = 0, mpar = 0;
        const uint8_t *wsrc = wimg + static_cast<size_t>(mb) * kp * 256 * nsplit;
        for (int piece = 0; piece < nsplit; ++piece) {
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(bar_w, wbytes);
                for (uint32_t off = 0; off < wbytes; off += 16384) {
                    const uint32_t n = wbytes - off < 16384u ? wbytes - off : 16384u;
                    bulk_g2s(smem + off, wsrc + static_cast<size_t>(piece) * wbytes + off, n, bar_w);
                }
            }
            mbar_wait(bar_w, wpar);
            wpar ^= 1;
            tcgen05_fence_after();
            if (warp == 0) {
                if (elect_one()) {
                    for (int k = 0; k < kp / 16; ++k)
                        tmem_cp_128x256b(tmem_base + piece * (kp / 2) + k * 8, make_smem_desc(sbase + k * 2 * kLboW, kLboW, kSbo));
                    umma_commit(bar_wm);
                }
                __syncwarp();
            }
            mbar_wait(bar_wm, mpar);
            mpar ^= 1;
            __syncthreads();
        }
    }

    const long long ntiles = (rows + NT - 1) / NT;
    const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(nsplit) * (kp / 2);
    if (warp == 0) {
        // ------------------------------------------------------------------ issuer: TMA ring + MMAs
        auto fetch = [&](long long tile, int slot) {
            const long long r0 = tile * NT;
            const uint32_t valid = static_cast<uint32_t>(rows - r0 < NT ? rows - r0 : NT);
            const uint32_t bytes = valid * static_cast<uint32_t>(k_real) * 4;
            if (lane == 0 && (dbg & 8)) {
                mbar_arrive_expect_tx(ring_full + slot, 0);
            } else if (lane == 0) {
                mbar_arrive_expect_tx(ring_full + slot, bytes + pool_bytes);
                const uint8_t *src = reinterpret_cast<const uint8_t *>((FUSED ? S.z : x) + r0 * k_real);
                bulk_g2s_block(ringbuf + slot * slot_bytes, src, bytes, ring_full + slot);
                if (FUSED) {
                    const size_t go = static_cast<size_t>(r0 / S.gs) * k_real;
                    uint8_t *pd = ringbuf + slot * slot_bytes + NT * static_cast<uint32_t>(k_real) * 4;
                    const uint32_t kb = static_cast<uint32_t>(k_real) * 4;
                    bulk_g2s(pd, S.pooled + go, kb, ring_full + slot);
                    bulk_g2s(pd + kb, S.gpool + go, kb, ring_full + slot);
                    bulk_g2s(pd + 2 * kb, S.inv + go, kb, ring_full + slot);
                }
            }
            __syncwarp();
        };
        for (int s = 0; s < nring; ++s)
            if (cta + s * ncta < ntiles) fetch(cta + s * ncta, s);
        const uint32_t idesc = make_idesc(1, 128, NT);
        long long it = 0;
        for (long long tile = cta; tile < ntiles; tile += ncta, ++it) {
            const int b = static_cast<int>(it & 1);
            mbar_wait(img_full + b, static_cast<uint32_t>((it >> 1) & 1));          // image b converted, its ring slot consumed
            if (it >= 2) mbar_wait(d_free + b, static_cast<uint32_t>(((it >> 1) - 1) & 1));  // accumulator b drained
            tcgen05_fence_after();
            if (tile + nring * ncta < ntiles) fetch(tile + nring * ncta, static_cast<int>(it % nring));
            if (elect_one()) {
                constexpr int nterms = nsplit == 3 ? 6 : 3;
                const uint32_t d = tmem_d + b * NT;
                uint32_t acc = 0;
#pragma unroll
                for (int term = 0; term < ((dbg & 2) ? 0 : nterms); ++term) {
                    const int ws = term == 2 ? 1 : term == 4 ? 2 : term == 5 ? 1 : 0;
                    const int xs = term == 1 ? 1 : term == 3 ? 2 : term == 5 ? 1 : 0;
                    const uint32_t wa = tmem_base + ws * (kp / 2);
                    const uint32_t xb = sbase + b * img_bytes + xs * split;
                    for (int k = 0; k < kp / 16; ++k) {
                        umma_f16_ts(d, wa + k * 8, make_smem_desc(xb + k * 2 * kLbo, kLbo, kSbo), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(mma_done + b);
            }
            __syncwarp();
        }
    } else if (warp <= lp::kConvWarps) {
        // ------------------------------------------------------------------ converters: ring slot -> operand image
        constexpr int kRowsPass = lp::kConvWarps * 4;  // rows converted per pass of the converter warps
        static_assert(NT % kRowsPass == 0, "tile rows must be a multiple of the converters' rows per pass");
        const int wc = warp - 1;
        const int rsub = lane >> 3, c4 = (lane >> 1) & 3, h = lane & 1;
        const uint32_t row_bytes = static_cast<uint32_t>(k_real) * 4;
        // FUSED: channel quad / first row / row step of this thread (k_real a multiple of 16 with (converter threads) % (k_real / 4) == 0)
        const int fquads = k_real >> 2, ft = threadIdx.x - 32;
        const int fq = ft % fquads, frow = ft / fquads, frows_pass = (lp::kConvWarps * 32) / fquads;
        DzCoefRegs KF{};
        if (FUSED) {
            auto ld = [&](int i) { return *reinterpret_cast<const float4 *>(coef_s + i * k_real + fq * 4); };
            KF = DzCoefRegs{ld(0), ld(1), ld(2), ld(3), ld(4), ld(5), ld(6)};
        }
        long long it = 0;
        for (long long tile = cta; tile < ntiles; tile += ncta, ++it) {
            const int b = static_cast<int>(it & 1), slot = static_cast<int>(it % nring);
            const long long r0 = tile * NT;
            mbar_wait(ring_full + slot, static_cast<uint32_t>((it / nring) & 1));
            if (it >= 2) mbar_wait(mma_done + b, static_cast<uint32_t>(((it >> 1) - 1) & 1));  // MMAs of tile it-2 have read image b
            uint8_t *img = smem + b * img_bytes;
            if (dbg & 1) {
            } else if constexpr (!FUSED) {
                for (int c = c4; c < kp / 8; c += 4) {
                    // X.coef: the rows in the ring are the previous layer's z; its BN + ReLU is applied here (XSource, dz_source.cuh)
                    float4 xsc = make_float4(1.f, 1.f, 1.f, 1.f), xsh = make_float4(0.f, 0.f, 0.f, 0.f);
                    const bool chok = c * 8 + h * 4 < k_real;
                    if (X.coef && chok) {
                        xsc = __ldg(reinterpret_cast<const float4 *>(X.coef + c * 8 + h * 4));
                        xsh = __ldg(reinterpret_cast<const float4 *>(X.coef + k_real + c * 8 + h * 4));
                    }
#pragma unroll
                    for (int p = 0; p < NT / kRowsPass; ++p) {
                        const int r = p * kRowsPass + wc * 4 + rsub;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r0 + r < rows && chok) {
                            a = *reinterpret_cast<const float4 *>(ringbuf + slot * slot_bytes + r * row_bytes + h * 16 + c * 32);
                            if (X.coef) a = x_value(a, xsc, xsh, X.relu);
                        }
                        uint8_t *dst = img + c * kLbo + r * 16 + h * 8;
#pragma unroll
                        for (int sp = 0; sp < nsplit; ++sp) {
                            const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
                            *reinterpret_cast<uint2 *>(dst + sp * split) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                            bf16x2_residual(a.x, a.y, *reinterpret_cast<const uint32_t *>(&h0), a.x, a.y);
                            bf16x2_residual(a.z, a.w, *reinterpret_cast<const uint32_t *>(&h1), a.z, a.w);
                        }
                    }
                }
            } else {
                // thread t of the converters owns channel quad t % quads for the whole kernel (its seven coefficient quads stay in
                // registers, KF below) and rows t / quads, + rows_pass, ... of every tile: one LDS.128 per four values instead of fourteen
                // (the coefficient and pooled rows used to be re-read for every row: the converters were shared-memory-bound)
                const float *pool = reinterpret_cast<const float *>(ringbuf + slot * slot_bytes + NT * row_bytes);
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 pm = *reinterpret_cast<const float4 *>(pool + fq * 4), gp = *reinterpret_cast<const float4 *>(pool + k_real + fq * 4),
                             iv = *reinterpret_cast<const float4 *>(pool + 2 * k_real + fq * 4);
                const float4 gsc = make_float4(__fmul_rn(gp.x, iv.x), __fmul_rn(gp.y, iv.y), __fmul_rn(gp.z, iv.z), __fmul_rn(gp.w, iv.w));
                float4 acc = z4;
                uint8_t *dst0 = img + (fq >> 1) * kLbo + (fq & 1) * 8;
                for (int r = frow; r < NT; r += frows_pass) {
                    float4 a = z4;
                    if (r0 + r < rows) {
                        const float4 zz = *reinterpret_cast<const float4 *>(ringbuf + slot * slot_bytes + r * row_bytes + fq * 16);
                        a = dz_value4(zz, KF, pm, gsc, S.relu);
                    }
                    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                    uint8_t *dst = dst0 + r * 16;
#pragma unroll
                    for (int sp = 0; sp < nsplit; ++sp) {
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
                        *reinterpret_cast<uint2 *>(dst + sp * split) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                        bf16x2_residual(a.x, a.y, *reinterpret_cast<const uint32_t *>(&h0), a.x, a.y);
                            bf16x2_residual(a.z, a.w, *reinterpret_cast<const uint32_t *>(&h1), a.z, a.w);
                    }
                }
                if (dgb) *reinterpret_cast<float4 *>(gred + frow * k_real + fq * 4) = acc;  // this thread's rows of the tile (= one group)
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(img_full + b);
            if (FUSED && dgb) {
                asm volatile("bar.sync 1, %0;" ::"n"(lp::kConvWarps * 32) : "memory");  // the converter warps
                float *dst = dgb + static_cast<size_t>(r0 / S.gs) * k_real;
                for (int ch = wc * 32 + lane; ch < k_real; ch += lp::kConvWarps * 32) {
                    float t = 0.f;
                    for (int g = 0; g < frows_pass; ++g) t += gred[g * k_real + ch];  // fixed order
                    dst[ch] = t;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(lp::kConvWarps * 32) : "memory");  // gred is rewritten by the next tile
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: accumulator -> global, BN statistics
        constexpr int NH = NT / 2;  // rows (accumulator columns) per epilogue warp: the two warps of a lane quarter split the tile
        const int q = warp & 3, half = (warp - 1 - lp::kConvWarps) >> 2;
        const int ch = q * 32 + lane;
        const int gch = mb * 128 + ch;
        const bool ch_ok = gch < nout;
        const float bb = (bias && ch_ok) ? __ldg(bias + gch) : 0.0f;
        float s1 = 0.0f, s2 = 0.0f;
        long long it = 0;
        for (long long tile = cta; tile < ntiles; tile += ncta, ++it) {
            const int b = static_cast<int>(it & 1);
            const long long r0 = tile * NT + half * NH;
            mbar_wait(mma_done + b, static_cast<uint32_t>((it >> 1) & 1));
            tcgen05_fence_after();
            uint32_t r[NH];
            if constexpr (NH == 32)
                tmem_ld32(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + b * NT + half * NH, r);
            else
                tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + b * NT + half * NH, r);
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_free + b);
            if (ch_ok && r0 < rows) {
                if (gbias) {
                    const long long g0 = r0 / gs;
                    const int rem0 = static_cast<int>(r0 - g0 * gs);
                    const long long gmax = (rows - 1) / gs;
                    if (rem0 + NH - 1 < gs) {
                        const float gb = __ldg(gbias + g0 * nout + gch);
#pragma unroll
                        for (int j = 0; j < NH; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + gb);
                    } else {
#pragma unroll
                        for (int j = 0; j < NH; ++j) {
                            long long grp = g0 + (rem0 + j) / gs;
                            grp = grp < gmax ? grp : gmax;
                            r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldg(gbias + grp * nout + gch));
                        }
                    }
                }
                float *o = out + r0 * nout + gch;
                const long long left = rows - r0;
                const int nvalid = (dbg & 4) ? 0 : left < NH ? static_cast<int>(left) : NH;
                float m1 = -INFINITY, c1 = 0.f, m2 = -INFINITY;
                if (NH == 32 && PE.zext) {  // forward of a pool-only layer (full tiles only: rows % 64 == 0)
                    const uint32_t sgn = __ldg(PE.gamma + gch) >= 0.f ? 0u : 0x80000000u;
                    store_channel_rows<NH, true, true>(o, nout, nvalid, r, bb, s1, s2, sgn, m1, c1, m2);
                    float *ze = PE.zext + static_cast<size_t>(r0 >> 5) * 3 * nout + gch;
                    ze[0] = m1;
                    ze[nout] = c1;
                    ze[2 * nout] = m2;
                } else if (part) {
                    store_channel_rows<NH, true, false>(o, nout, nvalid, r, bb, s1, s2, 0u, m1, c1, m2);
                } else {
                    store_channel_rows<NH, false, false>(o, nout, nvalid, r, bb, s1, s2, 0u, m1, c1, m2);
                }
            }
        }
        if (part && ch_ok) {  // two partial slots per CTA like lin_tc_kernel: one per half of the tiles' rows
            float *p = part + (static_cast<size_t>(cta) * 2 + half) * 2 * nout;
            p[gch] = s1;
            p[nout + gch] = s2;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------------
// wgrad: partW[cta][cin][cout] = sum over the CTA's rows of x[r][ci] * dz[r][co].
// D[ci (128, zero-padded) x co] accumulates in tensor memory over the whole row range of the CTA; a stage is 32 rows
// (two K = 16 MMA steps x 3 split terms).  The fp32 rows of a stage are two contiguous blocks (x and dz): bulk-TMA'd into
// a 2-deep ring, one stage ahead.  Both operands need K = row contiguous, so the conversion transposes 8 rows x 4 channels
// per thread: LDS.128 of 32 consecutive channel quads of one row (conflict-free), bf16 hi/lo split, four 16-byte chunks
// (8 rows of one channel) stored into the operand image whose core-matrix stride is padded to 144 B (conflict-free).
// The operand images are double-buffered and the roles are split: warp 0 only issues TMA fetches and MMAs, the other 15
// warps only convert; mbarriers (ring filled / image converted / MMAs done) connect them, so the MMAs and the fetch of
// stage s overlap the conversion of stage s+1 and nobody waits for the issuer.
namespace wg {
constexpr int kRows = 32;            // rows per stage
constexpr uint32_t kSboP = 144;      // stride between 8-channel core matrices (128 B + 16 B padding)
constexpr int kChunks = kRows / 8;   // K chunks per stage
constexpr int kThreads = 512;        // 16 warps: the conversion is the long phase of a stage
constexpr int kMaxRing = 4;          // ring slots of fp32 stages (as many as fit next to the operand images: the fetch of stage s + ring
                                     // is issued when stage s has been converted, so a 2-deep ring leaves the copy one stage of cover)
}  // namespace wg

// 8 rows x 4 channels of fp32 -> four 16-byte chunks (8 rows of one channel each) of the hi and the lo image
__device__ __forceinline__ void wgrad_store_unit(uint8_t *dst, uint32_t split, const float4 (&v)[8]) {
    uint4 hi, lo;
    {
        const float col[8] = {v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x};
        split8(col, hi, lo);
        *reinterpret_cast<uint4 *>(dst) = hi;
        *reinterpret_cast<uint4 *>(dst + split) = lo;
    }
    {
        const float col[8] = {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y};
        split8(col, hi, lo);
        *reinterpret_cast<uint4 *>(dst + 16) = hi;
        *reinterpret_cast<uint4 *>(dst + split + 16) = lo;
    }
    {
        const float col[8] = {v[0].z, v[1].z, v[2].z, v[3].z, v[4].z, v[5].z, v[6].z, v[7].z};
        split8(col, hi, lo);
        *reinterpret_cast<uint4 *>(dst + 32) = hi;
        *reinterpret_cast<uint4 *>(dst + split + 32) = lo;
    }
    {
        const float col[8] = {v[0].w, v[1].w, v[2].w, v[3].w, v[4].w, v[5].w, v[6].w, v[7].w};
        split8(col, hi, lo);
        *reinterpret_cast<uint4 *>(dst + 48) = hi;
        *reinterpret_cast<uint4 *>(dst + split + 48) = lo;
    }
}

__device__ __forceinline__ void wgrad_convert(uint8_t *img, uint32_t lbo, uint32_t split, const uint8_t *__restrict__ stage, int c, int valid_rows,
                                              const XSource X = XSource{nullptr, 0}) {
    // unit u -> (row chunk rc, channel quad c4); a warp handles 32 consecutive quads of one row chunk; converter threads only
    const int quads = c >> 2;
    for (int u = threadIdx.x - 32; u < wg::kChunks * quads; u += wg::kThreads - 32) {
        const int rc = u / quads, c4 = u - rc * quads;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rc * 8 + i;
            v[i] = r < valid_rows ? *reinterpret_cast<const float4 *>(stage + (static_cast<size_t>(r) * c + c4 * 4) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (X.coef) {  // the staged rows are the previous layer's z: its BN + ReLU is applied here (XSource, dz_source.cuh)
            const float4 xsc = __ldg(reinterpret_cast<const float4 *>(X.coef + c4 * 4)), xsh = __ldg(reinterpret_cast<const float4 *>(X.coef + c + c4 * 4));
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (rc * 8 + i < valid_rows) v[i] = x_value(v[i], xsc, xsh, X.relu);
        }
        wgrad_store_unit(img + rc * lbo + (c4 >> 1) * wg::kSboP + (c4 & 1) * 64, split, v);
    }
}

// Both operands of a stage in ONE pass over a combined unit index, with units of 8 rows x W channels (W = 4, 2 or 1): narrow layers
// have few 4-channel units (96 for 32 + 64 channels) and converting A and then B left the same three warps with two dependent units
// each while twelve warps idled -- the stage time was that latency (tools/wgrad_tc_phases.py).  W is chosen by the launcher so that
// about 384 units exist per stage.
template <int W>
__device__ __forceinline__ void wgrad_convert_both(uint8_t *img_a, uint32_t lbo_a, uint32_t split_a, uint8_t *img_b, uint32_t lbo_b, uint32_t split_b,
                                                   const uint8_t *__restrict__ stage_a, const uint8_t *__restrict__ stage_b, int cin, int cout,
                                                   int valid_rows, const XSource X, int dbg = 0) {
    const int ga = cin / W, gb = cout / W;
    const int ua = wg::kChunks * ga, ut = ua + wg::kChunks * gb;
    for (int u = threadIdx.x - 32; u < ut; u += wg::kThreads - 32) {
        const bool is_a = u < ua;
        const int uu = is_a ? u : u - ua, g = is_a ? ga : gb, c = is_a ? cin : cout;
        const int rc = uu / g, cg = uu - rc * g;
        const uint8_t *src = (is_a ? stage_a : stage_b) + static_cast<size_t>(cg) * W * 4;
        float v[8][W];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rc * 8 + i;
            const uint8_t *p = src + static_cast<size_t>(r) * c * 4;
            if (r < valid_rows && !(dbg & 4)) {
                if constexpr (W == 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(p);
                    v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
                } else if constexpr (W == 2) {
                    const float2 t = *reinterpret_cast<const float2 *>(p);
                    v[i][0] = t.x; v[i][1] = t.y;
                } else {
                    v[i][0] = *reinterpret_cast<const float *>(p);
                }
            } else {
#pragma unroll
                for (int j = 0; j < W; ++j) v[i][j] = 0.f;
            }
        }
        if (is_a && X.coef) {  // the staged rows are the previous layer's z: its BN + ReLU is applied here (XSource, dz_source.cuh)
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const float sc = __ldg(X.coef + cg * W + j), sh = __ldg(X.coef + cin + cg * W + j);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (rc * 8 + i < valid_rows) {
                        const float y = __fmaf_rn(v[i][j], sc, sh);
                        v[i][j] = X.relu ? fmaxf(y, 0.f) : y;
                    }
            }
        }
        uint8_t *img = is_a ? img_a : img_b;
        const uint32_t lbo = is_a ? lbo_a : lbo_b, split = is_a ? split_a : split_b;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const int ch = cg * W + j;
            const float col[8] = {v[0][j], v[1][j], v[2][j], v[3][j], v[4][j], v[5][j], v[6][j], v[7][j]};
            uint4 hi, lo;
            split8(col, hi, lo);
            uint8_t *dst = img + rc * lbo + (ch >> 3) * wg::kSboP + (ch & 7) * 16;
            if (!(dbg & 8)) {
                *reinterpret_cast<uint4 *>(dst) = hi;
                *reinterpret_cast<uint4 *>(dst + split) = lo;
            }
        }
    }
}

// The dz operand of a pool-only layer, formed on the fly (DzSource): `stage` holds the z rows of the stage, `pool` the pooled maximum /
// pooled gradient / 1 / ties rows of the stage's group, `coef` the per-channel table.  ALL converter threads take part: thread t owns the
// channel quad t % quads for the whole kernel (its seven coefficient quads live in registers) and rows t / quads, + rows_pass, ... of the
// stage; z is overwritten IN PLACE by dz = dz_value(z), after which the ordinary transposing conversion (wgrad_convert) runs on the slot.
// (One 8-row x 4-channel unit per thread, as the conversion itself is organised, left 128-256 of the 480 threads with ~400 dependent
// instructions each: the stage time was that thread's latency.)  `dbsum` is the thread's share of db = column sums of dz.
__device__ __forceinline__ void wgrad_dz_in_place(uint8_t *__restrict__ stage, int c, int valid_rows, const DzCoefRegs &K, const float *__restrict__ pool,
                                                  int relu, int q, int row0, int rows_pass, float4 &dbsum) {
    const float4 pm = *reinterpret_cast<const float4 *>(pool + q * 4), gp = *reinterpret_cast<const float4 *>(pool + c + q * 4),
                 iv = *reinterpret_cast<const float4 *>(pool + 2 * c + q * 4);
    const float4 gsc = make_float4(__fmul_rn(gp.x, iv.x), __fmul_rn(gp.y, iv.y), __fmul_rn(gp.z, iv.z), __fmul_rn(gp.w, iv.w));
    for (int r = row0; r < valid_rows; r += rows_pass) {
        float4 *p = reinterpret_cast<float4 *>(stage + (static_cast<size_t>(r) * c + q * 4) * 4);
        const float4 zz = *p;
        float4 d;
        d = dz_value4(zz, K, pm, gsc, relu);
        *p = d;
        dbsum.x += d.x; dbsum.y += d.y; dbsum.z += d.z; dbsum.w += d.w;
    }
}

// cin % 8 == 0, cin <= 128, cout % 16 == 0, cout <= 256.  One CTA per SM.
// FUSED: dz is not read but formed on the fly from z (DzSource, S.gs % 32 == 0 so that a stage lies inside one group); the column sums
// of dz (= db) come out as per-CTA partials partB[cta][cout].
template <bool FUSED>
__global__ void __launch_bounds__(wg::kThreads, 1)
wgrad_tc_kernel(long long rows, int cin, int cout, long long rows_per_cta, uint32_t tmem_cols, const float *__restrict__ x,
                const float *__restrict__ dz, float *__restrict__ partW, int dbg, DzSource S, float *__restrict__ partB, XSource X, int nring,
                int wunit) {
    using namespace ttc;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t lbo_a = 16 * wg::kSboP, lbo_b = static_cast<uint32_t>(cout / 8) * wg::kSboP;
    const uint32_t split_a = wg::kChunks * lbo_a, split_b = wg::kChunks * lbo_b;
    const uint32_t img_bytes = 2 * split_a + 2 * split_b;                 // one image buffer: [A hi | A lo | B hi | B lo]
    const uint32_t xs_bytes = wg::kRows * static_cast<uint32_t>(cin) * 4, ds_bytes = wg::kRows * static_cast<uint32_t>(cout) * 4;
    const uint32_t pool_bytes = FUSED ? 3u * static_cast<uint32_t>(cout) * 4 : 0u;  // pooled max | pooled gradient | 1 / ties of the stage's group
    const uint32_t stage_bytes = xs_bytes + ds_bytes + pool_bytes;
    uint8_t *ring = smem + 2 * img_bytes;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + 2 * img_bytes + nring * stage_bytes);  // [kMaxRing] ring stage filled
    uint64_t *bar_mma = bar_full + wg::kMaxRing;                                                   // [2] MMAs of an image buffer done
    uint64_t *bar_img = bar_mma + 2;                                                               // [2] image buffer converted
    uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(bar_img + 2);
    float *coef_s = reinterpret_cast<float *>(smem + 2 * img_bytes + nring * stage_bytes + 128);  // FUSED: [7][cout], then dbred [(480 / (cout / 4))][cout]
    float *dbred = coef_s + kDzCoefs * cout;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    if (FUSED) {
        for (int i = threadIdx.x; i < kDzCoefs * cout; i += wg::kThreads) coef_s[i] = __ldg(S.coef + i);
    }
    // measurement aid (f3d_debug_lin_tc_trace): clock64() stamps of CTA 0, thread 0 (issuer) and thread 32 (a converter), 16 slots per stage
    long long *tr = nullptr;
    if (g_lin_trace && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32)) tr = g_lin_trace + (threadIdx.x ? kTraceTiles * 16 : 0);
#define F3D_WT(i) if (tr && s < kTraceTiles) tr[s * 16 + (i)] = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < wg::kMaxRing; ++i) mbar_init(bar_full + i, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_mma + i, 1);
            mbar_init(bar_img + i, wg::kThreads / 32 - 1);  // one arrival per converter warp
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_base_s, tmem_cols);
        tmem_relinquish();
    }
    // the rows of the A images that belong to channels >= cin stay zero for the whole kernel
    for (uint32_t o = threadIdx.x * 16; o < 2 * img_bytes; o += wg::kThreads * 16) *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_s;
    const uint32_t idesc = make_idesc(1, 128, static_cast<uint32_t>(cout));
    const long long rbeg = blockIdx.x * rows_per_cta, rend = rbeg + rows_per_cta < rows ? rbeg + rows_per_cta : rows;
    const int nstages = static_cast<int>((rend - rbeg + wg::kRows - 1) / wg::kRows);

    if (warp == 0) {
        // ---- issuer warp: TMA fetches (one stage ahead) and the MMAs; never touches the data itself
        auto fetch = [&](int s) {  // the fp32 rows of stage s -> ring slot s % nring
            const long long r0 = rbeg + static_cast<long long>(s) * wg::kRows;
            const uint32_t valid = static_cast<uint32_t>(rend - r0 < wg::kRows ? rend - r0 : wg::kRows);
            if (lane == 0) {
                const int slot = s % nring;
                uint8_t *dst = ring + slot * stage_bytes;
                const uint32_t xb = valid * static_cast<uint32_t>(cin) * 4, db = valid * static_cast<uint32_t>(cout) * 4;
                mbar_arrive_expect_tx(bar_full + slot, xb + db + pool_bytes);
                const uint8_t *xsrc = reinterpret_cast<const uint8_t *>(x + r0 * cin);
                const uint8_t *dsrc = reinterpret_cast<const uint8_t *>((FUSED ? S.z : dz) + r0 * cout);
                bulk_g2s_block(dst, xsrc, xb, bar_full + slot);
                bulk_g2s_block(dst + xs_bytes, dsrc, db, bar_full + slot);
                if (FUSED) {
                    const size_t go = static_cast<size_t>(r0 / S.gs) * cout;
                    const uint32_t cb = static_cast<uint32_t>(cout) * 4;
                    bulk_g2s(dst + xs_bytes + ds_bytes, S.pooled + go, cb, bar_full + slot);
                    bulk_g2s(dst + xs_bytes + ds_bytes + cb, S.gpool + go, cb, bar_full + slot);
                    bulk_g2s(dst + xs_bytes + ds_bytes + 2 * cb, S.inv + go, cb, bar_full + slot);
                }
            }
            __syncwarp();
        };
        if (!(dbg & 1)) {
            for (int s = 0; s < nring && s < nstages; ++s) fetch(s);
        }
        uint32_t acc = 0;
        for (int s = 0; s < nstages; ++s) {
            const int b = s & 1;
            F3D_WT(0)
            mbar_wait(bar_img + b, static_cast<uint32_t>((s >> 1) & 1));  // image b converted, ring slot b consumed
            tcgen05_fence_after();
            F3D_WT(1)
            if (!(dbg & 1) && s + nring < nstages) fetch(s + nring);  // image b converted <=> ring slot s % nring consumed
            F3D_WT(3)
            if (elect_one()) {
                const uint32_t a0 = sbase + b * img_bytes, b0 = a0 + 2 * split_a;
                for (int pass = 0; pass < ((dbg & 2) ? 0 : 3); ++pass) {
                    const uint32_t a = a0 + (pass == 2 ? split_a : 0);
                    const uint32_t bb = b0 + (pass == 1 ? split_b : 0);
                    for (int k = 0; k < wg::kRows / 16; ++k) {
                        umma_f16(tmem_base, make_smem_desc(a + k * 2 * lbo_a, lbo_a, wg::kSboP), make_smem_desc(bb + k * 2 * lbo_b, lbo_b, wg::kSboP), idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar_mma + b);
            }
            __syncwarp();
            F3D_WT(2)
        }
    } else {
        // ---- converter warps: ring slot (fp32, row-major) -> operand image (bf16 hi/lo, K = row major)
        float4 dbsum = make_float4(0.f, 0.f, 0.f, 0.f);
        // FUSED: channel quad and first row of this thread in the dz pass (threads beyond rows_pass * quads sit it out)
        const int quads = cout >> 2, ct = threadIdx.x - 32;
        const int rows_pass = (wg::kThreads - 32) / quads;
        const int dq = ct % quads, drow = ct / quads;
        DzCoefRegs K{};
        if (FUSED && drow < rows_pass) {
            auto ld = [&](int i) { return *reinterpret_cast<const float4 *>(coef_s + i * cout + dq * 4); };
            K = DzCoefRegs{ld(0), ld(1), ld(2), ld(3), ld(4), ld(5), ld(6)};
        }
        for (int s = 0; s < nstages; ++s) {
            const int b = s & 1;
            uint8_t *img = smem + b * img_bytes;
            F3D_WT(0)
            if (s >= 2) mbar_wait(bar_mma + b, static_cast<uint32_t>(((s >> 1) - 1) & 1));  // MMAs of stage s-2 have read this image
            F3D_WT(1)
            if (!(dbg & 1)) {
                const int slot = s % nring;
                mbar_wait(bar_full + slot, static_cast<uint32_t>((s / nring) & 1));
                F3D_WT(2)
                const long long r0 = rbeg + static_cast<long long>(s) * wg::kRows;
                const int valid = static_cast<int>(rend - r0 < wg::kRows ? rend - r0 : wg::kRows);
                uint8_t *stage = ring + slot * stage_bytes;
                if constexpr (FUSED) {
                    if (drow < rows_pass)
                        wgrad_dz_in_place(stage + xs_bytes, cout, valid, K, reinterpret_cast<const float *>(stage + xs_bytes + ds_bytes), S.relu, dq, drow,
                                          rows_pass, dbsum);
                }
                if constexpr (FUSED) asm volatile("bar.sync 2, %0;" ::"n"(wg::kThreads - 32) : "memory");  // dz of the whole stage is in place
                if (dbg & 16) {
                } else if (wunit == 4)
                    wgrad_convert_both<4>(img, lbo_a, split_a, img + 2 * split_a, lbo_b, split_b, stage, stage + xs_bytes, cin, cout, valid, X, dbg);
                else if (wunit == 2)
                    wgrad_convert_both<2>(img, lbo_a, split_a, img + 2 * split_a, lbo_b, split_b, stage, stage + xs_bytes, cin, cout, valid, X, dbg);
                else
                    wgrad_convert_both<1>(img, lbo_a, split_a, img + 2 * split_a, lbo_b, split_b, stage, stage + xs_bytes, cin, cout, valid, X, dbg);
            }
            F3D_WT(3)
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_img + b);
            F3D_WT(4)
        }
        if (FUSED && drow < rows_pass)  // this thread's share of db: rows drow, drow + rows_pass, ... of every stage, channel quad dq
            *reinterpret_cast<float4 *>(dbred + drow * cout + dq * 4) = dbsum;
    }
    if (FUSED) {
        __syncthreads();
        const int rows_pass = (wg::kThreads - 32) / (cout >> 2);
        for (int ch = threadIdx.x; ch < cout; ch += wg::kThreads) {
            float t = 0.f;
            for (int r = 0; r < rows_pass; ++r) t += dbred[r * cout + ch];  // fixed order
            partB[static_cast<size_t>(blockIdx.x) * cout + ch] = t;
        }
    }
    // drain: the last MMAs of both image buffers
    if (nstages >= 1) mbar_wait(bar_mma + ((nstages - 1) & 1), static_cast<uint32_t>(((nstages - 1) >> 1) & 1));
    if (nstages >= 2) mbar_wait(bar_mma + ((nstages - 2) & 1), static_cast<uint32_t>(((nstages - 2) >> 1) & 1));
    tcgen05_fence_after();
    const int q = warp & 3;
    const int ci = q * 32 + lane;
    float *pw = partW + (static_cast<size_t>(blockIdx.x) * cin + ci) * cout;
    for (int cc = (warp >> 2) * 32; cc < cout; cc += (wg::kThreads / 128) * 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cc, r);
        tmem_ld_wait();
        if (ci < cin) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (cc + j < cout)
                    *reinterpret_cast<float4 *>(pw + cc + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

static int ttc_num_sms() {
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    return num_sms;
}

static uint32_t pow2_cols(uint32_t need) {
    uint32_t c = 32;
    while (c < need) c <<= 1;
    return c;
}

int lin_tc_kp(int k_real) { return (k_real + 15) / 16 * 16; }
bool lin_tc_supported(int k_real, int nout) { return k_real >= 1 && k_real <= 256 && nout >= 1; }
size_t lin_tc_weight_bytes(int k_real, int nout) { return static_cast<size_t>((nout + 127) / 128) * lin_tc_kp(k_real) * 256 * 3; }

static int lin_tc_ring(int k_real, int nsplit) {
    if (k_real % 8 != 0) return 0;
    const int kp = lin_tc_kp(k_real);
    const size_t op = static_cast<size_t>(nsplit) * (kp / 8) * ttc::lbo_x(nsplit);
    return op + static_cast<size_t>(kRingMax) * ttc::kTile * k_real * 4 + 64 <= 226 * 1024 ? kRingMax : 1;
}

static size_t lin_tc_smem(int k_real, int nsplit) {
    const int kp = lin_tc_kp(k_real);
    const size_t op = static_cast<size_t>(nsplit) * (kp / 8) * ttc::lbo_x(nsplit);
    return op + static_cast<size_t>(lin_tc_ring(k_real, nsplit)) * ttc::kTile * k_real * 4 + 64;
}

// launch plan of a lin_tc call
struct LinPlan {
    bool pipe;     // warp-specialised kernel
    int nt;        // rows per tile
    int nring;     // ring depth
    uint32_t cols; // TMEM columns
    size_t smem;
    int grid;      // row CTAs (the stats partials are 2 per CTA)
};

static int kPipeMinK = 128;  // (f3d_debug_set_lin_tc_pipe_min_k: measurement aid)
static LinPlan lin_tc_plan(long long rows, int k_real, int nsplit, bool fused = false) {
    LinPlan P{};
    const int kp = lin_tc_kp(k_real);
    // fused dz source: three pooled rows ride in every ring slot, the coefficient table and the group-sum scratch follow the barriers
    const size_t slot_extra = fused ? static_cast<size_t>(3) * k_real * 4 : 0, tail_extra = fused ? static_cast<size_t>(kDzCoefs) * k_real * 4 + lp::kConvWarps * 32 * 16 : 0;  // coefficient table + one float4 per converter thread
    // the warp-specialised kernel pays off when the operand conversion + MMAs are the long phases (K >= 128); for narrow
    // inputs the tile is store-bound and the all-warps epilogue of lin_tc_kernel at 3-4 CTAs/SM is faster (measured)
    if (k_real % 8 == 0 && kp >= kPipeMinK) {
        for (int nt = 64; nt >= 32 && !P.pipe; nt -= 32) {
            const size_t img = static_cast<size_t>(nsplit) * (kp / 8) * lp::lbo(nsplit, nt, fused);
            const size_t slot = static_cast<size_t>(nt) * k_real * 4 + slot_extra;
            for (int nring = lp::kMaxRing; nring >= 2; --nring) {
                const size_t smem = 2 * img + nring * slot + 128 + tail_extra;
                if (smem <= 226 * 1024 && static_cast<size_t>(256) * kp <= 2 * img + nring * slot) {
                    P.pipe = true;
                    P.nt = nt;
                    P.nring = nring;
                    P.smem = smem;
                    P.cols = pow2_cols(nsplit * (kp / 2) + 2 * nt);
                    break;
                }
            }
        }
    }
    if (P.pipe && P.cols <= 512) {
        int per_sm = static_cast<int>(512 / P.cols);
        const int by_smem = static_cast<int>((228 * 1024) / (P.smem + 1024));
        if (per_sm > by_smem) per_sm = by_smem;
        if (per_sm > 2) per_sm = 2;
        const long long ntiles = (rows + P.nt - 1) / P.nt;
        const long long g = static_cast<long long>(ttc_num_sms()) * per_sm;
        P.grid = static_cast<int>(ntiles < g ? ntiles : g);
        return P;
    }
    P.pipe = false;
    P.nt = ttc::kTile;
    P.nring = lin_tc_ring(k_real, nsplit) > 0 ? lin_tc_ring(k_real, nsplit) : 1;
    P.cols = pow2_cols(nsplit * (kp / 2) + ttc::kTile);
    P.smem = lin_tc_smem(k_real, nsplit);
    int per_sm = static_cast<int>(512 / P.cols);
    const int by_smem = static_cast<int>((228 * 1024) / (P.smem + 1024));  // 228 KB per SM, 1 KB reserved per CTA
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const long long ntiles = (rows + ttc::kTile - 1) / ttc::kTile;
    const long long g = static_cast<long long>(ttc_num_sms()) * per_sm;
    P.grid = static_cast<int>(ntiles < g ? ntiles : g);
    return P;
}

int lin_tc_grid(long long rows, int k_real, int nsplit) { return lin_tc_plan(rows, k_real, nsplit).grid; }
// rows per tile of the kernel lin_tc() launches for this shape (the pooling epilogue needs 64 = two half groups of 32)
int lin_tc_tile(long long rows, int k_real, int nsplit) {
    const LinPlan P = lin_tc_plan(rows, k_real, nsplit);
    return P.pipe ? P.nt : ttc::kTile;
}

// out (rows, nout) = x (rows, k_real) * A^T (+ bias) (+ gbias[row / gs]) with A[m][k] = src[m*sm + k*sk]; nsplit = 2 (bf16x3)
// or 3 (six product terms, fp32-grade);  part: 2*lin_tc_grid() partials of
// {sum, sum of squares} per channel, or NULL.  wimg: lin_tc_weight_bytes() of scratch.
// fused dz source (S != NULL, nsplit 2): can the contraction read z and form dz itself?  Needs the warp-specialised kernel, tiles inside
// one group, and -- when the per-group sums of dz are wanted -- tiles that ARE groups.
bool lin_tc_dz_supported(long long rows, int k_real, int gs, bool need_group_sums) {
    // the converters give every thread one channel quad: 16 | k_real and (k_real / 4) | 256 threads
    if (k_real % 16 != 0 || (lp::kConvWarps * 32) % (k_real / 4) != 0 || gs <= 0 || rows % gs != 0) return false;
    const LinPlan P = lin_tc_plan(rows, k_real, 2, true);
    if (!P.pipe || gs % P.nt != 0) return false;
    return !need_group_sums || gs == P.nt;
}

int lin_tc(long long rows, int k_real, int nout, const float *x, const float *src, long long sm, long long sk, const float *bias,
           const float *gbias, int gs, float *out, float *part, uint8_t *wimg, int nsplit, cudaStream_t st, const DzSource *S, float *dgb,
           const float *xcoef, int xrelu, float *zext, const float *gamma) {
    const int kp = lin_tc_kp(k_real);
    const XSource X{xcoef, xrelu};
    const PoolEpilogue PE{zext, gamma};
    if (zext && (S != nullptr || rows % 64 != 0 || !gamma || lin_tc_tile(rows, k_real, nsplit) != 64))
        return fail(F3D_ERR_UNSUPPORTED, "lin_tc: the pooling epilogue needs 64-row tiles and rows % 64 == 0");
    if (xcoef && (S != nullptr || k_real % 8 != 0)) return fail(F3D_ERR_UNSUPPORTED, "lin_tc: activation source needs k % 8 == 0 and no dz source");
    const int mblocks = (nout + 127) / 128;
    const long long total = 128LL * kp * mblocks;
    const bool fused = S != nullptr;
    if (fused && (nsplit != 2 || !lin_tc_dz_supported(rows, k_real, S->gs, dgb != nullptr) || mblocks != 1))
        return fail(F3D_ERR_UNSUPPORTED, "lin_tc: fused dz source not supported for this shape");
    lin_prep_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, sm, sk, nout, k_real, kp, mblocks, nsplit, wimg);
    int rc = check_launch("lin_prep_kernel");
    if (rc) return rc;
    const LinPlan P = lin_tc_plan(rows, k_real, nsplit, fused);
    const dim3 grid(mblocks, P.grid);
    cudaError_t e = cudaSuccess;
    const DzSource S0 = fused ? *S : DzSource{};
    // algorithmic bytes: the fp32 rows in and out (each read / written once), weights negligible
    ktimer_begin(nsplit == 3 ? "lin_tc (3-way split, forward)" : fused ? "lin_tc (2-way split, dgrad, dz formed from z)" : "lin_tc (2-way split, dgrad)",
                 4.0 * static_cast<double>(rows) * (k_real + nout), st);
#define F3D_LAUNCH_PIPE(NS, NT, FU)                                                                                                      \
    e = cudaFuncSetAttribute(lin_tc_pipe_kernel<NS, NT, FU>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(P.smem));    \
    if (e == cudaSuccess)                                                                                                                \
        lin_tc_pipe_kernel<NS, NT, FU><<<grid, lp::kThreads, P.smem, st>>>(rows, k_real, kp, nout, P.nring, P.cols, x, wimg, bias, gbias, gs, out, \
                                                                           part, S0, dgb, X, PE);
    if (P.pipe) {
        if (fused && P.nt == 64) { F3D_LAUNCH_PIPE(2, 64, true) }
        else if (fused) { F3D_LAUNCH_PIPE(2, 32, true) }
        else if (nsplit == 3 && P.nt == 64) { F3D_LAUNCH_PIPE(3, 64, false) }
        else if (nsplit == 3) { F3D_LAUNCH_PIPE(3, 32, false) }
        else if (P.nt == 64) { F3D_LAUNCH_PIPE(2, 64, false) }
        else { F3D_LAUNCH_PIPE(2, 32, false) }
        ktimer_end(st);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "lin_tc: cudaFuncSetAttribute");
        return check_launch("lin_tc_pipe_kernel");
    }
#undef F3D_LAUNCH_PIPE
    if (nsplit == 3) {
        e = cudaFuncSetAttribute(lin_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(P.smem));
        if (e == cudaSuccess) lin_tc_kernel<3><<<grid, ttc::kThreads, P.smem, st>>>(rows, k_real, kp, nout, P.nring, P.cols, x, wimg, bias, gbias, gs, out, part, X, PE);
    } else {
        e = cudaFuncSetAttribute(lin_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(P.smem));
        if (e == cudaSuccess) lin_tc_kernel<2><<<grid, ttc::kThreads, P.smem, st>>>(rows, k_real, kp, nout, P.nring, P.cols, x, wimg, bias, gbias, gs, out, part, X, PE);
    }
    ktimer_end(st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "lin_tc: cudaFuncSetAttribute");
    return check_launch("lin_tc_kernel");
}

bool wgrad_tc_supported(int cin, int cout) { return cin % 8 == 0 && cin >= 8 && cin <= 128 && cout % 16 == 0 && cout >= 16 && cout <= 256; }

void wgrad_tc_plan(long long rows, int *grid, long long *rows_per_cta) {
    const long long nst = (rows + wg::kRows - 1) / wg::kRows;
    long long g = ttc_num_sms();
    if (g > nst) g = nst;
    const long long per = (nst + g - 1) / g;
    *rows_per_cta = per * wg::kRows;
    *grid = static_cast<int>((nst + per - 1) / per);
}

static size_t wgrad_tc_smem(int cin, int cout, bool fused, int nring) {
    const size_t img = 2 * static_cast<size_t>(wg::kChunks) * 16 * wg::kSboP + 2 * static_cast<size_t>(wg::kChunks) * (cout / 8) * wg::kSboP;
    return 2 * img + nring * (static_cast<size_t>(wg::kRows) * (cin + cout) * 4 + (fused ? static_cast<size_t>(3) * cout * 4 : 0)) + 128 +
           (fused ? static_cast<size_t>(kDzCoefs) * cout * 4 + static_cast<size_t>(wg::kThreads - 32) * 16 : 0);
}

// as many ring slots (2 .. kMaxRing) as fit into shared memory
static int wgrad_tc_ring(int cin, int cout, bool fused) {
    int nring = wg::kMaxRing;
    while (nring > 2 && wgrad_tc_smem(cin, cout, fused, nring) > 227 * 1024) --nring;
    return nring;
}

// fused dz source: a 32-row stage must lie inside one group and everything must fit in shared memory
bool wgrad_tc_dz_supported(long long rows, int cin, int cout, int gs) {
    return wgrad_tc_supported(cin, cout) && gs > 0 && gs % wg::kRows == 0 && rows % gs == 0 && wgrad_tc_smem(cin, cout, true, 2) <= 227 * 1024;
}

// partW: grid x cin x cout floats.  S != NULL: dz formed on the fly from z (see the kernel), partB: grid x cout column sums of dz.
int wgrad_tc(long long rows, int cin, int cout, const float *x, const float *dz, float *partW, cudaStream_t st, int dbg, const DzSource *S,
             float *partB, const float *xcoef, int xrelu) {
    const XSource X{xcoef, xrelu};
    int grid = 0;
    long long per = 0;
    wgrad_tc_plan(rows, &grid, &per);
    const uint32_t cols = pow2_cols(static_cast<uint32_t>(cout));
    const bool fused = S != nullptr;
    if (fused && (!wgrad_tc_dz_supported(rows, cin, cout, S->gs) || !partB)) return fail(F3D_ERR_UNSUPPORTED, "wgrad_tc: fused dz source not supported for this shape");
    const int nring = wgrad_tc_ring(cin, cout, fused);
    const int units4 = wg::kChunks * (cin + cout) / 4;  // 8-row x 4-channel units per stage; about 384 units keep the 480 converter threads busy
    const int wunit = units4 >= 360 ? 4 : 2 * units4 >= 360 ? 2 : 1;
    const size_t smem = wgrad_tc_smem(cin, cout, fused, nring);
    cudaError_t e = fused ? cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                          : cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "wgrad_tc: cudaFuncSetAttribute");
    ktimer_begin(fused ? "wgrad_tc_kernel (dz formed from z)" : "wgrad_tc_kernel", 4.0 * static_cast<double>(rows) * (cin + cout), st);
    if (fused)
        wgrad_tc_kernel<true><<<grid, wg::kThreads, smem, st>>>(rows, cin, cout, per, cols, x, nullptr, partW, dbg, *S, partB, X, nring, wunit);
    else
        wgrad_tc_kernel<false><<<grid, wg::kThreads, smem, st>>>(rows, cin, cout, per, cols, x, dz, partW, dbg, DzSource{}, nullptr, X, nring, wunit);
    ktimer_end(st);
    return check_launch("wgrad_tc_kernel");
}

}  // namespace f3d

// Measurement aids: which phases of the lin_tc kernels run (see g_lin_dbg; 0 = all), and the contraction alone:
// out (rows, nout) = x (rows, k) * A^T with A[m][k] = W[m * k + kk]; wimg: f3d_debug_lin_tc_weight_bytes(k, nout) of scratch.
F3D_API int f3d_debug_set_lin_tc_phases(int skip_mask) {
    const cudaError_t e = cudaMemcpyToSymbol(f3d::g_lin_dbg, &skip_mask, sizeof(int));
    return e == cudaSuccess ? 0 : f3d::fail(static_cast<int>(e), "debug_set_lin_tc_phases");
}
// buf: device memory of 2 * 64 * 16 long long (threads 0 and 255 of CTA (0,0) of lin_tc_kernel: 16 clock64() slots per tile), NULL = off
F3D_API int f3d_debug_lin_tc_trace(void *buf) {
    const cudaError_t e = cudaMemcpyToSymbol(f3d::g_lin_trace, &buf, sizeof(void *));
    return e == cudaSuccess ? 0 : f3d::fail(static_cast<int>(e), "debug_lin_tc_trace");
}
F3D_API int f3d_debug_set_lin_tc_pipe_min_k(int k) {
    const int prev = f3d::kPipeMinK;
    f3d::kPipeMinK = k;
    return prev;
}
F3D_API size_t f3d_debug_lin_tc_weight_bytes(int k, int nout) { return f3d::lin_tc_weight_bytes(k, nout); }
F3D_API int f3d_debug_lin_tc(long long rows, int k, int nout, const float *x, const float *W, float *out, float *part, void *wimg, int nsplit,
                             void *stream) {
    if (!f3d::lin_tc_supported(k, nout) || (nsplit != 2 && nsplit != 3)) return f3d::fail(F3D_ERR_UNSUPPORTED, "debug_lin_tc: unsupported shape");
    return f3d::lin_tc(rows, k, nout, x, W, k, 1, nullptr, nullptr, 0, out, part, static_cast<uint8_t *>(wimg), nsplit, f3d::as_stream(stream), nullptr,
                       nullptr, nullptr, 0, nullptr, nullptr);
}

// Bring-up / micro-benchmark entry: the wgrad contraction alone.  dbg bit 0 skips the operand staging, bit 1 the MMAs.
F3D_API int f3d_debug_wgrad_tc(long long rows, int cin, int cout, const float *x, const float *dz, float *partW, int dbg, void *stream) {
    if (!f3d::wgrad_tc_supported(cin, cout)) return f3d::fail(F3D_ERR_UNSUPPORTED, "debug_wgrad_tc: unsupported shape");
    return f3d::wgrad_tc(rows, cin, cout, x, dz, partW, f3d::as_stream(stream), dbg, nullptr, nullptr, nullptr, 0);
}
