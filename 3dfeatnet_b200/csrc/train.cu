// train.cu -- training-step pieces of Feat3dNet: the attention-weighted triplet loss (forward + backward) and TF-1 Adam.
//
// Replaces Feat3dNet.get_loss (models/feat3dnet.py:315-357 + models/layers.py:49-62) and get_train_op
// (models/feat3dnet.py:359-375).  The reference materialises two (B,M,M,F) broadcast tensors (201 MB at B=6, M=512, F=32)
// for the pairwise distances and runs one ApplyAdam kernel per variable (~40 launches).  Here:
//   * loss forward  : one warp per anchor descriptor scans the M positive and M negative descriptors of its cloud from
//                     shared memory, keeps the minima (and how many columns tie for them), nothing (B,M,M) reaches HBM;
//                     one CTA per cloud then forms the attention-normalised hinge;
//   * loss backward : the same scan replayed; every anchor i pushes 2 w_i g (f_a - f_p[k*]) to itself and the opposite to
//                     its arg-min column(s).  Column gradients are accumulated WITHOUT atomics: one warp per column
//                     gathers, in ascending anchor order, the anchors that selected it (from a (B,M,M/32) bitmask of the
//                     tied minima the anchor pass records) => bit-reproducible;
//   * Adam          : one launch over a table of (param, grad, m, v, n) records -- lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
//                     theta -= lr_t*m/(sqrt(v)+eps) (the TF-1 form, eps outside the bias correction), grad pre-scaled by
//                     1/world so the data-parallel mean needs no extra pass.
// tf.reduce_min splits the gradient equally among tied minima; so does this code.
#include "common.cuh"

namespace f3d {

constexpr int kLossThreads = 256;

// (m, f) row-major global -> shared memory with row stride f + 1 (conflict-free column walks): a warp copies a row, no index division
__device__ __forceinline__ void stage_rows_padded(float *__restrict__ so, const float *__restrict__ o, int m, int f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = lane; c < f; c += 32) {
        int r = warp;
        for (; r + 7 * nw < m; r += 8 * nw) {  // eight independent loads in flight per thread
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(o + static_cast<size_t>(r + u * nw) * f + c);
#pragma unroll
            for (int u = 0; u < 8; ++u) so[(r + u * nw) * (f + 1) + c] = v[u];
        }
        for (; r < m; r += nw) so[r * (f + 1) + c] = __ldg(o + static_cast<size_t>(r) * f + c);
    }
}

// best[b,i] = min_k |fa[b,i]-fo[b,k]|^2 (squared_difference summed over F, layers.py:60), arg[b,i] = first arg-min,
// ties[b,i] = number of columns attaining the minimum.  One warp per anchor; `other` staged in shared memory.
__global__ void __launch_bounds__(kLossThreads)
loss_min_kernel(int m, int f, const float *__restrict__ fa, const float *__restrict__ fo, float *__restrict__ best,
                int *__restrict__ arg, int *__restrict__ ties) {
    extern __shared__ float so[];  // [m][f+1]
    const int b = blockIdx.y;
    const float *o = fo + static_cast<size_t>(b) * m * f;
    stage_rows_padded(so, o, m, f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m) return;
    const float *a = fa + (static_cast<size_t>(b) * m + i) * f;
    float bd = 3.0e38f;
    int bk = 0x7fffffff, nt = 0;
    for (int k = lane; k < m; k += 32) {
        float d = 0.0f;
        for (int c = 0; c < f; ++c) {
            const float df = a[c] - so[k * (f + 1) + c];
            d += df * df;
        }
        if (d < bd) { bd = d; bk = k; nt = 1; }
        else if (d == bd) { ++nt; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float od = __shfl_xor_sync(kFull, bd, s);
        const int ok = __shfl_xor_sync(kFull, bk, s);
        const int on = __shfl_xor_sync(kFull, nt, s);
        if (od < bd) { bd = od; bk = ok; nt = on; }
        else if (od == bd) { nt += on; bk = min(bk, ok); }
    }
    if (lane == 0) {
        best[static_cast<size_t>(b) * m + i] = bd;
        arg[static_cast<size_t>(b) * m + i] = bk;
        ties[static_cast<size_t>(b) * m + i] = nt;
    }
}

// per cloud: w = att/sum(att) (or 1/M), cost = max(0, sum w*best_p - sum w*best_n + margin); loss = mean over clouds.
// Also emits what the backward needs: gw[b,i] = dL/d(best_p[b,i]) = w_i * 1[cost>0]/B  and  datt.
__global__ void __launch_bounds__(kLossThreads)
loss_reduce_kernel(int nb, int m, float margin, const float *__restrict__ att, const float *__restrict__ best_p,
                   const float *__restrict__ best_n, float *__restrict__ cost, float *__restrict__ gw, float *__restrict__ datt) {
    __shared__ float red[3][kLossThreads / 32];
    __shared__ float tot[3];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *a = att ? att + static_cast<size_t>(b) * m : nullptr;
    const float *bp = best_p + static_cast<size_t>(b) * m, *bn = best_n + static_cast<size_t>(b) * m;
    float sa = 0.f, sp = 0.f, sn = 0.f;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float w = a ? a[i] : 1.0f;
        sa += w;
        sp += w * bp[i];
        sn += w * bn[i];
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        sa += __shfl_xor_sync(kFull, sa, s);
        sp += __shfl_xor_sync(kFull, sp, s);
        sn += __shfl_xor_sync(kFull, sn, s);
    }
    if (lane == 0) { red[0][warp] = sa; red[1][warp] = sp; red[2][warp] = sn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float x = 0.f, y = 0.f, z = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) { x += red[0][w]; y += red[1][w]; z += red[2][w]; }
        tot[0] = x; tot[1] = y; tot[2] = z;
    }
    __syncthreads();
    const float S = tot[0];
    const float sum_p = tot[1] / S, sum_n = tot[2] / S;  // attention_sm * best, summed (feat3dnet.py:341-343); 1/M weights otherwise
    const float c = sum_p - sum_n + margin;
    const float g = c > 0.0f ? 1.0f / static_cast<float>(nb) : 0.0f;  // d(mean_b max(0,c_b))/dc_b
    if (threadIdx.x == 0) cost[b] = fmaxf(c, 0.0f);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const float w = (a ? a[i] : 1.0f) / S;
        gw[static_cast<size_t>(b) * m + i] = g * w;
        if (datt) datt[static_cast<size_t>(b) * m + i] = a ? g * ((bp[i] - bn[i]) - (sum_p - sum_n)) / S : 0.0f;
    }
}

// d fa[b,i,:] (+)= sgn * 2 gw_i / ties_i * sum over the tied arg-min columns k of (fa_i - fo_k).  One warp per anchor replays
// the scan of loss_min_kernel (same arithmetic => `d == best` is exact) and records the tied columns as a bitmask
// sel[b,i,k/32] for the column-side kernel.
__global__ void __launch_bounds__(kLossThreads)
loss_grad_anchor_kernel(int m, int f, float sgn, int accumulate, const float *__restrict__ fa, const float *__restrict__ fo,
                        const float *__restrict__ gw, const float *__restrict__ best, const int *__restrict__ ties,
                        unsigned *__restrict__ sel, float *__restrict__ dfa) {
    extern __shared__ float so[];  // [m][f+1]
    const int b = blockIdx.y;
    const float *o = fo + static_cast<size_t>(b) * m * f;
    stage_rows_padded(so, o, m, f);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m) return;
    const size_t bi = static_cast<size_t>(b) * m + i;
    const float *a = fa + bi * f;
    const float bd = best[bi];
    const int words = (m + 31) >> 5;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < m; k0 += 32) {
        const int k = k0 + lane;
        bool tied = false;
        if (k < m) {
            float d = 0.0f;
            for (int c = 0; c < f; ++c) {
                const float df = a[c] - so[k * (f + 1) + c];
                d += df * df;
            }
            tied = d == bd;
        }
        unsigned mask = __ballot_sync(kFull, tied);
        if (lane == 0) sel[bi * words + (k0 >> 5)] = mask;
        while (mask) {
            const int kk = k0 + __ffs(mask) - 1;
            mask &= mask - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = lane + 32 * q;
                if (c < f) acc[q] += a[c] - so[kk * (f + 1) + c];
            }
        }
    }
    const float scale = sgn * 2.0f * gw[bi] / static_cast<float>(ties[bi]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = lane + 32 * q;
        if (c < f) dfa[bi * f + c] = accumulate ? dfa[bi * f + c] + scale * acc[q] : scale * acc[q];
    }
}

// d fo[b,k,:] = -sgn * sum over the anchors i that selected column k (ascending i) of 2 gw_i (fa_i - fo_k) / ties_i.
// One warp per column: the lanes test 32 anchors at a time against the tie bitmask and the warp then visits only the
// anchors that selected the column (about one per column on average), channels across lanes.
__global__ void __launch_bounds__(kLossThreads)
loss_grad_other_kernel(int m, int f, long long columns, float sgn, const float *__restrict__ fa, const float *__restrict__ fo,
                       const float *__restrict__ gw, const int *__restrict__ ties, const unsigned *__restrict__ sel, float *__restrict__ dfo) {
    const int lane = threadIdx.x & 31;
    const long long bk = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bk >= columns) return;
    const long long b = bk / m;
    const int k = static_cast<int>(bk - b * m);
    const int words = (m + 31) >> 5;
    float ov[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 4; ++q) ov[q] = lane + 32 * q < f ? fo[bk * f + lane + 32 * q] : 0.0f;
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false;
        if (i < m) hit = (__ldg(sel + (b * m + i) * words + (k >> 5)) >> (k & 31)) & 1u;
        unsigned mask = __ballot_sync(kFull, hit);
        while (mask) {
            const long long bi = b * m + i0 + __ffs(mask) - 1;
            mask &= mask - 1;
            const float w = sgn * 2.0f * gw[bi] / static_cast<float>(ties[bi]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = lane + 32 * q;
                if (c < f) acc[q] -= w * (fa[bi * f + c] - ov[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < f) dfo[bk * f + lane + 32 * q] = acc[q];
}

__global__ void loss_mean_kernel(int nb, const float *__restrict__ cost, float *__restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.0f;
        for (int b = 0; b < nb; ++b) s += cost[b];
        loss[0] = s / static_cast<float>(nb);
    }
}

struct AdamRecord {
    float *param;
    const float *grad;
    float *m;
    float *v;
    long long n;
};

// step_dev != NULL: the update count lives on the device (a captured CUDA graph replays the same launch every step):
// t = *step_dev + 1 and the bias-corrected rate is formed here; adam_tick_kernel then advances the counter.
__global__ void adam_kernel(int num_records, const AdamRecord *__restrict__ recs, float lr_t_host, float lr, float b1, float b2, float eps,
                            float grad_scale, const long long *__restrict__ step_dev) {
    __shared__ float lr_s;
    if (threadIdx.x == 0) {
        float v = lr_t_host;
        if (step_dev) {
            const double t = static_cast<double>(*step_dev + 1);
            v = static_cast<float>(static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), t)) / (1.0 - pow(static_cast<double>(b1), t)));
        }
        lr_s = v;
    }
    __syncthreads();
    const float lr_t = lr_s;
    for (int r = blockIdx.y; r < num_records; r += gridDim.y) {
        const AdamRecord R = recs[r];
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < R.n;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const float g = R.grad[i] * grad_scale;
            const float mm = b1 * R.m[i] + (1.0f - b1) * g;
            const float vv = b2 * R.v[i] + (1.0f - b2) * g * g;
            R.m[i] = mm;
            R.v[i] = vv;
            R.param[i] -= lr_t * mm / (sqrtf(vv) + eps);
        }
    }
}

__global__ void adam_tick_kernel(long long *step_dev) { *step_dev += 1; }

}  // namespace f3d

using namespace f3d;

F3D_API size_t f3d_triplet_loss_workspace_bytes(int b, int m) {
    if (b <= 0 || m <= 0) return 256;
    return static_cast<size_t>(b) * m * (4 * 2 + 4 * 2 + 4 * 2 + 4) + static_cast<size_t>(b) * 4 + 2 * static_cast<size_t>(b) * m * ((m + 31) / 32) * 4 + 256;
}

// Forward: loss (1 float).  Backward pieces are produced in the same call when the d* pointers are non-NULL (they are the
// gradients of the LOSS, i.e. already multiplied by 1): dfa, dfp, dfn (b,m,f) and datt (b,m; NULL when att is NULL).
F3D_API int f3d_triplet_loss(int b, int m, int f, float margin, const float *fa, const float *fp, const float *fn, const float *att,
                             float *loss, float *dfa, float *dfp, float *dfn, float *datt, void *workspace, size_t workspace_bytes,
                             void *stream) {
    if (b <= 0 || m <= 0 || f <= 0 || !fa || !fp || !fn || !loss) return fail(F3D_ERR_INVALID_ARGUMENT, "triplet_loss: bad arguments");
    if (!workspace || workspace_bytes < f3d_triplet_loss_workspace_bytes(b, m)) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "triplet_loss: workspace too small");
    const size_t smem = static_cast<size_t>(m) * (f + 1) * sizeof(float);
    if (smem > 200 * 1024) return fail(F3D_ERR_UNSUPPORTED, "triplet_loss: m*(f+1) floats must fit 200 KB of shared memory");
    cudaStream_t st = as_stream(stream);
    const size_t bm = static_cast<size_t>(b) * m;
    float *best_p = static_cast<float *>(workspace), *best_n = best_p + bm, *gw = best_n + bm, *cost = gw + bm;
    int *arg_p = reinterpret_cast<int *>(cost + b), *arg_n = arg_p + bm, *ties_p = arg_n + bm, *ties_n = ties_p + bm;
    cudaError_t e = cudaFuncSetAttribute(loss_min_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "triplet_loss: cudaFuncSetAttribute");
    const dim3 grid((m + kLossThreads / 32 - 1) / (kLossThreads / 32), b);
    loss_min_kernel<<<grid, kLossThreads, smem, st>>>(m, f, fa, fp, best_p, arg_p, ties_p);
    int rc = check_launch("loss_min_kernel");
    if (rc) return rc;
    loss_min_kernel<<<grid, kLossThreads, smem, st>>>(m, f, fa, fn, best_n, arg_n, ties_n);
    rc = check_launch("loss_min_kernel");
    if (rc) return rc;
    loss_reduce_kernel<<<b, kLossThreads, 0, st>>>(b, m, margin, att, best_p, best_n, cost, gw, datt);
    rc = check_launch("loss_reduce_kernel");
    if (rc) return rc;
    loss_mean_kernel<<<1, 32, 0, st>>>(b, cost, loss);
    rc = check_launch("loss_mean_kernel");
    if (rc) return rc;
    if (dfa && dfp && dfn) {
        if (f > 128) return fail(F3D_ERR_UNSUPPORTED, "triplet_loss: the gradient supports descriptors of up to 128 dimensions");
        unsigned *sel_p = reinterpret_cast<unsigned *>(ties_n + bm), *sel_n = sel_p + bm * ((m + 31) / 32);
        e = cudaFuncSetAttribute(loss_grad_anchor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "triplet_loss: cudaFuncSetAttribute");
        loss_grad_anchor_kernel<<<grid, kLossThreads, smem, st>>>(m, f, 1.0f, 0, fa, fp, gw, best_p, ties_p, sel_p, dfa);
        rc = check_launch("loss_grad_anchor_kernel");
        if (rc) return rc;
        loss_grad_anchor_kernel<<<grid, kLossThreads, smem, st>>>(m, f, -1.0f, 1, fa, fn, gw, best_n, ties_n, sel_n, dfa);
        rc = check_launch("loss_grad_anchor_kernel");
        if (rc) return rc;
        const long long columns = static_cast<long long>(bm);
        const unsigned blocks = static_cast<unsigned>((columns + kLossThreads / 32 - 1) / (kLossThreads / 32));
        loss_grad_other_kernel<<<blocks, kLossThreads, 0, st>>>(m, f, columns, 1.0f, fa, fp, gw, ties_p, sel_p, dfp);
        rc = check_launch("loss_grad_other_kernel");
        if (rc) return rc;
        loss_grad_other_kernel<<<blocks, kLossThreads, 0, st>>>(m, f, columns, -1.0f, fa, fn, gw, ties_n, sel_n, dfn);
        rc = check_launch("loss_grad_other_kernel");
        if (rc) return rc;
    }
    return 0;
}

// records: DEVICE array of {param*, grad*, m*, v*, n} (5 x 8 bytes each).  grad_scale folds the 1/world of the data-parallel mean.
// step_dev (device int64, may be NULL): when given, the update count is read from (and then advanced on) the device and
// `step` is ignored -- the form a CUDA graph of the training step needs.
F3D_API int f3d_adam_step(int num_records, const void *records, long long max_n, float lr, float beta1, float beta2, float eps,
                          long long step, float grad_scale, long long *step_dev, void *stream) {
    if (num_records <= 0 || !records || (!step_dev && step <= 0)) return fail(F3D_ERR_INVALID_ARGUMENT, "adam_step: bad arguments");
    double lr_t = 0.0;
    if (!step_dev)
        lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))) /
               (1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
    const unsigned gx = static_cast<unsigned>(max_n <= 0 ? 1 : (max_n + 255) / 256 > 64 ? 64 : (max_n + 255) / 256);
    const dim3 grid(gx, static_cast<unsigned>(num_records < 64 ? num_records : 64));
    adam_kernel<<<grid, 256, 0, as_stream(stream)>>>(num_records, static_cast<const AdamRecord *>(records), static_cast<float>(lr_t), lr,
                                                     beta1, beta2, eps, grad_scale, step_dev);
    int rc = check_launch("adam_kernel");
    if (rc || !step_dev) return rc;
    adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
    return check_launch("adam_tick_kernel");
}

// ---- detector heads (training) --------------------------------------------------------------------------------------
// feature_detection_module's per-cluster heads (models/feat3dnet.py:142-149): attention = softplus(h w_a + b_a),
// orientation = atan2 of the l2-normalised (h W_o + b_o).  As torch ops they were ~30 launches per step (two cuBLAS GEMVs with
// a 9216-long reduction for the weight gradients among them); here: one forward kernel, one backward kernel + a fixed-order
// reduction of the per-CTA weight-gradient partials (deterministic).  One warp per row of h (k channels, k a multiple of 32).
namespace f3d {
constexpr int kHeadWarps = 8;

__device__ __forceinline__ void heads_dots(const float *__restrict__ hrow, const float *__restrict__ wa, const float *__restrict__ wo, int k,
                                           int lane, float &a, float &ox, float &oy) {
    float sa = 0.f, sx = 0.f, sy = 0.f;
    for (int c = lane; c < k; c += 32) {
        const float v = __ldg(hrow + c);
        sa = fmaf(v, __ldg(wa + c), sa);
        sx = fmaf(v, __ldg(wo + 2 * c), sx);
        sy = fmaf(v, __ldg(wo + 2 * c + 1), sy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(kFull, sa, o);
        sx += __shfl_xor_sync(kFull, sx, o);
        sy += __shfl_xor_sync(kFull, sy, o);
    }
    a = sa; ox = sx; oy = sy;
}

__global__ void __launch_bounds__(kHeadWarps * 32)
heads_fwd_kernel(long long rows, int k, const float *__restrict__ h, const float *__restrict__ wa, const float *__restrict__ ba,
                 const float *__restrict__ wo, const float *__restrict__ bo, float *__restrict__ attention, float *__restrict__ orientation) {
    const int lane = threadIdx.x & 31;
    const long long r = static_cast<long long>(blockIdx.x) * kHeadWarps + (threadIdx.x >> 5);
    if (r >= rows) return;
    float a, ox, oy;
    heads_dots(h + r * k, wa, wo, k, lane, a, ox, oy);
    if (lane == 0) {
        a += __ldg(ba);
        ox += __ldg(bo);
        oy += __ldg(bo + 1);
        attention[r] = a > 20.0f ? a : log1pf(expf(a));                       // F.softplus (beta 1, threshold 20)
        const float inv = rsqrtf(fmaxf(ox * ox + oy * oy, 1e-8f));            // tf.nn.l2_normalize(eps = 1e-8)
        orientation[r] = atan2f(oy * inv, ox * inv);
    }
}

// part: [gridDim.x][3 k + 3] = per-CTA sums of {h^T da, h^T dox, h^T doy (interleaved like W_o), sum da, sum dox, sum doy}
__global__ void __launch_bounds__(kHeadWarps * 32)
heads_bwd_kernel(long long rows, int k, const float *__restrict__ h, const float *__restrict__ wa, const float *__restrict__ ba,
                 const float *__restrict__ wo, const float *__restrict__ bo, const float *__restrict__ g_att, const float *__restrict__ g_ori,
                 float *__restrict__ dh, float *__restrict__ part) {
    extern __shared__ float hsm[];  // [kHeadWarps][3 k + 3]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = k / 32;         // channels per lane (k <= 128)
    float aw[4] = {0, 0, 0, 0}, axw[4] = {0, 0, 0, 0}, ayw[4] = {0, 0, 0, 0};
    float sda = 0.f, sdx = 0.f, sdy = 0.f;
    const long long stride = static_cast<long long>(gridDim.x) * kHeadWarps;
    for (long long r = static_cast<long long>(blockIdx.x) * kHeadWarps + warp; r < rows; r += stride) {  // fixed row -> warp map
        const float *hrow = h + r * k;
        float a, ox, oy;
        heads_dots(hrow, wa, wo, k, lane, a, ox, oy);
        a += __ldg(ba);
        ox += __ldg(bo);
        oy += __ldg(bo + 1);
        const float ga = g_att ? __ldg(g_att + r) : 0.f, go = g_ori ? __ldg(g_ori + r) : 0.f;
        const float da = ga * (a > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-a)));  // d softplus = sigmoid
        const float ss = ox * ox + oy * oy;
        const float is = ss > 0.f ? 1.0f / ss : 0.f;                           // d atan2(oy, ox) = (-oy, ox) / (ox^2 + oy^2)
        const float dox = -go * oy * is, doy = go * ox * is;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= per) break;
            const int c = lane + 32 * j;
            const float v = __ldg(hrow + c);
            if (dh) dh[r * k + c] = fmaf(da, __ldg(wa + c), fmaf(dox, __ldg(wo + 2 * c), doy * __ldg(wo + 2 * c + 1)));
            aw[j] = fmaf(v, da, aw[j]);
            axw[j] = fmaf(v, dox, axw[j]);
            ayw[j] = fmaf(v, doy, ayw[j]);
        }
        sda += da; sdx += dox; sdy += doy;
    }
    float *mine = hsm + warp * (3 * k + 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= per) break;
        const int c = lane + 32 * j;
        mine[c] = aw[j];
        mine[k + 2 * c] = axw[j];
        mine[k + 2 * c + 1] = ayw[j];
    }
    if (lane == 0) { mine[3 * k] = sda; mine[3 * k + 1] = sdx; mine[3 * k + 2] = sdy; }
    __syncthreads();
    for (int e = threadIdx.x; e < 3 * k + 3; e += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < kHeadWarps; ++w) s += hsm[w * (3 * k + 3) + e];  // fixed order
        part[static_cast<size_t>(blockIdx.x) * (3 * k + 3) + e] = s;
    }
}

__global__ void heads_reduce_kernel(int nparts, int k, const float *__restrict__ part, float *__restrict__ dwa, float *__restrict__ dba,
                                    float *__restrict__ dwo, float *__restrict__ dbo) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * k + 3) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += part[static_cast<size_t>(p) * (3 * k + 3) + e];  // fixed order
    if (e < k) dwa[e] = s;
    else if (e < 3 * k) dwo[e - k] = s;
    else if (e == 3 * k) dba[0] = s;
    else dbo[e - 3 * k - 1] = s;
}
constexpr int kHeadParts = 296;
}  // namespace f3d

F3D_API size_t f3d_detector_heads_workspace_bytes(int k) { return static_cast<size_t>(f3d::kHeadParts) * (3 * (k > 0 ? k : 0) + 3) * sizeof(float); }

F3D_API int f3d_detector_heads_forward(long long rows, int k, const float *h, const float *w_att, const float *b_att, const float *w_ori,
                                       const float *b_ori, float *attention, float *orientation, void *stream) {
    if (rows < 0 || k <= 0 || k % 32 || k > 128 || !h || !w_att || !b_att || !w_ori || !b_ori || !attention || !orientation)
        return fail(F3D_ERR_INVALID_ARGUMENT, "detector_heads_forward: bad arguments (k must be 32, 64, 96 or 128)");
    if (rows == 0) return 0;
    heads_fwd_kernel<<<static_cast<unsigned>((rows + kHeadWarps - 1) / kHeadWarps), kHeadWarps * 32, 0, as_stream(stream)>>>(
        rows, k, h, w_att, b_att, w_ori, b_ori, attention, orientation);
    return check_launch("heads_fwd_kernel");
}

// g_att / g_ori: gradients of the loss w.r.t. attention / orientation (either may be NULL = zero).  dh may be NULL.
F3D_API int f3d_detector_heads_backward(long long rows, int k, const float *h, const float *w_att, const float *b_att, const float *w_ori,
                                        const float *b_ori, const float *g_att, const float *g_ori, float *dh, float *dw_att, float *db_att,
                                        float *dw_ori, float *db_ori, void *workspace, size_t workspace_bytes, void *stream) {
    if (rows <= 0 || k <= 0 || k % 32 || k > 128 || !h || !w_att || !b_att || !w_ori || !b_ori || !dw_att || !db_att || !dw_ori || !db_ori)
        return fail(F3D_ERR_INVALID_ARGUMENT, "detector_heads_backward: bad arguments (k must be 32, 64, 96 or 128)");
    if (!workspace || workspace_bytes < f3d_detector_heads_workspace_bytes(k))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "detector_heads_backward: workspace too small");
    cudaStream_t st = as_stream(stream);
    const long long want = (rows + kHeadWarps - 1) / kHeadWarps;
    const int grid = static_cast<int>(want < kHeadParts ? want : kHeadParts);
    float *part = static_cast<float *>(workspace);
    const size_t smem = static_cast<size_t>(kHeadWarps) * (3 * k + 3) * sizeof(float);
    heads_bwd_kernel<<<grid, kHeadWarps * 32, smem, st>>>(rows, k, h, w_att, b_att, w_ori, b_ori, g_att, g_ori, dh, part);
    int rc = check_launch("heads_bwd_kernel");
    if (rc) return rc;
    heads_reduce_kernel<<<(3 * k + 3 + 127) / 128, 128, 0, st>>>(grid, k, part, dw_att, db_att, dw_ori, db_ori);
    return check_launch("heads_reduce_kernel");
}
