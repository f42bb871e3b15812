// train_layers.cu -- training-mode 1x1 conv + bias + batch-norm (batch statistics) + ReLU, forward and backward.
//
// Replaces, for is_training=True, the op chain of models/layers.py:11-46 (slim.conv2d with bias) + :225-272
// (tf.nn.moments over [0,1,2], tf.nn.batch_normalization with eps 1e-3, EMA decay 0.9) + the ReLU, and the backward
// graph TensorFlow derives from it.  All eleven conv layers of the detector / descriptor (models/feat3dnet.py:90-187)
// are instances: rows = B*M*S grouped points (per-point layers) or B*M clusters (per-cluster layers), channels last.
//
//   forward : z = x W + b                         conv_fwd_kernel   (fp32 register tiles, per-tile column sums of z, z^2)
//             mean, var = moments(z)              bn_stats_finalize (fixed-order fp64 reduction of the tile partials)
//             y = relu(gamma (z-mean) rsqrt(var+eps) + beta)        bn_apply_kernel
//   backward: g = gy * [y > 0];  sum g, sum g*zhat                  bn_bwd_reduce_kernel + bn_bwd_finalize
//             dz = gamma istd (g - mean(g) - zhat mean(g zhat))     bn_bwd_apply_kernel
//             dW = x^T dz, db = sum dz                              conv_wgrad_kernel (split over rows) + partial_reduce
//             dx = dz W^T                                           conv_fwd_kernel on the transposed weights / dgrad3
// Every reduction runs in a fixed order (no atomics): two runs give identical bits.  The three contractions run either as
// the fp32 FFMA kernels of this file (precision 0) or on the tensor cores (precision 2, train_tc.cu).
#include "dz_source.cuh"
#include "mlp_tile.cuh"

namespace f3d {

constexpr int kLdT = 132;      // leading dimension of the transposed activation tile (128 rows + 4: 16-byte aligned rows)
constexpr int kWgKC = 32;      // rows staged per wgrad iteration
constexpr int kRedBlocks = 592;  // row-chunks of the BN backward reduction (4 per SM)
constexpr int kApplyBlocks = 444;  // row-chunks of the BN backward apply pass (3 resident CTAs per SM at 85 registers)
constexpr int kFwdSplit = 3;     // forward contraction on the tensor cores: hi/mid/lo bf16 split, 6 product terms (fp32-grade)

__device__ __forceinline__ void stage_rows_transposed(float *in_t, const float *__restrict__ x, long long row0, long long rows, int cin) {
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    const long long row = row0 + r;
    const bool ok = row < rows;
    const float *p = x + row * cin;
    if ((cin & 3) == 0) {
        for (int k = h * 4; k < cin; k += 8) {
            const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(p + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            in_t[(k + 0) * kLdT + r] = v.x;
            in_t[(k + 1) * kLdT + r] = v.y;
            in_t[(k + 2) * kLdT + r] = v.z;
            in_t[(k + 3) * kLdT + r] = v.w;
        }
    } else {
        for (int k = h; k < cin; k += 2) in_t[k * kLdT + r] = ok ? __ldg(p + k) : 0.0f;
    }
}

// z[rows x cout] = x[rows x cin] * W[cin x cout] (+ bias); optional per-tile column sums of z and z^2 into
// part[(tile*2+{0,1})*cout + col].  grid = (row tiles, cout / (16*CT)).
template <int CT>
__global__ void __launch_bounds__(kMlpThreads)
conv_fwd_kernel(long long rows, int cin, int cout, const float *__restrict__ x, const float *__restrict__ W,
                const float *__restrict__ bias, const float *__restrict__ gbias, int gs, float *__restrict__ z, float *__restrict__ part) {
    extern __shared__ __align__(16) float sm[];
    float *in_t = sm;
    float *red = sm + static_cast<size_t>(cin) * kLdT;  // [2][16][NC]
    constexpr int NC = 16 * CT;
    const long long tile = blockIdx.x;
    const int col0 = blockIdx.y * NC;
    stage_rows_transposed(in_t, x, tile * kTileRows, rows, cin);
    __syncthreads();
    float acc[8][CT];
    dense_tile<CT>(in_t, cin, W, cout, col0, acc, kLdT);
    const int rg = threadIdx.x & 15, cg = threadIdx.x >> 4;
    const int col = col0 + cg * CT;
    if (bias) {
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            const float bb = __ldg(bias + col + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][j] += bb;
        }
    }
    const long long r0 = tile * kTileRows + rg * 8;
    if (gbias) {  // per-group additive term (see lin_tc_kernel)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + i < rows) {
                const float *gp = gbias + ((r0 + i) / gs) * cout + col;
#pragma unroll
                for (int j = 0; j < CT; ++j) acc[i][j] += __ldg(gp + j);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (r0 + i < rows) {
            float *o = z + (r0 + i) * cout + col;
            if constexpr (CT >= 4) {
#pragma unroll
                for (int j = 0; j < CT; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < CT; ++j) o[j] = acc[i][j];
            }
        }
    }
    if (part) {
#pragma unroll
        for (int j = 0; j < CT; ++j) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (r0 + i < rows) { s += acc[i][j]; q = fmaf(acc[i][j], acc[i][j], q); }
            red[rg * NC + cg * CT + j] = s;
            red[16 * NC + rg * NC + cg * CT + j] = q;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * NC; e += kMlpThreads) {
            const int which = e / NC, c = e - which * NC;
            float s = 0.f;
            for (int g = 0; g < 16; ++g) s += red[which * 16 * NC + g * NC + c];
            part[(tile * 2 + which) * cout + col0 + c] = s;
        }
    }
}

// The rows a thread of a chunk-owning streaming kernel visits.  A block does NOT own one contiguous 1/grid of the rows: with 592 blocks
// streaming equal contiguous chunks in lockstep the kernels ran at 4.4 TB/s (every stream at the same relative offset: the same few
// DRAM channels at a time); walking SMALL chunks of kWalk * rl rows round-robin over the grid they run at 5.4 TB/s.  The order in which a
// thread meets its rows is fixed by (grid, rl), so the partial sums stay deterministic.
constexpr int kWalk = 8;
__device__ int g_row_walk = 1;  // f3d_debug_set_row_walk(0): one contiguous chunk per block (the round-1 walk), for A/B measurements
struct RowWalk {
    long long r, c0, cend, rows, cstep;
    int rl, rlane;
    long long sr;
    __device__ __forceinline__ RowWalk(long long rows_, int rl_, int rlane_) : rows(rows_), rl(rl_), rlane(rlane_) {
        sr = g_row_walk ? static_cast<long long>(kWalk) * rl : (rows + gridDim.x - 1) / gridDim.x;
        c0 = static_cast<long long>(blockIdx.x) * sr;
        cstep = static_cast<long long>(gridDim.x) * sr;
        r = c0 + rlane;
        cend = min(rows, c0 + sr);
    }
    __device__ __forceinline__ bool valid() const { return r < cend; }
    // -> true when the walk jumped to another chunk (group bookkeeping has to be re-derived)
    __device__ __forceinline__ bool next() {
        r += rl;
        if (r < cend) return false;
        c0 += cstep;
        r = c0 + rlane;
        cend = min(rows, c0 + sr);
        return true;
    }
};

// The xyz layers (3 input channels: detector / descriptor conv0, models/feat3dnet.py:43,119 on the grouped coordinates): z = x W + b in
// plain fp32 FMAs.  A streaming kernel bound by the write of z: 3 of 16 K-slots of an MMA would be used, and the tensor-core kernel spent
// its time converting and draining for nothing.  Thread = (channel quad, row lane); per-block column sums of z and z^2 for the BN
// statistics in the layout of the other forward kernels: part[(blk*2 + {0,1})*c + ch].
__global__ void __launch_bounds__(256)
conv3_fwd_kernel(long long rows, int c, const float *__restrict__ x, const float *__restrict__ W, const float *__restrict__ bias,
                 float *__restrict__ z, float *__restrict__ part) {
    __shared__ float4 red[2][256];
    const int cvec = c >> 2, rl = 256 / cvec;
    const int cv = threadIdx.x % cvec, rlane = threadIdx.x / cvec;
    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(W) + cv), w1 = __ldg(reinterpret_cast<const float4 *>(W + c) + cv),
                 w2 = __ldg(reinterpret_cast<const float4 *>(W + 2 * c) + cv);
    const float4 bb = bias ? __ldg(reinterpret_cast<const float4 *>(bias) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (rlane < rl) {
        for (RowWalk wk(rows, rl, rlane); wk.valid(); wk.next()) {
            const long long r = wk.r;
            const float x0 = __ldg(x + r * 3), x1 = __ldg(x + r * 3 + 1), x2 = __ldg(x + r * 3 + 2);
            float4 v;
            v.x = __fmaf_rn(x2, w2.x, __fmaf_rn(x1, w1.x, __fmaf_rn(x0, w0.x, bb.x)));
            v.y = __fmaf_rn(x2, w2.y, __fmaf_rn(x1, w1.y, __fmaf_rn(x0, w0.y, bb.y)));
            v.z = __fmaf_rn(x2, w2.z, __fmaf_rn(x1, w1.z, __fmaf_rn(x0, w0.z, bb.z)));
            v.w = __fmaf_rn(x2, w2.w, __fmaf_rn(x1, w1.w, __fmaf_rn(x0, w0.w, bb.w)));
            reinterpret_cast<float4 *>(z)[static_cast<size_t>(r) * cvec + cv] = v;
            s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
            s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
        }
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < 2 * cvec) {
        const int which = threadIdx.x / cvec, v = threadIdx.x - which * cvec;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < rl; ++l) {
            const float4 u = red[which][l * cvec + v];
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        reinterpret_cast<float4 *>(part + (static_cast<size_t>(blockIdx.x) * 2 + which) * c)[v] = t;
    }
}

// out[w] = sum over the parts of part[p*width + w] in fp64 and a fixed order: 32 slices (p mod 32) per column, each
// ascending, then the slices ascending.  Block = 32 columns x 32 slices.
// OutT = double keeps the total unrounded for a consumer that subtracts two such sums (the BN batch variance below).
template <typename OutT>
__global__ void __launch_bounds__(1024)
partial_reduce_kernel(int nparts, long long width, const float *__restrict__ part, OutT *__restrict__ out) {
    __shared__ double red[32][33];
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const long long w = static_cast<long long>(blockIdx.x) * 32 + lane;
    double s = 0.0;
    if (w < width)
        for (int p = slice; p < nparts; p += 32) s += static_cast<double>(__ldg(part + static_cast<size_t>(p) * width + w));
    red[slice][lane] = s;
    __syncthreads();
    if (slice == 0 && w < width) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += red[k][lane];
        out[w] = static_cast<OutT>(t);
    }
}

// scale / shift of the batch-norm affine map y = z*scale + shift, written with explicit roundings so that the forward
// (bn_stats_finalize -> bn_apply) and the backward kernels, which RECOMPUTE y from z instead of reading it back, agree bit
// for bit
__device__ __forceinline__ void bn_scale_shift(float gamma, float beta, float mean, float var, float eps, float &sc, float &sh) {
    sc = __fmul_rn(gamma, rsqrtf(__fadd_rn(var, eps)));
    sh = __fmaf_rn(-mean, sc, beta);
}

// sums[2][c] -> mean, var (population), coef = {scale = gamma*rsqrt(var+eps), shift = beta - mean*scale}.  The two sums
// arrive in fp64 (partial_reduce_kernel<double>): var = E[z^2] - mean^2 cancels when |mean| >> std, and a sum of squares
// already rounded to fp32 would leave var with a relative error of 2^-24 * mean^2 / var.
__global__ void bn_stats_finalize_kernel(int c, double inv_rows, float eps, const double *__restrict__ sums, const float *__restrict__ gamma,
                                         const float *__restrict__ beta, float *__restrict__ mean, float *__restrict__ var, float *__restrict__ coef) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    const double mu = sums[ch] * inv_rows;
    double v = sums[c + ch] * inv_rows - mu * mu;
    if (v < 0.0) v = 0.0;
    mean[ch] = static_cast<float>(mu);
    var[ch] = static_cast<float>(v);
    float sc, sh;
    bn_scale_shift(gamma[ch], beta[ch], static_cast<float>(mu), static_cast<float>(v), eps, sc, sh);
    coef[ch] = sc;
    coef[c + ch] = sh;
}

__global__ void bn_apply_kernel(long long n4, int c, const float *__restrict__ z, const float *__restrict__ coef, int relu, float *__restrict__ y) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n4) return;
    const int ch = static_cast<int>((e * 4) % c);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(z) + e);
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(coef + ch));
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(coef + c + ch));
    float4 o = make_float4(__fmaf_rn(v.x, sc.x, sh.x), __fmaf_rn(v.y, sc.y, sh.y), __fmaf_rn(v.z, sc.z, sh.z), __fmaf_rn(v.w, sc.w, sh.w));
    if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
    reinterpret_cast<float4 *>(y)[e] = o;
}

// Where the upstream gradient of a layer comes from: a dense (rows, c) tensor, or -- when the layer's output only feeds
// the max-pool over the sample axis -- the pooled maxima, the gradient of the pooled tensor and 1/(number of tied maxima):
// g[row][ch] = y == max ? gpool/ties : 0 is then formed on the fly and the (rows, c) gradient never exists in HBM.
struct GradSource {
    const float *gy;      // dense: (rows, c); pooled: (rows / s, c)
    const float *pooled;  // (rows / s, c) or NULL
    const float *inv;     // (rows / s, c) 1 / ties
    int s;                // samples per group, 0 = dense
    const float *dense2;  // pooled mode only: an additional dense (rows, c) gradient (the activation feeds the pool AND a dense consumer), or NULL
};

// the dense gradient tensor a BN-backward pass has to stream next to z (NULL: none, pooled-only layer)
__device__ __forceinline__ const float *dense_grad(const GradSource &G) { return G.s == 0 ? G.gy : G.dense2; }

// per-thread cache of the pooled maximum and the scaled pooled gradient of the group the thread is walking through
struct PoolCache {
    long long grp = -1;
    float4 pm, gs;
};

// row cursor of a thread that walks rows r, r + step, ...: the group index advances without a 64-bit division per row
struct GroupCursor {
    long long grp = 0;
    int rem = 0;
    __device__ __forceinline__ void init(long long r, int s) {
        if (s > 0) { grp = r / s; rem = static_cast<int>(r - grp * s); }
    }
    __device__ __forceinline__ void advance(int step, int s) {
        if (s > 0) { rem += step; while (rem >= s) { rem -= s; ++grp; } }
    }
};

__device__ __forceinline__ float4 load_grad(const GradSource &G, PoolCache &C, const float4 &dense, long long grp, int cv, int cvec, const float4 &yy) {
    if (G.s == 0) return dense;  // dense mode: the caller has already loaded gy[row] (one iteration ahead)
    if (grp != C.grp) {
        const size_t po = static_cast<size_t>(grp) * cvec + cv;
        C.grp = grp;
        C.pm = __ldg(reinterpret_cast<const float4 *>(G.pooled) + po);
        const float4 gp = __ldg(reinterpret_cast<const float4 *>(G.gy) + po);
        const float4 ic = __ldg(reinterpret_cast<const float4 *>(G.inv) + po);
        C.gs = make_float4(__fmul_rn(gp.x, ic.x), __fmul_rn(gp.y, ic.y), __fmul_rn(gp.z, ic.z), __fmul_rn(gp.w, ic.w));
    }
    float4 g = make_float4(yy.x == C.pm.x ? C.gs.x : 0.f, yy.y == C.pm.y ? C.gs.y : 0.f, yy.z == C.pm.z ? C.gs.z : 0.f, yy.w == C.pm.w ? C.gs.w : 0.f);
    if (G.dense2) { g.x += dense.x; g.y += dense.y; g.z += dense.z; g.w += dense.w; }  // same order as maxpool_bwd followed by the add of the two gradients
    return g;
}

// per row-chunk partial sums of g and g*zhat, g = gy * [y > 0] (ReLU) or gy.  part[(blk*2+{0,1})*c + ch]
// POOL = false: dense gradient only (G.s == 0), no group bookkeeping in registers.
template <bool POOL>
__global__ void __launch_bounds__(256, 4)  // 592 blocks = one wave at 4 blocks per SM: keep the kernel at 64 registers
bn_bwd_reduce_kernel(long long rows, int c, float eps, GradSource G, const float *__restrict__ gamma, const float *__restrict__ beta,
                     const float *__restrict__ z, const float *__restrict__ mean, const float *__restrict__ var, int relu,
                     float *__restrict__ part) {
    __shared__ float4 red[2][256];
    const int cvec = c >> 2, rl = 256 / cvec;
    const int cv = threadIdx.x % cvec, rlane = threadIdx.x / cvec;
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean) + cv);
    const float4 vv = __ldg(reinterpret_cast<const float4 *>(var) + cv);
    const float4 is = make_float4(rsqrtf(vv.x + eps), rsqrtf(vv.y + eps), rsqrtf(vv.z + eps), rsqrtf(vv.w + eps));
    const float4 bga = __ldg(reinterpret_cast<const float4 *>(gamma) + cv), bbe = __ldg(reinterpret_cast<const float4 *>(beta) + cv);
    float4 bsc, bsh;
    bn_scale_shift(bga.x, bbe.x, mu.x, vv.x, eps, bsc.x, bsh.x);
    bn_scale_shift(bga.y, bbe.y, mu.y, vv.y, eps, bsc.y, bsh.y);
    bn_scale_shift(bga.z, bbe.z, mu.z, vv.z, eps, bsc.z, bsh.z);
    bn_scale_shift(bga.w, bbe.w, mu.w, vv.w, eps, bsc.w, bsh.w);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sz = sg;
    PoolCache pc;
    if (rlane < rl) {
        RowWalk wk(rows, rl, rlane);
        GroupCursor gc;
        if (POOL) gc.init(wk.r, G.s);
        for (; wk.valid();) {
            const long long r = wk.r;
            const size_t o = static_cast<size_t>(r) * cvec + cv;
            const float4 zz = __ldg(reinterpret_cast<const float4 *>(z) + o);
            // the layer's activation is recomputed from z (bit-identical to what bn_apply stored) instead of read back
            float4 yy = make_float4(__fmaf_rn(zz.x, bsc.x, bsh.x), __fmaf_rn(zz.y, bsc.y, bsh.y), __fmaf_rn(zz.z, bsc.z, bsh.z), __fmaf_rn(zz.w, bsc.w, bsh.w));
            if (relu) yy = make_float4(fmaxf(yy.x, 0.f), fmaxf(yy.y, 0.f), fmaxf(yy.z, 0.f), fmaxf(yy.w, 0.f));
            float4 g;
            if (POOL) {
                const float4 gd = G.dense2 ? __ldg(reinterpret_cast<const float4 *>(G.dense2) + o) : make_float4(0.f, 0.f, 0.f, 0.f);
                g = load_grad(G, pc, gd, gc.grp, cv, cvec, yy);
            } else {
                g = __ldg(reinterpret_cast<const float4 *>(G.gy) + o);
            }
            if (wk.next()) {
                if (POOL) gc.init(wk.r, G.s);
            } else if (POOL) {
                gc.advance(rl, G.s);
            }
            if (relu) {
                g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
            }
            sg.x += g.x; sg.y += g.y; sg.z += g.z; sg.w += g.w;
            sz.x = fmaf(g.x, (zz.x - mu.x) * is.x, sz.x);
            sz.y = fmaf(g.y, (zz.y - mu.y) * is.y, sz.y);
            sz.z = fmaf(g.z, (zz.z - mu.z) * is.z, sz.z);
            sz.w = fmaf(g.w, (zz.w - mu.w) * is.w, sz.w);
        }
    }
    red[0][threadIdx.x] = sg;
    red[1][threadIdx.x] = sz;
    __syncthreads();
    if (threadIdx.x < 2 * cvec) {
        const int which = threadIdx.x / cvec, v = threadIdx.x - which * cvec;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < rl; ++l) {
            const float4 t = red[which][l * cvec + v];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        reinterpret_cast<float4 *>(part + (static_cast<size_t>(blockIdx.x) * 2 + which) * c)[v] = s;
    }
}

// Pooled mode: g is non-zero only where y equals its group maximum, the tied rows share gpool equally and all have the
// same zhat = (max - beta) / gamma, so  sum g = sum_groups gpool  and  sum g*zhat = sum_groups gpool * zhat_max  (masked by
// max > 0 under ReLU): the two reductions need the (groups, c) tensors only, not a pass over the rows.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_pooled_kernel(long long groups, int c, const float *__restrict__ gpool, const float *__restrict__ pooled,
                            const float *__restrict__ gamma, const float *__restrict__ beta, int relu, float *__restrict__ part) {
    __shared__ float4 red[2][256];
    const int cvec = c >> 2, rl = 256 / cvec;
    const int cv = threadIdx.x % cvec, rlane = threadIdx.x / cvec;
    const long long chunk = (groups + gridDim.x - 1) / gridDim.x;
    const long long gbeg = blockIdx.x * chunk, gend = min(groups, gbeg + chunk);
    const float4 ga = __ldg(reinterpret_cast<const float4 *>(gamma) + cv);
    const float4 be = __ldg(reinterpret_cast<const float4 *>(beta) + cv);
    const float4 ig = make_float4(ga.x != 0.f ? 1.0f / ga.x : 0.f, ga.y != 0.f ? 1.0f / ga.y : 0.f, ga.z != 0.f ? 1.0f / ga.z : 0.f,
                                  ga.w != 0.f ? 1.0f / ga.w : 0.f);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sz = sg;
    if (rlane < rl) {
        for (long long g = gbeg + rlane; g < gend; g += rl) {
            const size_t o = static_cast<size_t>(g) * cvec + cv;
            float4 gp = __ldg(reinterpret_cast<const float4 *>(gpool) + o);
            const float4 pm = __ldg(reinterpret_cast<const float4 *>(pooled) + o);
            if (relu) {
                gp.x = pm.x > 0.f ? gp.x : 0.f; gp.y = pm.y > 0.f ? gp.y : 0.f; gp.z = pm.z > 0.f ? gp.z : 0.f; gp.w = pm.w > 0.f ? gp.w : 0.f;
            }
            sg.x += gp.x; sg.y += gp.y; sg.z += gp.z; sg.w += gp.w;
            sz.x = fmaf(gp.x, (pm.x - be.x) * ig.x, sz.x);
            sz.y = fmaf(gp.y, (pm.y - be.y) * ig.y, sz.y);
            sz.z = fmaf(gp.z, (pm.z - be.z) * ig.z, sz.z);
            sz.w = fmaf(gp.w, (pm.w - be.w) * ig.w, sz.w);
        }
    }
    red[0][threadIdx.x] = sg;
    red[1][threadIdx.x] = sz;
    __syncthreads();
    if (threadIdx.x < 2 * cvec) {
        const int which = threadIdx.x / cvec, v = threadIdx.x - which * cvec;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < rl; ++l) {
            const float4 t = red[which][l * cvec + v];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        reinterpret_cast<float4 *>(part + (static_cast<size_t>(blockIdx.x) * 2 + which) * c)[v] = s;
    }
}

// sums[2][c] (sum g, sum g*zhat) -> dbeta, dgamma, coef2 = {s = gamma*istd, k1 = mean(g), k2 = mean(g*zhat)}
// coef7 (optional): the per-channel table of DzSource (dz_source.cuh): bsc, bsh, s, k1, mu, istd, k2
__global__ void bn_bwd_finalize_kernel(int c, double inv_rows, float eps, const float *__restrict__ sums, const float *__restrict__ gamma,
                                       const float *__restrict__ var, float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ coef2,
                                       const float *__restrict__ beta, const float *__restrict__ mean, float *__restrict__ coef7) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    dbeta[ch] = sums[ch];
    dgamma[ch] = sums[c + ch];
    const float s = gamma[ch] * rsqrtf(var[ch] + eps);
    const float k1 = static_cast<float>(static_cast<double>(sums[ch]) * inv_rows);
    const float k2 = static_cast<float>(static_cast<double>(sums[c + ch]) * inv_rows);
    coef2[ch] = s;
    coef2[c + ch] = k1;
    coef2[2 * c + ch] = k2;
    if (coef7) {
        float bsc, bsh;
        bn_scale_shift(gamma[ch], beta[ch], mean[ch], var[ch], eps, bsc, bsh);
        coef7[ch] = bsc;
        coef7[c + ch] = bsh;
        coef7[2 * c + ch] = s;
        coef7[3 * c + ch] = k1;
        coef7[4 * c + ch] = mean[ch];
        coef7[5 * c + ch] = rsqrtf(var[ch] + eps);
        coef7[6 * c + ch] = k2;
    }
}

// dz = s (g - k1 - zhat k2) and, per row chunk, the column sums of dz (= the bias gradient): partB[blk*c + ch]
template <bool POOL>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(long long rows, int c, float eps, GradSource G, const float *__restrict__ gamma, const float *__restrict__ beta,
                    const float *__restrict__ z,
                    const float *__restrict__ mean, const float *__restrict__ var, const float *__restrict__ coef2, int relu,
                    float *__restrict__ dz, float *__restrict__ partB, const float *__restrict__ x3, float *__restrict__ partW3) {
    // x3 != NULL: the layer has 3 input channels (the xyz layers); dW (3 x c) = x^T dz is accumulated in the same pass:
    // partW3[(blk*3 + k)*c + ch]
    __shared__ float4 red[256];
    const int cvec = c >> 2, rl = 256 / cvec;
    const int cv = threadIdx.x % cvec, rlane = threadIdx.x / cvec;
    const float4 mu = __ldg(reinterpret_cast<const float4 *>(mean) + cv);
    const float4 vv = __ldg(reinterpret_cast<const float4 *>(var) + cv);
    const float4 is = make_float4(rsqrtf(vv.x + eps), rsqrtf(vv.y + eps), rsqrtf(vv.z + eps), rsqrtf(vv.w + eps));
    const float4 bga = __ldg(reinterpret_cast<const float4 *>(gamma) + cv), bbe = __ldg(reinterpret_cast<const float4 *>(beta) + cv);
    float4 bsc, bsh;
    bn_scale_shift(bga.x, bbe.x, mu.x, vv.x, eps, bsc.x, bsh.x);
    bn_scale_shift(bga.y, bbe.y, mu.y, vv.y, eps, bsc.y, bsh.y);
    bn_scale_shift(bga.z, bbe.z, mu.z, vv.z, eps, bsc.z, bsh.z);
    bn_scale_shift(bga.w, bbe.w, mu.w, vv.w, eps, bsc.w, bsh.w);
    const float4 s = __ldg(reinterpret_cast<const float4 *>(coef2) + cv);
    const float4 k1 = __ldg(reinterpret_cast<const float4 *>(coef2 + c) + cv);
    const float4 k2 = __ldg(reinterpret_cast<const float4 *>(coef2 + 2 * c) + cv);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), w0 = sb, w1 = sb, w2 = sb;
    PoolCache pc;
    if (rlane < rl) {
        // the loads of row r + rl are issued before row r is processed (twice the bytes in flight per thread)
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        RowWalk wk(rows, rl, rlane);
        float4 zz_n = zero4, gd_n = zero4;
        const float *gdn = POOL ? G.dense2 : G.gy;
        GroupCursor gc;
        if (POOL) gc.init(wk.r, G.s);
        if (wk.valid()) {
            const size_t o = static_cast<size_t>(wk.r) * cvec + cv;
            zz_n = __ldg(reinterpret_cast<const float4 *>(z) + o);
            if (gdn) gd_n = __ldg(reinterpret_cast<const float4 *>(gdn) + o);
        }
        while (wk.valid()) {
            const long long r = wk.r;
            const size_t o = static_cast<size_t>(r) * cvec + cv;
            const float4 zz = zz_n, gd = gd_n;
            const bool jumped = wk.next();  // wk: the row after r in this thread's walk
            if (wk.valid()) {
                const size_t on = static_cast<size_t>(wk.r) * cvec + cv;
                zz_n = __ldg(reinterpret_cast<const float4 *>(z) + on);
                if (gdn) gd_n = __ldg(reinterpret_cast<const float4 *>(gdn) + on);
            }
            // the layer's activation is recomputed from z (bit-identical to what bn_apply stored) instead of read back
            float4 yy = make_float4(__fmaf_rn(zz.x, bsc.x, bsh.x), __fmaf_rn(zz.y, bsc.y, bsh.y), __fmaf_rn(zz.z, bsc.z, bsh.z), __fmaf_rn(zz.w, bsc.w, bsh.w));
            if (relu) yy = make_float4(fmaxf(yy.x, 0.f), fmaxf(yy.y, 0.f), fmaxf(yy.z, 0.f), fmaxf(yy.w, 0.f));
            float4 g = gd;
            if (POOL) {
                g = load_grad(G, pc, gd, gc.grp, cv, cvec, yy);
                if (jumped) gc.init(wk.r, G.s); else gc.advance(rl, G.s);
            }
            if (relu) {
                g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f; g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
            }
            float4 d;
            // explicit roundings, the same expression as dz_value() (dz_source.cuh): the fused contractions reproduce these bits
            d.x = __fmul_rn(s.x, __fsub_rn(__fsub_rn(g.x, k1.x), __fmul_rn(__fmul_rn(__fsub_rn(zz.x, mu.x), is.x), k2.x)));
            d.y = __fmul_rn(s.y, __fsub_rn(__fsub_rn(g.y, k1.y), __fmul_rn(__fmul_rn(__fsub_rn(zz.y, mu.y), is.y), k2.y)));
            d.z = __fmul_rn(s.z, __fsub_rn(__fsub_rn(g.z, k1.z), __fmul_rn(__fmul_rn(__fsub_rn(zz.z, mu.z), is.z), k2.z)));
            d.w = __fmul_rn(s.w, __fsub_rn(__fsub_rn(g.w, k1.w), __fmul_rn(__fmul_rn(__fsub_rn(zz.w, mu.w), is.w), k2.w)));
            if (dz) reinterpret_cast<float4 *>(dz)[o] = d;
            sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
            if (x3) {
                const float x0 = __ldg(x3 + r * 3), x1 = __ldg(x3 + r * 3 + 1), x2 = __ldg(x3 + r * 3 + 2);
                w0.x = fmaf(x0, d.x, w0.x); w0.y = fmaf(x0, d.y, w0.y); w0.z = fmaf(x0, d.z, w0.z); w0.w = fmaf(x0, d.w, w0.w);
                w1.x = fmaf(x1, d.x, w1.x); w1.y = fmaf(x1, d.y, w1.y); w1.z = fmaf(x1, d.z, w1.z); w1.w = fmaf(x1, d.w, w1.w);
                w2.x = fmaf(x2, d.x, w2.x); w2.y = fmaf(x2, d.y, w2.y); w2.z = fmaf(x2, d.z, w2.z); w2.w = fmaf(x2, d.w, w2.w);
            }
        }
    }
    const int npass = x3 ? 4 : 1;
    for (int pass = 0; pass < npass; ++pass) {
        if (pass) __syncthreads();
        red[threadIdx.x] = pass == 0 ? sb : pass == 1 ? w0 : pass == 2 ? w1 : w2;
        __syncthreads();
        if (threadIdx.x < cvec) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int l = 0; l < rl; ++l) {
                const float4 u = red[l * cvec + threadIdx.x];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            float *dst = pass == 0 ? partB + static_cast<size_t>(blockIdx.x) * c : partW3 + (static_cast<size_t>(blockIdx.x) * 3 + (pass - 1)) * c;
            reinterpret_cast<float4 *>(dst)[threadIdx.x] = t;
        }
    }
}

// dW partials.  CTA (bx, by): rows [bx*rows_per_cta, ...), weight block (ib, jb) = (by % nib, by / nib) of cin_t x cout_t.
// 256 threads = TI x TJ x KQ, each holding an 8x8 block (two 4-wide halves per side so that a warp's LDS.128 are
// contiguous); the KQ groups take alternate staged rows and write separate partials.
// partW[(bx*KQ+kq)][cin][cout], partB[(bx*KQ+kq)][cout] (column sums of dz, written by the ib == 0, ti == 0 threads).
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(long long rows, int cin, int cout, int cin_t, int cout_t, long long rows_per_cta, const float *__restrict__ x,
                  const float *__restrict__ dz, float *__restrict__ partW, float *__restrict__ partB) {
    extern __shared__ __align__(16) float sm[];
    float *xs = sm;                     // [kWgKC][cin_t]
    float *ds = sm + kWgKC * cin_t;     // [kWgKC][cout_t]
    const int nib = (cin + cin_t - 1) / cin_t;
    const int ib = blockIdx.y % nib, jb = blockIdx.y / nib;
    const int TI = cin_t >> 3, TJ = cout_t >> 3, KQ = 256 / (TI * TJ);
    const int ti = threadIdx.x % TI, tj = (threadIdx.x / TI) % TJ, kq = threadIdx.x / (TI * TJ);
    const long long rbeg = blockIdx.x * rows_per_cta, rend = min(rows, rbeg + rows_per_cta);
    float acc[8][8];
    float accb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        accb[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const int hi = cin_t >> 1, hj = cout_t >> 1;
    const bool vec_x = (cin & 3) == 0;
    for (long long r0 = rbeg; r0 < rend; r0 += kWgKC) {
        __syncthreads();
        if (vec_x) {
            const int v = cin_t >> 2;
            for (int e = threadIdx.x; e < kWgKC * v; e += 256) {
                const int r = e / v, cq = e - r * v;
                const long long row = r0 + r;
                const int ch = ib * cin_t + cq * 4;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rend && ch < cin) val = __ldg(reinterpret_cast<const float4 *>(x + row * cin + ch));
                reinterpret_cast<float4 *>(xs)[e] = val;
            }
        } else {
            for (int e = threadIdx.x; e < kWgKC * cin_t; e += 256) {
                const int r = e / cin_t, cq = e - r * cin_t;
                const long long row = r0 + r;
                const int ch = ib * cin_t + cq;
                xs[e] = (row < rend && ch < cin) ? __ldg(x + row * cin + ch) : 0.0f;
            }
        }
        {
            const int v = cout_t >> 2;
            for (int e = threadIdx.x; e < kWgKC * v; e += 256) {
                const int r = e / v, cq = e - r * v;
                const long long row = r0 + r;
                float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rend) val = __ldg(reinterpret_cast<const float4 *>(dz + row * cout + jb * cout_t + cq * 4));
                reinterpret_cast<float4 *>(ds)[e] = val;
            }
        }
        __syncthreads();
        if (kq < KQ) {
            for (int r = kq; r < kWgKC; r += KQ) {
                const float4 a0 = *reinterpret_cast<const float4 *>(xs + r * cin_t + ti * 4);
                const float4 a1 = *reinterpret_cast<const float4 *>(xs + r * cin_t + hi + ti * 4);
                const float4 b0 = *reinterpret_cast<const float4 *>(ds + r * cout_t + tj * 4);
                const float4 b1 = *reinterpret_cast<const float4 *>(ds + r * cout_t + hj + tj * 4);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
                for (int j = 0; j < 8; ++j) accb[j] += b[j];
            }
        }
    }
    if (kq >= KQ) return;
    const size_t p = static_cast<size_t>(blockIdx.x) * KQ + kq;
    float *pw = partW + p * cin * cout;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ci = ib * cin_t + (i < 4 ? ti * 4 + i : hi + ti * 4 + i - 4);
        if (ci >= cin) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cj = jb * cout_t + (half ? hj : 0) + tj * 4;
            *reinterpret_cast<float4 *>(pw + static_cast<size_t>(ci) * cout + cj) =
                make_float4(acc[i][half * 4], acc[i][half * 4 + 1], acc[i][half * 4 + 2], acc[i][half * 4 + 3]);
        }
    }
    if (ib == 0 && ti == 0) {
        float *pb = partB + p * cout;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int cj = jb * cout_t + (half ? hj : 0) + tj * 4;
            *reinterpret_cast<float4 *>(pb + cj) = make_float4(accb[half * 4], accb[half * 4 + 1], accb[half * 4 + 2], accb[half * 4 + 3]);
        }
    }
}

// dx[rows x 3] = dz[rows x cout] * W[3 x cout]^T  (first layer of the descriptor: the gradient continues into the rotation)
__global__ void conv_dgrad3_kernel(long long rows, int cout, const float *__restrict__ dz, const float *__restrict__ W, float *__restrict__ dx) {
    extern __shared__ float ws[];  // [3][cout]
    for (int e = threadIdx.x; e < 3 * cout; e += blockDim.x) ws[e] = W[e];
    __syncthreads();
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float4 *p = reinterpret_cast<const float4 *>(dz + r * cout);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = 0; k < cout / 4; ++k) {
        const float4 v = __ldg(p + k);
        const float *w0 = ws + 4 * k, *w1 = ws + cout + 4 * k, *w2 = ws + 2 * cout + 4 * k;
        a0 = fmaf(v.x, w0[0], a0); a0 = fmaf(v.y, w0[1], a0); a0 = fmaf(v.z, w0[2], a0); a0 = fmaf(v.w, w0[3], a0);
        a1 = fmaf(v.x, w1[0], a1); a1 = fmaf(v.y, w1[1], a1); a1 = fmaf(v.z, w1[2], a1); a1 = fmaf(v.w, w1[3], a1);
        a2 = fmaf(v.x, w2[0], a2); a2 = fmaf(v.y, w2[1], a2); a2 = fmaf(v.z, w2[2], a2); a2 = fmaf(v.w, w2[3], a2);
    }
    dx[r * 3] = a0;
    dx[r * 3 + 1] = a1;
    dx[r * 3 + 2] = a2;
}

// tf.reduce_max over the sample axis: x (groups, s, c) -> out (groups, c), and inv = 1 / (number of samples attaining the
// maximum) in the same single pass.  One thread per (group, 4 channels).
__device__ __forceinline__ void max_count(float v, float &m, int &n) {
    if (v > m) { m = v; n = 1; }
    else if (v == m) ++n;
}

__global__ void maxpool_fwd_kernel(long long groups, int s, int c4, const float *__restrict__ x, float *__restrict__ out, float *__restrict__ inv) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= groups * c4) return;
    const long long g = e / c4;
    const int cv = static_cast<int>(e - g * c4);
    const float4 *p = reinterpret_cast<const float4 *>(x) + g * s * c4 + cv;
    float4 m = __ldg(p);
    int nx = 1, ny = 1, nz = 1, nw = 1;
    int k = 1;
    for (; k + 4 <= s; k += 4) {  // four independent loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(p + static_cast<size_t>(k + u) * c4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            max_count(v[u].x, m.x, nx); max_count(v[u].y, m.y, ny); max_count(v[u].z, m.z, nz); max_count(v[u].w, m.w, nw);
        }
    }
    for (; k < s; ++k) {
        const float4 v = __ldg(p + static_cast<size_t>(k) * c4);
        max_count(v.x, m.x, nx); max_count(v.y, m.y, ny); max_count(v.z, m.z, nz); max_count(v.w, m.w, nw);
    }
    reinterpret_cast<float4 *>(out)[e] = m;
    if (inv) reinterpret_cast<float4 *>(inv)[e] = make_float4(1.0f / nx, 1.0f / ny, 1.0f / nz, 1.0f / nw);
}

// bn_apply + maxpool_fwd in one pass for a layer whose activation ONLY feeds the max-pool over the sample axis (detector conv2,
// descriptor conv_mid): y = act(z * scale + shift) is formed on the fly with the roundings of bn_apply_kernel, so the pooled maxima and
// tie counts are bit-identical to the two-kernel path, and the (rows, c) activation is neither written nor re-read (604 MB each way
// for the detector's 256-channel layer at C4).  The backward recomputes y from z as well and never needed it.
__global__ void bn_apply_pool_kernel(long long groups, int s, int c4, const float *__restrict__ z, const float *__restrict__ coef, int relu,
                                     float *__restrict__ out, float *__restrict__ inv) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= groups * c4) return;
    const long long g = e / c4;
    const int cv = static_cast<int>(e - g * c4);
    const float4 *p = reinterpret_cast<const float4 *>(z) + g * s * c4 + cv;
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(coef) + cv);
    const float4 sh = __ldg(reinterpret_cast<const float4 *>(coef) + c4 + cv);
    auto act = [&](float4 v) -> float4 {
        float4 o = make_float4(__fmaf_rn(v.x, sc.x, sh.x), __fmaf_rn(v.y, sc.y, sh.y), __fmaf_rn(v.z, sc.z, sh.z), __fmaf_rn(v.w, sc.w, sh.w));
        if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
        return o;
    };
    float4 m = act(__ldg(p));
    int nx = 1, ny = 1, nz = 1, nw = 1;
    int k = 1;
    for (; k + 4 <= s; k += 4) {  // four independent loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(p + static_cast<size_t>(k + u) * c4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 y = act(v[u]);
            max_count(y.x, m.x, nx); max_count(y.y, m.y, ny); max_count(y.z, m.z, nz); max_count(y.w, m.w, nw);
        }
    }
    for (; k < s; ++k) {
        const float4 y = act(__ldg(p + static_cast<size_t>(k) * c4));
        max_count(y.x, m.x, nx); max_count(y.y, m.y, ny); max_count(y.z, m.z, nz); max_count(y.w, m.w, nw);
    }
    reinterpret_cast<float4 *>(out)[e] = m;
    reinterpret_cast<float4 *>(inv)[e] = make_float4(1.0f / nx, 1.0f / ny, 1.0f / nz, 1.0f / nw);
}

// The pooled maximum and the tie count of a pool-only layer from the statistics its forward contraction took in the epilogue
// (PoolEpilogue, dz_source.cuh), once the BN scale / shift exist: no pass over the (rows, c) tensor z.  Per (group of 64 rows, channel)
// the two half groups' {largest key m1, multiplicity c1, second largest distinct key m2} are merged; the activation
// y = act(z * scale + shift) is non-decreasing in the key, so max y = y(m1) and the rows with y == max are exactly the c1 rows at m1
// UNLESS y(m2) == y(m1) (two distinct z round to one activation: a handful of (group, channel) pairs per step) -- then the 64 values of
// the group are re-read and counted like bn_apply_pool_kernel does.  Same pooled bits as that kernel, same inv_ties bits wherever the
// pooled value is not a ReLU-clamped zero (there the gradient is masked and the count unused).
__global__ void pool_from_extremes_kernel(long long groups, int c, const float *__restrict__ zext, const float *__restrict__ coef, int relu,
                                          const float *__restrict__ z, float *__restrict__ out, float *__restrict__ inv) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= groups * c) return;
    const long long g = e / c;
    const int ch = static_cast<int>(e - g * c);
    const float sc = __ldg(coef + ch), sh = __ldg(coef + c + ch);
    const float *a = zext + (2 * g) * 3 * c + ch, *b = a + 3 * c;
    const float am1 = __ldg(a), ac1 = __ldg(a + c), am2 = __ldg(a + 2 * c), bm1 = __ldg(b), bc1 = __ldg(b + c), bm2 = __ldg(b + 2 * c);
    float m1, c1, m2;
    if (am1 > bm1) { m1 = am1; c1 = ac1; m2 = fmaxf(am2, bm1); }
    else if (bm1 > am1) { m1 = bm1; c1 = bc1; m2 = fmaxf(bm2, am1); }
    else { m1 = am1; c1 = ac1 + bc1; m2 = fmaxf(am2, bm2); }
    // key -> z: the epilogue negated z where gamma < 0; scale has gamma's sign (scale == 0: every row has y = shift, handled below)
    const bool neg = sc < 0.f || (sc == 0.f && __float_as_uint(sc) != 0u);
    auto act = [&](float key) {
        const float zz = neg ? -key : key;
        const float y = __fmaf_rn(zz, sc, sh);
        return relu ? fmaxf(y, 0.f) : y;
    };
    const float y1 = act(m1);
    float n = c1;
    // rounding ties rows below m1 with the maximum: count them all.  (A maximum clamped to 0 by the ReLU ties every inactive row, but its
    // gradient is masked by y > 0 everywhere it is used, so the count is never read: no recount for it.)
    if (m2 != -INFINITY && !(relu && y1 == 0.f) && act(m2) == y1) {
        const float *p = z + g * 64 * c + ch;
        n = 0.f;
        for (int k = 0; k < 64; ++k) {
            const float y = __fmaf_rn(__ldg(p + static_cast<size_t>(k) * c), sc, sh);
            n += ((relu ? fmaxf(y, 0.f) : y) == y1) ? 1.f : 0.f;
        }
    }
    out[e] = y1;
    inv[e] = 1.0f / n;
}

// its gradient: dx = gout / (number of samples attaining the maximum) where x == max, else 0 (TensorFlow's _MinOrMaxGrad)
__global__ void maxpool_bwd_kernel(long long groups, int s, int c4, const float *__restrict__ x, const float *__restrict__ mx,
                                   const float *__restrict__ inv, const float *__restrict__ gout, float *__restrict__ dx) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= groups * c4) return;
    const long long g = e / c4;
    const int cv = static_cast<int>(e - g * c4);
    const float4 *p = reinterpret_cast<const float4 *>(x) + g * s * c4 + cv;
    float4 *q = reinterpret_cast<float4 *>(dx) + g * s * c4 + cv;
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mx) + e);
    float4 go = __ldg(reinterpret_cast<const float4 *>(gout) + e);
    const float4 ic = __ldg(reinterpret_cast<const float4 *>(inv) + e);
    go.x *= ic.x; go.y *= ic.y; go.z *= ic.z; go.w *= ic.w;
    for (int k = 0; k < s; ++k) {
        const float4 v = __ldg(p + static_cast<size_t>(k) * c4);
        q[static_cast<size_t>(k) * c4] = make_float4(v.x == m.x ? go.x : 0.f, v.y == m.y ? go.y : 0.f, v.z == m.z ? go.z : 0.f, v.w == m.w ? go.w : 0.f);
    }
}

// out[g][ch] = sum over the s rows of group g of x[g*s + k][ch] (ascending k): the gradient of a per-group additive term
__global__ void group_sum_kernel(long long groups, int s, int c4, const float *__restrict__ x, float *__restrict__ out) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= groups * c4) return;
    const long long g = e / c4;
    const int cv = static_cast<int>(e - g * c4);
    const float4 *p = reinterpret_cast<const float4 *>(x) + g * s * c4 + cv;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < s; ++k) {
        const float4 v = __ldg(p + static_cast<size_t>(k) * c4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4 *>(out)[e] = a;
}

__global__ void transpose_kernel(int r, int c, const float *__restrict__ in, float *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= r * c) return;
    const int i = e / c, j = e - i * c;
    out[j * r + i] = in[e];
}

// tensor-core versions of the three contractions (train_tc.cu)
int lin_tc_kp(int k_real);
bool lin_tc_supported(int k_real, int nout);
size_t lin_tc_weight_bytes(int k_real, int nout);
int lin_tc_grid(long long rows, int k_real, int nsplit);
int lin_tc(long long rows, int k_real, int nout, const float *x, const float *src, long long sm, long long sk, const float *bias,
           const float *gbias, int gs, float *out, float *part, uint8_t *wimg, int nsplit, cudaStream_t st, const DzSource *S = nullptr,
           float *dgb = nullptr, const float *xcoef = nullptr, int xrelu = 0, float *zext = nullptr, const float *gamma = nullptr);
int lin_tc_tile(long long rows, int k_real, int nsplit);
bool lin_tc_dz_supported(long long rows, int k_real, int gs, bool need_group_sums);
bool wgrad_tc_supported(int cin, int cout);
bool wgrad_tc_dz_supported(long long rows, int cin, int cout, int gs);
void wgrad_tc_plan(long long rows, int *grid, long long *rows_per_cta);
int wgrad_tc(long long rows, int cin, int cout, const float *x, const float *dz, float *partW, cudaStream_t st, int dbg = 0,
             const DzSource *S = nullptr, float *partB = nullptr, const float *xcoef = nullptr, int xrelu = 0);

// Pool-only layers on the tensor-core path: dz is formed inside the two contractions that consume it instead of being written by
// bn_bwd_apply_kernel and read back twice (f3d_debug_set_fuse_dz(0) restores the three-kernel path; same bits in dW and dx).
static int g_fuse_dz = 1;
// Pool-only layers on the tensor-core path: the max-pool statistics ride in the forward contraction's epilogue (pool_from_extremes_kernel)
// instead of a second pass over z (bn_apply_pool_kernel); f3d_debug_set_epilogue_pool(0) restores that pass.  Same pooled / inv_ties bits.
static int g_epilogue_pool = 1;

static int pick_ct(int cout) { return cout % 128 == 0 ? 8 : cout % 64 == 0 ? 4 : cout % 32 == 0 ? 2 : cout % 16 == 0 ? 1 : 0; }

static int launch_conv_fwd(long long rows, int cin, int cout, const float *x, const float *W, const float *bias, const float *gbias, int gs,
                           float *z, float *part, cudaStream_t st) {
    const int ct = pick_ct(cout);
    if (!ct) return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train: output channels must be a multiple of 16");
    const long long tiles = (rows + kTileRows - 1) / kTileRows;
    const int nc = 16 * ct;
    const size_t smem = (static_cast<size_t>(cin) * kLdT + 2 * 16 * nc) * sizeof(float);
    if (smem > 220 * 1024) return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train: input channels exceed the shared-memory tile (max 384)");
    const dim3 grid(static_cast<unsigned>(tiles), cout / nc);
    cudaError_t e = cudaSuccess;
#define F3D_LAUNCH_FWD(CT)                                                                                              \
    e = cudaFuncSetAttribute(conv_fwd_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    if (e == cudaSuccess) conv_fwd_kernel<CT><<<grid, kMlpThreads, smem, st>>>(rows, cin, cout, x, W, bias, gbias, gs, z, part);
    switch (ct) {
        case 8: F3D_LAUNCH_FWD(8) break;
        case 4: F3D_LAUNCH_FWD(4) break;
        case 2: F3D_LAUNCH_FWD(2) break;
        default: F3D_LAUNCH_FWD(1) break;
    }
#undef F3D_LAUNCH_FWD
    if (e != cudaSuccess) return fail(static_cast<int>(e), "conv_bn_train: cudaFuncSetAttribute");
    return check_launch("conv_fwd_kernel");
}

struct WgradPlan {
    int cin_t, cout_t, kq, ny, gx;
    long long rows_per_cta;
    int nparts;
};

static WgradPlan plan_wgrad(long long rows, int cin, int cout) {
    WgradPlan p;
    p.cin_t = cin >= 128 ? 128 : cin >= 64 ? 64 : cin >= 32 ? 32 : cin >= 16 ? 16 : 8;
    p.cout_t = cout >= 128 ? 128 : cout;  // cout is a multiple of 16 and, below 128, one of 16/32/64
    while ((p.cin_t / 8) * (p.cout_t / 8) > 256) p.cout_t /= 2;
    p.kq = 256 / ((p.cin_t / 8) * (p.cout_t / 8));
    p.ny = ((cin + p.cin_t - 1) / p.cin_t) * (cout / p.cout_t);
    long long gx = (rows + 4 * kWgKC - 1) / (4 * kWgKC);
    const long long cap = (4 * 148 + p.ny - 1) / p.ny;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    p.rows_per_cta = ((rows + gx - 1) / gx + kWgKC - 1) / kWgKC * kWgKC;
    p.gx = static_cast<int>((rows + p.rows_per_cta - 1) / p.rows_per_cta);
    p.nparts = p.gx * p.kq;
    return p;
}

static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace f3d

using namespace f3d;

F3D_API size_t f3d_conv_bn_train_workspace_bytes(long long rows, int cin, int cout) {
    if (rows <= 0 || cin <= 0 || cout <= 0) return 256;
    const size_t tiles = static_cast<size_t>((rows + kTileRows - 1) / kTileRows);
    const size_t stat_parts = tiles > static_cast<size_t>(2 * lin_tc_grid(rows, cin, kFwdSplit)) ? tiles : static_cast<size_t>(2 * lin_tc_grid(rows, cin, kFwdSplit));
    const size_t wimg = lin_tc_weight_bytes(cin, cout) > lin_tc_weight_bytes(cout, cin) ? lin_tc_weight_bytes(cin, cout) : lin_tc_weight_bytes(cout, cin);
    const size_t fwd = align256(stat_parts * 2 * cout * 4) + align256(2 * cout * 8) + align256(2 * cout * 4) + align256(wimg) +
                       align256(static_cast<size_t>((rows + 31) / 32) * 3 * cout * 4);  // pooling statistics of the epilogue (per half group)
    const WgradPlan p = plan_wgrad(rows, cin, cout);
    int tcg = 0;
    long long tcper = 0;
    wgrad_tc_plan(rows, &tcg, &tcper);
    size_t wparts = static_cast<size_t>(p.nparts > tcg ? p.nparts : tcg);
    if (wparts < static_cast<size_t>(kRedBlocks)) wparts = kRedBlocks;
    const size_t bwd = align256(static_cast<size_t>(rows) * cout * 4)                       // dz
                       + align256(static_cast<size_t>(kRedBlocks) * 2 * cout * 4)           // BN reduction partials
                       + align256(2 * cout * 4) + align256(3 * cout * 4)                    // sums, coef2
                       + align256(kDzCoefs * cout * 4)                                      // coefficient table of the fused dz source
                       + align256(wparts * cin * cout * 4)                                  // dW partials
                       + align256(static_cast<size_t>(p.nparts) * cout * 4)                 // db partials (FFMA wgrad scratch)
                       + align256(static_cast<size_t>(kRedBlocks) * cout * 4)               // db partials
                       + align256(static_cast<size_t>(cin) * cout * 4)                      // W^T
                       + align256(wimg);
    return (fwd > bwd ? fwd : bwd) + 256;
}

// Any of three results of the layer on top of z and the batch moments:
//   y != NULL        the (rows, cout) activation is written;
//   pool_s > 0       its max over groups of pool_s rows is (pooled, inv_ties);
//   coef_out != NULL the BN scale / shift [2][cout] is, for a consumer that forms the activation from z itself (XSource).
// xcoef != NULL: x holds the PREVIOUS layer's z and xcoef its scale / shift: the input rows are act(x * scale + shift), formed inside
// the contraction (tensor-core path only).
static int conv_bn_forward_impl(long long rows, int cin, int cout, const float *x, const float *xcoef, int xrelu, const float *W, const float *bias,
                                const float *group_bias, int group_s, const float *gamma, const float *beta, int relu, float eps,
                                float *z, float *y, int pool_s, float *pooled, float *inv_ties, float *coef_out, float *mean, float *var,
                                int precision, void *workspace, size_t workspace_bytes, void *stream) {
    if (group_bias && (group_s <= 0 || rows % group_s != 0))
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: group_bias needs group_s > 0 dividing rows");
    if (rows <= 0 || cin <= 0 || cout <= 0 || !x || !W || !gamma || !beta || !z || !mean || !var)
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: bad arguments");
    if (pool_s < 0 || (pool_s > 0 && (rows % pool_s != 0 || !pooled || !inv_ties)))
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: pool_s must divide rows and comes with pooled / inv_ties");
    if (!y && pool_s == 0 && !coef_out)
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: nothing to produce (y, pooled or coef_out)");
    if (xcoef && !(precision == 2 && lin_tc_supported(cin, cout) && cin % 8 == 0))
        return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train_forward: an activation source needs the tensor-core path and cin % 8 == 0");
    if (precision != 0 && precision != 2) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: precision must be 0 (fp32) or 2 (bf16x3)");
    if (cout % 4 != 0) return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train_forward: output channels must be a multiple of 4");
    if (!workspace || workspace_bytes < f3d_conv_bn_train_workspace_bytes(rows, cin, cout))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "conv_bn_train_forward: workspace too small");
    cudaStream_t st = as_stream(stream);
    const size_t tiles = static_cast<size_t>((rows + kTileRows - 1) / kTileRows);
    // the xyz layers on either precision: plain fp32 FMAs (exact fp32, and faster than padding K = 3 to an MMA)
    const bool xyz3 = cin == 3 && !group_bias && cout % 4 == 0 && 256 % (cout / 4) == 0;
    const int grid3 = static_cast<int>(tiles < static_cast<size_t>(kRedBlocks) ? tiles : static_cast<size_t>(kRedBlocks));
    const bool tc = !xyz3 && precision == 2 && lin_tc_supported(cin, cout);
    const size_t nparts = xyz3 ? static_cast<size_t>(grid3) : tc ? static_cast<size_t>(2 * lin_tc_grid(rows, cin, kFwdSplit)) : tiles;
    const size_t stat_parts = tiles > static_cast<size_t>(2 * lin_tc_grid(rows, cin, kFwdSplit)) ? tiles : static_cast<size_t>(2 * lin_tc_grid(rows, cin, kFwdSplit));
    char *w = static_cast<char *>(workspace);
    float *part = reinterpret_cast<float *>(w);
    w += align256(stat_parts * 2 * cout * 4);
    double *sums = reinterpret_cast<double *>(w);  // fp64 through bn_stats_finalize
    w += align256(2 * cout * 8);
    float *coef = coef_out ? coef_out : reinterpret_cast<float *>(w);
    w += align256(2 * cout * 4);
    uint8_t *wimg = reinterpret_cast<uint8_t *>(w);
    {
        const size_t wimg_b = lin_tc_weight_bytes(cin, cout) > lin_tc_weight_bytes(cout, cin) ? lin_tc_weight_bytes(cin, cout) : lin_tc_weight_bytes(cout, cin);
        w += align256(wimg_b);
    }
    float *zext = reinterpret_cast<float *>(w);
    // pool-only layer on the tensor cores with 64-row tiles = groups: the pool's statistics are taken in the contraction's epilogue
    const bool epi_pool = g_epilogue_pool && tc && pool_s == 64 && rows % 64 == 0 && cin % 8 == 0 && lin_tc_tile(rows, cin, kFwdSplit) == 64;
    int rc = 0;
    if (xyz3) {
        ktimer_begin("conv3_fwd_kernel", 4.0 * static_cast<double>(rows) * (3 + cout), st);
        conv3_fwd_kernel<<<grid3, 256, 0, st>>>(rows, cout, x, W, bias, z, part);
        ktimer_end(st);
        rc = check_launch("conv3_fwd_kernel");
    } else {
        rc = tc ? lin_tc(rows, cin, cout, x, W, 1, cout, bias, group_bias, group_s, z, part, wimg, kFwdSplit, st, nullptr, nullptr, xcoef, xrelu,
                         epi_pool ? zext : nullptr, gamma)
                : launch_conv_fwd(rows, cin, cout, x, W, bias, group_bias, group_s, z, part, st);
    }
    if (rc) return rc;
    partial_reduce_kernel<double><<<(2 * cout + 31) / 32, 1024, 0, st>>>(static_cast<int>(nparts), 2 * cout, part, sums);
    rc = check_launch("partial_reduce_kernel");
    if (rc) return rc;
    bn_stats_finalize_kernel<<<(cout + 127) / 128, 128, 0, st>>>(cout, 1.0 / static_cast<double>(rows), eps, sums, gamma, beta, mean, var, coef);
    rc = check_launch("bn_stats_finalize_kernel");
    if (rc) return rc;
    if (pool_s > 0 && epi_pool) {
        const long long np = (rows / pool_s) * cout;
        pool_from_extremes_kernel<<<static_cast<unsigned>((np + 255) / 256), 256, 0, st>>>(rows / pool_s, cout, zext, coef, relu, z, pooled, inv_ties);
        rc = check_launch("pool_from_extremes_kernel");
        if (rc) return rc;
    } else if (pool_s > 0) {
        const long long groups = rows / pool_s, np = groups * (cout / 4);
        ktimer_begin("bn_apply_pool_kernel", 4.0 * static_cast<double>(rows) * cout, st);  // z read once; the pooled output is 1/pool_s of it
        bn_apply_pool_kernel<<<static_cast<unsigned>((np + 127) / 128), 128, 0, st>>>(groups, pool_s, cout / 4, z, coef, relu, pooled, inv_ties);
        ktimer_end(st);
        rc = check_launch("bn_apply_pool_kernel");
        if (rc) return rc;
    }
    if (!y) return 0;
    const long long n4 = rows * cout / 4;
    ktimer_begin("bn_apply_kernel", 8.0 * static_cast<double>(rows) * cout, st);  // z in, y out
    bn_apply_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(n4, cout, z, coef, relu, y);
    ktimer_end(st);
    return check_launch("bn_apply_kernel");
}

// x (rows,cin), W (cin,cout), bias/gamma/beta (cout) -> z (rows,cout) pre-BN, y (rows,cout) post BN(+ReLU), mean/var (cout)
// = the batch moments (population variance) the caller feeds to the EMA update.
F3D_API int f3d_conv_bn_train_forward(long long rows, int cin, int cout, const float *x, const float *W, const float *bias,
                                      const float *group_bias, int group_s, const float *gamma, const float *beta, int relu, float eps,
                                      float *z, float *y, float *mean, float *var, int precision, void *workspace, size_t workspace_bytes,
                                      void *stream) {
    if (!y) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward: y is NULL");
    return conv_bn_forward_impl(rows, cin, cout, x, nullptr, 0, W, bias, group_bias, group_s, gamma, beta, relu, eps, z, y, 0, nullptr, nullptr, nullptr,
                                mean, var, precision, workspace, workspace_bytes, stream);
}

// The same layer when its activation only feeds tf.reduce_max over groups of pool_s consecutive rows (the sample axis): returns
// pooled (rows/pool_s, cout) and inv_ties = 1 / (number of rows attaining the maximum) instead of y, which is never materialised.
F3D_API int f3d_conv_bn_train_forward_pooled(long long rows, int cin, int cout, const float *x, const float *W, const float *bias,
                                             const float *group_bias, int group_s, const float *gamma, const float *beta, int relu, float eps,
                                             float *z, int pool_s, float *pooled, float *inv_ties, float *mean, float *var, int precision,
                                             void *workspace, size_t workspace_bytes, void *stream) {
    if (pool_s <= 0) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_forward_pooled: pool_s must be positive");
    return conv_bn_forward_impl(rows, cin, cout, x, nullptr, 0, W, bias, group_bias, group_s, gamma, beta, relu, eps, z, nullptr, pool_s, pooled, inv_ties,
                                nullptr, mean, var, precision, workspace, workspace_bytes, stream);
}

// The layer inside a CHAIN of layers whose activations are never materialised.  Inputs: x_coef != NULL says that x is the previous
// layer's pre-BN tensor z and x_coef [2][cin] its BN scale / shift (that layer's coef_out): the rows act(x * scale + shift) (ReLU when
// x_relu) are formed inside the contraction.  Results, any subset (at least one): y (rows, cout) written; pool_s > 0: pooled / inv_ties
// (rows / pool_s, cout); coef_out [2][cout]: scale / shift of THIS layer for the next one in the chain.
F3D_API int f3d_conv_bn_train_forward_chain(long long rows, int cin, int cout, const float *x, const float *x_coef, int x_relu, const float *W,
                                            const float *bias, const float *group_bias, int group_s, const float *gamma, const float *beta, int relu,
                                            float eps, float *z, float *y, int pool_s, float *pooled, float *inv_ties, float *coef_out, float *mean,
                                            float *var, int precision, void *workspace, size_t workspace_bytes, void *stream) {
    return conv_bn_forward_impl(rows, cin, cout, x, x_coef, x_relu, W, bias, group_bias, group_s, gamma, beta, relu, eps, z, y, pool_s, pooled, inv_ties,
                                coef_out, mean, var, precision, workspace, workspace_bytes, stream);
}

// gy (rows, cout) = dL/dy (dense, may be NULL when the pooled gradient is given); pool_s > 0: gpool (rows / pool_s, cout) = gradient of the
// pooled tensor, with pooled / inv_ties of the forward.  Both may be present (the activation feeds the pool and a dense consumer).
// xcoef != NULL: x holds the previous layer's z (see conv_bn_forward_impl).  Outputs: dx (rows,cin; NULL to skip), dW, db, dgamma, dbeta.
static int conv_bn_backward_impl(long long rows, int cin, int cout, const float *x, const float *xcoef, int xrelu, const float *W, const float *gamma,
                                 const float *beta, const float *z, const float *mean, const float *var, int relu, float eps,
                                 const float *gy, int pool_s, const float *pooled, const float *gpool, const float *inv_ties, float *dx, float *dW,
                                 float *db, float *dgamma, float *dbeta, float *dgroup_bias, int group_s, int precision,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    if (dgroup_bias && (group_s <= 0 || rows % group_s != 0))
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: dgroup_bias needs group_s > 0 dividing rows");
    if (rows <= 0 || cin <= 0 || cout <= 0 || !x || !W || !gamma || !beta || !z || !mean || !var || !dW || !db || !dgamma || !dbeta)
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: bad arguments");
    if (pool_s < 0 || (pool_s > 0 && (!pooled || !gpool || !inv_ties || rows % pool_s != 0)))
        return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: pooled gradient needs pooled, gpool, inv_ties and rows % pool_s == 0");
    if (pool_s == 0 && !gy) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: no gradient given (gy or the pooled set)");
    const bool mixed = pool_s > 0 && gy != nullptr;  // pooled + dense gradient
    const GradSource G{pool_s > 0 ? gpool : gy, pool_s > 0 ? pooled : nullptr, pool_s > 0 ? inv_ties : nullptr, pool_s, mixed ? gy : nullptr};
    if (xcoef && !(precision == 2 && cin % 8 == 0 && wgrad_tc_supported(cin, cout)))
        return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train_backward: an activation source needs the tensor-core weight gradient (cin % 8 == 0, <= 128)");
    if (precision != 0 && precision != 2) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: precision must be 0 (fp32) or 2 (bf16x3)");
    if (cout % 16 != 0 || 256 % (cout / 4) != 0)
        return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train_backward: output channels must be 16, 32, 64, 128, 256, 512 or 1024");
    const bool tc_dgrad = precision == 2 && cin != 3 && lin_tc_supported(cout, cin);
    if (dx && cin != 3 && !tc_dgrad && pick_ct(cin) == 0)
        return fail(F3D_ERR_UNSUPPORTED, "conv_bn_train_backward: fp32 dx needs cin == 3 or a multiple of 16");
    if (!workspace || workspace_bytes < f3d_conv_bn_train_workspace_bytes(rows, cin, cout))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "conv_bn_train_backward: workspace too small");
    cudaStream_t st = as_stream(stream);
    const WgradPlan p = plan_wgrad(rows, cin, cout);
    int tcg = 0;
    long long tcper = 0;
    wgrad_tc_plan(rows, &tcg, &tcper);
    size_t wparts = static_cast<size_t>(p.nparts > tcg ? p.nparts : tcg);
    if (wparts < static_cast<size_t>(kRedBlocks)) wparts = kRedBlocks;
    const size_t wimg_bytes = lin_tc_weight_bytes(cin, cout) > lin_tc_weight_bytes(cout, cin) ? lin_tc_weight_bytes(cin, cout) : lin_tc_weight_bytes(cout, cin);
    (void)wimg_bytes;
    char *w = static_cast<char *>(workspace);
    float *dz = reinterpret_cast<float *>(w);
    w += align256(static_cast<size_t>(rows) * cout * 4);
    float *part = reinterpret_cast<float *>(w);
    w += align256(static_cast<size_t>(kRedBlocks) * 2 * cout * 4);
    float *sums = reinterpret_cast<float *>(w);
    w += align256(2 * cout * 4);
    float *coef2 = reinterpret_cast<float *>(w);
    w += align256(3 * cout * 4);
    float *coef7 = reinterpret_cast<float *>(w);
    w += align256(kDzCoefs * cout * 4);
    float *partW = reinterpret_cast<float *>(w);
    w += align256(wparts * cin * cout * 4);
    float *partB_scratch = reinterpret_cast<float *>(w);
    w += align256(static_cast<size_t>(p.nparts) * cout * 4);
    float *partB = reinterpret_cast<float *>(w);
    w += align256(static_cast<size_t>(kRedBlocks) * cout * 4);
    float *Wt = reinterpret_cast<float *>(w);
    w += align256(static_cast<size_t>(cin) * cout * 4);
    uint8_t *wimg = reinterpret_cast<uint8_t *>(w);

    const int nred = static_cast<int>(rows < kRedBlocks ? rows : kRedBlocks);
    int rc = 0;
    int nred1 = nred;
    if (pool_s > 0 && !mixed) {
        const long long groups = rows / pool_s;
        nred1 = static_cast<int>(groups < kRedBlocks ? groups : kRedBlocks);
        bn_bwd_reduce_pooled_kernel<<<nred1, 256, 0, st>>>(groups, cout, gpool, pooled, gamma, beta, relu, part);
        rc = check_launch("bn_bwd_reduce_pooled_kernel");
    } else {
        ktimer_begin("bn_bwd_reduce_kernel", 8.0 * static_cast<double>(rows) * cout, st);  // gy and z in
        if (pool_s > 0)
            bn_bwd_reduce_kernel<true><<<nred, 256, 0, st>>>(rows, cout, eps, G, gamma, beta, z, mean, var, relu, part);
        else
            bn_bwd_reduce_kernel<false><<<nred, 256, 0, st>>>(rows, cout, eps, G, gamma, beta, z, mean, var, relu, part);
        ktimer_end(st);
        rc = check_launch("bn_bwd_reduce_kernel");
    }
    if (rc) return rc;
    partial_reduce_kernel<float><<<(2 * cout + 31) / 32, 1024, 0, st>>>(nred1, 2 * cout, part, sums);
    rc = check_launch("partial_reduce_kernel");
    if (rc) return rc;
    const bool w3 = cin == 3;  // the xyz layers: dW rides along with the dz pass (partials in the dW-partials area: 3*cout per chunk)
    // pool-only layer on the tensor-core path: dz is formed inside wgrad / dgrad from z and the pooled tensors, never stored
    const bool fuse_dz = g_fuse_dz && pool_s > 0 && !mixed && precision == 2 && !w3 && dx && tc_dgrad && wgrad_tc_dz_supported(rows, cin, cout, pool_s) &&
                         lin_tc_dz_supported(rows, cout, pool_s, dgroup_bias != nullptr) && (!dgroup_bias || group_s == pool_s) && cin <= 128;
    bn_bwd_finalize_kernel<<<(cout + 127) / 128, 128, 0, st>>>(cout, 1.0 / static_cast<double>(rows), eps, sums, gamma, var, dgamma, dbeta, coef2, beta,
                                                               mean, fuse_dz ? coef7 : nullptr);
    rc = check_launch("bn_bwd_finalize_kernel");
    if (rc) return rc;
    if (fuse_dz) {
        const DzSource S{z, coef7, pooled, gpool, inv_ties, pool_s, relu};
        rc = wgrad_tc(rows, cin, cout, x, nullptr, partW, st, 0, &S, partB, xcoef, xrelu);
        if (rc) return rc;
        const long long nw = static_cast<long long>(cin) * cout;
        partial_reduce_kernel<float><<<static_cast<unsigned>((nw + 31) / 32), 1024, 0, st>>>(tcg, nw, partW, dW);
        rc = check_launch("partial_reduce_kernel");
        if (rc) return rc;
        partial_reduce_kernel<float><<<(cout + 31) / 32, 1024, 0, st>>>(tcg, cout, partB, db);
        rc = check_launch("partial_reduce_kernel");
        if (rc) return rc;
        return lin_tc(rows, cout, cin, nullptr, W, cout, 1, nullptr, nullptr, 0, dx, nullptr, wimg, 2, st, &S, dgroup_bias);
    }
    const int napp = static_cast<int>(rows < kApplyBlocks ? rows : kApplyBlocks);
    // gy (dense mode only) and z in, dz out
    ktimer_begin("bn_bwd_apply_kernel", (pool_s > 0 && !mixed ? 8.0 : 12.0) * static_cast<double>(rows) * cout -
                                            ((w3 && !dx && !dgroup_bias) ? 4.0 * static_cast<double>(rows) * cout : 0.0), st);
    // an xyz layer whose input needs no gradient (the detector's conv0): dW and db come out of this pass, nobody reads dz
    float *dz_out = (w3 && !dx && !dgroup_bias) ? nullptr : dz;
    if (pool_s > 0)
        bn_bwd_apply_kernel<true><<<napp, 256, 0, st>>>(rows, cout, eps, G, gamma, beta, z, mean, var, coef2, relu, dz_out, partB, w3 ? x : nullptr, w3 ? partW : nullptr);
    else
        bn_bwd_apply_kernel<false><<<napp, 256, 0, st>>>(rows, cout, eps, G, gamma, beta, z, mean, var, coef2, relu, dz_out, partB, w3 ? x : nullptr, w3 ? partW : nullptr);
    ktimer_end(st);
    rc = check_launch("bn_bwd_apply_kernel");
    if (rc) return rc;
    partial_reduce_kernel<float><<<(cout + 31) / 32, 1024, 0, st>>>(napp, cout, partB, db);
    rc = check_launch("partial_reduce_kernel");
    if (rc) return rc;

    if (dgroup_bias) {
        const long long n = (rows / group_s) * (cout / 4);
        group_sum_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, st>>>(rows / group_s, group_s, cout / 4, dz, dgroup_bias);
        rc = check_launch("group_sum_kernel");
        if (rc) return rc;
    }
    const long long nw = static_cast<long long>(cin) * cout;
    if (w3) {
        partial_reduce_kernel<float><<<static_cast<unsigned>((nw + 31) / 32), 1024, 0, st>>>(napp, nw, partW, dW);
    } else if (precision == 2 && wgrad_tc_supported(cin, cout)) {
        rc = wgrad_tc(rows, cin, cout, x, dz, partW, st, 0, nullptr, nullptr, xcoef, xrelu);
        if (rc) return rc;
        partial_reduce_kernel<float><<<static_cast<unsigned>((nw + 31) / 32), 1024, 0, st>>>(tcg, nw, partW, dW);
    } else {
        const size_t smem = static_cast<size_t>(kWgKC) * (p.cin_t + p.cout_t) * sizeof(float);
        conv_wgrad_kernel<<<dim3(p.gx, p.ny), 256, smem, st>>>(rows, cin, cout, p.cin_t, p.cout_t, p.rows_per_cta, x, dz, partW, partB_scratch);
        rc = check_launch("conv_wgrad_kernel");
        if (rc) return rc;
        partial_reduce_kernel<float><<<static_cast<unsigned>((nw + 31) / 32), 1024, 0, st>>>(p.nparts, nw, partW, dW);
    }
    rc = check_launch("partial_reduce_kernel");
    if (rc) return rc;
    if (dx) {
        if (cin == 3) {
            conv_dgrad3_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 3 * cout * sizeof(float), st>>>(rows, cout, dz, W, dx);
            rc = check_launch("conv_dgrad3_kernel");
        } else if (tc_dgrad) {
            rc = lin_tc(rows, cout, cin, dz, W, cout, 1, nullptr, nullptr, 0, dx, nullptr, wimg, 2, st);  // A[m = ci][k = co] = W[ci][co]
        } else {
            transpose_kernel<<<(cin * cout + 255) / 256, 256, 0, st>>>(cin, cout, W, Wt);
            rc = check_launch("transpose_kernel");
            if (rc) return rc;
            rc = launch_conv_fwd(rows, cout, cin, dz, Wt, nullptr, nullptr, 0, dx, nullptr, st);
        }
    }
    return rc;
}

F3D_API int f3d_conv_bn_train_backward(long long rows, int cin, int cout, const float *x, const float *W, const float *gamma,
                                       const float *beta, const float *z, const float *y, const float *mean, const float *var, int relu, float eps,
                                       const float *gy, int pool_s, const float *pooled, const float *inv_ties, float *dx, float *dW,
                                       float *db, float *dgamma, float *dbeta, float *dgroup_bias, int group_s, int precision,
                                       void *workspace, size_t workspace_bytes, void *stream) {
    (void)y;  // the activation is recomputed from z (bit-identical), not read back
    if (!gy) return fail(F3D_ERR_INVALID_ARGUMENT, "conv_bn_train_backward: gy is NULL");
    // gy is the dense gradient, or -- pool_s > 0 -- the gradient of the pooled tensor
    return conv_bn_backward_impl(rows, cin, cout, x, nullptr, 0, W, gamma, beta, z, mean, var, relu, eps, pool_s > 0 ? nullptr : gy, pool_s, pooled,
                                 pool_s > 0 ? gy : nullptr, inv_ties, dx, dW, db, dgamma, dbeta, dgroup_bias, group_s, precision, workspace,
                                 workspace_bytes, stream);
}

// Backward of a layer of a chain (f3d_conv_bn_train_forward_chain): x_coef as there; gy (dense, rows x cout) and / or the pooled set
// (pool_s, pooled, gpool, inv_ties): when both are given the layer's activation fed the pool and a dense consumer and the two gradients
// are summed on the fly (no maxpool backward pass, no add).
F3D_API int f3d_conv_bn_train_backward_chain(long long rows, int cin, int cout, const float *x, const float *x_coef, int x_relu, const float *W,
                                             const float *gamma, const float *beta, const float *z, const float *mean, const float *var, int relu,
                                             float eps, const float *gy, int pool_s, const float *pooled, const float *gpool, const float *inv_ties,
                                             float *dx, float *dW, float *db, float *dgamma, float *dbeta, float *dgroup_bias, int group_s,
                                             int precision, void *workspace, size_t workspace_bytes, void *stream) {
    return conv_bn_backward_impl(rows, cin, cout, x, x_coef, x_relu, W, gamma, beta, z, mean, var, relu, eps, gy, pool_s, pooled, gpool, inv_ties, dx,
                                 dW, db, dgamma, dbeta, dgroup_bias, group_s, precision, workspace, workspace_bytes, stream);
}

// out (rows, nout) = x (rows, k) W, W (k, nout) row-major, and its gradients, on the tensor-core contractions of the training layers: the
// per-cluster term of conv_mid -- pooled (B*M, C) times the lower rows of the layer's weights, the pooled half of concat([h, tile(max h)])
// (models/feat3dnet.py:60-69) -- and anything else that is a plain matrix product of row-major fp32 tensors.  Forward with the 3-way split
// (fp32-grade), gradients with the 2-way split, like the layers.  k % 8 == 0, k <= 128, nout % 16 == 0, nout <= 256 for dW.
F3D_API size_t f3d_linear_workspace_bytes(long long rows, int k, int nout) {
    if (rows <= 0 || k <= 0 || nout <= 0) return 256;
    const size_t wimg = lin_tc_weight_bytes(k, nout) > lin_tc_weight_bytes(nout, k) ? lin_tc_weight_bytes(k, nout) : lin_tc_weight_bytes(nout, k);
    int tcg = 0;
    long long per = 0;
    wgrad_tc_plan(rows, &tcg, &per);
    return align256(wimg) + align256(static_cast<size_t>(tcg) * k * nout * 4) + 256;
}

F3D_API int f3d_linear_forward(long long rows, int k, int nout, const float *x, const float *W, float *out, void *workspace, size_t workspace_bytes,
                               void *stream) {
    if (rows <= 0 || k <= 0 || nout <= 0 || !x || !W || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "linear_forward: bad arguments");
    if (!lin_tc_supported(k, nout)) return fail(F3D_ERR_UNSUPPORTED, "linear_forward: k must be <= 256");
    if (!workspace || workspace_bytes < f3d_linear_workspace_bytes(rows, k, nout)) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "linear_forward: workspace too small");
    // A[m = output channel][kk] = W[kk * nout + m]
    return lin_tc(rows, k, nout, x, W, 1, nout, nullptr, nullptr, 0, out, nullptr, static_cast<uint8_t *>(workspace), kFwdSplit, as_stream(stream));
}

// g (rows, nout) = dL/dout -> dx (rows, k; NULL to skip) = g W^T, dW (k, nout; NULL to skip) = x^T g
F3D_API int f3d_linear_backward(long long rows, int k, int nout, const float *x, const float *W, const float *g, float *dx, float *dW, void *workspace,
                                size_t workspace_bytes, void *stream) {
    if (rows <= 0 || k <= 0 || nout <= 0 || !x || !W || !g) return fail(F3D_ERR_INVALID_ARGUMENT, "linear_backward: bad arguments");
    if (!workspace || workspace_bytes < f3d_linear_workspace_bytes(rows, k, nout)) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "linear_backward: workspace too small");
    cudaStream_t st = as_stream(stream);
    char *w = static_cast<char *>(workspace);
    uint8_t *wimg = reinterpret_cast<uint8_t *>(w);
    const size_t wimg_b = lin_tc_weight_bytes(k, nout) > lin_tc_weight_bytes(nout, k) ? lin_tc_weight_bytes(k, nout) : lin_tc_weight_bytes(nout, k);
    float *partW = reinterpret_cast<float *>(w + align256(wimg_b));
    int rc = 0;
    if (dW) {
        if (!wgrad_tc_supported(k, nout)) return fail(F3D_ERR_UNSUPPORTED, "linear_backward: dW needs k % 8 == 0, k <= 128, nout % 16 == 0, nout <= 256");
        int tcg = 0;
        long long per = 0;
        wgrad_tc_plan(rows, &tcg, &per);
        rc = wgrad_tc(rows, k, nout, x, g, partW, st);
        if (rc) return rc;
        const long long nw = static_cast<long long>(k) * nout;
        partial_reduce_kernel<float><<<static_cast<unsigned>((nw + 31) / 32), 1024, 0, st>>>(tcg, nw, partW, dW);
        rc = check_launch("partial_reduce_kernel");
        if (rc) return rc;
    }
    if (dx) {
        if (!lin_tc_supported(nout, k)) return fail(F3D_ERR_UNSUPPORTED, "linear_backward: nout must be <= 256 for dx");
        rc = lin_tc(rows, nout, k, g, W, nout, 1, nullptr, nullptr, 0, dx, nullptr, wimg, 2, st);  // A[m = kk][co] = W[kk * nout + co]
    }
    return rc;
}

// Measurement / test aid: 0 = pool-only layers write dz with bn_bwd_apply_kernel and read it back in wgrad and dgrad (the path the fused
// one is checked against, bit for bit); 1 (default) = dz formed inside the contractions.  Returns the previous setting.
F3D_API int f3d_debug_set_fuse_dz(int on) {
    const int prev = g_fuse_dz;
    g_fuse_dz = on ? 1 : 0;
    return prev;
}

// Measurement aid: 0 = the chunk-owning streaming kernels give every block one contiguous 1/grid of the rows (round 1), 1 (default) =
// small chunks round-robin over the grid.  Different summation order, same determinism.  Returns 0.
F3D_API int f3d_debug_set_row_walk(int on) {
    const int v = on ? 1 : 0;
    const cudaError_t e = cudaMemcpyToSymbol(f3d::g_row_walk, &v, sizeof(int));
    return e == cudaSuccess ? 0 : fail(static_cast<int>(e), "debug_set_row_walk");
}

// Measurement / test aid: 0 = pool-only layers take the pooled maximum and the tie counts with a pass over z (bn_apply_pool_kernel),
// 1 (default) = from the statistics of the forward contraction's epilogue.  Both give the same bits.  Returns the previous setting.
F3D_API int f3d_debug_set_epilogue_pool(int on) {
    const int prev = g_epilogue_pool;
    g_epilogue_pool = on ? 1 : 0;
    return prev;
}

// tf.reduce_max(x, axis=[2]) of models/feat3dnet.py:138,147,182 on a channels-last (groups, s, c) tensor, c % 4 == 0.
// inv_ties (groups, c; may be NULL) receives 1 / (number of samples attaining the maximum), which the gradient needs.
F3D_API int f3d_maxpool_samples_forward(long long groups, int s, int c, const float *x, float *out, float *inv_ties, void *stream) {
    if (groups <= 0 || s <= 0 || c <= 0 || (c & 3) || !x || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "maxpool_samples_forward: bad arguments");
    const long long n = groups * (c / 4);
    maxpool_fwd_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, as_stream(stream)>>>(groups, s, c / 4, x, out, inv_ties);
    return check_launch("maxpool_fwd_kernel");
}

// gradient of the above: the samples that attain the maximum share gout equally.
F3D_API int f3d_maxpool_samples_backward(long long groups, int s, int c, const float *x, const float *out, const float *inv_ties,
                                         const float *gout, float *dx, void *stream) {
    if (groups <= 0 || s <= 0 || c <= 0 || (c & 3) || !x || !out || !inv_ties || !gout || !dx)
        return fail(F3D_ERR_INVALID_ARGUMENT, "maxpool_samples_backward: bad arguments");
    const long long n = groups * (c / 4);
    maxpool_bwd_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, as_stream(stream)>>>(groups, s, c / 4, x, out, inv_ties, gout, dx);
    return check_launch("maxpool_bwd_kernel");
}
