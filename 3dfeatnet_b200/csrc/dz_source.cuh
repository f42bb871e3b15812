// dz_source.cuh -- operands formed on the fly inside the contractions.
//
// (1) DzSource: the BN-backward gradient dz of a POOL-ONLY layer (detector conv2, descriptor conv_mid), formed on the fly.
//
// For such a layer the upstream gradient lives on the pooled tensor: g[row][ch] = (y == max of the group) ? gpool / ties : 0, and
//     dz = s (g - k1 - (z - mu) istd k2),   s = gamma istd, k1 = mean(g), k2 = mean(g zhat)          (train_layers.cu)
// is a function of z[row][ch], seven per-channel coefficients and three per-(group, channel) values.  bn_bwd_apply_kernel writes it
// to HBM for the two contractions that consume it (dW = x^T dz, dx = dz W^T); the fused path hands this description to the
// contractions instead, whose operand converters evaluate dz_value() on the z tile they have just fetched: the (rows, c) gradient is
// then neither written nor read.  One function, explicit roundings: both paths produce the same bits.
#pragma once
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

namespace f3d {

struct DzSource {
    const float *z;       // (rows, c) pre-BN activation
    const float *coef;    // [7][c]: bsc, bsh (y = z bsc + bsh), s, k1, mu, istd, k2
    const float *pooled;  // (rows / gs, c) group maxima of y
    const float *gpool;   // (rows / gs, c) gradient of the pooled tensor
    const float *inv;     // (rows / gs, c) 1 / number of tied maxima
    int gs;               // rows per group
    int relu;
};

constexpr int kDzCoefs = 7;

// the seven coefficient quads of one channel quad, kept in registers by a converter thread that owns that quad for the whole kernel
struct DzCoefRegs {
    float4 bsc, bsh, ss, k1, mu, is, k2;
};

// The INPUT rows of a contraction when they are the activation of the previous layer, y = act(z * scale + shift) (bn_apply): the
// consumers (forward lin_tc, wgrad) fetch the previous layer's pre-BN tensor z and evaluate the activation in their operand converters
// with bn_apply's own rounding, so y is neither written by the producer nor read by the consumers: the tensor never exists in HBM.
struct XSource {
    const float *coef;  // NULL: the rows are read as they are.  Else [2][k]: scale, shift of the producing layer (bn_stats_finalize)
    int relu;
};

// Max-pool statistics taken in the forward contraction's epilogue (pool-only layers): per HALF group of 32 consecutive rows and per
// channel, in "key" space key = z for gamma >= 0, -z otherwise (the BN scale has gamma's sign, so the activation is non-decreasing in the
// key): zext[(half_group * 3 + {0,1,2}) * c + ch] = {largest key m1, number of rows attaining it, largest key below m1 (or -inf)}.
// After the batch statistics are known, pool_from_extremes_kernel turns them into the pooled maximum and the tie count of the ACTIVATION
// -- exactly: ties among activations are the rows at m1 unless the second key rounds to the same activation, which it checks.
struct PoolEpilogue {
    float *zext;         // NULL: off
    const float *gamma;  // (c) BN scale parameter (its sign selects max or min of z)
};

// (packed fp32 pairs, tc_ptx.cuh: the same IEEE operations as the scalar forms, two per issued instruction -- the operand converters of the
// contractions are paced by their instruction count)
__device__ __forceinline__ float4 x_value(const float4 &z, const float4 &sc, const float4 &sh, int relu) {
    float4 y;
    tc::upk2(tc::fma2(tc::pk2(z.x, z.y), tc::pk2(sc.x, sc.y), tc::pk2(sh.x, sh.y)), y.x, y.y);
    tc::upk2(tc::fma2(tc::pk2(z.z, z.w), tc::pk2(sc.z, sc.w), tc::pk2(sh.z, sh.w)), y.z, y.w);
    if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    return y;
}

__device__ __forceinline__ float dz_value(float z, float bsc, float bsh, float s, float k1, float mu, float is, float k2, float pm, float gsc,
                                          int relu) {
    float y = __fmaf_rn(z, bsc, bsh);  // bit-identical to what bn_apply computes
    if (relu) y = fmaxf(y, 0.f);
    float g = y == pm ? gsc : 0.f;
    if (relu) g = y > 0.f ? g : 0.f;
    return __fmul_rn(s, __fsub_rn(__fsub_rn(g, k1), __fmul_rn(__fmul_rn(__fsub_rn(z, mu), is), k2)));
}

// dz_value() of two channels at once: same expression, same roundings
__device__ __forceinline__ void dz_value2(float z0, float z1, float bsc0, float bsc1, float bsh0, float bsh1, float s0, float s1, float k10, float k11,
                                          float mu0, float mu1, float is0, float is1, float k20, float k21, float pm0, float pm1, float gsc0,
                                          float gsc1, int relu, float &d0, float &d1) {
    float y0, y1;
    tc::upk2(tc::fma2(tc::pk2(z0, z1), tc::pk2(bsc0, bsc1), tc::pk2(bsh0, bsh1)), y0, y1);
    if (relu) {
        y0 = fmaxf(y0, 0.f);
        y1 = fmaxf(y1, 0.f);
    }
    float g0 = y0 == pm0 ? gsc0 : 0.f, g1 = y1 == pm1 ? gsc1 : 0.f;
    if (relu) {
        g0 = y0 > 0.f ? g0 : 0.f;
        g1 = y1 > 0.f ? g1 : 0.f;
    }
    // ptxas contracts a packed multiply that feeds a packed add into one FFMA2 even when both carry an explicit .rn (measured:
    // tools/dz_bits.cu; the scalar forms are left alone), which would round differently from bn_bwd_apply_kernel: the one subtraction
    // whose operand is a product stays scalar
    float t0, t1, u0, u1;
    tc::upk2(tc::mul2(tc::mul2(tc::sub2(tc::pk2(z0, z1), tc::pk2(mu0, mu1)), tc::pk2(is0, is1)), tc::pk2(k20, k21)), t0, t1);
    tc::upk2(tc::sub2(tc::pk2(g0, g1), tc::pk2(k10, k11)), u0, u1);
    tc::upk2(tc::mul2(tc::pk2(s0, s1), tc::pk2(__fsub_rn(u0, t0), __fsub_rn(u1, t1))), d0, d1);
}
// ... and of a channel quad with its coefficient quads
template <class Coef>
__device__ __forceinline__ float4 dz_value4(const float4 &z, const Coef &K, const float4 &pm, const float4 &gsc, int relu) {
    float4 d;
    dz_value2(z.x, z.y, K.bsc.x, K.bsc.y, K.bsh.x, K.bsh.y, K.ss.x, K.ss.y, K.k1.x, K.k1.y, K.mu.x, K.mu.y, K.is.x, K.is.y, K.k2.x, K.k2.y, pm.x, pm.y,
              gsc.x, gsc.y, relu, d.x, d.y);
    dz_value2(z.z, z.w, K.bsc.z, K.bsc.w, K.bsh.z, K.bsh.w, K.ss.z, K.ss.w, K.k1.z, K.k1.w, K.mu.z, K.mu.w, K.is.z, K.is.w, K.k2.z, K.k2.w, pm.z, pm.w,
              gsc.z, gsc.w, relu, d.z, d.w);
    return d;
}

}  // namespace f3d
