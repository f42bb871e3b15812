// frames.cu -- the neighbourhoods of the cluster centres in their local frames, one kernel each way.
//
// Replaces the op chain of models/pointnet_common.py:42-54 (query_and_group_points) and :104-119 (sample_and_group):
//     grouped = group_point(xyz, idx) - centre            (tf_grouping_g.cu:93-111 + tile + sub)
//     grouped /= radius                                   (normalize_radius)
//     grouped = rotate about z by the cluster's angle     (cos / sin / stack, or matmul with R)
// which is a gather and a dozen element-wise passes over (b, m, nsample, 3) tensors in the reference and in the op-by-op
// statement of this repository.  Same arithmetic, same roundings (a true division by the radius; x c - y s as two rounded
// products and a rounded difference), so the rows carry the bits of the op-by-op statement.
// The gradient that the training step needs is the one with respect to the ANGLES (the descriptor's clusters are rotated by the
// detector's orientation, models/feat3dnet.py:301-305): d angle = sum over the samples of (gy x' - gx y') (counter-clockwise;
// negated for the clockwise convention), one warp per cluster, fixed-order reduction.
#include "common.cuh"

namespace f3d {

// one thread per sample; clockwise != 0: x' = x c + y s, y' = -x s + y c (query_and_group_points), else x' = x c - y s, y' = x s + y c
__global__ void __launch_bounds__(256)
local_frames_kernel(long long total, int n, int m, int s, const float *__restrict__ xyz, const float *__restrict__ centres,
                    const int *__restrict__ idx, const float *__restrict__ angles, int clockwise, float radius, int normalize,
                    float *__restrict__ before, float *__restrict__ rotation, float *__restrict__ out) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long long cl = e / s;  // cluster index b * m + j
    const long long b = cl / m;
    const int i = __ldg(idx + e);
    const float *p = xyz + (b * n + i) * 3;
    const float *c = centres + cl * 3;
    float x = __fsub_rn(__ldg(p), __ldg(c)), y = __fsub_rn(__ldg(p + 1), __ldg(c + 1)), z = __fsub_rn(__ldg(p + 2), __ldg(c + 2));
    if (normalize) {
        x = __fdiv_rn(x, radius);
        y = __fdiv_rn(y, radius);
        z = __fdiv_rn(z, radius);
    }
    if (before) {
        before[e * 3] = x;
        before[e * 3 + 1] = y;
        before[e * 3 + 2] = z;
    }
    if (angles) {
        const float a = __ldg(angles + cl);
        const float cs = cosf(a);
        float sn = sinf(a);
        if (clockwise) sn = -sn;
        if (rotation && e == cl * s) {  // the cluster's matrix R (sample_and_group's end_points['rotation']): rows (c, s, 0), (-s, c, 0), (0, 0, 1)
            float *R = rotation + cl * 9;
            R[0] = cs; R[1] = sn; R[2] = 0.f;
            R[3] = -sn; R[4] = cs; R[5] = 0.f;
            R[6] = 0.f; R[7] = 0.f; R[8] = 1.f;
        }
        const float xr = __fsub_rn(__fmul_rn(x, cs), __fmul_rn(y, sn));
        const float yr = __fadd_rn(__fmul_rn(x, sn), __fmul_rn(y, cs));
        x = xr;
        y = yr;
    }
    out[e * 3] = x;
    out[e * 3 + 1] = y;
    out[e * 3 + 2] = z;
}

// one warp per cluster: d angle = sum_s (gy x' - gx y') with (x', y') the rotated coordinates the forward produced
__global__ void __launch_bounds__(256)
local_frames_angle_grad_kernel(long long clusters, int s, const float *__restrict__ out, const float *__restrict__ gout, int clockwise,
                               float *__restrict__ dangle) {
    const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= clusters) return;
    const float *o = out + w * s * 3, *g = gout + w * s * 3;
    float acc = 0.f;
    for (int k = lane; k < s; k += 32) {
        const float xr = __ldg(o + 3 * k), yr = __ldg(o + 3 * k + 1);
        const float gx = __ldg(g + 3 * k), gy = __ldg(g + 3 * k + 1);
        acc += __fsub_rn(__fmul_rn(gy, xr), __fmul_rn(gx, yr));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(kFull, acc, d);
    if (lane == 0) dangle[w] = clockwise ? -acc : acc;
}

}  // namespace f3d

using namespace f3d;

// xyz (b,n,3), centres (b,m,3), idx (b,m,s) -> out (b,m,s,3) = rotate_z((xyz[idx] - centre) [/ radius], angle[b,m]).
// angles NULL: no rotation.  before (b,m,s,3) or NULL: the rows before the rotation (sample_and_group's 'grouped_xyz_before');
// rotation (b,m,3,3) or NULL: the matrices R the rows are multiplied with (end_points['rotation']).
F3D_API int f3d_group_local_frames(int b, int n, int m, int s, const float *xyz, const float *centres, const int *idx, const float *angles,
                                   int clockwise, float radius, int normalize_radius, float *before, float *rotation, float *out, void *stream) {
    if (b <= 0 || n <= 0 || m <= 0 || s <= 0 || !xyz || !centres || !idx || !out)
        return fail(F3D_ERR_INVALID_ARGUMENT, "group_local_frames: bad arguments");
    if (normalize_radius && !(radius > 0.f)) return fail(F3D_ERR_INVALID_ARGUMENT, "group_local_frames: radius must be positive");
    const long long total = static_cast<long long>(b) * m * s;
    local_frames_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(total, n, m, s, xyz, centres, idx, angles, clockwise,
                                                                                                  radius, normalize_radius, before, rotation, out);
    return check_launch("local_frames_kernel");
}

// gradient of the above with respect to the angles: out = the forward's result, gout = dL/dout -> dangle (b,m)
F3D_API int f3d_group_local_frames_angle_grad(int b, int m, int s, const float *out, const float *gout, int clockwise, float *dangle, void *stream) {
    if (b <= 0 || m <= 0 || s <= 0 || !out || !gout || !dangle) return fail(F3D_ERR_INVALID_ARGUMENT, "group_local_frames_angle_grad: bad arguments");
    const long long clusters = static_cast<long long>(b) * m;
    local_frames_angle_grad_kernel<<<static_cast<unsigned>((clusters * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(clusters, s, out, gout, clockwise,
                                                                                                                     dangle);
    return check_launch("local_frames_angle_grad_kernel");
}
