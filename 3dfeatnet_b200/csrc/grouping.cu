// grouping.cu -- ball query, group_point, selection-sort top-k and knn_point for sm_100a.
//
// Replaces tf_ops/grouping/tf_grouping_g.cu:3-111,137-177.
//
// The reference launches ONE block of 256 threads per batch element and lets each thread scan all n
// points serially for its centres (1 SM busy at b = 1).  Here one WARP owns kCentresPerWarp centres and its 32
// lanes own 32 consecutive points per step: hits are compacted in index order with ballot + popc, so the
// "first nsample indices in ascending order" contract holds by construction, every centre stops on its own
// 32-point boundary, and b*m/4 warps fill the 148 SMs at any batch size.
// The squared distance uses the reference's FMA association and is compared against the exact threshold
// T = min{s : sqrt_rn(s) >= radius} (common.cuh: ball_threshold), which is the same predicate as
// fmaxf(sqrtf(s),1e-20f) < radius without the square root.
// Empty balls (tf_grouping_g.cu:43-47) are resolved by a second kernel that replays the reference's carried
// "nearest" state for exactly those centres (see bq_fallback_kernel).
#include "common.cuh"

namespace f3d {

constexpr int kCentresPerWarp = 4;
constexpr int kBqWarps = 8;  // warps per CTA
constexpr int kRefStride = 256;  // blockDim.x of the reference launch (tf_grouping_g.cu:180): centre j belongs to thread j % 256

// PER_CENTRE_RADIUS = false: query_ball_point  (radius uniform)    tf_grouping_g.cu:3-52
// PER_CENTRE_RADIUS = true : query_ball_point2 (radii[b,m])        tf_grouping_g.cu:56-90 (empty rows untouched)
template <bool PER_CENTRE_RADIUS>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(int b, int n, int m, float radius, const float *__restrict__ radii, int nsample,
                  const float *__restrict__ xyz1, const float *__restrict__ xyz2, int *__restrict__ idx,
                  int *__restrict__ pts_cnt) {
    const int lane = threadIdx.x & 31;
    const int groups_per_batch = (m + kCentresPerWarp - 1) / kCentresPerWarp;
    const long long wg = static_cast<long long>(blockIdx.x) * kBqWarps + (threadIdx.x >> 5);
    if (wg >= static_cast<long long>(b) * groups_per_batch) return;
    const int batch = static_cast<int>(wg / groups_per_batch);
    const int j0 = static_cast<int>(wg - static_cast<long long>(batch) * groups_per_batch) * kCentresPerWarp;

    const float *p1 = xyz1 + static_cast<size_t>(batch) * n * 3;
    const float *p2 = xyz2 + static_cast<size_t>(batch) * m * 3;
    int *row0 = idx + (static_cast<size_t>(batch) * m + j0) * nsample;
    const unsigned lt = lanemask_lt();

    float cx[kCentresPerWarp], cy[kCentresPerWarp], cz[kCentresPerWarp], T[kCentresPerWarp];
    int cnt[kCentresPerWarp], first[kCentresPerWarp];
    unsigned active = 0;
#pragma unroll
    for (int c = 0; c < kCentresPerWarp; ++c) {
        const int j = min(j0 + c, m - 1);
        cx[c] = __ldg(p2 + 3 * j);
        cy[c] = __ldg(p2 + 3 * j + 1);
        cz[c] = __ldg(p2 + 3 * j + 2);
        const float r = PER_CENTRE_RADIUS ? __ldg(radii + static_cast<size_t>(batch) * m + j) : radius;
        T[c] = ball_threshold(r);
        cnt[c] = 0;
        first[c] = -1;
        if (j0 + c < m && T[c] > 0.0f) active |= 1u << c;
    }

    // software-pipelined scan: the next 32 points are in flight while the current 32 are tested
    float nx = 0.f, ny = 0.f, nz = 0.f;
    if (lane < n) {
        nx = __ldg(p1 + 3 * lane);
        ny = __ldg(p1 + 3 * lane + 1);
        nz = __ldg(p1 + 3 * lane + 2);
    }
    for (int base = 0; base < n && active; base += 32) {
        const float px = nx, py = ny, pz = nz;
        const int k = base + lane;
        const bool valid = k < n;
        const int kn = k + 32;
        if (kn < n) {
            nx = __ldg(p1 + 3 * kn);
            ny = __ldg(p1 + 3 * kn + 1);
            nz = __ldg(p1 + 3 * kn + 2);
        }
#pragma unroll
        for (int c = 0; c < kCentresPerWarp; ++c) {
            if (active & (1u << c)) {  // warp-uniform
                const float s = sqdist_ref(cx[c] - px, cy[c] - py, cz[c] - pz);
                const bool hit = valid && !(s >= T[c]);
                const unsigned ball = __ballot_sync(kFull, hit);
                if (ball) {
                    const int pos = cnt[c] + __popc(ball & lt);
                    if (hit && pos < nsample) row0[c * nsample + pos] = k;
                    if (cnt[c] == 0) first[c] = base + __ffs(ball) - 1;
                    cnt[c] += __popc(ball);
                    if (cnt[c] >= nsample) active &= ~(1u << c);
                }
            }
        }
    }
    // pad with the first hit (tf_grouping_g.cu:29-32) and publish the count (:49)
#pragma unroll
    for (int c = 0; c < kCentresPerWarp; ++c) {
        if (j0 + c < m) {
            const int cc = min(cnt[c], nsample);
            if (cc > 0)
                for (int s = cc + lane; s < nsample; s += 32) row0[c * nsample + s] = first[c];
            if (lane == 0) pts_cnt[static_cast<size_t>(batch) * m + j0 + c] = cc;
        }
    }
}

// Empty balls.  The reference declares nearest_d / nearest_k OUTSIDE its centre loop (tf_grouping_g.cu:13-14), so
// thread t = j % 256 carries them through its centres t, t+256, ...; they are updated with a strict '<' for every
// EXAMINED point (:36-39), i.e. for k in [0, k_exit(j')] where k_exit is the index of the nsample-th hit (the scan
// breaks at the top of the next iteration, :18-19) or n-1.  An empty centre j is filled with the carried value
// after its own full scan (:43-47) = first strict minimum of d over the concatenation of the examined prefixes of
// j' = t, t+256, ..., j.  One warp replays that sequence for one empty centre; "first strict minimum" in scan
// order is the lexicographic minimum of (d, ordinal of j', k).  Runs after ball_query_kernel on the same stream.
__global__ void __launch_bounds__(kBqWarps * 32)
bq_fallback_kernel(int b, int n, int m, int nsample, const float *__restrict__ xyz1,
                   const float *__restrict__ xyz2, int *__restrict__ idx, const int *__restrict__ pts_cnt) {
    const int lane = threadIdx.x & 31;
    const long long w = static_cast<long long>(blockIdx.x) * kBqWarps + (threadIdx.x >> 5);
    if (w >= static_cast<long long>(b) * m) return;
    const int batch = static_cast<int>(w / m);
    const int j = static_cast<int>(w - static_cast<long long>(batch) * m);
    const int *cntb = pts_cnt + static_cast<size_t>(batch) * m;
    if (cntb[j] != 0) return;

    const float *p1 = xyz1 + static_cast<size_t>(batch) * n * 3;
    const float *p2 = xyz2 + static_cast<size_t>(batch) * m * 3;
    int *idxb = idx + static_cast<size_t>(batch) * m * nsample;

    float ld = __int_as_float(0x7f800000);  // (float)1.0e99 == +inf
    int lk = -1;
    unsigned lord = 0xffffffffu;
    unsigned ord = 0;
    for (int jj = j % kRefStride; jj <= j; jj += kRefStride, ++ord) {
        const int cj = cntb[jj];
        const int kexit = (cj >= nsample) ? idxb[static_cast<size_t>(jj) * nsample + nsample - 1] : n - 1;
        const float x2 = __ldg(p2 + 3 * jj), y2 = __ldg(p2 + 3 * jj + 1), z2 = __ldg(p2 + 3 * jj + 2);
        for (int k = lane; k <= kexit; k += 32) {
            const float x1 = __ldg(p1 + 3 * k), y1 = __ldg(p1 + 3 * k + 1), z1 = __ldg(p1 + 3 * k + 2);
            const float d = fmaxf(__fsqrt_rn(sqdist_ref(x2 - x1, y2 - y1, z2 - z1)), 1e-20f);
            if (d < ld) {
                ld = d;
                lk = k;
                lord = ord;
            }
        }
    }
    // lexicographic min over lanes of (d, ord, k); d >= 1e-20 > 0 so its bits order like unsigned ints
    const unsigned db = __float_as_uint(ld);
    const unsigned dmin = __reduce_min_sync(kFull, db);
    const unsigned omin = __reduce_min_sync(kFull, db == dmin ? lord : 0xffffffffu);
    const unsigned kmin = __reduce_min_sync(kFull, (db == dmin && lord == omin) ? static_cast<unsigned>(lk) : 0xffffffffu);
    const int fill = static_cast<int>(kmin);  // -1 (0xffffffff) when nothing was ever nearer than +inf
    for (int s = lane; s < nsample; s += 32) idxb[static_cast<size_t>(j) * nsample + s] = fill;
}

// group_point: out[b,j,k,:] = points[b,idx[b,j,k],:]  (tf_grouping_g.cu:94-111).
// One thread per output element of VEC floats: writes are fully coalesced, each gathered row is read as
// contiguous VEC-wide pieces.  VEC = 4 when c % 4 == 0 and both pointers are 16-byte aligned.
template <int VEC>
__global__ void group_point_kernel(int n, int c, long long slots_per_batch, long long total,
                                   const float *__restrict__ points, const int *__restrict__ idx,
                                   float *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int cv = c / VEC;
    const long long r = i / cv;
    const int l = static_cast<int>(i - r * cv);
    const long long bb = r / slots_per_batch;
    const int ii = __ldg(idx + r);
    const float *src = points + (bb * n + ii) * c + l * VEC;
    if (VEC == 4) {
        reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(src));
    } else {
        out[i] = __ldg(src);
    }
}

// selection sort, one WARP per (b,j) row (the reference: one thread per row, tf_grouping_g.cu:137-177).
// Round s: arg-min over positions [s,n) by (value, position) -- identical to the reference's
// "min=s; if (p[t]<p[min]) min=t" scan -- then the same swap, so the leftover order (and therefore the tie
// behaviour of later rounds) is the reference's.
__global__ void __launch_bounds__(256)
selection_sort_kernel(long long rows, int n, int k, const float *__restrict__ dist, int *__restrict__ outi,
                      float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *d = dist + row * n;
    float *p = out + row * n;
    int *pi = outi + row * n;
    for (int s = lane; s < n; s += 32) {
        p[s] = d[s];
        pi[s] = s;
    }
    __syncwarp();
    const int rounds = min(k, n);
    for (int s = 0; s < rounds; ++s) {
        float bv = __int_as_float(0x7f800000);
        int bt = 0x7fffffff;
        for (int t = s + lane; t < n; t += 32) {
            const float v = p[t];
            if (v < bv || (v == bv && t < bt)) {  // lanes walk ascending t: keeps the lowest t per value
                bv = v;
                bt = t;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(kFull, bv, o);
            const int ot = __shfl_xor_sync(kFull, bt, o);
            if (ov < bv || (ov == bv && ot < bt)) {
                bv = ov;
                bt = ot;
            }
        }
        if (bt == 0x7fffffff) bt = s;  // row of +inf / NaN: nothing is '<' p[s], the reference keeps min = s
        if (lane == 0 && bt != s) {
            const float tv = p[bt];
            p[bt] = p[s];
            p[s] = tv;
            const int ti = pi[bt];
            pi[bt] = pi[s];
            pi[s] = ti;
        }
        __syncwarp();
    }
}

// Squared-L2 matrix of knn_point (tf_grouping.py:79-81): sequential over c, separate multiply and add
// (the association is stated in oracle/ops_oracle.c: oracle_knn_dist).  dist is (b,m,n).
__global__ void knn_dist_kernel(int n, int m, int c, long long total, const float *__restrict__ xyz1,
                                const float *__restrict__ xyz2, float *__restrict__ dist) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long row = i / n;  // b*m + j
    const int s = static_cast<int>(i - row * n);
    const long long bb = row / m;
    const float *p = xyz1 + (bb * n + s) * c;
    const float *q = xyz2 + row * c;
    float acc = 0.0f;
    for (int l = 0; l < c; ++l) {
        const float df = __fsub_rn(__ldg(p + l), __ldg(q + l));
        acc = __fadd_rn(acc, __fmul_rn(df, df));
    }
    dist[i] = acc;
}

__global__ void knn_slice_kernel(int n, int k, long long total, const float *__restrict__ out,
                                 const int *__restrict__ outi, float *__restrict__ val, int *__restrict__ idx) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long row = i / k;
    const int s = static_cast<int>(i - row * k);
    val[i] = s < n ? out[row * n + s] : 0.0f;
    idx[i] = s < n ? outi[row * n + s] : 0;
}

static inline unsigned blocks_for(long long total, int per_block) {
    return static_cast<unsigned>((total + per_block - 1) / per_block);
}

}  // namespace f3d

using namespace f3d;

static int ball_query_common(bool per_centre, int b, int n, int m, float radius, const float *radii, int nsample,
                             const float *xyz1, const float *xyz2, int *idx, int *pts_cnt, cudaStream_t st) {
    const long long warps = static_cast<long long>(b) * ((m + kCentresPerWarp - 1) / kCentresPerWarp);
    if (warps == 0) return 0;
    if (per_centre)
        ball_query_kernel<true><<<blocks_for(warps, kBqWarps), kBqWarps * 32, 0, st>>>(b, n, m, 0.f, radii, nsample, xyz1, xyz2, idx, pts_cnt);
    else
        ball_query_kernel<false><<<blocks_for(warps, kBqWarps), kBqWarps * 32, 0, st>>>(b, n, m, radius, nullptr, nsample, xyz1, xyz2, idx, pts_cnt);
    return check_launch("ball_query_kernel");
}

F3D_API int f3d_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                                 int *idx, int *pts_cnt, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f) || !xyz1 || !xyz2 || !idx || !pts_cnt)
        return fail(F3D_ERR_INVALID_ARGUMENT, "query_ball_point: bad arguments");
    cudaStream_t st = as_stream(stream);
    int rc = ball_query_common(false, b, n, m, radius, nullptr, nsample, xyz1, xyz2, idx, pts_cnt, st);
    if (rc) return rc;
    const long long w = static_cast<long long>(b) * m;
    if (w == 0) return 0;
    bq_fallback_kernel<<<blocks_for(w, kBqWarps), kBqWarps * 32, 0, st>>>(b, n, m, nsample, xyz1, xyz2, idx, pts_cnt);
    return check_launch("bq_fallback_kernel");
}

F3D_API int f3d_query_ball_point2(int b, int n, int m, int nsample, const float *xyz1, const float *xyz2,
                                  const float *radii, int *idx, int *pts_cnt, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !xyz1 || !xyz2 || !radii || !idx || !pts_cnt)
        return fail(F3D_ERR_INVALID_ARGUMENT, "query_ball_point2: bad arguments");
    return ball_query_common(true, b, n, m, 0.f, radii, nsample, xyz1, xyz2, idx, pts_cnt, as_stream(stream));
}

F3D_API int f3d_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                            void *stream) {
    if (b < 0 || n <= 0 || c <= 0 || m < 0 || nsample <= 0 || !points || !idx || !out)
        return fail(F3D_ERR_INVALID_ARGUMENT, "group_point: bad arguments");
    const long long slots = static_cast<long long>(m) * nsample;
    const bool vec4 = (c % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) | reinterpret_cast<uintptr_t>(out)) % 16 == 0);
    const long long total = static_cast<long long>(b) * slots * (vec4 ? c / 4 : c);
    if (total == 0) return 0;
    if (vec4)
        group_point_kernel<4><<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(n, c, slots, total, points, idx, out);
    else
        group_point_kernel<1><<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(n, c, slots, total, points, idx, out);
    return check_launch("group_point_kernel");
}

F3D_API int f3d_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || k <= 0 || !dist || !outi || !out)
        return fail(F3D_ERR_INVALID_ARGUMENT, "selection_sort: bad arguments");
    const long long rows = static_cast<long long>(b) * m;
    if (rows == 0) return 0;
    selection_sort_kernel<<<blocks_for(rows, 8), 256, 0, as_stream(stream)>>>(rows, n, k, dist, outi, out);
    return check_launch("selection_sort_kernel");
}

F3D_API size_t f3d_knn_workspace_bytes(int b, int n, int m, int c, int k) {
    (void)c;
    (void)k;
    return static_cast<size_t>(b) * m * n * 12 + 256;
}

F3D_API int f3d_knn_point(int b, int n, int m, int c, int k, const float *xyz1, const float *xyz2, float *val, int *idx,
                          void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || c <= 0 || k <= 0 || !xyz1 || !xyz2 || !val || !idx)
        return fail(F3D_ERR_INVALID_ARGUMENT, "knn_point: bad arguments");
    if (!workspace || workspace_bytes < f3d_knn_workspace_bytes(b, n, m, c, k))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "knn_point: workspace too small");
    const long long rows = static_cast<long long>(b) * m;
    if (rows == 0) return 0;
    cudaStream_t st = as_stream(stream);
    float *dist = static_cast<float *>(workspace);
    float *out = dist + rows * n;
    int *outi = reinterpret_cast<int *>(out + rows * n);
    knn_dist_kernel<<<blocks_for(rows * n, 256), 256, 0, st>>>(n, m, c, rows * n, xyz1, xyz2, dist);
    int rc = check_launch("knn_dist_kernel");
    if (rc) return rc;
    selection_sort_kernel<<<blocks_for(rows, 8), 256, 0, st>>>(rows, n, k, dist, outi, out);
    rc = check_launch("selection_sort_kernel");
    if (rc) return rc;
    knn_slice_kernel<<<blocks_for(rows * k, 256), 256, 0, st>>>(n, k, rows * k, out, outi, val, idx);
    return check_launch("knn_slice_kernel");
}
