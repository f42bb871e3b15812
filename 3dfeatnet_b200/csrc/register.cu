// register.cu -- descriptor matching and RANSAC rigid registration on the device.
//
// Replaces the MATLAB evaluation step of the reference:
//   scripts/computeAndVisualizeMatches.m:43-44   [~, m] = pdist2(desc2, desc1, 'euclidean', 'smallest', 1)
//   scripts/external/ransacfitRt.m:42-75         ransac() on 3-point samples, then a least-squares refit on the inliers
//   scripts/external/ransac.m:103-214            adaptive number of trials (p = 0.99), ">=" best-score rule, maxTrials 10000
//   scripts/external/estimateRigidTransform.m    closed-form quaternion fit: smallest singular vector of B = sum A_i' A_i
//   scripts/external/quat2rot.m
// MATLAB's generator (randsample on the reset global stream, ransac.m:128-141) is not reproducible here, so the sample
// triples are an INPUT: given the same triples the trial sequence, the stopping trial, the inlier set and Rt follow the
// reference's sequential algorithm exactly -- every hypothesis is scored in parallel (one CTA per trial) and a scan then
// replays the sequential bookkeeping over the scores.  All geometry is fp64 like MATLAB's.
#include "common.cuh"

namespace f3d {

// ---------------------------------------------------------------------------------------------- matching
// One warp per query row; the candidate rows are staged through shared memory in tiles of 64.
constexpr int kMatchWarps = 4;
constexpr int kMatchTile = 64;

__global__ void __launch_bounds__(kMatchWarps * 32)
match_kernel(int n1, int n2, int dim, const float *__restrict__ d1, const float *__restrict__ d2, int *__restrict__ match,
             float *__restrict__ dist2) {
    extern __shared__ float msm[];
    float *tile = msm;                                   // [kMatchTile][dim + 1]
    float *qrow = msm + kMatchTile * (dim + 1);          // [kMatchWarps][dim]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * kMatchWarps + warp;
    if (i < n1)
        for (int k = lane; k < dim; k += 32) qrow[warp * dim + k] = __ldg(d1 + static_cast<size_t>(i) * dim + k);
    float best = 3.0e38f;
    int besti = 0x7fffffff;
    for (int j0 = 0; j0 < n2; j0 += kMatchTile) {
        __syncthreads();
        const int rows = min(kMatchTile, n2 - j0);
        for (int e = threadIdx.x; e < rows * dim; e += blockDim.x) {
            const int r = e / dim, k = e - r * dim;
            tile[r * (dim + 1) + k] = __ldg(d2 + static_cast<size_t>(j0 + r) * dim + k);
        }
        __syncthreads();
        if (i < n1) {
            for (int r = lane; r < rows; r += 32) {
                float acc = 0.0f;
                for (int k = 0; k < dim; ++k) {
                    const float df = qrow[warp * dim + k] - tile[r * (dim + 1) + k];
                    acc = __fmaf_rn(df, df, acc);
                }
                if (acc < best) {  // strict: a lane keeps its lowest index among equal distances (j ascends)
                    best = acc;
                    besti = j0 + r;
                }
            }
        }
    }
    // warp arg-min, lowest index on ties ('smallest' keeps the first)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float ob = __shfl_xor_sync(kFull, best, s);
        const int oi = __shfl_xor_sync(kFull, besti, s);
        if (ob < best || (ob == best && oi < besti)) {
            best = ob;
            besti = oi;
        }
    }
    if (i < n1 && lane == 0) {
        match[i] = besti;
        if (dist2) dist2[i] = best;
    }
}

// ---------------------------------------------------------------------------------------------- rigid fit
// Smallest-eigenvalue eigenvector of a symmetric 4x4 matrix (cyclic Jacobi, fp64).
__device__ void smallest_eigenvector4(double a[4][4], double q[4]) {
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < 4; ++p) {
            diag += fabs(a[p][p]);
            for (int r = p + 1; r < 4; ++r) off += fabs(a[p][r]);
        }
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 3; ++p)
            for (int r = p + 1; r < 4; ++r) {
                const double apr = a[p][r];
                if (fabs(apr) <= 1e-300) continue;
                const double theta = (a[r][r] - a[p][p]) / (2.0 * apr);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {  // A <- A J
                    const double akp = a[k][p], akr = a[k][r];
                    a[k][p] = c * akp - s * akr;
                    a[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 4; ++k) {  // A <- J' A
                    const double apk = a[p][k], ark = a[r][k];
                    a[p][k] = c * apk - s * ark;
                    a[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = v[k][p], vkr = v[k][r];
                    v[k][p] = c * vkp - s * vkr;
                    v[k][r] = s * vkp + c * vkr;
                }
            }
    }
    int m = 0;
    for (int k = 1; k < 4; ++k)
        if (a[k][k] < a[m][m]) m = k;
    for (int k = 0; k < 4; ++k) q[k] = v[k][m];
}

// B += A' A for one correspondence (estimateRigidTransform.m:58-75): x, y already centred
__device__ __forceinline__ void accumulate_B(double B[4][4], const double x[3], const double y[3]) {
    const double a[3] = {y[0] - x[0], y[1] - x[1], y[2] - x[2]};  // R12 row; R21 = -a
    const double s[3] = {y[0] + x[0], y[1] + x[1], y[2] + x[2]};  // R22_1
    const double A[4][4] = {{0.0, a[0], a[1], a[2]},
                            {-a[0], 0.0, -s[2], s[1]},   // crossTimesMatrix.m
                            {-a[1], s[2], 0.0, -s[0]},
                            {-a[2], -s[1], s[0], 0.0}};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double acc = 0.0;
            for (int k = 0; k < 4; ++k) acc += A[k][r] * A[k][c];
            B[r][c] += acc;
        }
}

// rot = quat2rot(q) (quat2rot.m), t = xc - rot yc  =>  x ~ rot y + t;  Rt row-major 3x4
__device__ void model_from_B(double B[4][4], const double xc[3], const double yc[3], double Rt[12]) {
    double q[4];
    smallest_eigenvector4(B, q);
    const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    const double R[3][3] = {{q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)},
                            {2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1)},
                            {2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3}};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Rt[r * 4 + c] = R[r][c];
        Rt[r * 4 + 3] = xc[r] - (R[r][0] * yc[0] + R[r][1] * yc[1] + R[r][2] * yc[2]);
    }
}

__device__ __forceinline__ bool is_inlier(const double *Rt, const float *p1, const float *p2, int k, double thr) {
    const double y0 = p2[3 * k], y1 = p2[3 * k + 1], y2 = p2[3 * k + 2];
    double d2 = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double e = static_cast<double>(p1[3 * k + r]) - (Rt[r * 4] * y0 + Rt[r * 4 + 1] * y1 + Rt[r * 4 + 2] * y2 + Rt[r * 4 + 3]);
        d2 += e * e;
    }
    return sqrt(d2) < thr;  // ransacfitRt.m:91-93: abs(d) < t
}

// One CTA per trial: fit the 3-point model (thread 0), count the inliers of all correspondences (ransac.m:147-167).
__global__ void __launch_bounds__(128)
ransac_score_kernel(int npts, const float *__restrict__ p1, const float *__restrict__ p2, int ntrials, const int *__restrict__ triples,
                    double thr, int *__restrict__ ninl, double *__restrict__ models) {
    __shared__ double Rt[12];
    __shared__ int cnt[4];
    const int trial = blockIdx.x;
    if (threadIdx.x == 0) {
        double x[3][3], y[3][3], xc[3] = {0, 0, 0}, yc[3] = {0, 0, 0};
        for (int j = 0; j < 3; ++j) {
            const int k = min(max(triples[trial * 3 + j], 0), npts - 1);
            for (int c = 0; c < 3; ++c) {
                x[j][c] = p1[3 * k + c];
                y[j][c] = p2[3 * k + c];
                xc[c] += x[j][c];
                yc[c] += y[j][c];
            }
        }
        for (int c = 0; c < 3; ++c) {
            xc[c] /= 3.0;
            yc[c] /= 3.0;
        }
        double B[4][4] = {};
        for (int j = 0; j < 3; ++j) {
            const double xx[3] = {x[j][0] - xc[0], x[j][1] - xc[1], x[j][2] - xc[2]};
            const double yy[3] = {y[j][0] - yc[0], y[j][1] - yc[1], y[j][2] - yc[2]};
            accumulate_B(B, xx, yy);
        }
        model_from_B(B, xc, yc, Rt);
        for (int e = 0; e < 12; ++e) models[static_cast<size_t>(trial) * 12 + e] = Rt[e];
    }
    __syncthreads();
    int c = 0;
    for (int k = threadIdx.x; k < npts; k += blockDim.x) c += is_inlier(Rt, p1, p2, k, thr) ? 1 : 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(kFull, c, s);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) ninl[trial] = cnt[0] + cnt[1] + cnt[2] + cnt[3];
}

// The sequential bookkeeping of ransac.m:110-205 replayed over the per-trial scores.  info = {inliers of the chosen model,
// trialcount, chosen trial, status (0 ok, 1 = the supplied triples ran out before N <= trialcount)}.
__global__ void ransac_select_kernel(int npts, int ntrials, int max_trials, const int *__restrict__ ninl, int *__restrict__ info) {
    if (threadIdx.x || blockIdx.x) return;
    const double p = 0.99, eps = 2.220446049250313e-16;
    double N = 1.0;
    int trialcount = 0, bestscore = 0, best = -1, status = 0;
    while (N > trialcount) {
        if (trialcount >= ntrials) {
            status = 1;
            break;
        }
        const int n = ninl[trialcount];
        if (n >= bestscore) {  // ransac.m:170 ("I change it from > to >=")
            bestscore = n;
            best = trialcount;
            const double frac = static_cast<double>(n) / npts;
            double pno = 1.0 - frac * frac * frac;
            pno = fmax(eps, pno);
            pno = fmin(1.0 - eps, pno);
            N = log(1.0 - p) / log(pno);
            N = fmax(N, 10.0);
        }
        ++trialcount;
        if (trialcount > max_trials) break;  // ransac.m:190-196
    }
    info[0] = bestscore;
    info[1] = trialcount;
    info[2] = best;
    info[3] = status;
}

__global__ void ransac_mask_kernel(int npts, const float *__restrict__ p1, const float *__restrict__ p2, const double *__restrict__ models,
                                   const int *__restrict__ info, double thr, unsigned char *__restrict__ mask) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npts) return;
    const int best = info[2];
    mask[k] = best >= 0 && is_inlier(models + static_cast<size_t>(best) * 12, p1, p2, k, thr) ? 1 : 0;
}

// Least-squares fit over the masked correspondences (ransacfitRt.m:66-73 / estimateRt.m): one CTA, fixed-order fp64
// reductions (deterministic).  mask == NULL: all points.  Fewer than 3 points: Rt = NaN (the reference returns []).
__global__ void __launch_bounds__(256)
rigid_fit_kernel(int npts, const float *__restrict__ p1, const float *__restrict__ p2, const unsigned char *__restrict__ mask,
                 double *__restrict__ Rt_out, double *__restrict__ eps_out) {
    __shared__ double red[256][16];  // both reductions
    __shared__ double cen[7];
    const int tid = threadIdx.x;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int k = tid; k < npts; k += 256)
        if (!mask || mask[k]) {
            for (int c = 0; c < 3; ++c) {
                acc[c] += p1[3 * k + c];
                acc[3 + c] += p2[3 * k + c];
            }
            acc[6] += 1.0;
        }
    for (int e = 0; e < 7; ++e) red[tid][e] = acc[e];
    __syncthreads();
    if (tid < 7) {
        double s = 0.0;
        for (int t = 0; t < 256; ++t) s += red[t][tid];
        cen[tid] = s;
    }
    __syncthreads();
    const double cnt = cen[6];
    if (cnt < 3.0) {
        if (tid < 12) Rt_out[tid] = nan("");
        if (tid == 0 && eps_out) *eps_out = nan("");
        return;
    }
    const double xc[3] = {cen[0] / cnt, cen[1] / cnt, cen[2] / cnt}, yc[3] = {cen[3] / cnt, cen[4] / cnt, cen[5] / cnt};
    double B[4][4] = {};
    for (int k = tid; k < npts; k += 256)
        if (!mask || mask[k]) {
            const double xx[3] = {p1[3 * k] - xc[0], p1[3 * k + 1] - xc[1], p1[3 * k + 2] - xc[2]};
            const double yy[3] = {p2[3 * k] - yc[0], p2[3 * k + 1] - yc[1], p2[3 * k + 2] - yc[2]};
            accumulate_B(B, xx, yy);
        }
    for (int e = 0; e < 16; ++e) red[tid][e] = B[e >> 2][e & 3];
    __syncthreads();
    __shared__ double Bs[16];
    if (tid < 16) {
        double s = 0.0;
        for (int t = 0; t < 256; ++t) s += red[t][tid];
        Bs[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double Bm[4][4], Rt[12];
        for (int e = 0; e < 16; ++e) Bm[e >> 2][e & 3] = 0.5 * (Bs[e] + Bs[(e & 3) * 4 + (e >> 2)]);
        model_from_B(Bm, xc, yc, Rt);
        for (int e = 0; e < 12; ++e) Rt_out[e] = Rt[e];
        if (eps_out) {  // Eps = S(4,4): the smallest eigenvalue left on the diagonal by the Jacobi sweeps
            double m = Bm[0][0];
            for (int k = 1; k < 4; ++k) m = fmin(m, Bm[k][k]);
            *eps_out = m;
        }
    }
}

// npts == 3 (ransacfitRt.m:49-54): the three correspondences are the inliers
__global__ void ransac_all_inliers_kernel(int npts, unsigned char *mask, int *info) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < npts) mask[k] = 1;
    if (k == 0) {
        info[0] = npts;
        info[1] = 0;
        info[2] = -1;
        info[3] = 0;
    }
}

}  // namespace f3d

using namespace f3d;

F3D_API int f3d_match_descriptors(int n1, int n2, int dim, const float *desc1, const float *desc2, int *match, float *dist2,
                                  void *stream) {
    if (n1 < 0 || n2 <= 0 || dim <= 0 || dim > 1024 || !desc1 || !desc2 || !match)
        return fail(F3D_ERR_INVALID_ARGUMENT, "match_descriptors: bad arguments");
    if (n1 == 0) return 0;
    const size_t smem = (static_cast<size_t>(kMatchTile) * (dim + 1) + static_cast<size_t>(kMatchWarps) * dim) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "match_descriptors: cudaFuncSetAttribute");
    match_kernel<<<(n1 + kMatchWarps - 1) / kMatchWarps, kMatchWarps * 32, smem, as_stream(stream)>>>(n1, n2, dim, desc1, desc2, match, dist2);
    return check_launch("match_kernel");
}

F3D_API size_t f3d_ransac_workspace_bytes(int npts, int ntrials) {
    (void)npts;
    const size_t t = ntrials > 0 ? ntrials : 0;
    return t * 12 * sizeof(double) + t * sizeof(int) + 256;
}

F3D_API int f3d_rigid_fit(int npts, const float *pts1, const float *pts2, const unsigned char *mask, double *Rt, double *eps, void *stream) {
    if (npts < 0 || !pts1 || !pts2 || !Rt) return fail(F3D_ERR_INVALID_ARGUMENT, "rigid_fit: bad arguments");
    rigid_fit_kernel<<<1, 256, 0, as_stream(stream)>>>(npts, pts1, pts2, mask, Rt, eps);
    return check_launch("rigid_fit_kernel");
}

F3D_API int f3d_ransac_fit_rt(int npts, const float *pts1, const float *pts2, int ntrials, const int *triples, float threshold,
                              int max_trials, double *Rt, unsigned char *inlier_mask, int *info, void *workspace,
                              size_t workspace_bytes, void *stream) {
    if (npts < 0 || !pts1 || !pts2 || ntrials < 0 || (ntrials > 0 && !triples) || !(threshold > 0.0f) || !Rt || !inlier_mask || !info)
        return fail(F3D_ERR_INVALID_ARGUMENT, "ransac_fit_rt: bad arguments");
    if (npts < 3) return fail(F3D_ERR_INVALID_ARGUMENT, "ransac_fit_rt: fewer than 3 correspondences (the reference returns an empty model)");
    cudaStream_t st = as_stream(stream);
    if (npts == 3) {
        ransac_all_inliers_kernel<<<1, 32, 0, st>>>(npts, inlier_mask, info);
        int rc = check_launch("ransac_all_inliers_kernel");
        if (rc) return rc;
        return f3d_rigid_fit(npts, pts1, pts2, nullptr, Rt, nullptr, stream);
    }
    if (ntrials == 0) return fail(F3D_ERR_INVALID_ARGUMENT, "ransac_fit_rt: no sample triples");
    if (!workspace || workspace_bytes < f3d_ransac_workspace_bytes(npts, ntrials) || reinterpret_cast<uintptr_t>(workspace) % 8)
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "ransac_fit_rt: workspace missing, too small or misaligned");
    double *models = static_cast<double *>(workspace);
    int *ninl = reinterpret_cast<int *>(models + static_cast<size_t>(ntrials) * 12);
    ransac_score_kernel<<<ntrials, 128, 0, st>>>(npts, pts1, pts2, ntrials, triples, static_cast<double>(threshold), ninl, models);
    int rc = check_launch("ransac_score_kernel");
    if (rc) return rc;
    ransac_select_kernel<<<1, 32, 0, st>>>(npts, ntrials, max_trials > 0 ? max_trials : 10000, ninl, info);
    rc = check_launch("ransac_select_kernel");
    if (rc) return rc;
    ransac_mask_kernel<<<(npts + 255) / 256, 256, 0, st>>>(npts, pts1, pts2, models, info, static_cast<double>(threshold), inlier_mask);
    rc = check_launch("ransac_mask_kernel");
    if (rc) return rc;
    return f3d_rigid_fit(npts, pts1, pts2, inlier_mask, Rt, nullptr, stream);
}
