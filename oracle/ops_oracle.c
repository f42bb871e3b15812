/*
 * ops_oracle.c -- CPU restatement of the 3DFeat-Net sample-and-group operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and there only as the checker (or the reported CPU baseline), never as the
 * thing shipped.  The product path is 3dfeatnet_b200/csrc/ (CUDA, sm_100a).
 *
 * Every function follows the reference kernel it cites *literally* (same thread ownership,
 * same tie rules, same floating-point association).  The association of the squared
 * distance is the one nvcc 12.9 emits for the reference source on sm_100a (checked in the
 * SASS of oracle/_ref/libref_{sampling,grouping}.so: FMUL dy,dy ; FFMA dx,dx ; FFMA dz,dz),
 * i.e.  fmaf(dz,dz, fmaf(dx,dx, dy*dy)).  Build with -ffp-contract=off so that gcc adds no
 * contraction of its own (oracle/Makefile).
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - selection sort   : known-answer vector of tf_ops/grouping/test/selection_sort.cpp:65-93
 *   - ball query/group : the reference's own CPU loops (tf_ops/grouping/test/query_ball_point.cpp:19-84)
 *                        compiled as-is into oracle/_ref/libref_cpu_grouping.so
 *   - everything       : the reference CUDA kernels compiled as-is for sm_100a
 *                        (oracle/_ref/libref_{sampling,grouping}.so), run on the GPU box;
 *                        their outputs on seeded inputs are committed under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* squared distance with the association observed in the reference SASS */
static inline float sqdist_ref(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling.  Reference: tf_ops/sampling/tf_sampling_g.cu:105-170
 * (farthestpointsamplingKernel, launched <<<32,512>>> at :204).
 * One block of 512 threads per cloud; thread t owns k = t, t+512, ...; it keeps its FIRST strict
 * maximum (best=-1 start, `d2>best`, :146); a 9-level tree keeps the LOWER slot on ties
 * (`dists[i1]<dists[i2]`, :158).  idx[0]=0 (:114-116); running distance starts at 1e38 (:118).
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_farthest_point_sample(int b, int n, int m, const float *inp, int *out) {
    if (m <= 0) return; /* :106-107 */
    enum { BS = 512 };
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < b; ++i) {
        const float *p = inp + (size_t)i * n * 3;
        float *td = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float dists[BS];
        int dists_i[BS];
        for (int k = 0; k < n; ++k) td[k] = 1e38f;
        int old = 0;
        out[(size_t)i * m + 0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < BS; ++t) {
                int besti = 0;
                float best = -1.0f;
                for (int k = t; k < n; k += BS) {
                    const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
                    const float d = sqdist_ref(x2 - x1, y2 - y1, z2 - z1);
                    const float d2 = fminf(d, td[k]);
                    if (d2 != td[k]) td[k] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = k;
                    }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int u = 0; (1 << u) < BS; ++u) {
                for (int t = 0; t < (BS >> (u + 1)); ++t) {
                    const int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (dists[i1] < dists[i2]) {
                        dists[i1] = dists[i2];
                        dists_i[i1] = dists_i[i2];
                    }
                }
            }
            old = dists_i[0];
            out[(size_t)i * m + j] = old;
        }
        free(td);
    }
}

/* gather_point.  Reference: tf_sampling_g.cu:172-181 (c fixed at 3). */
ORACLE_API void oracle_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
#pragma omp parallel for
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l) out[((size_t)i * m + j) * 3 + l] = inp[((size_t)i * n + a) * 3 + l];
        }
}

/* gather_point_grad (scatter-add).  Reference: tf_sampling_g.cu:183-192; the caller zero-fills
 * (tf_sampling.cpp:174).  The reference adds with atomics in an unspecified order; here the
 * order is ascending j, which is also the order the deterministic CUDA path uses. */
ORACLE_API void oracle_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
#pragma omp parallel for
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const int a = idx[(size_t)i * m + j];
            for (int l = 0; l < 3; ++l) inp_g[((size_t)i * n + a) * 3 + l] += out_g[((size_t)i * m + j) * 3 + l];
        }
}

/* ------------------------------------------------------------------------------------------
 * query_ball_point.  Reference: tf_ops/grouping/tf_grouping_g.cu:3-52, launched <<<b,256>>> (:180).
 * Thread t owns centres j = t, t+256, ...  nearest_d / nearest_k are declared OUTSIDE the centre
 * loop (:13-14) and therefore carry from one centre of a thread to its next one; an empty ball
 * is filled with the carried nearest_k (:43-47).  This literal restatement keeps that.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1,
                                        const float *xyz2, int *idx, int *pts_cnt) {
    enum { STRIDE = 256 };
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int i = 0; i < b; ++i)
        for (int t = 0; t < STRIDE; ++t) {
            const float *p1 = xyz1 + (size_t)i * n * 3;
            const float *p2 = xyz2 + (size_t)i * m * 3;
            int *pidx = idx + (size_t)i * m * nsample;
            int *pcnt = pts_cnt + (size_t)i * m;
            float nearest_d = (float)1.0e99; /* +inf as float, :13 */
            int nearest_k = -1;
            for (int j = t; j < m; j += STRIDE) {
                int cnt = 0;
                const float x2 = p2[j * 3 + 0], y2 = p2[j * 3 + 1], z2 = p2[j * 3 + 2];
                for (int k = 0; k < n; ++k) {
                    if (cnt == nsample) break;
                    const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
                    const float d = fmaxf(sqrtf(sqdist_ref(x2 - x1, y2 - y1, z2 - z1)), 1e-20f);
                    if (d < radius) {
                        if (cnt == 0)
                            for (int l = 0; l < nsample; ++l) pidx[j * nsample + l] = k;
                        pidx[j * nsample + cnt] = k;
                        cnt += 1;
                    }
                    if (d < nearest_d) {
                        nearest_d = d;
                        nearest_k = k;
                    }
                }
                if (cnt == 0)
                    for (int l = 0; l < nsample; ++l) pidx[j * nsample + l] = nearest_k;
                pcnt[j] = cnt;
            }
        }
}

/* query_ball_point2 (per-centre radius, no fallback).  Reference: tf_grouping_g.cu:56-90.
 * Empty rows are left untouched (the reference leaves them uninitialised). */
ORACLE_API void oracle_query_ball_point2(int b, int n, int m, int nsample, const float *xyz1, const float *xyz2,
                                         const float *radii, int *idx, int *pts_cnt) {
#pragma omp parallel for collapse(2) schedule(dynamic, 16)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *p1 = xyz1 + (size_t)i * n * 3;
            const float *p2 = xyz2 + (size_t)i * m * 3;
            int *pidx = idx + (size_t)i * m * nsample;
            const float r = radii[(size_t)i * m + j];
            int cnt = 0;
            const float x2 = p2[j * 3 + 0], y2 = p2[j * 3 + 1], z2 = p2[j * 3 + 2];
            for (int k = 0; k < n; ++k) {
                if (cnt == nsample) break;
                const float x1 = p1[k * 3 + 0], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
                const float d = fmaxf(sqrtf(sqdist_ref(x2 - x1, y2 - y1, z2 - z1)), 1e-20f);
                if (d < r) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) pidx[j * nsample + l] = k;
                    pidx[j * nsample + cnt] = k;
                    cnt += 1;
                }
            }
            pts_cnt[(size_t)i * m + j] = cnt;
        }
}

/* group_point.  Reference: tf_grouping_g.cu:94-111. */
ORACLE_API void oracle_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                                   float *out) {
#pragma omp parallel for collapse(2)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                const int ii = idx[((size_t)i * m + j) * nsample + k];
                memcpy(out + (((size_t)i * m + j) * nsample + k) * c, points + ((size_t)i * n + ii) * c,
                       sizeof(float) * c);
            }
}

/* group_point_grad.  Reference: tf_grouping_g.cu:115-132 (atomics, caller zero-fills,
 * tf_grouping.cpp:270) and the CPU statement tf_ops/grouping/test/query_ball_point.cpp:68-84,
 * whose (j,k)-ascending accumulation order is the one used here. */
ORACLE_API void oracle_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out,
                                        const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
#pragma omp parallel for
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                const int ii = idx[((size_t)i * m + j) * nsample + k];
                const float *g = grad_out + (((size_t)i * m + j) * nsample + k) * c;
                float *o = grad_points + ((size_t)i * n + ii) * c;
                for (int l = 0; l < c; ++l) o[l] += g[l];
            }
}

/* selection sort top-k.  Reference: tf_grouping_g.cu:137-177 (also test/selection_sort.cpp:20-63).
 * Outputs are full (b,m,n) arrays; only the first k of each row are meaningful. */
ORACLE_API void oracle_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
#pragma omp parallel for collapse(2)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *d = dist + ((size_t)i * m + j) * n;
            float *p = out + ((size_t)i * m + j) * n;
            int *pi = outi + ((size_t)i * m + j) * n;
            for (int s = 0; s < n; ++s) {
                p[s] = d[s];
                pi[s] = s;
            }
            for (int s = 0; s < k && s < n; ++s) {
                int mn = s;
                for (int t = s + 1; t < n; ++t)
                    if (p[t] < p[mn]) mn = t;
                if (mn != s) {
                    float tmp = p[mn];
                    p[mn] = p[s];
                    p[s] = tmp;
                    int ti = pi[mn];
                    pi[mn] = pi[s];
                    pi[s] = ti;
                }
            }
        }
}

/* Squared-L2 distance matrix of knn_point.  Reference: tf_ops/grouping/tf_grouping.py:79-81
 * (tile, subtract, square, reduce_sum over c -- TensorFlow arithmetic, so the association is
 * OURS, stated once here and used identically by the CUDA path: sequential over c, separate
 * multiply and add, no FMA).  dist is (b,m,n). */
ORACLE_API void oracle_knn_dist(int b, int n, int m, int c, const float *xyz1, const float *xyz2, float *dist) {
#pragma omp parallel for collapse(2)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *q = xyz2 + ((size_t)i * m + j) * c;
            for (int s = 0; s < n; ++s) {
                const float *p = xyz1 + ((size_t)i * n + s) * c;
                float acc = 0.0f;
                for (int l = 0; l < c; ++l) {
                    const float df = p[l] - q[l];
                    const float sq = df * df;
                    acc = acc + sq;
                }
                dist[((size_t)i * m + j) * n + s] = acc;
            }
        }
}

/* knn_point = distance matrix + selection sort + slice.  Reference: tf_grouping.py:63-88.
 * val/idx are (b,m,k). */
ORACLE_API void oracle_knn_point(int b, int n, int m, int c, int k, const float *xyz1, const float *xyz2, float *val,
                                 int *idx) {
#pragma omp parallel for collapse(2)
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            float *p = (float *)malloc(sizeof(float) * (size_t)n);
            int *pi = (int *)malloc(sizeof(int) * (size_t)n);
            const float *q = xyz2 + ((size_t)i * m + j) * c;
            for (int s = 0; s < n; ++s) {
                const float *pp = xyz1 + ((size_t)i * n + s) * c;
                float acc = 0.0f;
                for (int l = 0; l < c; ++l) {
                    const float df = pp[l] - q[l];
                    const float sq = df * df;
                    acc = acc + sq;
                }
                p[s] = acc;
                pi[s] = s;
            }
            for (int s = 0; s < k && s < n; ++s) {
                int mn = s;
                for (int t = s + 1; t < n; ++t)
                    if (p[t] < p[mn]) mn = t;
                if (mn != s) {
                    float tmp = p[mn];
                    p[mn] = p[s];
                    p[s] = tmp;
                    int ti = pi[mn];
                    pi[mn] = pi[s];
                    pi[s] = ti;
                }
            }
            for (int s = 0; s < k; ++s) {
                val[((size_t)i * m + j) * k + s] = s < n ? p[s] : 0.0f;
                idx[((size_t)i * m + j) * k + s] = s < n ? pi[s] : 0;
            }
            free(p);
            free(pi);
        }
}

/* ------------------------------------------------------------------------------------------
 * cumsum + prob_sample.  Reference: tf_sampling_g.cu:7-104 (cumsumKernel <<<32,512>>>,
 * binarysearchKernel).  Literal restatement of the blocked scan: chunks of 8192 values, groups of
 * four summed as v1, v1+v2, v3+(v1+v2), (v3+v4)+(v1+v2) (:20-33), group totals scanned with an
 * up-sweep / down-sweep tree over a padded buffer (:46-67), compensated carry between chunks
 * (:81-84).  Levels of the tree touch disjoint slots, so a sequential walk per level is exact.
 * ---------------------------------------------------------------------------------------- */
ORACLE_API void oracle_cumsum(int b, int n, const float *inp, float *out) {
    enum { BlockSize = 2048, PL = 5 };
#pragma omp parallel for
    for (int i = 0; i < b; ++i) {
        float *buffer4 = (float *)malloc(sizeof(float) * BlockSize * 4);
        float *buffer = (float *)malloc(sizeof(float) * (BlockSize + (BlockSize >> PL)));
        float runningsum = 0, runningsum2 = 0;
        for (int j = 0; j < n; j += BlockSize * 4) {
            const int n24_i = (n - j) < BlockSize * 4 ? (n - j) : BlockSize * 4;
            const int n24 = (n24_i + 3) & ~3;
            const int n2 = n24 >> 2;
            for (int k = 0; k < n24_i; k += 4) {
                if (k + 3 < n24_i) {
                    float v1 = inp[(size_t)i * n + j + k];
                    float v2 = inp[(size_t)i * n + j + k + 1];
                    v2 += v1;
                    float v3 = inp[(size_t)i * n + j + k + 2];
                    float v4 = inp[(size_t)i * n + j + k + 3];
                    v4 += v3;
                    v3 += v2;
                    v4 += v2;
                    buffer4[k] = v1;
                    buffer4[k + 1] = v2;
                    buffer4[k + 2] = v3;
                    buffer4[k + 3] = v4;
                    buffer[(k >> 2) + (k >> (2 + PL))] = v4;
                } else {
                    float v = 0;
                    for (int k2 = k; k2 < n24_i; k2++) {
                        v += inp[(size_t)i * n + j + k2];
                        buffer4[k2] = v;
                    }
                    for (int k2 = n24_i; k2 < n24; k2++) buffer4[k2] = v;
                    buffer[(k >> 2) + (k >> (2 + PL))] = v;
                }
            }
            int u = 0;
            for (; (2 << u) <= n2; u++) {
                for (int k = 0; k < (int)(n2 >> (u + 1)); ++k) {
                    int i1 = (((k << 1) + 2) << u) - 1;
                    int i2 = (((k << 1) + 1) << u) - 1;
                    i1 += i1 >> PL;
                    i2 += i2 >> PL;
                    buffer[i1] += buffer[i2];
                }
            }
            u--;
            for (; u >= 0; u--) {
                for (int k = 0; k < (int)((n2 - (1 << u)) >> (u + 1)); ++k) {
                    int i1 = (((k << 1) + 3) << u) - 1;
                    int i2 = (((k << 1) + 2) << u) - 1;
                    i1 += i1 >> PL;
                    i2 += i2 >> PL;
                    buffer[i1] += buffer[i2];
                }
            }
            for (int k = 4; k < n24; k += 4) {
                const int k2 = ((k >> 2) - 1) + (((k >> 2) - 1) >> PL);
                buffer4[k] += buffer[k2];
                buffer4[k + 1] += buffer[k2];
                buffer4[k + 2] += buffer[k2];
                buffer4[k + 3] += buffer[k2];
            }
            for (int k = 0; k < n24_i; ++k) out[(size_t)i * n + j + k] = buffer4[k] + runningsum;
            const float t = buffer[(n2 - 1) + ((n2 - 1) >> PL)] + runningsum2;
            const float r2 = runningsum + t;
            runningsum2 = t - (r2 - runningsum);
            runningsum = r2;
        }
        free(buffer4);
        free(buffer);
    }
}

/* Reference: tf_sampling_g.cu:90-104 (binarysearchKernel) on top of the cumsum (probsampleLauncher :197-201).
 * temp is (b,n) working space holding the cumsum. */
ORACLE_API void oracle_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp,
                                   int *out) {
    oracle_cumsum(b, n, inp_p, temp);
    int base = 1;
    while (base < n) base <<= 1;
#pragma omp parallel for
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float q = inp_r[(size_t)i * m + j] * temp[(size_t)i * n + n - 1];
            int r = n - 1;
            for (int k = base; k >= 1; k >>= 1)
                if (r >= k && temp[(size_t)i * n + r - k] >= q) r -= k;
            out[(size_t)i * m + j] = r;
        }
}
