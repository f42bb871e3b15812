// probsample.cu -- cumsum + inverse-CDF sampling (prob_sample) for sm_100a.
//
// Replaces tf_sampling_g.cu:7-104 (cumsumKernel, binarysearchKernel; launchers :194-201).  Unused by the model
// (tf_sampling.py:60-89 only), kept for API completeness -- but bit-exact: the sampled index depends on every rounding
// of the prefix sum, so the kernel reproduces the reference's summation DAG, stated here in closed form rather than as
// the reference's in-place sweeps:
//   * the row is cut into chunks of 8192 values and groups of 4; inside a group the running sums are
//     v1, v1+v2, v3+(v1+v2), (v3+v4)+(v1+v2)                                                    (:20-33)
//   * group totals are combined pairwise into a pyramid L_{u+1}[i] = L_u[2i+1] + L_u[2i]        (up-sweep :46-56)
//   * the inclusive prefix of group i is the pyramid blocks of the binary digits of i+1, added from the largest block
//     to the smallest: acc = B_next + acc                                                        (down-sweep :57-67)
//   * out = (group-local sum + prefix of the previous group) + carry, the carry being updated per chunk with the
//     reference's compensated two-term form                                                      (:69-84)
#include "common.cuh"

namespace f3d {

constexpr int kCsChunk = 8192;
constexpr int kCsGroups = kCsChunk / 4;
constexpr int kCsThreads = 1024;

__global__ void __launch_bounds__(kCsThreads, 1)
cumsum_kernel(int n, const float *__restrict__ inp, float *__restrict__ out) {
    extern __shared__ float cs_smem[];
    float *g4 = cs_smem;                   // [kCsChunk]      running sums inside each group of four
    float *pyr = g4 + kCsChunk;            // [2*kCsGroups]   pyramid levels back to back
    float *pre = pyr + 2 * kCsGroups;      // [kCsGroups]     inclusive prefix over groups
    const float *x = inp + static_cast<size_t>(blockIdx.x) * n;
    float *y = out + static_cast<size_t>(blockIdx.x) * n;
    const int tid = threadIdx.x;
    float runningsum = 0.0f, runningsum2 = 0.0f;
    for (int j = 0; j < n; j += kCsChunk) {
        const int len = min(n - j, kCsChunk);
        const int n2 = (len + 3) >> 2;
        for (int g = tid; g < n2; g += kCsThreads) {
            const int k = g * 4;
            float tot;
            if (k + 3 < len) {
                const float v1 = x[j + k];
                const float v2 = x[j + k + 1] + v1;
                const float v34 = x[j + k + 3] + x[j + k + 2];
                const float v3 = x[j + k + 2] + v2;
                const float v4 = v34 + v2;
                g4[k] = v1; g4[k + 1] = v2; g4[k + 2] = v3; g4[k + 3] = v4;
                tot = v4;
            } else {
                float v = 0.0f;
                for (int k2 = k; k2 < len; ++k2) {
                    v += x[j + k2];
                    g4[k2] = v;
                }
                for (int k2 = len; k2 < k + 4; ++k2) g4[k2] = v;
                tot = v;
            }
            pyr[g] = tot;
        }
        __syncthreads();
        // pyramid: level u has n2 >> u complete blocks
        int off = 0, cnt = n2, nlev = 1;
        while ((cnt >> 1) >= 1) {
            const int nxt = cnt >> 1;
            for (int i = tid; i < nxt; i += kCsThreads) pyr[off + cnt + i] = pyr[off + 2 * i + 1] + pyr[off + 2 * i];
            __syncthreads();
            off += cnt;
            cnt = nxt;
            ++nlev;
        }
        for (int g = tid; g < n2; g += kCsThreads) {
            // blocks of the binary digits of g+1, largest first
            const int target = g + 1;
            int pos = 0, o = 0, c = n2;
            // offsets of the levels: recompute on the fly (level u starts at sum_{v<u} (n2 >> v))
            int lev_off[14];
            for (int u = 0; u < nlev; ++u) {
                lev_off[u] = o;
                o += c;
                c >>= 1;
            }
            float acc = 0.0f;
            bool first = true;
            for (int u = nlev - 1; u >= 0; --u) {
                if (target & (1 << u)) {
                    const float blk = pyr[lev_off[u] + (pos >> u)];
                    acc = first ? blk : blk + acc;
                    first = false;
                    pos += 1 << u;
                }
            }
            pre[g] = acc;
        }
        __syncthreads();
        for (int k = tid; k < len; k += kCsThreads) {
            const int g = k >> 2;
            const float v = g == 0 ? g4[k] : g4[k] + pre[g - 1];
            y[j + k] = v + runningsum;
        }
        const float t = pre[n2 - 1] + runningsum2;
        const float r2 = runningsum + t;
        runningsum2 = t - (r2 - runningsum);
        runningsum = r2;
        __syncthreads();
    }
}

// r = first index with cumsum[r] >= q * cumsum[n-1], branch-free binary search (tf_sampling_g.cu:90-104)
__global__ void prob_search_kernel(int n, int m, long long total, const float *__restrict__ cdf, const float *__restrict__ query,
                                   int *__restrict__ result) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bb = i / m;
    const float *d = cdf + bb * n;
    int base = 1;
    while (base < n) base <<= 1;
    const float q = query[i] * d[n - 1];
    int r = n - 1;
    for (int k = base; k >= 1; k >>= 1)
        if (r >= k && d[r - k] >= q) r -= k;
    result[i] = r;
}

}  // namespace f3d

using namespace f3d;

F3D_API int f3d_cumsum(int b, int n, const float *inp, float *out, void *stream) {
    if (b < 0 || n <= 0 || !inp || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "cumsum: bad arguments");
    if (b == 0) return 0;
    const size_t smem = sizeof(float) * (kCsChunk + 3 * kCsGroups);
    cudaError_t e = cudaFuncSetAttribute(cumsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "cumsum: cudaFuncSetAttribute");
    cumsum_kernel<<<b, kCsThreads, smem, as_stream(stream)>>>(n, inp, out);
    return check_launch("cumsum_kernel");
}

F3D_API int f3d_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || !inp_p || !inp_r || !temp || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "prob_sample: bad arguments");
    if (b == 0) return 0;
    int rc = f3d_cumsum(b, n, inp_p, temp, stream);
    if (rc) return rc;
    const long long total = static_cast<long long>(b) * m;
    if (total == 0) return 0;
    prob_search_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(n, m, total, temp, inp_r, out);
    return check_launch("prob_search_kernel");
}
