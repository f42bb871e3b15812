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

#include <algorithm>
#include <cstdlib>

namespace f3d {

constexpr int kCentresPerWarp = 4;
constexpr int kBqWarps = 8;  // warps per CTA
constexpr int kBqSortCentres = 4096;  // centres per cloud from which the grid query walks them in spatially binned order
constexpr int kRefStride = 256;  // blockDim.x of the reference launch (tf_grouping_g.cu:180): centre j belongs to thread j % 256

// per-cloud header of the grid path's workspace (bq_grid_build_kernel)
struct BqGridInfo {
    float x0, y0, inv_c;
    int ncx, ncy;
    int has_empty;  // set by the query kernel when a centre of this cloud has no point in its ball (the fallback kernel's cue)
    int pad[2];
};

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
                   const float *__restrict__ xyz2, int *__restrict__ idx, const int *__restrict__ pts_cnt,
                   const BqGridInfo *__restrict__ info) {
    const int lane = threadIdx.x & 31;
    // grid-stride over the centres; with the grid path's per-cloud flag a cloud without empty balls costs one load
    for (long long w = static_cast<long long>(blockIdx.x) * kBqWarps + (threadIdx.x >> 5); w < static_cast<long long>(b) * m;
         w += static_cast<long long>(gridDim.x) * kBqWarps) {
    const int batch = static_cast<int>(w / m);
    if (info && info[batch].has_empty == 0) continue;
    const int j = static_cast<int>(w - static_cast<long long>(batch) * m);
    const int *cntb = pts_cnt + static_cast<size_t>(batch) * m;
    if (cntb[j] != 0) continue;

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
}

// ------------------------------------------------------------------------------------------------------------------
// Grid-accelerated ball query (uniform radius).  The brute-force scan above tests every (centre, point) pair up to the
// nsample-th hit; on LiDAR clouds 95 % of those tests are against points tens of metres away.  Here the cloud is
// binned once per call into an xy grid of cell size c >= 1.002 r (counting sort in shared memory, one CTA per cloud);
// a centre then tests only the points of its 3x3 cell neighbourhood -- three contiguous runs of the sorted array --
// with the SAME distance expression and threshold as the scan, so the hit set is identical.  (A point that tests as a
// hit has |dx|,|dy| <= r (1 + 2^-22); the 0.2 % slack in c is two orders above the rounding of the cell coordinates,
// so it can never lie two cells away.)  The reference's contract is the first nsample hits in ASCENDING INDEX order:
// hits are recorded in a per-warp bitmap over the point indices (shared memory, n/8 bytes) and read back in index
// order with popc prefix sums, which also yields the first hit for padding.  Empty balls go to bq_fallback_kernel.
constexpr int kBqMaxCells = 4096;
constexpr int kBqBuildThreads = 1024;
// Large clouds (n >= kBqWindowedFrom) are binned per INDEX WINDOW: the point indices are cut into <= 32 windows of `win` consecutive
// indices and every window gets its own cell table over the common grid -- sorted order = [window][cell].  The reference's contract is
// the first nsample hits in ascending index order, and a KITTI-shape scan has thousands of points in the 3 x 3 cells of a centre
// (inference.py scores every point of a 131 072-point scan): a centre walks its windows in ascending order and stops at nsample hits,
// i.e. after ~1/30 of its candidates, and the hits of a window are ordered through 4 bitmap words per lane.
constexpr int kBqWindowedFrom = 32768;
constexpr int kBqMaxWpl = 8;  // bitmap words of one window per lane: window <= 8192 indices (n <= 262144, the grid path's limit)
constexpr int kBqSparse = 1024;  // centres with at most this many candidates in all windows together take them in one pass
__host__ __device__ inline int bq_window(int n) {  // indices per window: 1024 (= 32 lanes x 32 bits) x a power of two, 17 .. 32 windows
    if (n < kBqWindowedFrom) return n;
    int win = 1024;
    while (32LL * win < n) win <<= 1;  // (a power of two: a group of the grouped kernel holds a whole number of windows)
    return win;
}
__host__ __device__ inline int bq_num_windows(int n) { return n < kBqWindowedFrom ? 1 : (n + bq_window(n) - 1) / bq_window(n); }


__device__ __forceinline__ int bq_cell_coord(float v, float v0, float inv_c, int nc) {
    const float t = floorf((v - v0) * inv_c);
    return static_cast<int>(fminf(fmaxf(t, -2.0f), static_cast<float>(nc + 1)));
}

// grid (clouds, windows): CTA (cloud, w) sorts the points with indices [w win, (w+1) win) into sorted[w win ...] by cell and writes the
// cell table of its window; every CTA of a cloud derives the same grid from the whole cloud's bounding box.  One window = the whole cloud
// for n < kBqWindowedFrom and for the spatial binning of the centres.
__global__ void __launch_bounds__(kBqBuildThreads, 1)
bq_grid_build_kernel(int n, int win, float radius, const float *__restrict__ xyz1, float4 *__restrict__ sorted,
                     int *__restrict__ cell_start, BqGridInfo *__restrict__ info) {
    __shared__ int cursor[kBqMaxCells];
    __shared__ float red[4][32];
    __shared__ int wsum[32];
    __shared__ BqGridInfo gi;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz1 + static_cast<size_t>(blockIdx.x) * n * 3;
    float4 *out = sorted + static_cast<size_t>(blockIdx.x) * n;
    int *cs = cell_start + (static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * (kBqMaxCells + 1);
    const int k0 = blockIdx.y * win, k1 = min(n, k0 + win);  // this CTA's index window

    float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
    for (int k = tid; k < n; k += kBqBuildThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1);
        mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, s));
        mny = fminf(mny, __shfl_xor_sync(kFull, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, s));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mxx; red[2][warp] = mny; red[3][warp] = mxy; }
    for (int i = tid; i < kBqMaxCells; i += kBqBuildThreads) cursor[i] = 0;
    __syncthreads();
    if (warp == 0) {
        mnx = red[0][lane]; mxx = red[1][lane]; mny = red[2][lane]; mxy = red[3][lane];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, s));
            mny = fminf(mny, __shfl_xor_sync(kFull, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, s));
        }
        if (lane == 0) {
            const float ex = fmaxf(mxx - mnx, 0.0f), ey = fmaxf(mxy - mny, 0.0f);
            float c = radius * 1.002f;
            int ncx = 1, ncy = 1;
            for (int it = 0; it < 64; ++it) {  // grow the cell until the grid fits the shared-memory histogram
                const float inv = 1.0f / c;
                const float fx = floorf(ex * inv), fy = floorf(ey * inv);
                if (fx < 4000.0f && fy < 4000.0f) {
                    ncx = static_cast<int>(fx) + 1;
                    ncy = static_cast<int>(fy) + 1;
                    if (ncx * ncy <= kBqMaxCells) break;
                }
                c *= 1.25f;
            }
            if (!(ncx * ncy <= kBqMaxCells)) { ncx = 1; ncy = 1; c = 3.0e38f; }  // non-finite extents: a single cell
            gi.x0 = mnx; gi.y0 = mny; gi.inv_c = 1.0f / c; gi.ncx = ncx; gi.ncy = ncy; gi.has_empty = 0; gi.pad[0] = gi.pad[1] = 0;
            if (blockIdx.y == 0) info[blockIdx.x] = gi;
        }
    }
    __syncthreads();
    const float x0 = gi.x0, y0 = gi.y0, inv_c = gi.inv_c;
    const int ncx = gi.ncx, ncy = gi.ncy, ncells = ncx * ncy;
    auto cell_of = [&](float x, float y) {
        const int cx = min(max(bq_cell_coord(x, x0, inv_c, ncx), 0), ncx - 1);
        const int cy = min(max(bq_cell_coord(y, y0, inv_c, ncy), 0), ncy - 1);
        return cy * ncx + cx;
    };
    for (int k = k0 + tid; k < k1; k += kBqBuildThreads) atomicAdd(&cursor[cell_of(__ldg(p + 3 * k), __ldg(p + 3 * k + 1))], 1);
    __syncthreads();
    {  // exclusive scan of kBqMaxCells counters: 4 per thread, warp shuffle scan, 32 warp totals
        constexpr int PER = kBqMaxCells / kBqBuildThreads;
        int c[PER], tot = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { c[i] = cursor[tid * PER + i]; tot += c[i]; }
        int inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, inc, s); if (lane >= s) inc += v; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int w = wsum[lane];
            int winc = w;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, winc, s); if (lane >= s) winc += v; }
            wsum[lane] = winc - w;
        }
        __syncthreads();
        int run = k0 + wsum[warp] + inc - tot;  // positions are absolute: window w owns sorted[w win, ...)
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            cursor[tid * PER + i] = run;
            if (tid * PER + i <= ncells) cs[tid * PER + i] = run;
            run += c[i];
        }
        if (tid == kBqBuildThreads - 1 && ncells == kBqMaxCells) cs[kBqMaxCells] = run;
    }
    __syncthreads();
    for (int k = k0 + tid; k < k1; k += kBqBuildThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        const int pos = atomicAdd(&cursor[cell_of(x, y)], 1);
        out[pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

// one warp per centre; dynamic shared memory per warp: ceil(n/32) bitmap words + one SUMMARY bit per bitmap word.
// Hits land all over the index range (~110 of 16384 at the bench shape), so reading the bitmap back word by word meant
// 10-16 rounds of load + popc + 5-step warp scan over mostly empty words.  With the summary, lane l owns the contiguous
// block of bitmap words [l*32*SPL, (l+1)*32*SPL) and walks only its NON-EMPTY words (3-4 of them): one warp scan of the
// per-lane hit counts gives every lane its output offset, and ascending lane order = ascending index order.
__global__ void __launch_bounds__(256)
bq_grid_query_kernel(int b, int n, int m, float radius, int nsample, const float4 *__restrict__ sorted,
                     const int *__restrict__ cell_start, BqGridInfo *info,
                     const float *__restrict__ xyz2, int *__restrict__ idx, int *__restrict__ pts_cnt, const float4 *__restrict__ centres_sorted) {
    extern __shared__ unsigned bq_bitmap[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int nwords = (n + 31) >> 5;
    const int spl = (((nwords + 31) >> 5) + 31) >> 5;  // summary words per lane
    const int per_warp = nwords + 32 * spl;
    unsigned *bm = bq_bitmap + static_cast<size_t>(wl) * per_warp;
    unsigned *sm = bm + nwords;                        // summary word s covers bitmap words [32 s, 32 s + 32)
    const long long w = static_cast<long long>(blockIdx.x) * wpc + wl;
    if (w >= static_cast<long long>(b) * m) return;
    const int batch = static_cast<int>(w / m);
    int j = static_cast<int>(w - static_cast<long long>(batch) * m);
    float cx, cy, cz;
    if (centres_sorted) {
        // many centres per cloud (attention at every point, inference.py:99-131): the warps take the centres in the order of a spatial
        // binning of the centres themselves, so that the warps of a CTA (and of the CTAs that share its SM) walk the same candidate
        // cells and find them in L1 instead of each gathering its own 3 x 3 cells from L2.  .w = the centre's index (row of idx).
        const float4 rec = __ldg(centres_sorted + static_cast<size_t>(batch) * m + j);
        cx = rec.x; cy = rec.y; cz = rec.z;
        j = __float_as_int(rec.w);
    } else {
        const float *c = xyz2 + (static_cast<size_t>(batch) * m + j) * 3;
        cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
    }
    const float T = ball_threshold(radius);
    const BqGridInfo gi = info[batch];
    const float4 *pts = sorted + static_cast<size_t>(batch) * n;
    const int *cs = cell_start + static_cast<size_t>(batch) * (kBqMaxCells + 1);
    int *row = idx + (static_cast<size_t>(batch) * m + j) * nsample;

    for (int i = lane; i < per_warp; i += 32) bm[i] = 0;
    __syncwarp();
    const int gx = bq_cell_coord(cx, gi.x0, gi.inv_c, gi.ncx), gy = bq_cell_coord(cy, gi.y0, gi.inv_c, gi.ncy);
    const int c0 = max(gx - 1, 0), c1 = min(gx + 1, gi.ncx - 1);
    const int r0 = max(gy - 1, 0), r1 = min(gy + 1, gi.ncy - 1);
    if (c0 <= c1 && T > 0.0f) {
        for (int r = r0; r <= r1; ++r) {
            const int s = cs[r * gi.ncx + c0], e = cs[r * gi.ncx + c1 + 1];
            for (int i = s + lane; i < e; i += 32) {
                const float4 q = __ldg(pts + i);
                if (!(sqdist_ref(cx - q.x, cy - q.y, cz - q.z) >= T)) {
                    const int k = __float_as_int(q.w);
                    atomicOr(&bm[k >> 5], 1u << (k & 31));
                    atomicOr(&sm[k >> 10], 1u << ((k >> 5) & 31));
                }
            }
        }
    }
    __syncwarp();
    // pass 1: hits owned by this lane, and its lowest hit
    int mine = 0, low = 0x7fffffff;
    for (int sp = 0; sp < spl; ++sp) {
        const int s = lane * spl + sp;
        unsigned sw = sm[s];
        while (sw) {
            const int wi = s * 32 + __ffs(sw) - 1;
            sw &= sw - 1;
            const unsigned word = bm[wi];
            mine += __popc(word);
            low = min(low, wi * 32 + __ffs(word) - 1);  // words are visited in ascending order: min keeps the first
        }
    }
    int inc = mine;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, inc, s); if (lane >= s) inc += v; }
    const int H = __shfl_sync(kFull, inc, 31);
    const unsigned owners = __ballot_sync(kFull, mine > 0);
    const int first = owners ? __shfl_sync(kFull, low, __ffs(owners) - 1) : -1;
    // pass 2: this lane's hits go to row[off ...] in ascending index order, up to nsample
    int pos = inc - mine;
    if (mine > 0 && pos < nsample) {
        for (int sp = 0; sp < spl && pos < nsample; ++sp) {
            const int s = lane * spl + sp;
            unsigned sw = sm[s];
            while (sw && pos < nsample) {
                const int wi = s * 32 + __ffs(sw) - 1;
                sw &= sw - 1;
                unsigned word = bm[wi];
                while (word && pos < nsample) {
                    row[pos++] = wi * 32 + __ffs(word) - 1;
                    word &= word - 1;
                }
            }
        }
    }
    const int cc = min(H, nsample);
    if (cc > 0)
        for (int s = cc + lane; s < nsample; s += 32) row[s] = first;
    if (lane == 0) {
        pts_cnt[static_cast<size_t>(batch) * m + j] = cc;
        if (cc == 0) info[batch].has_empty = 1;  // every writer stores the same value
    }
}

// Windowed form of the grid query for n >= kBqWindowedFrom (see bq_window).  Persistent warps: warp g takes centres g, g + G, ...;
// its bitmap is zeroed once and every centre clears the words it set.  Lane w holds the three row ranges of window w.
//  * at most kBqSparse candidates over all windows: lane w tests the candidates of window w, the hits are ordered by the
//    summary walk of bq_grid_query_kernel;
//  * more: the windows are visited in ascending order with the whole warp striding over each row range, the hits of a window sit
//    in the window's own bitmap words (win / 1024 per lane, ascending lane order = ascending index order) and are emitted
//    before the next window is touched; the walk stops at nsample hits -- exactly the reference's break at cnt == nsample.
__global__ void __launch_bounds__(256)
bq_grid_query_win_kernel(int b, int n, int m, int win, int nwin, float radius, int nsample, const float4 *__restrict__ sorted,
                         const int *__restrict__ cell_start, BqGridInfo *info, const float *__restrict__ xyz2, int *__restrict__ idx,
                         int *__restrict__ pts_cnt, const float4 *__restrict__ centres_sorted, int alias_off, int alias_span) {
    extern __shared__ unsigned bq_bitmap[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int nwords = (n + 31) >> 5;
    const int spl = (((nwords + 31) >> 5) + 31) >> 5;  // summary words per lane
    const int per_warp = nwords + 32 * spl;
    unsigned *bm = bq_bitmap + static_cast<size_t>(wl) * per_warp;
    unsigned *sm = bm + nwords;                        // summary word s covers bitmap words [32 s, 32 s + 32)
    for (int i = lane; i < per_warp; i += 32) bm[i] = 0;
    __syncwarp();
    const float T = ball_threshold(radius);
    const int wpl = win >> 10;                         // bitmap words of one window per lane
    // alias_off >= 0 (one cloud): the centres ARE points alias_off .. alias_off + m - 1 of the cloud (same memory), centres_sorted is the
    // cloud's own binning from the first window that holds one of them on, alias_span records long: records of other points are skipped
    const long long total_warps = static_cast<long long>(gridDim.x) * wpc,
                    centres = alias_off >= 0 ? static_cast<long long>(alias_span) : static_cast<long long>(b) * m;
    // centre record (x, y, z, row): from the centres' own spatial binning when there is one (.w = the centre's index), else in order
    auto load_centre = [&](long long w) -> float4 {
        if (w >= centres) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (centres_sorted) return __ldg(centres_sorted + w);
        const float *c = xyz2 + w * 3;
        return make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), __int_as_float(static_cast<int>(w % m)));
    };
    float4 rec_next = load_centre(static_cast<long long>(blockIdx.x) * wpc + wl);
    for (long long w = static_cast<long long>(blockIdx.x) * wpc + wl; w < centres; w += total_warps) {
        const int batch = alias_off >= 0 ? 0 : static_cast<int>(w / m);
        const float4 rec = rec_next;
        rec_next = load_centre(w + total_warps);  // in flight under this centre's walk
        const float cx = rec.x, cy = rec.y, cz = rec.z;
        int j = __float_as_int(rec.w);
        if (alias_off >= 0) {
            j -= alias_off;
            if (j < 0 || j >= m) continue;  // (the whole warp: one centre per warp)
        }
        const BqGridInfo gi = info[batch];
        const float4 *pts = sorted + static_cast<size_t>(batch) * n;
        int *row = idx + (static_cast<size_t>(batch) * m + j) * nsample;
        const int gx = bq_cell_coord(cx, gi.x0, gi.inv_c, gi.ncx), gy = bq_cell_coord(cy, gi.y0, gi.inv_c, gi.ncy);
        const int c0 = max(gx - 1, 0), c1 = min(gx + 1, gi.ncx - 1);
        const int r0 = max(gy - 1, 0), r1 = min(gy + 1, gi.ncy - 1);
        // lane w: the (up to three) row ranges of window w
        int rs[3] = {0, 0, 0}, re[3] = {0, 0, 0};
        if (lane < nwin && c0 <= c1 && T > 0.0f) {
            const int *cs = cell_start + (static_cast<size_t>(batch) * nwin + lane) * (kBqMaxCells + 1);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (r0 + k <= r1) {
                    rs[k] = __ldg(cs + (r0 + k) * gi.ncx + c0);
                    re[k] = __ldg(cs + (r0 + k) * gi.ncx + c1 + 1);
                }
        }
        const int cw = (re[0] - rs[0]) + (re[1] - rs[1]) + (re[2] - rs[2]);
        int C = cw;
#pragma unroll
        for (int s2 = 16; s2 > 0; s2 >>= 1) C += __shfl_xor_sync(kFull, C, s2);
        int H = 0, first = -1;  // hits found (all of them below nsample, at least nsample otherwise), lowest hit
        if (C > 0 && C <= kBqSparse) {
            {  // the lane's three row ranges as one list, four loads in flight (a lone dependent load per step left the lane waiting on L2)
                const int l0 = re[0] - rs[0], l1 = re[1] - rs[1];
                auto at = [&](int i) { return i < l0 ? rs[0] + i : (i < l0 + l1 ? rs[1] + (i - l0) : rs[2] + (i - l0 - l1)); };
                for (int i0 = 0; i0 < cw; i0 += 4) {
                    float4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u < cw) q[u] = __ldg(pts + at(i0 + u));
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u < cw && !(sqdist_ref(cx - q[u].x, cy - q[u].y, cz - q[u].z) >= T)) {
                            const int kk = __float_as_int(q[u].w);
                            atomicOr(&bm[kk >> 5], 1u << (kk & 31));
                            atomicOr(&sm[kk >> 10], 1u << ((kk >> 5) & 31));
                        }
                }
            }
            __syncwarp();
            int mine = 0, low = 0x7fffffff;
            for (int sp = 0; sp < spl; ++sp) {
                const int s2 = lane * spl + sp;
                unsigned sw = sm[s2];
                while (sw) {
                    const int wi = s2 * 32 + __ffs(sw) - 1;
                    sw &= sw - 1;
                    const unsigned word = bm[wi];
                    mine += __popc(word);
                    low = min(low, wi * 32 + __ffs(word) - 1);
                }
            }
            int inc = mine;
#pragma unroll
            for (int s2 = 1; s2 < 32; s2 <<= 1) { const int v = __shfl_up_sync(kFull, inc, s2); if (lane >= s2) inc += v; }
            H = __shfl_sync(kFull, inc, 31);
            const unsigned owners = __ballot_sync(kFull, mine > 0);
            if (owners) first = __shfl_sync(kFull, low, __ffs(owners) - 1);
            int pos = inc - mine;
            for (int sp = 0; sp < spl; ++sp) {  // emit in ascending order up to nsample, clear every touched word
                const int s2 = lane * spl + sp;
                unsigned sw = sm[s2];
                if (!sw) continue;
                sm[s2] = 0;
                while (sw) {
                    const int wi = s2 * 32 + __ffs(sw) - 1;
                    sw &= sw - 1;
                    unsigned word = bm[wi];
                    bm[wi] = 0;
                    while (word && pos < nsample) {
                        row[pos++] = wi * 32 + __ffs(word) - 1;
                        word &= word - 1;
                    }
                }
            }
        } else if (C > 0) {
            unsigned todo = __ballot_sync(kFull, cw > 0);
            while (todo && H < nsample) {
                const int wn = __ffs(todo) - 1;
                todo &= todo - 1;
                // the window's three row ranges as ONE candidate list, eight candidates per lane in flight (each step of the walk is a
                // round trip to L2: the per-warp bitmaps leave the SM little L1)
                const int s0 = __shfl_sync(kFull, rs[0], wn), s1 = __shfl_sync(kFull, rs[1], wn), s2 = __shfl_sync(kFull, rs[2], wn);
                const int l0 = __shfl_sync(kFull, re[0], wn) - s0, l1 = __shfl_sync(kFull, re[1], wn) - s1;
                const int lt = l0 + l1 + __shfl_sync(kFull, re[2], wn) - s2;
                auto at = [&](int i) { return i < l0 ? s0 + i : (i < l0 + l1 ? s1 + (i - l0) : s2 + (i - l0 - l1)); };
                for (int i0 = lane; i0 < lt; i0 += 256) {
                    float4 q[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (i0 + 32 * u < lt) q[u] = __ldg(pts + at(i0 + 32 * u));
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (i0 + 32 * u < lt && !(sqdist_ref(cx - q[u].x, cy - q[u].y, cz - q[u].z) >= T)) {
                            const int kk = __float_as_int(q[u].w);
                            atomicOr(&bm[kk >> 5], 1u << (kk & 31));
                        }
                }
                __syncwarp();
                unsigned *wb = bm + static_cast<size_t>(wn) * (win >> 5) + lane * wpl;  // this lane's words of the window
                const int wbase = wn * win + lane * wpl * 32;
                int mine = 0, low = 0x7fffffff;
                unsigned wd[kBqMaxWpl];  // the lane's words of the window, in registers for the count and the emission
#pragma unroll
                for (int t = 0; t < kBqMaxWpl; ++t) {
                    wd[t] = (t < wpl && wbase + t * 32 < n) ? wb[t] : 0u;
                    mine += __popc(wd[t]);
                    if (wd[t] && low == 0x7fffffff) low = wbase + t * 32 + __ffs(wd[t]) - 1;
                }
                int inc = mine;
#pragma unroll
                for (int s2 = 1; s2 < 32; s2 <<= 1) { const int v = __shfl_up_sync(kFull, inc, s2); if (lane >= s2) inc += v; }
                const int Hw = __shfl_sync(kFull, inc, 31);
                const unsigned owners = __ballot_sync(kFull, mine > 0);
                if (first < 0 && owners) first = __shfl_sync(kFull, low, __ffs(owners) - 1);
                int pos = H + inc - mine;
#pragma unroll
                for (int t = 0; t < kBqMaxWpl; ++t) {
                    unsigned word = wd[t];
                    if (!word) continue;
                    wb[t] = 0;
                    while (word && pos < nsample) {
                        row[pos++] = wbase + t * 32 + __ffs(word) - 1;
                        word &= word - 1;
                    }
                }
                H += Hw;
                __syncwarp();
            }
        }
        const int cc = min(H, nsample);
        if (cc > 0)
            for (int s2 = cc + lane; s2 < nsample; s2 += 32) row[s2] = first;
        if (lane == 0) {
            pts_cnt[static_cast<size_t>(batch) * m + j] = cc;
            if (cc == 0) info[batch].has_empty = 1;  // every writer stores the same value
        }
        __syncwarp();
    }
}

// Grouped form of the windowed query (the default).  bq_grid_query_win_kernel gives every warp a bitmap of the whole cloud (16 KB at
// 131 072 points): 12 warps per SM, and ncu shows them waiting on their candidate loads (L2 round trips) with 22 % of the issue slots used.
// Here a warp's bitmap covers kBqGroupBits = 32 768 indices -- a GROUP of 32 768 / win consecutive windows, 4 KB + one summary word per lane
// whatever n is -- and a centre walks its groups in ascending order, stopping at nsample hits like the window walk does:
//  * a group with at most kBqSparse candidates: win / 1024 lanes per window, each testing every (win / 1024)-th candidate of that window
//    with four loads in flight; lane l then owns bitmap words [32 l, 32 l + 32) of the group and finds its non-empty ones in summary
//    word l (ascending lane order = ascending index order);
//  * a denser group: its windows one at a time with the whole warp striding over each, as in bq_grid_query_win_kernel.
// Static shared memory (8 warps x 4224 B) and <= 85 registers: 24 warps per SM.
constexpr int kBqGroupBits = 32768;
__global__ void __launch_bounds__(256, 3)
bq_grid_query_grp_kernel(int b, int n, int m, int win, int nwin, float radius, int nsample, const float4 *__restrict__ sorted,
                         const int *__restrict__ cell_start, BqGridInfo *info, const float *__restrict__ xyz2, int *__restrict__ idx,
                         int *__restrict__ pts_cnt, const float4 *__restrict__ centres_sorted, int alias_off, int alias_span, int mode) {
    __shared__ unsigned bq_group_bitmap[8][kBqGroupBits / 32 + 32];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    unsigned *bm = bq_group_bitmap[wl];
    unsigned *sm = bm + kBqGroupBits / 32;             // summary word l covers bitmap words [32 l, 32 l + 32)
    for (int i = lane; i < kBqGroupBits / 32 + 32; i += 32) bm[i] = 0;
    __syncwarp();
    const float T = ball_threshold(radius);
    const int wpl = win >> 10;                         // bitmap words of one window per lane = lanes per window of a sparse group pass
    const int G = kBqGroupBits / win;                  // windows per group (32 / wpl when win is a power of two)
    const int ngroups = (nwin + G - 1) / G;
    const int lane_win = lane / wpl, lane_sub = lane - lane_win * wpl;  // sparse pass: the lane serves window w0 + lane_win, candidates lane_sub, + wpl ...
    // (the host takes this kernel for b * m < 2^31 only: 32-bit centre numbers, no 64-bit division subroutine in the loop)
    const unsigned total_warps = gridDim.x * wpc, centres = alias_off >= 0 ? static_cast<unsigned>(alias_span) : static_cast<unsigned>(b) * m;
    auto load_centre = [&](unsigned w) -> float4 {
        if (w >= centres) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (centres_sorted) return __ldg(centres_sorted + w);
        const float *c = xyz2 + static_cast<size_t>(w) * 3;
        return make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), __int_as_float(static_cast<int>(w % static_cast<unsigned>(m))));
    };
    float4 rec_next = load_centre(blockIdx.x * wpc + wl);
    for (unsigned w = blockIdx.x * wpc + wl; w < centres; w += total_warps) {
        const int batch = alias_off >= 0 ? 0 : static_cast<int>(w / static_cast<unsigned>(m));
        const float4 rec = rec_next;
        rec_next = load_centre(w + total_warps);  // in flight under this centre's walk
        const float cx = rec.x, cy = rec.y, cz = rec.z;
        int j = __float_as_int(rec.w);
        if (alias_off >= 0) {
            j -= alias_off;
            if (j < 0 || j >= m) continue;  // (the whole warp: one centre per warp)
        }
        const BqGridInfo gi = info[batch];
        const float4 *pts = sorted + static_cast<size_t>(batch) * n;
        int *row = idx + (static_cast<size_t>(batch) * m + j) * nsample;
        const int gx = bq_cell_coord(cx, gi.x0, gi.inv_c, gi.ncx), gy = bq_cell_coord(cy, gi.y0, gi.inv_c, gi.ncy);
        const int c0 = max(gx - 1, 0), c1 = min(gx + 1, gi.ncx - 1);
        const int r0 = max(gy - 1, 0), r1 = min(gy + 1, gi.ncy - 1);
        // lane w: the (up to three) row ranges of window w
        int rs[3] = {0, 0, 0}, re[3] = {0, 0, 0};
        if (lane < nwin && c0 <= c1 && T > 0.0f) {
            const int *cs = cell_start + (static_cast<size_t>(batch) * nwin + lane) * (kBqMaxCells + 1);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (r0 + k <= r1) {
                    rs[k] = __ldg(cs + (r0 + k) * gi.ncx + c0);
                    re[k] = __ldg(cs + (r0 + k) * gi.ncx + c1 + 1);
                }
        }
        const int cw = (re[0] - rs[0]) + (re[1] - rs[1]) + (re[2] - rs[2]);
        const unsigned nonempty = __ballot_sync(kFull, cw > 0);
        int H = 0, first = -1;  // hits found (all of them below nsample, at least nsample otherwise), lowest hit
        for (int g = 0; g < ngroups && H < nsample; ++g) {
            const int w0 = g * G, gbase = w0 * win;  // first window of the group, index of bit 0 of its bitmap
            const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << w0) & nonempty;
            if (!gmask) continue;
            int Cg = (gmask >> lane) & 1u ? cw : 0;
#pragma unroll
            for (int s2 = 16; s2 > 0; s2 >>= 1) Cg += __shfl_xor_sync(kFull, Cg, s2);
            if (mode == 1 || (mode == 0 && Cg <= kBqSparse)) {
                // wpl lanes per window: the window's three row ranges as one list, every wpl-th candidate, four loads in flight
                // (win need not divide kBqGroupBits -- 3072, 5120 ... -- : lanes beyond the group's G windows sit the pass out)
                const bool serves = lane_win < G && w0 + lane_win < 32;
                const int wsrc = serves ? w0 + lane_win : 0, sub = lane_sub;
                const int a0 = __shfl_sync(kFull, rs[0], wsrc), a1 = __shfl_sync(kFull, rs[1], wsrc), a2 = __shfl_sync(kFull, rs[2], wsrc);
                const int l0 = __shfl_sync(kFull, re[0], wsrc) - a0, l1 = __shfl_sync(kFull, re[1], wsrc) - a1;
                const int e2 = __shfl_sync(kFull, re[2], wsrc);  // (by every lane: inside `serves ? ... : 0` the lanes that sit out skipped
                const int lt = serves ? l0 + l1 + e2 - a2 : 0;   //  a full-mask shuffle -- the fault of the 3072 / 6144-point windows)
                auto at = [&](int i) { return i < l0 ? a0 + i : (i < l0 + l1 ? a1 + (i - l0) : a2 + (i - l0 - l1)); };
                for (int i0 = sub; i0 < lt; i0 += 4 * wpl) {
                    float4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * wpl < lt) q[u] = __ldg(pts + at(i0 + u * wpl));
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + u * wpl < lt && !(sqdist_ref(cx - q[u].x, cy - q[u].y, cz - q[u].z) >= T)) {
                            const int kl = __float_as_int(q[u].w) - gbase;
                            atomicOr(&bm[kl >> 5], 1u << (kl & 31));
                            atomicOr(&sm[kl >> 10], 1u << ((kl >> 5) & 31));
                        }
                }
                __syncwarp();
                unsigned sw = sm[lane];
                int mine = 0, low = 0x7fffffff;
                for (unsigned tmp = sw; tmp; tmp &= tmp - 1) {
                    const int wi = lane * 32 + __ffs(tmp) - 1;
                    const unsigned word = bm[wi];
                    mine += __popc(word);
                    low = min(low, gbase + wi * 32 + __ffs(word) - 1);
                }
                int inc = mine;
#pragma unroll
                for (int s2 = 1; s2 < 32; s2 <<= 1) { const int v = __shfl_up_sync(kFull, inc, s2); if (lane >= s2) inc += v; }
                const int Hg = __shfl_sync(kFull, inc, 31);
                const unsigned owners = __ballot_sync(kFull, mine > 0);
                if (first < 0 && owners) first = __shfl_sync(kFull, low, __ffs(owners) - 1);
                int pos = H + inc - mine;
                if (sw) {  // emit in ascending order up to nsample, clear every touched word
                    sm[lane] = 0;
                    for (; sw; sw &= sw - 1) {
                        const int wi = lane * 32 + __ffs(sw) - 1;
                        unsigned word = bm[wi];
                        bm[wi] = 0;
                        while (word && pos < nsample) {
                            row[pos++] = gbase + wi * 32 + __ffs(word) - 1;
                            word &= word - 1;
                        }
                    }
                }
                H += Hg;
                __syncwarp();
            } else {
                unsigned todo = gmask;
                while (todo && H < nsample) {
                    const int wn = __ffs(todo) - 1;
                    todo &= todo - 1;
                    // the window's three row ranges as ONE candidate list, four candidates per lane in flight
                    const int s0 = __shfl_sync(kFull, rs[0], wn), s1 = __shfl_sync(kFull, rs[1], wn), s2 = __shfl_sync(kFull, rs[2], wn);
                    const int l0 = __shfl_sync(kFull, re[0], wn) - s0, l1 = __shfl_sync(kFull, re[1], wn) - s1;
                    const int lt = l0 + l1 + __shfl_sync(kFull, re[2], wn) - s2;
                    auto at = [&](int i) { return i < l0 ? s0 + i : (i < l0 + l1 ? s1 + (i - l0) : s2 + (i - l0 - l1)); };
                    for (int i0 = lane; i0 < lt; i0 += 128) {
                        float4 q[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (i0 + 32 * u < lt) q[u] = __ldg(pts + at(i0 + 32 * u));
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (i0 + 32 * u < lt && !(sqdist_ref(cx - q[u].x, cy - q[u].y, cz - q[u].z) >= T)) {
                                const int kl = __float_as_int(q[u].w) - gbase;
                                atomicOr(&bm[kl >> 5], 1u << (kl & 31));
                            }
                    }
                    __syncwarp();
                    unsigned *wb = bm + static_cast<size_t>(wn - w0) * (win >> 5) + lane * wpl;  // this lane's words of the window
                    const int wbase = wn * win + lane * wpl * 32;
                    int mine = 0, low = 0x7fffffff;
                    unsigned wd[kBqMaxWpl];
#pragma unroll
                    for (int t = 0; t < kBqMaxWpl; ++t) {
                        wd[t] = t < wpl ? wb[t] : 0u;
                        mine += __popc(wd[t]);
                        if (wd[t] && low == 0x7fffffff) low = wbase + t * 32 + __ffs(wd[t]) - 1;
                    }
                    int inc = mine;
#pragma unroll
                    for (int s3 = 1; s3 < 32; s3 <<= 1) { const int v = __shfl_up_sync(kFull, inc, s3); if (lane >= s3) inc += v; }
                    const int Hw = __shfl_sync(kFull, inc, 31);
                    const unsigned owners = __ballot_sync(kFull, mine > 0);
                    if (first < 0 && owners) first = __shfl_sync(kFull, low, __ffs(owners) - 1);
                    int pos = H + inc - mine;
#pragma unroll
                    for (int t = 0; t < kBqMaxWpl; ++t) {
                        unsigned word = wd[t];
                        if (!word) continue;
                        wb[t] = 0;
                        while (word && pos < nsample) {
                            row[pos++] = wbase + t * 32 + __ffs(word) - 1;
                            word &= word - 1;
                        }
                    }
                    H += Hw;
                    __syncwarp();
                }
            }
        }
        const int cc = min(H, nsample);
        if (cc > 0)
            for (int s2 = cc + lane; s2 < nsample; s2 += 32) row[s2] = first;
        if (lane == 0) {
            pts_cnt[static_cast<size_t>(batch) * m + j] = cc;
            if (cc == 0) info[batch].has_empty = 1;  // every writer stores the same value
        }
        __syncwarp();
    }
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
    bq_fallback_kernel<<<blocks_for(w, kBqWarps), kBqWarps * 32, 0, st>>>(b, n, m, nsample, xyz1, xyz2, idx, pts_cnt, nullptr);
    return check_launch("bq_fallback_kernel");
}

F3D_API size_t f3d_query_ball_point_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 256;
    return static_cast<size_t>(b) * n * sizeof(float4) + static_cast<size_t>(b) * bq_num_windows(n) * (kBqMaxCells + 1) * sizeof(int) +
           static_cast<size_t>(b) * sizeof(BqGridInfo) + 512;
}

// the grid path needs a workspace, a finite radius and a per-warp index bitmap that fits shared memory
static bool bq_grid_ok(int b, int n, float radius, const void *workspace, size_t workspace_bytes) {
    return workspace && workspace_bytes >= f3d_query_ball_point_workspace_bytes(b, n) && n <= 262144 && radius < 1.0e18f &&
           (reinterpret_cast<uintptr_t>(workspace) % 16 == 0);
}
struct BqWorkspace {
    float4 *sorted;
    int *cell_start;
    BqGridInfo *info;
};
static BqWorkspace bq_workspace(int b, int n, void *workspace) {
    BqWorkspace w;
    w.sorted = static_cast<float4 *>(workspace);
    w.cell_start = reinterpret_cast<int *>(w.sorted + static_cast<size_t>(b) * n);
    w.info = reinterpret_cast<BqGridInfo *>(w.cell_start + static_cast<size_t>(b) * bq_num_windows(n) * (kBqMaxCells + 1));
    return w;
}

// First half of f3d_query_ball_point_ws: bins xyz1 into the workspace.  It depends on the cloud only (not on the centres),
// so a caller may run it on a second stream while the centres are still being sampled (3dfeatnet_b200/pipeline.py does).
F3D_API int f3d_ball_grid_build(int b, int n, float radius, const float *xyz1, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || !(radius > 0.0f) || !xyz1) return fail(F3D_ERR_INVALID_ARGUMENT, "ball_grid_build: bad arguments");
    if (!bq_grid_ok(b, n, radius, workspace, workspace_bytes))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "ball_grid_build: workspace missing, too small or misaligned (or n > 262144)");
    if (b == 0) return 0;
    const BqWorkspace w = bq_workspace(b, n, workspace);
    bq_grid_build_kernel<<<dim3(b, bq_num_windows(n)), kBqBuildThreads, 0, as_stream(stream)>>>(n, bq_window(n), radius, xyz1, w.sorted,
                                                                                              w.cell_start, w.info);
    return check_launch("bq_grid_build_kernel");
}

// Second half: query_ball_point over a grid built by f3d_ball_grid_build with the same (b, n, radius, xyz1, workspace).
F3D_API int f3d_ball_grid_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                                int *pts_cnt, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f) || !xyz1 || !xyz2 || !idx || !pts_cnt)
        return fail(F3D_ERR_INVALID_ARGUMENT, "ball_grid_query: bad arguments");
    if (!bq_grid_ok(b, n, radius, workspace, workspace_bytes))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "ball_grid_query: workspace missing, too small or misaligned (or n > 262144)");
    const long long w = static_cast<long long>(b) * m;
    if (w == 0 || b == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const BqWorkspace ws = bq_workspace(b, n, workspace);
    const int nwords = (n + 31) / 32;
    const int wpc = nwords <= 1024 ? 8 : (nwords <= 4096 ? 4 : 2);
    const int spl = (((nwords + 31) / 32) + 31) / 32;  // summary words per lane (bq_grid_query_kernel)
    const size_t smem = static_cast<size_t>(wpc) * (nwords + 32 * spl) * sizeof(unsigned);
    cudaError_t e = cudaFuncSetAttribute(bq_grid_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "bq_grid_query: cudaFuncSetAttribute");
    // algorithmic bytes of the whole ball query: B * (12 N + 12 M + 4 M S + 4 M)
    ktimer_begin("bq_grid_query_kernel", static_cast<double>(b) * (12.0 * n + 12.0 * m + 4.0 * m * nsample + 4.0 * m), st);
    // with at least kBqSortCentres centres per cloud and room for a second binning in the workspace, the centres are binned too
    const float4 *centres_sorted = nullptr;
    const size_t off2 = (f3d_query_ball_point_workspace_bytes(b, n) + 255) & ~static_cast<size_t>(255);
    // centres that are a slice of the cloud itself (inference.py:118-131 scores every point of a scan, MAX_POINTS centres at a time): the
    // cloud's binning already holds them in a spatial order, window by window -- no second binning (26 us per chunk of a KITTI-shape scan)
    int alias_off = -1, alias_span = 0;
    if (b == 1 && n >= kBqWindowedFrom && m >= kBqSortCentres) {
        const uintptr_t p1 = reinterpret_cast<uintptr_t>(xyz1), p2 = reinterpret_cast<uintptr_t>(xyz2);
        if (p2 >= p1 && (p2 - p1) % 12 == 0 && (p2 - p1) / 12 + static_cast<uintptr_t>(m) <= static_cast<uintptr_t>(n)) {
            alias_off = static_cast<int>((p2 - p1) / 12);
            const int win = bq_window(n), first = alias_off / win * win;
            const long long last = (static_cast<long long>(alias_off) + m - 1) / win * win + win;
            alias_span = static_cast<int>((last < n ? last : n) - first);
            centres_sorted = ws.sorted + first;
        }
    }
    if (alias_off < 0 && m >= kBqSortCentres && m <= 262144 && workspace_bytes >= off2 + f3d_query_ball_point_workspace_bytes(b, m)) {
        const BqWorkspace w2 = bq_workspace(b, m, static_cast<char *>(workspace) + off2);
        bq_grid_build_kernel<<<b, kBqBuildThreads, 0, st>>>(m, m, radius, xyz2, w2.sorted, w2.cell_start, w2.info);  // one window: a spatial order
        const int rc2 = check_launch("bq_grid_build_kernel");
        if (rc2) return rc2;
        centres_sorted = w2.sorted;
    }
    if (n >= kBqWindowedFrom) {
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
            if (num_sms <= 0) num_sms = 148;
        }
        static const bool whole_cloud_bitmaps = std::getenv("F3D_BQ_WHOLE_CLOUD_BITMAPS") != nullptr;  // A/B: the kernel before the grouped one
        // bq_window() is a power of two, so win divides kBqGroupBits.  (With windows of 3072 / 6144 points -- the sizing before: any multiple
        // of 1024 -- the sparse pass of the grouped kernel, then 3 or 6 lanes per window, ended in "illegal instruction": one of its full-mask
        // shuffles sat inside `serves ? ... : 0` and the two lanes that serve no window skipped it -- found with progress markers in
        // host-mapped memory, tools/bq_fault_core.py, and hoisted.  The check stays as a guard: a window that does not divide the group
        // would take bq_grid_query_win_kernel.)
        if (!whole_cloud_bitmaps && w < (1LL << 31) - (1LL << 20) && kBqGroupBits % bq_window(n) == 0) {
            static const int grp_mode = std::getenv("F3D_BQ_GRP_MODE") ? std::atoi(std::getenv("F3D_BQ_GRP_MODE")) : 0;  // diagnosis: 1 sparse pass only, 2 window walk only
            const unsigned need = blocks_for(alias_off >= 0 ? alias_span : w, 8), cap = static_cast<unsigned>(num_sms) * 3u;
            bq_grid_query_grp_kernel<<<need < cap ? need : cap, 256, 0, st>>>(b, n, m, bq_window(n), bq_num_windows(n), radius, nsample, ws.sorted,
                                                                             ws.cell_start, ws.info, xyz2, idx, pts_cnt, centres_sorted,
                                                                             alias_off, alias_span, grp_mode);
            ktimer_end(st);
            int rcg = check_launch("bq_grid_query_grp_kernel");
            if (rcg) return rcg;
            const unsigned fbg = blocks_for(w, kBqWarps);
            bq_fallback_kernel<<<fbg < 296u ? fbg : 296u, kBqWarps * 32, 0, st>>>(b, n, m, nsample, xyz1, xyz2, idx, pts_cnt, ws.info);
            return check_launch("bq_fallback_kernel");
        }
        e = cudaFuncSetAttribute(bq_grid_query_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "bq_grid_query_win: cudaFuncSetAttribute");
        const unsigned per_sm = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024))));
        const unsigned need = blocks_for(alias_off >= 0 ? alias_span : w, wpc), cap = static_cast<unsigned>(num_sms) * per_sm;
        bq_grid_query_win_kernel<<<need < cap ? need : cap, wpc * 32, smem, st>>>(b, n, m, bq_window(n), bq_num_windows(n), radius, nsample,
                                                                               ws.sorted, ws.cell_start, ws.info, xyz2, idx, pts_cnt,
                                                                               centres_sorted, alias_off, alias_span);
    } else {
        bq_grid_query_kernel<<<blocks_for(w, wpc), wpc * 32, smem, st>>>(b, n, m, radius, nsample, ws.sorted, ws.cell_start, ws.info, xyz2,
                                                                        idx, pts_cnt, centres_sorted);
    }
    ktimer_end(st);
    int rc = check_launch("bq_grid_query_kernel");
    if (rc) return rc;
    const unsigned fb_blocks = blocks_for(w, kBqWarps);
    bq_fallback_kernel<<<fb_blocks < 296u ? fb_blocks : 296u, kBqWarps * 32, 0, st>>>(b, n, m, nsample, xyz1, xyz2, idx, pts_cnt, ws.info);
    return check_launch("bq_fallback_kernel");
}

F3D_API int f3d_query_ball_point_ws(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                                    int *idx, int *pts_cnt, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f) || !xyz1 || !xyz2 || !idx || !pts_cnt)
        return fail(F3D_ERR_INVALID_ARGUMENT, "query_ball_point: bad arguments");
    if (!bq_grid_ok(b, n, radius, workspace, workspace_bytes))
        return f3d_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
    if (static_cast<long long>(b) * m == 0) return 0;
    const int rc = f3d_ball_grid_build(b, n, radius, xyz1, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return f3d_ball_grid_query(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, workspace, workspace_bytes, stream);
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
