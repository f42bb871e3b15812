// nms.cu -- keypoint non-maximum suppression on the device.
//
// Replaces nms() of inference.py:226-261, which the reference runs on the CPU with a scikit-learn BallTree:
//   for every point: its 50 nearest neighbours (itself first, float64 distances); neighbours farther than nms_radius
//   are ignored; the point survives iff no consulted neighbour has a strictly larger attention (argmax == position 0);
//   survivors with attention <= max(attention) * min_response_ratio are dropped; the rest are sorted by
//   (attention, index) descending, the first max_keypoints kept, the tail padded with the best one.
// Device statement (identical whenever no two neighbours of a point are at exactly the same float64 distance -- the
// BallTree's order among exact ties is unspecified):
//   * distances are evaluated like the tree does: float64, d2 = (dx*dx + dy*dy) + dz*dz without FMA, d = sqrt(d2);
//     "d > radius" is tested as d2 > T with T = max{s : sqrt_rn(s) <= radius} (monotone IEEE sqrt => same predicate);
//   * the cloud is binned into an xy grid of cells >= nms_radius wide (counting sort), so a query only visits the 3 x 3 cells
//     around it -- O(N) instead of the O(N^2) all-pairs scan (612 ms -> a few ms at N = 131 072);
//   * a point with at most 49 other points inside the radius consults all of them (they are its nearest ones); a point
//     with more than 49 dies iff its nearest larger-attention neighbour, in (d2, index) order, is among the first 49;
//   * the threshold is computed in float64 like NumPy 1.19 does for float32_scalar * python_float
//     (requirements.txt:29);
//   * the descending (attention, index) order is produced by counting, for every survivor, the survivors that precede
//     it -- no sort, deterministic.
#include "common.cuh"

namespace f3d {


__device__ __forceinline__ double nms_d2(double ax, double ay, double az, double bx, double by, double bz) {
    const double dx = ax - bx, dy = ay - by, dz = az - bz;
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// largest double s with sqrt_rn(s) <= r
__device__ __forceinline__ double nms_threshold(double r) {
    if (!(r >= 0.0)) return -1.0;
    double t = __dmul_rn(r, r);
    for (int it = 0; it < 8; ++it) {
        if (sqrt(t) > r) t = __longlong_as_double(__double_as_longlong(t) - 1); else break;
    }
    for (int it = 0; it < 8; ++it) {
        const double u = __longlong_as_double(__double_as_longlong(t) + 1);
        if (sqrt(u) <= r) t = u; else break;
    }
    return t;
}

// ---- spatial binning: an xy grid of cells >= radius wide, so that every in-radius neighbour of a point lies in the 3 x 3
// cells around it.  Points are counting-sorted by cell (order inside a cell is irrelevant: the keep rule is a max / count,
// the slow path orders by (d2, index) itself).
struct NmsGrid {
    double x0, y0, inv_h;
    int nx, ny;
};

__device__ __forceinline__ NmsGrid nms_grid(const float *__restrict__ bbox, double radius, int max_cells) {
    NmsGrid g;
    g.x0 = bbox[0];
    g.y0 = bbox[1];
    double h = (radius > 1e-9 ? radius : 1e-9) * 1.000001 + 1e-12;
    const double ex = static_cast<double>(bbox[2]) - g.x0, ey = static_cast<double>(bbox[3]) - g.y0;
    for (int it = 0; it < 64; ++it) {
        const double nx = floor(ex / h) + 1.0, ny = floor(ey / h) + 1.0;
        if (nx * ny <= static_cast<double>(max_cells)) break;
        h *= 1.5;  // a cloud too wide for the cell budget: coarser cells (still >= radius), more candidates per query
    }
    g.inv_h = 1.0 / h;
    g.nx = static_cast<int>(floor(ex / h)) + 1;
    g.ny = static_cast<int>(floor(ey / h)) + 1;
    return g;
}

__device__ __forceinline__ int nms_cell_coord(double v, double v0, double inv_h, int n) {
    const int c = static_cast<int>(floor((v - v0) * inv_h));
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// bbox[b] = {min x, min y, max x, max y} and maxatt[b] = max attention of cloud b in ONE launch of gridDim.x CTAs per cloud (two launches of
// one CTA per cloud took 25 us each at 131 072 points): every CTA reduces a stripe, the last one to finish (a ticket per cloud in `done`,
// zeroed by the caller) folds the stripes' results.  min / max are order-independent: same bits whatever the finishing order.
__global__ void __launch_bounds__(1024)
nms_bbox_max_kernel(int n, const float *__restrict__ xyz, const float *__restrict__ attention, float *__restrict__ stripes,
                    int *__restrict__ done, float *__restrict__ bbox, float *__restrict__ maxatt) {
    __shared__ float red[5][32];
    __shared__ int s_last;
    const int batch = blockIdx.y, G = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p = xyz + static_cast<size_t>(batch) * n * 3;
    const float *att = attention + static_cast<size_t>(batch) * n;
    float v[5] = {3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.4e38f};  // min x, min y, max x, max y, max attention
    for (int k = blockIdx.x * 1024 + threadIdx.x; k < n; k += G * 1024) {
        const float x = p[3 * k], y = p[3 * k + 1];
        v[0] = fminf(v[0], x); v[1] = fminf(v[1], y); v[2] = fmaxf(v[2], x); v[3] = fmaxf(v[3], y);
        v[4] = fmaxf(v[4], att[k]);
    }
    auto fold = [&](int s) {
        v[0] = fminf(v[0], __shfl_xor_sync(kFull, v[0], s)); v[1] = fminf(v[1], __shfl_xor_sync(kFull, v[1], s));
        v[2] = fmaxf(v[2], __shfl_xor_sync(kFull, v[2], s)); v[3] = fmaxf(v[3], __shfl_xor_sync(kFull, v[3], s));
        v[4] = fmaxf(v[4], __shfl_xor_sync(kFull, v[4], s));
    };
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) fold(s);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 5; ++c) red[c][warp] = v[c];
    }
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = red[c][lane];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) fold(s);
    if (G > 1) {
        float *mine = stripes + (static_cast<size_t>(batch) * G + blockIdx.x) * 5;
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 5; ++c) __stcg(mine + c, v[c]);
            __threadfence();
            s_last = atomicAdd(done + batch, 1) == G - 1;
        }
        __syncwarp();
        if (!s_last) return;
        __threadfence();
        for (int c = lane; c < G; c += 32) {  // (the last CTA's own stripe is in v already; folding it again changes nothing)
            const float *o = stripes + (static_cast<size_t>(batch) * G + c) * 5;
            v[0] = fminf(v[0], __ldcg(o)); v[1] = fminf(v[1], __ldcg(o + 1)); v[2] = fmaxf(v[2], __ldcg(o + 2));
            v[3] = fmaxf(v[3], __ldcg(o + 3)); v[4] = fmaxf(v[4], __ldcg(o + 4));
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) fold(s);
    }
    if (lane == 0) {
        float *o = bbox + batch * 4;
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
        maxatt[batch] = v[4];
    }
}

__global__ void nms_cell_count_kernel(int n, double radius, int max_cells, const float *__restrict__ xyz, const float *__restrict__ bbox,
                                      int *__restrict__ cell_count) {
    const int batch = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const NmsGrid g = nms_grid(bbox + batch * 4, radius, max_cells);
    const float *p = xyz + (static_cast<size_t>(batch) * n + k) * 3;
    const int cell = nms_cell_coord(p[1], g.y0, g.inv_h, g.ny) * g.nx + nms_cell_coord(p[0], g.x0, g.inv_h, g.nx);
    atomicAdd(cell_count + static_cast<size_t>(batch) * (max_cells + 1) + cell, 1);
}

// exclusive scan of the cell counts (in place -> cell starts, entry [cells] = n) and a copy as the fill cursors, over the cells the
// cloud's grid really has (nx * ny + 1 entries of the max_cells + 1 the table is sized for).  Two launches over chunks of kScanChunk
// cells: chunk sums, then every CTA adds up the sums in front of its chunk and scans its own cells -- a 131 072-point scan has
// up to 524 289 cells, which took 356 us as a loop of one CTA.
constexpr int kScanChunk = 4096;  // 1024 threads x 4 consecutive cells
__device__ __forceinline__ int nms_block_sum(int v, int *warp_sum) {  // 1024 threads; result in every thread
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(kFull, v, s);
    __syncthreads();
    if (lane == 0) warp_sum[warp] = v;
    __syncthreads();
    int w = warp_sum[lane];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) w += __shfl_xor_sync(kFull, w, s);
    return w;
}
__global__ void __launch_bounds__(1024)
nms_cell_partial_kernel(double radius, int max_cells, const float *__restrict__ bbox, const int *__restrict__ cell_count,
                        int *__restrict__ partial) {
    __shared__ int warp_sum[32];
    const int batch = blockIdx.y;
    const NmsGrid g = nms_grid(bbox + batch * 4, radius, max_cells);
    const int last = g.nx * g.ny;  // entries 0 .. last
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * 4;
    if (blockIdx.x * kScanChunk > last) return;
    const int *cs = cell_count + static_cast<size_t>(batch) * (max_cells + 1);
    int v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + k <= last) v += cs[i0 + k];
    v = nms_block_sum(v, warp_sum);
    if (threadIdx.x == 0) partial[batch * gridDim.x + blockIdx.x] = v;
}
__global__ void __launch_bounds__(1024)
nms_cell_scan_kernel(double radius, int max_cells, const float *__restrict__ bbox, const int *__restrict__ partial,
                     int *__restrict__ cell_start, int *__restrict__ cursor) {
    __shared__ int warp_sum[32];
    const int batch = blockIdx.y;
    const NmsGrid g = nms_grid(bbox + batch * 4, radius, max_cells);
    const int last = g.nx * g.ny;
    if (blockIdx.x * kScanChunk > last) return;
    int *cs = cell_start + static_cast<size_t>(batch) * (max_cells + 1);
    int *cu = cursor + static_cast<size_t>(batch) * (max_cells + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int before = 0;  // cells of the chunks in front of this one
    for (int c = threadIdx.x; c < static_cast<int>(blockIdx.x); c += 1024) before += partial[batch * gridDim.x + c];
    before = nms_block_sum(before, warp_sum);
    const int i0 = blockIdx.x * kScanChunk + threadIdx.x * 4;
    int v[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = i0 + k <= last ? cs[i0 + k] : 0;
        tot += v[k];
    }
    int incl = tot;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, s);
        if (lane >= s) incl += t;
    }
    __syncthreads();
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sum[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int t = __shfl_up_sync(kFull, w, s);
            if (lane >= s) w += t;
        }
        warp_sum[lane] = w;
    }
    __syncthreads();
    int run = before + (warp ? warp_sum[warp - 1] : 0) + incl - tot;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k <= last) {
            cs[i0 + k] = run;
            cu[i0 + k] = run;
        }
        run += v[k];
    }
}

// sorted[pos] = index of the point at position pos of the cell-sorted order, sorted_pts[pos] = (x, y, z, attention) of that point: the
// query kernel walks the cloud in this order, so that the threads of a warp share their candidate cells (same loop bounds, broadcast loads)
// and read a candidate's record with one 16-byte load instead of index -> three coordinates -> attention
__global__ void nms_cell_fill_kernel(int n, double radius, int max_cells, const float *__restrict__ xyz, const float *__restrict__ attention,
                                     const float *__restrict__ bbox, int *__restrict__ cursor, int *__restrict__ sorted,
                                     float4 *__restrict__ sorted_pts) {
    const int batch = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const NmsGrid g = nms_grid(bbox + batch * 4, radius, max_cells);
    const float *p = xyz + (static_cast<size_t>(batch) * n + k) * 3;
    const int cell = nms_cell_coord(p[1], g.y0, g.inv_h, g.ny) * g.nx + nms_cell_coord(p[0], g.x0, g.inv_h, g.nx);
    const int pos = atomicAdd(cursor + static_cast<size_t>(batch) * (max_cells + 1) + cell, 1);
    sorted[static_cast<size_t>(batch) * n + pos] = k;
    sorted_pts[static_cast<size_t>(batch) * n + pos] = make_float4(p[0], p[1], p[2], attention[static_cast<size_t>(batch) * n + k]);
}

// keep[b,p] = 1 iff p is a local attention maximum among its (at most num_neighbors-1) nearest in-radius neighbours.
// One thread per query point scans the 3 x 3 cells around it (float64 distances, the tree's arithmetic) and records the
// number of in-radius neighbours and the NEAREST one with a larger attention (the "threat", ordered by (d2, index)).
// No threat: the point survives.  A threat and at most num_neighbors-1 neighbours: it is consulted, the point dies.
// Otherwise the tree only returns the num_neighbors-1 nearest: the point dies iff fewer than num_neighbors-1 neighbours
// precede the nearest threat -- a second scan that counts them (instead of materialising the sorted neighbour list).
//
// A float32 look at every candidate first (relative error of d2f < 3e-7): the float64 distance -- three conversions and seven
// double-precision operations per candidate -- is only evaluated where the float32 one cannot decide, i.e. within 1e-5 of the radius
// or for a larger-attention candidate that may be nearer than the nearest threat so far (a handful per query: the bound shrinks).
//
// The 32 queries of a warp are neighbours in the cell-sorted order.  When they sit in one grid row and span at most three cells (always
// the case where the cloud is dense, which is where the time goes: thousands of candidates per query near the sensor of a KITTI-shape
// scan) the warp walks the UNION of their candidate ranges together: 32 records per coalesced load into a shared-memory tile (the next
// tile's load in flight), every lane reading the tile by broadcast -- candidates outside a lane's own 3 x 3 cells are farther than the
// radius (cells are >= radius wide) and drop out at the float32 test, so the union changes nothing.  Other warps (row ends, sparse
// regions) keep the per-lane walk.
struct NmsQuery {
    float x, y, z, a;
    int t;
    float t_hi, t_lo;
    double T;
};
struct NmsThreat {
    int cnt;
    double td;    // nearest threat so far, (d2, index) order
    float td_hi;  // d2f > td_hi  =>  d > td
    int tk;
};
__device__ __forceinline__ void nms_visit_first(const NmsQuery &q, NmsThreat &s, const float4 c, int e, const int *__restrict__ srt) {
    const float fx = q.x - c.x, fy = q.y - c.y, fz = q.z - c.z;
    const float d2f = fx * fx + fy * fy + fz * fz;
    if (d2f > q.t_hi || e == q.t) return;  // outside for sure / the query itself
    const bool larger = c.w > q.a;
    if (d2f < q.t_lo && !(larger && d2f <= s.td_hi)) {  // inside for sure, and not the nearest threat for sure
        ++s.cnt;
        return;
    }
    const double d = nms_d2(q.x, q.y, q.z, c.x, c.y, c.z);
    if (d > q.T) return;
    ++s.cnt;
    if (larger && d <= s.td) {  // ties in attention: position 0 (self) wins the argmax
        const int k = __ldg(srt + e);
        if (d < s.td || k < s.tk) {
            s.td = d;
            s.tk = k;
            s.td_hi = __double2float_ru(d) * 1.00002f;
        }
    }
}
// second scan: the in-radius neighbours that precede the threat in (d2, index) order
__device__ __forceinline__ void nms_visit_second(const NmsQuery &q, const NmsThreat &s, float td_lo, int &before, const float4 c, int e,
                                                 const int *__restrict__ srt) {
    const float fx = q.x - c.x, fy = q.y - c.y, fz = q.z - c.z;
    const float d2f = fx * fx + fy * fy + fz * fz;
    if (d2f > q.t_hi || d2f > s.td_hi || e == q.t) return;  // outside, or behind the threat, for sure
    if (d2f < q.t_lo && d2f < td_lo) {                      // inside and in front of the threat for sure
        ++before;
        return;
    }
    const double d = nms_d2(q.x, q.y, q.z, c.x, c.y, c.z);
    if (d > q.T) return;
    before += (d < s.td || (d == s.td && __ldg(srt + e) < s.tk)) ? 1 : 0;
}
// (slice, slices): this warp's share of the candidates -- every slices-th candidate of a per-lane walk, every slices-th tile of a warp walk
template <bool kSecond>
__device__ __forceinline__ void nms_walk_lane(const NmsQuery &q, NmsThreat &s, float td_lo, int &before, const int *__restrict__ cs, int nx,
                                              int y_lo, int y_hi, int x_lo, int x_hi, const float4 *__restrict__ pts,
                                              const int *__restrict__ srt, int slice, int slices) {
    for (int yy = y_lo; yy <= y_hi; ++yy) {
        const int e1 = cs[yy * nx + x_hi + 1];
        for (int e = cs[yy * nx + x_lo] + slice; e < e1; e += slices) {  // the cells of one grid row are contiguous
            const float4 c = __ldg(pts + e);
            if (kSecond) nms_visit_second(q, s, td_lo, before, c, e, srt);
            else nms_visit_first(q, s, c, e, srt);
        }
    }
}
template <bool kSecond>
__device__ __forceinline__ void nms_walk_warp(const NmsQuery &q, NmsThreat &s, float td_lo, int &before, const int *__restrict__ cs, int nx,
                                              int y_lo, int y_hi, int x_lo, int x_hi, const float4 *__restrict__ pts,
                                              const int *__restrict__ srt, float4 *tile, int lane, int slice, int slices) {
    for (int yy = y_lo; yy <= y_hi; ++yy) {  // (bounds are the warp's: every lane walks the same candidates)
        const int e0 = cs[yy * nx + x_lo] + 32 * slice, e1 = cs[yy * nx + x_hi + 1];
        float4 next = e0 + lane < e1 ? __ldg(pts + e0 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = e0; base < e1; base += 32 * slices) {
            __syncwarp();
            tile[lane] = next;
            __syncwarp();
            if (base + 32 * slices + lane < e1) next = __ldg(pts + base + 32 * slices + lane);
            const int lim = min(32, e1 - base);
#pragma unroll 4
            for (int j = 0; j < lim; ++j) {
                const float4 c = tile[j];
                if (kSecond) nms_visit_second(q, s, td_lo, before, c, base + j, srt);
                else nms_visit_first(q, s, c, base + j, srt);
            }
        }
    }
}
// One CTA of kNmsSlices warps per 32 queries (lane = query in every warp): warp w takes every kNmsSlices-th tile of the candidates, the
// partial (count, nearest threat) and `before` counts are folded through shared memory.  The kernel's time is the walk of the densest
// group of queries (a single warp spent ~250 us on the thousands of candidates around the sensor of a KITTI-shape scan, twice when the
// 50-NN rule asks for the second scan, while most of the device had long finished).
constexpr int kNmsSlices = 4;
__global__ void __launch_bounds__(32 * kNmsSlices)
nms_keep_kernel(int n, double radius, int num_neighbors, int max_cells, const float *__restrict__ bbox, const int *__restrict__ cell_start,
                const int *__restrict__ sorted, const float4 *__restrict__ sorted_pts, unsigned char *__restrict__ keep) {
    __shared__ float4 tiles[kNmsSlices][32];
    __shared__ NmsGrid s_grid;
    __shared__ double s_T;
    __shared__ double p_td[kNmsSlices][32];
    __shared__ int p_tk[kNmsSlices][32], p_cnt[kNmsSlices][32];
    const int batch = blockIdx.y, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_grid = nms_grid(bbox + batch * 4, radius, max_cells);
        s_T = nms_threshold(radius);
    }
    __syncthreads();
    const int t0 = blockIdx.x * 32 + lane;  // position in the cell-sorted order: neighbouring lanes, neighbouring cells
    const int t = min(t0, n - 1);
    const int *cs = cell_start + static_cast<size_t>(batch) * (max_cells + 1);
    const int *srt = sorted + static_cast<size_t>(batch) * n;
    const float4 *pts = sorted_pts + static_cast<size_t>(batch) * n;
    const NmsGrid g = s_grid;
    const float4 me = __ldg(pts + t);
    NmsQuery q;
    q.x = me.x; q.y = me.y; q.z = me.z; q.a = me.w; q.t = t;
    q.T = s_T;
    q.t_hi = __double2float_ru(q.T * 1.00001);
    q.t_lo = __double2float_rd(q.T * 0.99999);
    const int cx = nms_cell_coord(me.x, g.x0, g.inv_h, g.nx), cy = nms_cell_coord(me.y, g.y0, g.inv_h, g.ny);
    const int y_lo = max(cy - 1, 0), y_hi = min(cy + 1, g.ny - 1), x_lo = max(cx - 1, 0), x_hi = min(cx + 1, g.nx - 1);
    // the warp's cells: one row, at most three cells wide -> walk the union together
    const int cx_min = __reduce_min_sync(kFull, cx), cx_max = __reduce_max_sync(kFull, cx);
    const bool together = __all_sync(kFull, cy == __shfl_sync(kFull, cy, 0)) && cx_max - cx_min <= 2;
    const int wx_lo = max(cx_min - 1, 0), wx_hi = min(cx_max + 1, g.nx - 1);
    float4 *tile = tiles[wl];
    NmsThreat s;
    s.cnt = 0; s.td = 1.0e300; s.td_hi = 3.4e38f; s.tk = 0x7fffffff;
    int before = 0;
    if (together) nms_walk_warp<false>(q, s, 0.f, before, cs, g.nx, y_lo, y_hi, wx_lo, wx_hi, pts, srt, tile, lane, wl, kNmsSlices);
    else nms_walk_lane<false>(q, s, 0.f, before, cs, g.nx, y_lo, y_hi, x_lo, x_hi, pts, srt, wl, kNmsSlices);
    p_cnt[wl][lane] = s.cnt; p_td[wl][lane] = s.td; p_tk[wl][lane] = s.tk;
    __syncthreads();
    s.cnt = 0; s.td = 1.0e300; s.tk = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < kNmsSlices; ++w) {  // every warp folds the same partials in the same order
        s.cnt += p_cnt[w][lane];
        const double d = p_td[w][lane];
        const int k = p_tk[w][lane];
        if (d < s.td || (d == s.td && k < s.tk)) { s.td = d; s.tk = k; }
    }
    unsigned char kp = 1;
    const bool again = s.tk != 0x7fffffff && s.cnt > num_neighbors - 1;  // the tree only returns the num_neighbors-1 nearest
    if (s.tk != 0x7fffffff) kp = 0;
    if (!__syncthreads_or(again)) {  // (uniform over the CTA)
        if (wl == 0 && t0 < n) keep[static_cast<size_t>(batch) * n + __ldg(srt + t)] = kp;
        return;
    }
    s.td_hi = again ? __double2float_ru(s.td) * 1.00002f : 0.f;
    const float td_lo = again ? __double2float_rd(s.td) * 0.99998f : 0.f;  // d2f < td_lo  =>  d < td
    if (together) {
        nms_walk_warp<true>(q, s, td_lo, before, cs, g.nx, y_lo, y_hi, wx_lo, wx_hi, pts, srt, tile, lane, wl, kNmsSlices);
    } else if (again) {
        nms_walk_lane<true>(q, s, td_lo, before, cs, g.nx, y_lo, y_hi, x_lo, x_hi, pts, srt, wl, kNmsSlices);
    }
    p_cnt[wl][lane] = before;  // (the first fold's reads are behind the barrier of __syncthreads_or)
    __syncthreads();
    if (wl != 0) return;
    before = 0;
#pragma unroll
    for (int w = 0; w < kNmsSlices; ++w) before += p_cnt[w][lane];
    if (again && before >= num_neighbors - 1) kp = 1;  // the threat is not among the neighbours the tree returns
    if (t0 < n) keep[static_cast<size_t>(batch) * n + __ldg(srt + t)] = kp;
}

// Survivors' attentions lie in (threshold, max]: a monotone map of that interval onto kNmsBuckets buckets (float subtraction, product and
// truncation are all monotone) lets the compaction histogram them on the side -- the top-K selection then reads which bucket the K-th best
// falls into instead of searching for it.
constexpr int kNmsBuckets = 2048;
__device__ __forceinline__ int nms_bucket(float a, float lo, float scale) {
    const int v = static_cast<int>((a - lo) * scale);
    return v < 0 ? 0 : (v >= kNmsBuckets ? kNmsBuckets - 1 : v);
}
__global__ void nms_compact_kernel(int n, double ratio, const float *__restrict__ attention, const unsigned char *__restrict__ keep,
                                   const float *__restrict__ maxatt, int *__restrict__ list, int *__restrict__ bucket, int *__restrict__ count,
                                   int *__restrict__ hist) {
    const int batch = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float mx = maxatt[batch];
    const double thresh = static_cast<double>(mx) * ratio;  // float32 scalar * python float -> float64 (NumPy 1.19)
    const float a = attention[static_cast<size_t>(batch) * n + k];
    if (keep[static_cast<size_t>(batch) * n + k] == 1 && static_cast<double>(a) > thresh) {
        const float lo = __double2float_rd(thresh);
        float scale = static_cast<float>(kNmsBuckets) / (mx - lo);
        if (!(scale > 0.0f && scale < 3.0e38f)) scale = 0.0f;  // (max == threshold, infinities: one bucket)
        const int bk = nms_bucket(a, lo, scale);
        const int pos = atomicAdd(count + batch, 1);
        list[static_cast<size_t>(batch) * n + pos] = k;
        bucket[static_cast<size_t>(batch) * n + pos] = bk;
        atomicAdd(hist + batch * kNmsBuckets + bk, 1);
    }
}

// The max_keypoints best survivors in (attention, index) DESCENDING order -- out_idx[rank] = point index -- without ranking every survivor
// against every other one (a KITTI-shape scan leaves tens of thousands; the all-pairs count took 146 us).  One CTA per cloud:
//  1. the bucket histogram of nms_compact_kernel, summed from the top, names the bucket the K-th best falls into; survivors of higher
//     buckets are kept outright, those of that bucket (a few dozen when the attentions differ) are the candidates for the places left;
//  2. among the candidates, the r-th best is found by an MSB-first radix select over the 64-bit value (order-preserving attention key,
//     point index) -- passes of 8 bits with warp-aggregated histogram updates, stopping as soon as the digit of the r-th holds exactly
//     what is still needed; ties of the K-th go on into the index bits, so that a cloud of equal attentions keeps exactly K too;
//  3. the K kept survivors are ranked among themselves with the reference's order.
__device__ __forceinline__ unsigned nms_okey(float a) {  // a < b  <=>  okey(a) < okey(b);  -0 counts as +0 like the float comparison
    const unsigned u = __float_as_uint(a + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(1024)
nms_topk_kernel(int n, int max_keypoints, const float *__restrict__ attention, const int *__restrict__ list, const int *__restrict__ bucket,
                const int *__restrict__ count, const int *__restrict__ bucket_hist, int *__restrict__ cand, unsigned *__restrict__ tie_key,
                int *__restrict__ tie_pos, int *__restrict__ out_idx) {
    constexpr int kStage = 2048;
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_r, s_m, s_t, s_done, s_cut, s_above;
    __shared__ __align__(16) unsigned long long tv[kStage];
    const int batch = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int cnt = count[batch];
    if (cnt == 0) return;
    const int *lst = list + static_cast<size_t>(batch) * n;
    const int *bkt = bucket + static_cast<size_t>(batch) * n;
    const float *att = attention + static_cast<size_t>(batch) * n;
    int *cd = cand + static_cast<size_t>(batch) * n;         // positions (in the survivor list) of the kept ones
    unsigned *ky = tie_key + static_cast<size_t>(batch) * n;  // candidates of the cut bucket: key ...
    int *tp = tie_pos + static_cast<size_t>(batch) * n;       // ... and position in the survivor list
    int *out = out_idx + static_cast<size_t>(batch) * max_keypoints;
    if (tid == 0) {
        s_m = 0;
        s_t = 0;
        s_cut = -1;  // cnt <= max_keypoints: every survivor is kept
        s_above = 0;
    }
    __syncthreads();
    if (cnt > max_keypoints && tid < 32) {  // 1. the bucket of the K-th best: lane l sums buckets [B - 64 (l + 1), B - 64 l) from the top
        const int *h = bucket_hist + batch * kNmsBuckets;
        constexpr int kPer = kNmsBuckets / 32;
        const int top = kNmsBuckets - 1 - lane * kPer;
        int tot = 0;
        for (int i = 0; i < kPer; ++i) tot += h[top - i];
        int inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_up_sync(kFull, inc, s);
            if (lane >= s) inc += v;
        }
        if (inc >= max_keypoints && inc - tot < max_keypoints) {  // exactly one lane: the sum crosses K inside its buckets
            int cum = inc - tot;
            for (int i = 0; i < kPer; ++i) {
                const int c = h[top - i];
                if (cum + c >= max_keypoints) {
                    s_cut = top - i;
                    s_above = cum;
                    break;
                }
                cum += c;
            }
        }
    }
    __syncthreads();
    const int cut = s_cut, r0 = max_keypoints - s_above;  // r0 places left for the candidates of bucket `cut`
    for (int e0 = tid; e0 < cnt; e0 += 4096) {
        int bb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) bb[u] = e0 + u * 1024 < cnt ? bkt[e0 + u * 1024] : -2;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (bb[u] > cut) cd[atomicAdd(&s_m, 1)] = e0 + u * 1024;
            else if (bb[u] == cut) tp[atomicAdd(&s_t, 1)] = e0 + u * 1024;
        }
    }
    __syncthreads();
    const int T = s_t;
    if (T > 0) {  // 2. the r0 best of the T candidates (T >= r0 by the choice of the bucket)
        for (int c = tid; c < T; c += 1024) ky[c] = nms_okey(att[lst[tp[c]]]);
        __syncthreads();
        unsigned long long prefix = 0, mask = 0, thr = 0;
        int r = r0;  // the r-th largest value among those that match the prefix
        if (T > r0) {
            for (int shift = 56; shift >= 0; shift -= 8) {
                for (int d = tid; d < 256; d += 1024) hist[d] = 0;
                __syncthreads();
                for (int c0 = 0; c0 < T; c0 += 1024) {  // (whole warps take part in the match)
                    const int c = c0 + tid;
                    unsigned long long v = c < T ? static_cast<unsigned long long>(ky[c]) << 32 : 0ull;
                    if (shift < 32 && c < T) v |= static_cast<unsigned>(lst[tp[c]]);
                    const bool in = c < T && (v & mask) == prefix;
                    const unsigned d = in ? static_cast<unsigned>(v >> shift) & 255u : 256u;
                    const unsigned same = __match_any_sync(kFull, d);
                    if (in && lane == __ffs(same) - 1) atomicAdd(&hist[d], __popc(same));
                }
                __syncthreads();
                if (tid == 0) {
                    int cum = 0;
                    for (int d = 255; d >= 0; --d) {
                        const int c = hist[d];
                        if (cum + c >= r) {
                            s_prefix = prefix | (static_cast<unsigned long long>(d) << shift);
                            s_r = r - cum;
                            s_done = c == r - cum;  // the whole digit is wanted: its lower bits need no look
                            break;
                        }
                        cum += c;
                    }
                }
                __syncthreads();
                prefix = s_prefix;
                r = s_r;
                mask |= 0xffull << shift;
                const bool done = s_done != 0;
                __syncthreads();
                if (done) break;
            }
            thr = prefix;
        }
        for (int c = tid; c < T; c += 1024)
            if ((static_cast<unsigned long long>(ky[c]) << 32 | static_cast<unsigned>(lst[tp[c]])) >= thr) cd[atomicAdd(&s_m, 1)] = tp[c];
        __syncthreads();
    }
    const int M = s_m;  // = min(cnt, max_keypoints)
    // 3. rank of a kept survivor = how many kept (key, index) values are larger than its own (the values are distinct).  The values are
    // staged in shared memory and read two per 16-byte broadcast load: with one 4-byte load per key and per index the 1024 x 1024
    // comparisons of a KITTI-shape scan spent 34 us on the shared-memory pipe alone.
    auto value_at = [&](int c) {
        const int i = lst[cd[c]];
        return static_cast<unsigned long long>(nms_okey(att[i])) << 32 | static_cast<unsigned>(i);
    };
    for (int c0 = 0; c0 < M; c0 += 1024) {  // (all threads walk the staging loop)
        const int c = c0 + tid;
        const unsigned long long vc = c < M ? value_at(c) : ~0ull;
        int rank = 0;
        for (int base = 0; base < M; base += kStage) {
            const int lim = min(kStage, M - base);
            __syncthreads();
            for (int j = tid; j < ((lim + 7) & ~7); j += 1024) tv[j] = j < lim ? value_at(base + j) : 0ull;  // (0 is larger than nothing)
            __syncthreads();
            const ulonglong2 *t2 = reinterpret_cast<const ulonglong2 *>(tv);
            for (int j = 0; j < (lim + 7) / 8 * 4; j += 4) {
                const ulonglong2 a0 = t2[j], a1 = t2[j + 1], a2 = t2[j + 2], a3 = t2[j + 3];
                rank += (a0.x > vc) + (a0.y > vc) + (a1.x > vc) + (a1.y > vc) + (a2.x > vc) + (a2.y > vc) + (a3.x > vc) + (a3.y > vc);
            }
        }
        if (c < M && rank < max_keypoints) out[rank] = static_cast<int>(vc & 0xffffffffu);
    }
}

__global__ void nms_finalize_kernel(int n, int max_keypoints, const float *__restrict__ xyz, const float *__restrict__ attention,
                                    const int *__restrict__ count, int *__restrict__ out_idx, float *__restrict__ out_xyz,
                                    float *__restrict__ out_att, int *__restrict__ num_keypoints) {
    const int batch = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= max_keypoints) return;
    const int num = min(count[batch], max_keypoints);
    int *oi = out_idx + static_cast<size_t>(batch) * max_keypoints;
    const int best = num > 0 ? oi[0] : 0;  // the reference raises when nothing survives; here: index 0, count 0
    const int k = s < num ? oi[s] : best;
    if (s >= num) oi[s] = k;  // oi[0] is never rewritten when num > 0, so reading it above races with nothing
    const float *p = xyz + (static_cast<size_t>(batch) * n + k) * 3;
    float *o = out_xyz + (static_cast<size_t>(batch) * max_keypoints + s) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    out_att[static_cast<size_t>(batch) * max_keypoints + s] = attention[static_cast<size_t>(batch) * n + k];
    if (s == 0) num_keypoints[batch] = num;
}

}  // namespace f3d

using namespace f3d;

static int nms_max_cells(int n) {
    long long c = 4LL * n;
    if (c < 4096) c = 4096;
    if (c > (1 << 22)) c = 1 << 22;
    return static_cast<int>(c);
}

F3D_API size_t f3d_nms_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 256;
    const size_t bn = static_cast<size_t>(b) * n;
    const size_t cells = static_cast<size_t>(nms_max_cells(n)) + 1;
    return bn * (1 + 4 + 4 + 4 + 16) + static_cast<size_t>(b) * cells * 8 + static_cast<size_t>(b) * (32 + kNmsBuckets * 4) + 2048 +
           static_cast<size_t>(b) * ((cells + kScanChunk - 1) / kScanChunk) * 4 + 64;
}

F3D_API int f3d_nms(int b, int n, const float *xyz, const float *attention, double nms_radius, double min_response_ratio,
                    int max_keypoints, int num_neighbors, int *out_idx, float *out_xyz, float *out_attention,
                    int *num_keypoints, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || max_keypoints <= 0 || num_neighbors <= 0 || !xyz || !attention || !out_idx || !out_xyz || !out_attention ||
        !num_keypoints || !(nms_radius >= 0.0))
        return fail(F3D_ERR_INVALID_ARGUMENT, "nms: bad arguments");
    if (n < num_neighbors)  // sklearn: "Expected n_neighbors <= n_samples"
        return fail(F3D_ERR_INVALID_ARGUMENT, "nms: fewer points than num_neighbors (the reference's BallTree query raises)");
    if (!workspace || workspace_bytes < f3d_nms_workspace_bytes(b, n)) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "nms: workspace too small");
    if (b == 0) return 0;
    cudaStream_t st = as_stream(stream);
    const size_t bn = static_cast<size_t>(b) * n;
    const int max_cells = nms_max_cells(n);
    const size_t cells = static_cast<size_t>(max_cells) + 1;
    // layout: [counts: 2b ints][maxatt: b floats][bbox: 4b floats][bucket histogram: b*kNmsBuckets ints][pad] [cell_start: b*cells ints][cursor: b*cells ints]
    //         [list: bn ints][dense_list: bn ints][sorted: bn ints][sorted_pts: bn float4][keep: bn bytes]
    char *base = static_cast<char *>(workspace);
    int *count = reinterpret_cast<int *>(base);
    int *dense_count = count + b;
    float *maxatt = reinterpret_cast<float *>(dense_count + b);
    float *bbox = maxatt + b;
    int *bucket_hist = reinterpret_cast<int *>(bbox + 4 * b);
    const size_t head = (static_cast<size_t>(b) * (28 + kNmsBuckets * 4) + 255) & ~static_cast<size_t>(255);
    int *cell_start = reinterpret_cast<int *>(base + head);
    int *cursor = cell_start + static_cast<size_t>(b) * cells;
    int *list = cursor + static_cast<size_t>(b) * cells;
    int *dense_list = list + bn;
    int *sorted = dense_list + bn;
    float4 *sorted_pts = reinterpret_cast<float4 *>((reinterpret_cast<uintptr_t>(sorted + bn) + 15) & ~static_cast<uintptr_t>(15));
    unsigned char *keep = reinterpret_cast<unsigned char *>(sorted_pts + bn);
    int *partial = reinterpret_cast<int *>((reinterpret_cast<uintptr_t>(keep + bn) + 15) & ~static_cast<uintptr_t>(15));  // chunk sums of the cell scan
    cudaError_t e = cudaMemsetAsync(base, 0, head + static_cast<size_t>(b) * cells * sizeof(int), st);  // counters + cell counts
    if (e != cudaSuccess) return fail(static_cast<int>(e), "nms: memset");
    const double grid_radius = nms_radius;
    // (the stripes' results sit in `list` until nms_compact fills it; the tickets in the zeroed dense_count)
    const int stripes = n >= 8192 ? min(64, n / 2048) : 1;
    nms_bbox_max_kernel<<<dim3(stripes, b), 1024, 0, st>>>(n, xyz, attention, reinterpret_cast<float *>(list), dense_count, bbox, maxatt);
    int rc = check_launch("nms_bbox_max_kernel");
    if (rc) return rc;
    nms_cell_count_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(n, grid_radius, max_cells, xyz, bbox, cell_start);
    rc = check_launch("nms_cell_count_kernel");
    if (rc) return rc;
    const int nchunks = (max_cells + 1 + kScanChunk - 1) / kScanChunk;
    nms_cell_partial_kernel<<<dim3(nchunks, b), 1024, 0, st>>>(grid_radius, max_cells, bbox, cell_start, partial);
    rc = check_launch("nms_cell_partial_kernel");
    if (rc) return rc;
    nms_cell_scan_kernel<<<dim3(nchunks, b), 1024, 0, st>>>(grid_radius, max_cells, bbox, partial, cell_start, cursor);
    rc = check_launch("nms_cell_scan_kernel");
    if (rc) return rc;
    nms_cell_fill_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(n, grid_radius, max_cells, xyz, attention, bbox, cursor, sorted, sorted_pts);
    rc = check_launch("nms_cell_fill_kernel");
    if (rc) return rc;
    nms_keep_kernel<<<dim3((n + 31) / 32, b), 32 * kNmsSlices, 0, st>>>(n, nms_radius, num_neighbors, max_cells, bbox, cell_start, sorted, sorted_pts, keep);
    rc = check_launch("nms_keep_kernel");
    if (rc) return rc;
    nms_compact_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(n, min_response_ratio, attention, keep, maxatt, list, dense_list, count,
                                                                 bucket_hist);
    rc = check_launch("nms_compact_kernel");
    if (rc) return rc;
    // (dense_list holds the survivors' buckets; once nms_keep has run, sorted and sorted_pts are free: kept positions, and keys /
    // positions of the cut bucket's candidates)
    nms_topk_kernel<<<b, 1024, 0, st>>>(n, max_keypoints, attention, list, dense_list, count, bucket_hist, sorted,
                                        reinterpret_cast<unsigned *>(sorted_pts) + bn, reinterpret_cast<int *>(sorted_pts), out_idx);
    rc = check_launch("nms_topk_kernel");
    if (rc) return rc;
    nms_finalize_kernel<<<dim3((max_keypoints + 1023) / 1024, b), min(1024, ((max_keypoints + 31) / 32) * 32), 0, st>>>(
        n, max_keypoints, xyz, attention, count, out_idx, out_xyz, out_attention, num_keypoints);
    return check_launch("nms_finalize_kernel");
}
