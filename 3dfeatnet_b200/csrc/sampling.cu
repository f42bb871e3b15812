// sampling.cu -- farthest point sampling + gather_point for sm_100a.
//
// Replaces tf_ops/sampling/tf_sampling_g.cu:105-181 (farthestpointsamplingKernel, gatherpointKernel).
//
// FPS is a serial chain of m-1 block-wide arg-max rounds; the only lever is the latency of one round.
// The reference keeps the running distances in GLOBAL memory (temp) and spends 10 __syncthreads per
// round on a 9-level shared-memory tree.  Here, per cloud (one 512-thread CTA, fps_group_kernel):
//   * coordinates are staged once into shared memory as SoA (12 B/point, <= 196 KB at n = 16384),
//   * the running distances live in REGISTERS for the whole kernel,
//   * points are binned spatially once (32 x 32 xy cells, Morton order) and culled in GROUPS of 128 neighbours: a group
//     whose bounding box is provably too far from the new sample to change any of its running distances is not touched
//     (exact: see fps_group_kernel),
//   * every level of the arg-max is a lexicographic maximum of (distance, key) on redux.sync, ONE __syncthreads per
//     round, and the key says where the winner's coordinates sit in shared memory: no dependent global load per round.
// The reference's tie rule is reproduced exactly: its thread t owns k = t (mod 512) and keeps its first
// strict maximum, the tree keeps the lower thread => the winner is the maximum distance with the lowest
// (k mod 512, k).  That order is folded into the key.
#include "common.cuh"
#include <type_traits>

namespace f3d {

constexpr int kFpsThreads = 1024;

__device__ __forceinline__ unsigned fps_tie_key(int k) {  // orders by (k mod 512, k); bijective for k < 2^32
    return (static_cast<unsigned>(k & 511) << 23) | (static_cast<unsigned>(k) >> 9);
}
__device__ __forceinline__ int fps_tie_key_inv(unsigned t) { return static_cast<int>(((t & 0x7fffffu) << 9) | (t >> 23)); }

struct FpsSlots {  // double-buffered per-warp winners
    int d[2][32];
    unsigned key[2][32];
    float x[2][32], y[2][32], z[2][32];
};

// Block-wide arg-max of (best, besti) under the reference tie rule.  Returns the winner index and its
// coordinates to every thread.  One __syncthreads.  `par` alternates 0/1 between consecutive calls.
template <class CoordFn>
__device__ __forceinline__ int fps_block_argmax(FpsSlots &S, int par, float best, int besti, CoordFn coords,
                                                float &ox, float &oy, float &oz) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // distances are >= +0 (or the -1 sentinel), so their bit patterns order like signed ints
    const int bi = __float_as_int(best);
    const unsigned tk = fps_tie_key(besti);
    const int wmax = __reduce_max_sync(kFull, bi);
    const unsigned wmin = __reduce_min_sync(kFull, bi == wmax ? tk : 0xffffffffu);
    const unsigned who = __ballot_sync(kFull, bi == wmax && tk == wmin);
    if (lane == __ffs(who) - 1) {  // only the warp's winner fetches its coordinates
        float bx, by, bz;
        coords(besti, bx, by, bz);
        S.d[par][warp] = wmax;
        S.key[par][warp] = wmin;
        S.x[par][warp] = bx;
        S.y[par][warp] = by;
        S.z[par][warp] = bz;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    const int d2 = lane < nw ? S.d[par][lane] : static_cast<int>(0x80000000);
    const unsigned k2 = lane < nw ? S.key[par][lane] : 0xffffffffu;
    const int bmax = __reduce_max_sync(kFull, d2);
    const unsigned bmin = __reduce_min_sync(kFull, d2 == bmax ? k2 : 0xffffffffu);
    const int src = __ffs(__ballot_sync(kFull, d2 == bmax && k2 == bmin)) - 1;
    ox = S.x[par][src];
    oy = S.y[par][src];
    oz = S.z[par][src];
    return fps_tie_key_inv(bmin);
}

// CL > 1: a thread-block CLUSTER of CL CTAs shares one cloud (n up to CL*16384): each CTA owns a contiguous chunk of
// the points and runs the same binned / culled update on it; per round the CL local winners are exchanged through
// distributed shared memory (one remote store per peer) and a cluster barrier, and every CTA reduces them identically.
struct FpsPeerSlots {  // double-buffered winners of the CTAs of one cluster
    int d[2][8];
    unsigned key[2][8];
    float x[2][8], y[2][8], z[2][8];
};
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(void *local_ptr, unsigned cta, unsigned v) {
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(local_ptr)), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(cta));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- main path: n <= 16384 per CTA, coordinates in shared memory, distances in registers, exact spatial culling ----
//
// Each round only the points near the newly selected sample can lower their running distance.  The kernel first BINS
// the cloud (32 x 32 xy grid in Morton order, counting sort in shared memory, two 16-bit cursors per word) and cuts the
// sorted order into GROUPS of 128 points (one float4 of x, y, z per lane).  Group q belongs to warp q mod W as its local
// group q / W: spatially neighbouring groups sit in DIFFERENT warps, so the few groups a sample touches spread over the
// schedulers instead of serialising in one warp.  Lane g of a warp keeps the bounding box of local group g and the
// maximum running distance inside it; one ballot tests all G boxes of the warp at once, and only groups with
//     box_dist^2 * 0.9999 < group max     (1e-4 safety margin; the fp32 arithmetic error is < 1e-6)
// are loaded and updated.  A skipped group has fmin(d, td) == td for every point, so skipping changes no value, hence
// no result.  ncu source-level sampling of the previous (warp-culled, 1024-thread) kernel showed the round to be a
// dependent chain in the ACTIVE warps (select chains to locate the maximum, vote / find-leader / shuffle sequences, a
// separate tie path): here the arg-max is a plain lexicographic maximum of (distance bits, key) at every level --
// element, group, lane, warp (2 redux.sync), block (2 redux.sync after the one barrier).  key = tie code << 14 | sorted
// position; the tie code is 0x7fff - ((k mod 512) << 6 | (k >> 9) - (k_first >> 9)) of the GLOBAL original index k, so a
// larger key is the reference's preferred point, the winner's coordinates are at xs/ys/zs[key & 0x3fff] and its index is
// decoded from key >> 14.  No tie detection, no divergent tie path, no original-index loads in the loop.
template <int LO, int N, class F>
__device__ __forceinline__ void fps_walk(unsigned act, F &f) {  // f(g) for every set bit g of act in [LO, LO + N)
    if constexpr (N == 1) {
        if ((act >> LO) & 1u) f(std::integral_constant<int, LO>{});
    } else {
        if (act & (((1u << N) - 1u) << LO)) {
            fps_walk<LO, N / 2>(act, f);
            fps_walk<LO + N / 2, N / 2>(act, f);
        }
    }
}

template <int W>
struct FpsWin {  // double-buffered per-warp winners: (distance bits, tie code << 14 | sorted position)
    int2 dp[2][W];
};

template <int W, int G, int CL>
__global__ void __launch_bounds__(W * 32, 1)
fps_group_kernel(int num_clouds, int n_total, int m, const float *__restrict__ inp, int *__restrict__ out, float *__restrict__ out_xyz) {
    constexpr int T = W * 32;
    constexpr int NP = T * 4 * G;
    constexpr int kGrid = 32, kCells = kGrid * kGrid;
    static_assert(G <= 16 && (G & (G - 1)) == 0 && NP <= 16384, "14-bit sorted positions / one box per lane / binary walk");
    extern __shared__ float4 fps_smem[];
    float *xs = reinterpret_cast<float *>(fps_smem);
    float *ys = xs + NP;
    float *zs = ys + NP;
    // per sorted position: 0x7fff - (15-bit code of the original index that orders like the reference tie rule); a LARGER
    // value is the preferred point, and the original index is recovered from it (see the output store)
    unsigned short *tkc = reinterpret_cast<unsigned short *>(zs + NP);
    __shared__ union {
        unsigned cell2[kCells / 2];  // setup: two 16-bit cursors per word (n <= 16384 per CTA: no carry between halves)
        FpsWin<W> win;               // main loop
    } u;
    __shared__ float red[4][32];
    __shared__ float bbox[4];
    static_assert(sizeof(FpsPeerSlots) <= sizeof(float) * 4 * 32, "peer slots must fit the reduction scratch");
    FpsPeerSlots &peers = *reinterpret_cast<FpsPeerSlots *>(&red[0][0]);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = CL > 1 ? cluster_ctarank() : 0u;
    // A CTA (or cluster) walks the clouds blockIdx.x / CL, + gridDim.x / CL, ...: with fewer CTAs than clouds the sampling of a
    // batch occupies a chosen number of SMs for longer (the pipeline runs it beside the contractions of the previous batch).
    for (int cloud = blockIdx.x / CL; cloud < num_clouds; cloud += gridDim.x / CL) {
    const int chunk = CL > 1 ? (n_total + CL - 1) / CL : n_total;
    const int k_first = static_cast<int>(rank) * chunk;
    const int n = max(min(chunk, n_total - k_first), 0);
    const float *p0 = inp + static_cast<size_t>(cloud) * n_total * 3;
    const float *p = p0 + static_cast<size_t>(k_first) * 3;
    int *o = out + static_cast<size_t>(cloud) * m;
    float *oc = out_xyz ? out_xyz + static_cast<size_t>(cloud) * m * 3 : nullptr;  // optional fused gather_point of the samples
    const int kbase = k_first >> 9;

    // ---- 1. xy bounding box of the chunk
    float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
    for (int k = tid; k < n; k += T) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1);
        mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, s));
        mny = fminf(mny, __shfl_xor_sync(kFull, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, s));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mxx; red[2][warp] = mny; red[3][warp] = mxy; }
    for (int k = tid; k < kCells / 2; k += T) u.cell2[k] = 0u;
    __syncthreads();
    if (warp == 0) {
        mnx = lane < W ? red[0][lane] : 3.0e38f; mxx = lane < W ? red[1][lane] : -3.0e38f;
        mny = lane < W ? red[2][lane] : 3.0e38f; mxy = lane < W ? red[3][lane] : -3.0e38f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, s));
            mny = fminf(mny, __shfl_xor_sync(kFull, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, s));
        }
        if (lane == 0) { bbox[0] = mnx; bbox[1] = mxx; bbox[2] = mny; bbox[3] = mxy; }
    }
    __syncthreads();
    const float x0 = bbox[0], y0 = bbox[2];
    const float ex = bbox[1] - bbox[0], ey = bbox[3] - bbox[2];
    const float sx = (ex > 0.0f && ex < 3.0e38f) ? static_cast<float>(kGrid) / ex : 0.0f;  // degenerate extents: one cell
    const float sy = (ey > 0.0f && ey < 3.0e38f) ? static_cast<float>(kGrid) / ey : 0.0f;
    auto cell_of = [&](float x, float y) -> int {
        int ix = static_cast<int>((x - x0) * sx), iy = static_cast<int>((y - y0) * sy);
        ix = min(max(ix, 0), kGrid - 1);
        iy = min(max(iy, 0), kGrid - 1);
        // Morton interleave of two 5-bit numbers
        ix = (ix | (ix << 4)) & 0x10f; ix = (ix | (ix << 2)) & 0x133; ix = (ix | (ix << 1)) & 0x155;
        iy = (iy | (iy << 4)) & 0x10f; iy = (iy | (iy << 2)) & 0x133; iy = (iy | (iy << 1)) & 0x155;
        return ix | (iy << 1);
    };
    // ---- 2. counting sort by cell: histogram, exclusive scan, scatter
    for (int k = tid; k < n; k += T) {
        const int c = cell_of(__ldg(p + 3 * k), __ldg(p + 3 * k + 1));
        atomicAdd(&u.cell2[c >> 1], 1u << ((c & 1) * 16));
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int WPL = kCells / 2 / 32;  // words per lane
        unsigned wv[WPL];
        int tot = 0;
#pragma unroll
        for (int i = 0; i < WPL; ++i) { wv[i] = u.cell2[lane * WPL + i]; tot += static_cast<int>((wv[i] & 0xffffu) + (wv[i] >> 16)); }
        int inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, inc, s); if (lane >= s) inc += v; }
        unsigned run = static_cast<unsigned>(inc - tot);
#pragma unroll
        for (int i = 0; i < WPL; ++i) {
            const unsigned lo = wv[i] & 0xffffu, hi = wv[i] >> 16;
            u.cell2[lane * WPL + i] = run | ((run + lo) << 16);
            run += lo + hi;
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += T) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        const int c = cell_of(x, y);
        const int sh = (c & 1) * 16;
        const int pos = static_cast<int>((atomicAdd(&u.cell2[c >> 1], 1u << sh) >> sh) & 0xffffu);
        xs[pos] = x; ys[pos] = y; zs[pos] = z;
        const int kg = k_first + k;  // GLOBAL original index: the reference tie rule orders by (kg mod 512, kg)
        tkc[pos] = static_cast<unsigned short>(0x7fff - (((kg & 511) << 6) | ((kg >> 9) - kbase)));
    }
    for (int k = n + tid; k < NP; k += T) { xs[k] = ys[k] = zs[k] = 0.0f; tkc[k] = 0; }
    __syncthreads();  // u.cell2 is dead from here on: u.win takes its place

    // ---- 3. per-thread state: running distances + per-group maxima (registers); lane g keeps the box of local group g
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    const float4 *ys4 = reinterpret_cast<const float4 *>(ys);
    const float4 *zs4 = reinterpret_cast<const float4 *>(zs);
    auto slot4 = [&](int g) -> int { return (g * W + warp) * 32 + lane; };  // float4 index of this lane in local group g
    float td[4 * G], gm[G];
    unsigned gkey[G];  // per local group: (tie code << 14 | sorted position) of the lane's preferred point at the group maximum
    float lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f, loz = 3.0e38f, hiz = -3.0e38f;
    const uint2 *tk2 = reinterpret_cast<const uint2 *>(tkc);
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 X = xs4[slot4(g)], Y = ys4[slot4(g)], Z = zs4[slot4(g)];
        const float xv[4] = {X.x, X.y, X.z, X.w}, yv[4] = {Y.x, Y.y, Y.z, Y.w}, zv[4] = {Z.x, Z.y, Z.z, Z.w};
        float ax = 3.0e38f, bx = -3.0e38f, ay = 3.0e38f, by = -3.0e38f, az = 3.0e38f, bz = -3.0e38f;
        float gmax = -1.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool real = 4 * slot4(g) + q < n;
            td[g * 4 + q] = real ? 1e38f : -1.0f;
            gmax = fmaxf(gmax, td[g * 4 + q]);
            if (real) {
                ax = fminf(ax, xv[q]); bx = fmaxf(bx, xv[q]);
                ay = fminf(ay, yv[q]); by = fmaxf(by, yv[q]);
                az = fminf(az, zv[q]); bz = fmaxf(bz, zv[q]);
            }
        }
        gm[g] = gmax;
        gkey[g] = static_cast<unsigned>(4 * slot4(g));  // round 1 updates every group that holds a real point
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            ax = fminf(ax, __shfl_xor_sync(kFull, ax, s)); bx = fmaxf(bx, __shfl_xor_sync(kFull, bx, s));
            ay = fminf(ay, __shfl_xor_sync(kFull, ay, s)); by = fmaxf(by, __shfl_xor_sync(kFull, by, s));
            az = fminf(az, __shfl_xor_sync(kFull, az, s)); bz = fmaxf(bz, __shfl_xor_sync(kFull, bz, s));
        }
        if (lane == g) { lox = ax; hix = bx; loy = ay; hiy = by; loz = az; hiz = bz; }
    }
    // lanes >= G (and empty groups) keep lo = 3e38, hi = -3e38: their box distance is +inf, never active
    float gthr = 3.0e38f;  // lane g: maximum running distance inside local group g (the culling threshold of that group)

    float ox = __ldg(p0 + 0), oy = __ldg(p0 + 1), oz = __ldg(p0 + 2);  // the first sample is point 0 (:114-116)
    if (tid == 0 && rank == 0) {
        o[0] = 0;
        if (oc) { oc[0] = ox; oc[1] = oy; oc[2] = oz; }
    }
    if (CL > 1) cluster_sync_all();  // every CTA of the cluster is resident before the first remote store
    int cw_d = __float_as_int(-1.0f);  // cached arg-max of this warp (uniform across its lanes): distance bits, key
    unsigned cw_key = 0u;
    FpsWin<W> &S2 = u.win;

    // Every level of the arg-max orders candidates by (distance, key): the key's high bits are the inverted tie code, so
    // "maximum distance, then the reference's lowest (k mod 512, k)" is a plain lexicographic maximum -- element, group,
    // lane, warp and block levels need no tie detection, and the key's low 14 bits say where the winner's coordinates are.
    for (int j = 1; j < m; ++j) {
        const float bx = fmaxf(fmaxf(lox - ox, ox - hix), 0.0f);
        const float by = fmaxf(fmaxf(loy - oy, oy - hiy), 0.0f);
        const float bz = fmaxf(fmaxf(loz - oz, oz - hiz), 0.0f);
        const float lbd = (bx * bx + by * by + bz * bz) * 0.9999f;
        const unsigned act = __ballot_sync(kFull, (j == 1 && lane < G) || !(lbd >= gthr));  // bit g: local group g may change
        if (act != 0u) {
            auto update = [&](auto gc) {
                constexpr int g = decltype(gc)::value;
                const float4 X = xs4[slot4(g)], Y = ys4[slot4(g)], Z = zs4[slot4(g)];
                const uint2 kk = tk2[slot4(g)];
                const float t0 = fminf(sqdist_ref(X.x - ox, Y.x - oy, Z.x - oz), td[g * 4 + 0]);
                const float t1 = fminf(sqdist_ref(X.y - ox, Y.y - oy, Z.y - oz), td[g * 4 + 1]);
                const float t2 = fminf(sqdist_ref(X.z - ox, Y.z - oy, Z.z - oz), td[g * 4 + 2]);
                const float t3 = fminf(sqdist_ref(X.w - ox, Y.w - oy, Z.w - oz), td[g * 4 + 3]);
                td[g * 4 + 0] = t0; td[g * 4 + 1] = t1; td[g * 4 + 2] = t2; td[g * 4 + 3] = t3;
                const float mx = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
                gm[g] = mx;
                const unsigned base = static_cast<unsigned>(4 * slot4(g));
                const unsigned k0 = ((kk.x & 0xffffu) << 14) | base, k1 = ((kk.x >> 16) << 14) | (base + 1u);
                const unsigned k2 = ((kk.y & 0xffffu) << 14) | (base + 2u), k3 = ((kk.y >> 16) << 14) | (base + 3u);
                unsigned gk = t0 == mx ? k0 : 0u;
                gk = t1 == mx ? max(gk, k1) : gk;
                gk = t2 == mx ? max(gk, k2) : gk;
                gk = t3 == mx ? max(gk, k3) : gk;
                gkey[g] = gk;
                const int gw = __reduce_max_sync(kFull, __float_as_int(mx));
                if (lane == g) gthr = __int_as_float(gw);
            };
            // binary tree over the bits of `act`: a test-and-skip costs ~28 cycles on the critical path of the round, and
            // usually one or two of the G groups are active
            if constexpr (G == 1) {
                update(std::integral_constant<int, 0>{});
            } else {
                fps_walk<0, G / 2>(act, update);
                fps_walk<G / 2, G / 2>(act, update);
            }
            float vmax = gm[0];
#pragma unroll
            for (int g = 1; g < G; ++g) vmax = fmaxf(vmax, gm[g]);
            const int bi = __float_as_int(vmax);
            const int wmax = __reduce_max_sync(kFull, bi);
            unsigned ck[G];  // independent selects + a max tree (not a G-long dependent chain)
#pragma unroll
            for (int g = 0; g < G; ++g) ck[g] = gm[g] == vmax ? gkey[g] : 0u;
#pragma unroll
            for (int w = 1; w < G; w <<= 1)
#pragma unroll
                for (int g = 0; g + w < G; g += 2 * w) ck[g] = max(ck[g], ck[g + w]);
            const unsigned lk = ck[0];
            cw_key = __reduce_max_sync(kFull, bi == wmax ? lk : 0u);
            cw_d = wmax;
        }
        const int par = j & 1;
        if (lane == 0) S2.dp[par][warp] = make_int2(cw_d, static_cast<int>(cw_key));
        __syncthreads();
        const int2 dp = lane < W ? S2.dp[par][lane] : make_int2(static_cast<int>(0x80000000), 0);
        const int bmax = __reduce_max_sync(kFull, dp.x);
        const unsigned bkey = __reduce_max_sync(kFull, dp.x == bmax ? static_cast<unsigned>(dp.y) : 0u);
        const int wpos = static_cast<int>(bkey & 0x3fffu);
        const float wx = xs[wpos], wy = ys[wpos], wz = zs[wpos];
        const int code = 0x7fff - static_cast<int>(bkey >> 14);
        const int widx = ((((code & 63) + kbase) << 9) | (code >> 6));  // GLOBAL original index of the winner
        if constexpr (CL == 1) {
            ox = wx;
            oy = wy;
            oz = wz;
            if (tid == 0) {
                o[j] = widx;
                if (oc) { oc[3 * j] = wx; oc[3 * j + 1] = wy; oc[3 * j + 2] = wz; }
            }
        } else {
            // publish this CTA's winner (tie key of the GLOBAL original index) to every CTA of the cluster
            if (warp == 0 && lane < CL) {
                const unsigned wk = bmax < 0 ? 0xffffffffu : fps_tie_key(widx);
                st_cluster_u32(&peers.d[par][rank], lane, static_cast<unsigned>(bmax));
                st_cluster_u32(&peers.key[par][rank], lane, wk);
                st_cluster_u32(&peers.x[par][rank], lane, __float_as_uint(wx));
                st_cluster_u32(&peers.y[par][rank], lane, __float_as_uint(wy));
                st_cluster_u32(&peers.z[par][rank], lane, __float_as_uint(wz));
            }
            cluster_sync_all();
            const int pd = lane < CL ? peers.d[par][lane] : static_cast<int>(0x80000000);
            const unsigned pk = lane < CL ? peers.key[par][lane] : 0xffffffffu;
            const int gmax = __reduce_max_sync(kFull, pd);
            const unsigned gmin = __reduce_min_sync(kFull, pd == gmax ? pk : 0xffffffffu);
            const int src3 = __ffs(__ballot_sync(kFull, pd == gmax && pk == gmin)) - 1;
            ox = peers.x[par][src3];
            oy = peers.y[par][src3];
            oz = peers.z[par][src3];
            if (tid == 0 && rank == 0) {
                o[j] = fps_tie_key_inv(gmin);
                if (oc) { oc[3 * j] = ox; oc[3 * j + 1] = oy; oc[3 * j + 2] = oz; }
            }
        }
    }
    if (CL > 1) cluster_sync_all();  // no CTA exits (or re-initialises) while a peer may still write into its shared memory
    else __syncthreads();            // the shared-memory cloud and the winner slots are rebuilt for the next cloud
    }
}

// ---- fallback for any n: running distances in a caller-provided (b,n) scratch, points from L2 ------------
// Thread t owns k = t, t+1024, ...: k mod 512 is constant per thread and k ascends, so the first strict
// maximum is again the thread's lowest (k mod 512, k).
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_global_kernel(int n, int m, const float *__restrict__ inp, float *__restrict__ temp, int *__restrict__ out) {
    __shared__ FpsSlots slots;
    const int tid = threadIdx.x;
    const float *p = inp + static_cast<size_t>(blockIdx.x) * n * 3;
    float *td = temp + static_cast<size_t>(blockIdx.x) * n;
    int *o = out + static_cast<size_t>(blockIdx.x) * m;
    for (int k = tid; k < n; k += kFpsThreads) td[k] = 1e38f;
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (n > 0) {
        ox = __ldg(p + 0);
        oy = __ldg(p + 1);
        oz = __ldg(p + 2);
    }
    if (tid == 0) o[0] = 0;
    for (int j = 1; j < m; ++j) {
        float best = -1.0f, bx = 0.f, by = 0.f, bz = 0.f;
        int besti = 0;
        for (int k = tid; k < n; k += kFpsThreads) {
            const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
            const float d2 = fminf(sqdist_ref(x - ox, y - oy, z - oz), td[k]);
            td[k] = d2;
            if (d2 > best) {
                best = d2;
                besti = k;
                bx = x;
                by = y;
                bz = z;
            }
        }
        const int old = fps_block_argmax(
            slots, j & 1, best, besti, [&](int, float &x, float &y, float &z) { x = bx; y = by; z = bz; }, ox, oy, oz);
        if (tid == 0) o[j] = old;
    }
}

// gather_point: out[b,j,:] = inp[b,idx[b,j],:]  (tf_sampling_g.cu:172-181).  One thread per output float.
__global__ void gather_point_kernel(int n, int m, long long total, const float *__restrict__ inp,
                                    const int *__restrict__ idx, float *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long r = i / 3;
    const int c = static_cast<int>(i - r * 3);
    const long long bb = r / m;
    const int a = __ldg(idx + r);
    out[i] = __ldg(inp + (bb * n + a) * 3 + c);
}

template <int W, int G, int CL>
static int launch_fps_group(int b, int n, int m, const float *inp, int *out, float *out_xyz, int max_ctas, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(W) * 32 * 4 * G * (3 * sizeof(float) + sizeof(unsigned short));
    cudaError_t e = cudaFuncSetAttribute(fps_group_kernel<W, G, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "fps: cudaFuncSetAttribute");
    // algorithmic bytes: B * (12 N + 4 M) (+ 12 M when the sampled coordinates are written too)
    ktimer_begin("fps_group_kernel", static_cast<double>(b) * (12.0 * n + 4.0 * m + (out_xyz ? 12.0 * m : 0.0)), st);
    if (CL == 1) {
        const int grid = (max_ctas > 0 && max_ctas < b) ? max_ctas : b;  // fewer CTAs than clouds: each walks several clouds
        fps_group_kernel<W, G, CL><<<grid, W * 32, smem, st>>>(b, n, m, inp, out, out_xyz);
        ktimer_end(st);
        return check_launch("fps_group_kernel");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(b) * CL);
    cfg.blockDim = dim3(W * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, fps_group_kernel<W, G, CL>, b, n, m, inp, out, out_xyz);
    ktimer_end(st);
    ++g_launches;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(static_cast<int>(e), "fps_group_kernel (cluster launch)");
    }
    return 0;
}

}  // namespace f3d

using namespace f3d;

static int fps_dispatch(int b, int n, int m, const float *inp, float *temp, int *out, float *out_xyz, int max_ctas, void *stream) {
    if (b < 0 || n <= 0 || m <= 0 || !inp || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample: bad arguments");
    if (b == 0) return 0;
    cudaStream_t st = as_stream(stream);
    // 512 threads (measured: 256 threads x 16 groups 0.308 ms, 1024 threads x 4 groups 0.325 ms, 512 x 8 0.266 ms at n = 16384)
    if (n <= 2048) return launch_fps_group<16, 1, 1>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (n <= 4096) return launch_fps_group<16, 2, 1>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (n <= 8192) return launch_fps_group<16, 4, 1>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (n <= 16384) return launch_fps_group<16, 8, 1>(b, n, m, inp, out, out_xyz, max_ctas, st);
    // larger clouds: a cluster of 2 / 4 / 8 CTAs per cloud, 16384 points each (KITTI-shape scans, 131072 points)
    if (n <= 2 * 16384) return launch_fps_group<16, 8, 2>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (n <= 4 * 16384) return launch_fps_group<16, 8, 4>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (n <= 8 * 16384) return launch_fps_group<16, 8, 8>(b, n, m, inp, out, out_xyz, max_ctas, st);
    if (!temp) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "farthest_point_sample: n > 131072 needs temp of b*n floats");
    fps_global_kernel<<<b, kFpsThreads, 0, st>>>(n, m, inp, temp, out);
    int rc = check_launch("fps_global_kernel");
    if (rc || !out_xyz) return rc;
    const long long total = 3LL * b * m;
    gather_point_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(n, m, total, inp, out, out_xyz);
    return check_launch("gather_point_kernel");
}

F3D_API int f3d_farthest_point_sample(int b, int n, int m, const float *inp, float *temp, int *out, void *stream) {
    return fps_dispatch(b, n, m, inp, temp, out, nullptr, 0, stream);
}

// farthest_point_sample + gather_point of the samples in one launch (sample_points, models/pointnet_common.py:14-29): the
// kernel already holds the winner's coordinates each round, so new_xyz costs three more stores and no second kernel.
F3D_API int f3d_farthest_point_sample_gather(int b, int n, int m, const float *inp, float *temp, int *out, float *new_xyz, void *stream) {
    if (!new_xyz) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample_gather: new_xyz is NULL");
    return fps_dispatch(b, n, m, inp, temp, out, new_xyz, 0, stream);
}

// Same, on at most `max_ctas` CTAs (n <= 16384: one CTA per cloud, each CTA walks ceil(b / max_ctas) clouds one after the other;
// larger clouds keep their cluster of CTAs per cloud and ignore the limit).  The sampling of a batch is a chain of m dependent
// rounds per cloud -- latency, not throughput -- so a caller that has other work for the remaining SMs (the contractions of the
// previous batch, 3dfeatnet_b200/pipeline.py) trades a longer sampling phase for SMs.  Results do not depend on max_ctas.
F3D_API int f3d_farthest_point_sample_gather_ctas(int b, int n, int m, const float *inp, float *temp, int *out, float *new_xyz, int max_ctas,
                                                  void *stream) {
    if (!new_xyz) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample_gather: new_xyz is NULL");
    if (max_ctas < 0) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample_gather: max_ctas < 0");
    return fps_dispatch(b, n, m, inp, temp, out, new_xyz, max_ctas, stream);
}

F3D_API int f3d_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || !inp || !idx || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "gather_point: bad arguments");
    const long long total = 3LL * b * m;
    if (total == 0) return 0;
    gather_point_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(n, m, total, inp, idx, out);
    return check_launch("gather_point_kernel");
}
