// sampling.cu -- farthest point sampling + gather_point for sm_100a.
//
// Replaces tf_ops/sampling/tf_sampling_g.cu:105-181 (farthestpointsamplingKernel, gatherpointKernel).
//
// FPS is a serial chain of m-1 block-wide arg-max rounds; the only lever is the latency of one round.
// The reference keeps the running distances in GLOBAL memory (temp) and spends 10 __syncthreads per
// round on a 9-level shared-memory tree.  Here, per cloud (one 1024-thread CTA):
//   * coordinates are staged once into shared memory as SoA (12 B/point, <= 196 KB at n = 16384),
//   * the running distances live in REGISTERS (PPT per thread) for the whole kernel,
//   * the arg-max is two redux.sync per warp + one 32-entry shared-memory hop: ONE __syncthreads per round,
//   * the winner's coordinates travel with the arg-max, so no dependent global load per round,
//   * points are binned spatially once, and a warp whose bounding box is provably too far from the new sample
//     to change any of its running distances skips the round (exact: see fps_cull_kernel).
// The reference's tie rule is reproduced exactly: its thread t owns k = t (mod 512) and keeps its first
// strict maximum, the tree keeps the lower thread => the winner is the maximum distance with the lowest
// (k mod 512, k).  That pair is packed into one 32-bit tie key so it rides through redux.sync.min.
#include "common.cuh"

namespace f3d {

constexpr int kFpsThreads = 1024;

__device__ __forceinline__ unsigned fps_tie_key(int k) {  // orders by (k mod 512, k); bijective for k < 2^32
    return (static_cast<unsigned>(k & 511) << 23) | (static_cast<unsigned>(k) >> 9);
}
__device__ __forceinline__ int fps_tie_key_inv(unsigned t) { return static_cast<int>(((t & 0x7fffffu) << 9) | (t >> 23)); }

struct FpsSlots2 {  // double-buffered per-warp winners of fps_cull_kernel: distance bits, sorted position, coordinates
    int d[2][32];
    int pos[2][32];
    float x[2][32], y[2][32], z[2][32];
};

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

// ---- main path: n <= 1024*PPT, coordinates in shared memory, distances in registers, exact spatial culling ------
//
// Each round only the points near the newly selected sample can lower their running distance.  The kernel therefore
// first BINS the cloud (16 x 16 xy grid in Morton order, counting sort in shared memory) so that a warp owns 32*PPT
// spatially neighbouring points, and keeps per warp an axis-aligned box and the maximum running distance `wmx`.
// A round skips a warp when the squared distance from the new sample to the warp's box (with a 1e-4 safety margin
// against fp32 rounding, the arithmetic error being < 1e-6) is >= wmx: then fmin(d, td) == td for every point of the
// warp, so its cached arg-max is still exact.  Skipping changes no value, hence no result.  Binning permutes the
// points, so the reference tie rule is applied on the ORIGINAL index (kept as u16 per point): thread, warp and block
// levels all select the maximum distance with the lowest tie key of that index -- a total order, independent of the
// (non-deterministic) order in which the counting sort places points inside a cell.
// Warp w owns sorted positions [w*32*PPT, (w+1)*32*PPT); lane l reads float4 #(l + 32 g) of that range (conflict-free).
__device__ long long *g_fps_dbg = nullptr;  // bring-up: [round][warp][4] clock64() stamps of CTA 0 (rounds < 256)

// CL > 1: a thread-block CLUSTER of CL CTAs shares one cloud (n up to CL*1024*PPT): each CTA owns a contiguous chunk of
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

template <int PPT, int CL>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_cull_kernel(int n_total, int m, const float *__restrict__ inp, int *__restrict__ out) {
    constexpr int G = PPT / 4;
    constexpr int NP = kFpsThreads * PPT;
    constexpr int kCells = 256;
    extern __shared__ float4 fps_smem[];
    float *xs = reinterpret_cast<float *>(fps_smem);
    float *ys = xs + NP;
    float *zs = ys + NP;
    unsigned short *oi = reinterpret_cast<unsigned short *>(zs + NP);  // original index of each sorted position
    __shared__ FpsSlots2 slots2;
    __shared__ int cell_cursor[kCells];
    __shared__ float red[4][32];
    __shared__ float bbox[4];
    // the peer slots reuse the bounding-box scratch (dead after the initial cluster barrier): 227 KB is tight at PPT = 16
    static_assert(sizeof(FpsPeerSlots) <= sizeof(float) * 4 * 32, "peer slots must fit the reduction scratch");
    FpsPeerSlots &peers = *reinterpret_cast<FpsPeerSlots *>(&red[0][0]);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = CL > 1 ? cluster_ctarank() : 0u;
    const int cloud = blockIdx.x / CL;
    const int chunk = CL > 1 ? (n_total + CL - 1) / CL : n_total;  // points per CTA (the last chunk may be shorter)
    const int k_first = static_cast<int>(rank) * chunk;
    const int n = max(min(chunk, n_total - k_first), 0);            // points owned by this CTA
    const float *p0 = inp + static_cast<size_t>(cloud) * n_total * 3;  // the whole cloud (point 0 = first sample)
    const float *p = p0 + static_cast<size_t>(k_first) * 3;            // this CTA's chunk
    int *o = out + static_cast<size_t>(cloud) * m;

    // ---- 1. xy bounding box of the chunk
    float mnx = 3.0e38f, mxx = -3.0e38f, mny = 3.0e38f, mxy = -3.0e38f;
    for (int k = tid; k < n; k += kFpsThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1);
        mnx = fminf(mnx, x); mxx = fmaxf(mxx, x); mny = fminf(mny, y); mxy = fmaxf(mxy, y);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        mnx = fminf(mnx, __shfl_xor_sync(kFull, mnx, s)); mxx = fmaxf(mxx, __shfl_xor_sync(kFull, mxx, s));
        mny = fminf(mny, __shfl_xor_sync(kFull, mny, s)); mxy = fmaxf(mxy, __shfl_xor_sync(kFull, mxy, s));
    }
    if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mxx; red[2][warp] = mny; red[3][warp] = mxy; }
    if (tid < kCells) cell_cursor[tid] = 0;
    __syncthreads();
    if (warp == 0) {
        mnx = red[0][lane]; mxx = red[1][lane]; mny = red[2][lane]; mxy = red[3][lane];
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
    const float sx = (ex > 0.0f && ex < 3.0e38f) ? 16.0f / ex : 0.0f;  // degenerate / non-finite extents: one cell
    const float sy = (ey > 0.0f && ey < 3.0e38f) ? 16.0f / ey : 0.0f;
    auto cell_of = [&](float x, float y) -> int {
        int ix = static_cast<int>((x - x0) * sx), iy = static_cast<int>((y - y0) * sy);
        ix = min(max(ix, 0), 15);
        iy = min(max(iy, 0), 15);
        // Morton interleave of two 4-bit numbers
        ix = (ix | (ix << 2)) & 0x33; ix = (ix | (ix << 1)) & 0x55;
        iy = (iy | (iy << 2)) & 0x33; iy = (iy | (iy << 1)) & 0x55;
        return ix | (iy << 1);
    };
    // ---- 2. counting sort by cell: histogram, exclusive scan, scatter
    for (int k = tid; k < n; k += kFpsThreads) atomicAdd(&cell_cursor[cell_of(__ldg(p + 3 * k), __ldg(p + 3 * k + 1))], 1);
    __syncthreads();
    if (warp == 0) {
        int c[8], tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i] = cell_cursor[lane * 8 + i]; tot += c[i]; }
        int inc = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, inc, s); if (lane >= s) inc += v; }
        int run = inc - tot;
#pragma unroll
        for (int i = 0; i < 8; ++i) { cell_cursor[lane * 8 + i] = run; run += c[i]; }
    }
    __syncthreads();
    for (int k = tid; k < n; k += kFpsThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        const int pos = atomicAdd(&cell_cursor[cell_of(x, y)], 1);
        xs[pos] = x; ys[pos] = y; zs[pos] = z;
        oi[pos] = static_cast<unsigned short>(k);
    }
    for (int k = n + tid; k < NP; k += kFpsThreads) { xs[k] = ys[k] = zs[k] = 0.0f; oi[k] = 0; }
    __syncthreads();

    // ---- 3. per-thread state: running distances (registers), per-warp box
    const int wbase = warp * 32 * PPT;  // first sorted position of this warp
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs + wbase);
    const float4 *ys4 = reinterpret_cast<const float4 *>(ys + wbase);
    const float4 *zs4 = reinterpret_cast<const float4 *>(zs + wbase);
    float td[PPT];
    float lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f, loz = 3.0e38f, hiz = -3.0e38f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 X = xs4[lane + 32 * g], Y = ys4[lane + 32 * g], Z = zs4[lane + 32 * g];
        const float xv[4] = {X.x, X.y, X.z, X.w}, yv[4] = {Y.x, Y.y, Y.z, Y.w}, zv[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool real = wbase + 4 * (lane + 32 * g) + q < n;
            td[g * 4 + q] = real ? 1e38f : -1.0f;
            if (real) {
                lox = fminf(lox, xv[q]); hix = fmaxf(hix, xv[q]);
                loy = fminf(loy, yv[q]); hiy = fmaxf(hiy, yv[q]);
                loz = fminf(loz, zv[q]); hiz = fmaxf(hiz, zv[q]);
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(kFull, lox, s)); hix = fmaxf(hix, __shfl_xor_sync(kFull, hix, s));
        loy = fminf(loy, __shfl_xor_sync(kFull, loy, s)); hiy = fmaxf(hiy, __shfl_xor_sync(kFull, hiy, s));
        loz = fminf(loz, __shfl_xor_sync(kFull, loz, s)); hiz = fmaxf(hiz, __shfl_xor_sync(kFull, hiz, s));
    }

    float ox = __ldg(p0 + 0), oy = __ldg(p0 + 1), oz = __ldg(p0 + 2);  // the first sample is point 0 (:114-116)
    if (tid == 0 && rank == 0) o[0] = 0;
    if (CL > 1) cluster_sync_all();  // every CTA of the cluster is resident before the first remote store
    // cached arg-max of this warp (uniform across its lanes): distance bits, sorted position, coordinates
    int cw_d = __float_as_int(-1.0f);
    int cw_pos = 0;
    float cw_x = 0.f, cw_y = 0.f, cw_z = 0.f, wmx = 3.0e38f;
    FpsSlots2 &S2 = slots2;

#ifdef F3D_FPS_TIMELINE  // bring-up instrumentation (tools/fps_timeline.py); costs ~25 % of the kernel, off by default
    long long *dbg = (blockIdx.x == 0 && lane == 0) ? g_fps_dbg : nullptr;
#define F3D_FPS_STAMP(slot) if (dbg && j < 256) dbg[(j * 32 + warp) * 4 + (slot)] = clock64()
#else
#define F3D_FPS_STAMP(slot)
#endif
    for (int j = 1; j < m; ++j) {
        F3D_FPS_STAMP(0);
        const float bx = fmaxf(fmaxf(lox - ox, ox - hix), 0.0f);
        const float by = fmaxf(fmaxf(loy - oy, oy - hiy), 0.0f);
        const float bz = fmaxf(fmaxf(loz - oz, oz - hiz), 0.0f);
        const float lbd = (bx * bx + by * by + bz * bz) * 0.9999f;
        if (j == 1 || !(lbd >= wmx)) {  // warp-uniform
            float vmax = -1.0f;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 X = xs4[lane + 32 * g], Y = ys4[lane + 32 * g], Z = zs4[lane + 32 * g];
                td[g * 4 + 0] = fminf(sqdist_ref(X.x - ox, Y.x - oy, Z.x - oz), td[g * 4 + 0]);
                td[g * 4 + 1] = fminf(sqdist_ref(X.y - ox, Y.y - oy, Z.y - oz), td[g * 4 + 1]);
                td[g * 4 + 2] = fminf(sqdist_ref(X.z - ox, Y.z - oy, Z.z - oz), td[g * 4 + 2]);
                td[g * 4 + 3] = fminf(sqdist_ref(X.w - ox, Y.w - oy, Z.w - oz), td[g * 4 + 3]);
                vmax = fmaxf(fmaxf(fmaxf(td[g * 4 + 0], td[g * 4 + 1]), fmaxf(td[g * 4 + 2], td[g * 4 + 3])), vmax);
            }
            const int bi = __float_as_int(vmax);
            const int wmax = __reduce_max_sync(kFull, bi);
            // position of the lane's maximum and how many of its points share it
            int bpos = 0, neq = 0;
#pragma unroll
            for (int g = G - 1; g >= 0; --g)
#pragma unroll
                for (int q = 3; q >= 0; --q)
                    if (td[g * 4 + q] == vmax) {
                        bpos = wbase + 4 * (lane + 32 * g) + q;
                        ++neq;
                    }
            const unsigned cand = __ballot_sync(kFull, bi == wmax);
            int src = __ffs(cand) - 1;
            // Ties (several points at exactly the warp maximum -- duplicated points) are the only case that needs the
            // reference's tie rule, i.e. the original indices; the common case skips those loads and the second redux.
            if (__any_sync(kFull, bi == wmax && (neq > 1 || (cand & (cand - 1)) != 0))) {
                unsigned tk = 0xffffffffu;
                if (bi == wmax) {
#pragma unroll
                    for (int g = 0; g < G; ++g)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (td[g * 4 + q] == vmax) {
                                const int pos = wbase + 4 * (lane + 32 * g) + q;
                                const unsigned k2 = fps_tie_key(vmax < 0.0f ? 0 : k_first + oi[pos])  /* tie rule on the GLOBAL index */;
                                if (k2 < tk) { tk = k2; bpos = pos; }
                            }
                }
                const unsigned wmin = __reduce_min_sync(kFull, tk);
                src = __ffs(__ballot_sync(kFull, tk == wmin)) - 1;
            }
            float vx = 0.f, vy = 0.f, vz = 0.f;
            if (lane == src) {  // only the winner touches shared memory
                vx = xs[bpos];
                vy = ys[bpos];
                vz = zs[bpos];
            }
            cw_d = wmax;
            cw_pos = __shfl_sync(kFull, bpos, src);
            cw_x = __shfl_sync(kFull, vx, src);
            cw_y = __shfl_sync(kFull, vy, src);
            cw_z = __shfl_sync(kFull, vz, src);
            wmx = __int_as_float(wmax);
        }
        const int par = j & 1;
        F3D_FPS_STAMP(1);
        if (lane == 0) {
            S2.d[par][warp] = cw_d;
            S2.pos[par][warp] = cw_pos;
            S2.x[par][warp] = cw_x;
            S2.y[par][warp] = cw_y;
            S2.z[par][warp] = cw_z;
        }
        __syncthreads();
        F3D_FPS_STAMP(2);
        const int d2 = S2.d[par][lane];
        const int bmax = __reduce_max_sync(kFull, d2);
        const unsigned cand2 = __ballot_sync(kFull, d2 == bmax);
        int src2 = __ffs(cand2) - 1;
        if ((cand2 & (cand2 - 1)) != 0) {  // several warps at the block maximum: the reference tie rule decides
            const unsigned k2 = d2 == bmax ? fps_tie_key(bmax < 0 ? 0 : k_first + oi[S2.pos[par][lane]]) : 0xffffffffu;
            const unsigned bmin = __reduce_min_sync(kFull, k2);
            src2 = __ffs(__ballot_sync(kFull, k2 == bmin)) - 1;
        }
        if constexpr (CL == 1) {
            ox = S2.x[par][src2];
            oy = S2.y[par][src2];
            oz = S2.z[par][src2];
            if (tid == 0) o[j] = oi[S2.pos[par][src2]];
        } else {
            // publish this CTA's winner (tie key of the GLOBAL original index) to every CTA of the cluster
            if (warp == 0 && lane < CL) {
                const int wd = S2.d[par][src2];
                const unsigned wk = wd < 0 ? 0xffffffffu : fps_tie_key(k_first + oi[S2.pos[par][src2]]);
                st_cluster_u32(&peers.d[par][rank], lane, static_cast<unsigned>(wd));
                st_cluster_u32(&peers.key[par][rank], lane, wk);
                st_cluster_u32(&peers.x[par][rank], lane, __float_as_uint(S2.x[par][src2]));
                st_cluster_u32(&peers.y[par][rank], lane, __float_as_uint(S2.y[par][src2]));
                st_cluster_u32(&peers.z[par][rank], lane, __float_as_uint(S2.z[par][src2]));
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
            if (tid == 0 && rank == 0) o[j] = fps_tie_key_inv(gmin);
        }
        F3D_FPS_STAMP(3);
    }
    if (CL > 1) cluster_sync_all();  // no CTA exits while a peer may still write into its shared memory
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

template <int PPT, int CL>
static int launch_fps_smem(int b, int n, int m, const float *inp, int *out, cudaStream_t st) {
    const size_t smem = static_cast<size_t>(kFpsThreads) * PPT * (3 * sizeof(float) + sizeof(unsigned short));
    cudaError_t e = cudaFuncSetAttribute(fps_cull_kernel<PPT, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "fps: cudaFuncSetAttribute");
    if (CL == 1) {
        fps_cull_kernel<PPT, CL><<<b, kFpsThreads, smem, st>>>(n, m, inp, out);
        return check_launch("fps_cull_kernel");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(b) * CL);
    cfg.blockDim = dim3(kFpsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, fps_cull_kernel<PPT, CL>, n, m, inp, out);
    ++g_launches;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(static_cast<int>(e), "fps_cull_kernel (cluster launch)");
    }
    return 0;
}

}  // namespace f3d

using namespace f3d;

F3D_API int f3d_farthest_point_sample(int b, int n, int m, const float *inp, float *temp, int *out, void *stream) {
    if (b < 0 || n <= 0 || m <= 0 || !inp || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "farthest_point_sample: bad arguments");
    if (b == 0) return 0;
    cudaStream_t st = as_stream(stream);
    if (n <= 1024 * 4) return launch_fps_smem<4, 1>(b, n, m, inp, out, st);
    if (n <= 1024 * 8) return launch_fps_smem<8, 1>(b, n, m, inp, out, st);
    if (n <= 1024 * 16) return launch_fps_smem<16, 1>(b, n, m, inp, out, st);
    // larger clouds: a cluster of 2 / 4 / 8 CTAs per cloud, 16384 points each (KITTI-shape scans, 131072 points)
    if (n <= 2 * 1024 * 16) return launch_fps_smem<16, 2>(b, n, m, inp, out, st);
    if (n <= 4 * 1024 * 16) return launch_fps_smem<16, 4>(b, n, m, inp, out, st);
    if (n <= 8 * 1024 * 16) return launch_fps_smem<16, 8>(b, n, m, inp, out, st);
    if (!temp) return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "farthest_point_sample: n > 131072 needs temp of b*n floats");
    fps_global_kernel<<<b, kFpsThreads, 0, st>>>(n, m, inp, temp, out);
    return check_launch("fps_global_kernel");
}

// Bring-up: device buffer of 256*32*4 int64 receiving CTA 0's per-round clock64() stamps of fps_cull_kernel (NULL disables).
F3D_API void f3d_debug_set_fps_timeline(void *buf) {
    long long *p = static_cast<long long *>(buf);
    cudaMemcpyToSymbol(f3d::g_fps_dbg, &p, sizeof(p));
}

F3D_API int f3d_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || !inp || !idx || !out) return fail(F3D_ERR_INVALID_ARGUMENT, "gather_point: bad arguments");
    const long long total = 3LL * b * m;
    if (total == 0) return 0;
    gather_point_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(n, m, total, inp, idx, out);
    return check_launch("gather_point_kernel");
}
