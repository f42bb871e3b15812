// scatter.cu -- deterministic scatter-add (the gradients of gather_point and group_point) for sm_100a.
//
// Replaces tf_sampling_g.cu:183-192 (scatteraddpointKernel) and tf_grouping_g.cu:115-132
// (group_point_grad_gpu).  The reference zero-fills on the host side and then issues one atomicAdd per
// element: the fp32 sum order -- hence the result -- changes from run to run, and popular points serialise.
// Here the (slot -> point) map is inverted once -- per cloud in shared memory by a stable counting sort
// (scatter_cloud_kernel, clouds of up to 16384 points / 65535 slots), otherwise with a stable LSD radix sort
// of (point id, slot id) pairs -- and every output element is then produced by exactly one thread that adds
// its rows in ascending slot order: no atomics on floats, bit-reproducible, and equal to the reference's CPU
// statement (tf_ops/grouping/test/query_ball_point.cpp:68-84), which accumulates in the same (j,k) order.
//
// The radix sort is hand-written (8-bit digits; per pass: block histograms -> one exclusive scan ->
// stable scatter with warp match_any ranking).
#include "common.cuh"

namespace f3d {

constexpr int kRsTile = 16384;    // keys per CTA (128 CTAs = one wave for the 2 M slots of a 64 x 512 x 64 batch)
constexpr int kRsThreads = 256;   // 8 warps, 2048 consecutive keys each
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsPerWarp = kRsTile / kRsWarps;

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const unsigned *__restrict__ keys, long long n, int shift, unsigned *__restrict__ hist, int nblocks) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long t0 = static_cast<long long>(blockIdx.x) * kRsTile;
    for (int i = threadIdx.x; i < kRsTile; i += kRsThreads) {
        const long long g = t0 + i;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];  // digit-major
}

// exclusive scan of `count` unsigned values in place, one CTA of 1024 threads.  Each WARP owns a contiguous chunk and walks it
// 32 values at a time (coalesced), twice: once for the chunk totals, once to write the prefix.
__global__ void __launch_bounds__(1024)
rs_scan_kernel(unsigned *__restrict__ data, long long count) {
    __shared__ unsigned warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long per = ((count + 31) / 32 + 31) / 32 * 32;  // chunk length, multiple of 32
    const long long lo = min(per * warp, count), hi = min(lo + per, count);
    unsigned sum = 0;
    for (long long i = lo + lane; i < hi; i += 32) sum += data[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    if (lane == 0) warp_tot[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = warp_tot[lane];
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(kFull, winc, o);
            if (lane >= o) winc += v;
        }
        warp_tot[lane] = winc - w;  // exclusive
    }
    __syncthreads();
    unsigned run = warp_tot[warp];
    for (long long i0 = lo; i0 < hi; i0 += 32) {
        const long long i = i0 + lane;
        const unsigned v = i < hi ? data[i] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += u;
        }
        if (i < hi) data[i] = run + inc - v;
        run += __shfl_sync(kFull, inc, 31);
    }
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const unsigned *__restrict__ keys_in, const unsigned *__restrict__ vals_in,
                  unsigned *__restrict__ keys_out, unsigned *__restrict__ vals_out, long long n, int shift,
                  const unsigned *__restrict__ offsets, int nblocks) {
    __shared__ unsigned wh[kRsWarps][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kRsWarps * 256; i += kRsThreads) (&wh[0][0])[i] = 0;
    __syncthreads();
    const long long w0 = static_cast<long long>(blockIdx.x) * kRsTile + static_cast<long long>(warp) * kRsPerWarp;
    for (int ch = 0; ch < kRsPerWarp; ch += 32) {
        const long long g = w0 + ch + lane;
        if (g < n) atomicAdd(&wh[warp][(keys_in[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    {  // per digit: turn the per-warp counts into per-warp global start offsets
        unsigned base = offsets[static_cast<size_t>(tid) * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const unsigned c = wh[w][tid];
            wh[w][tid] = base;
            base += c;
        }
    }
    __syncthreads();
    const unsigned lt = lanemask_lt();
    for (int ch = 0; ch < kRsPerWarp; ch += 32) {
        const long long g = w0 + ch + lane;
        const bool valid = g < n;
        const unsigned mask = __ballot_sync(kFull, valid);
        if (valid) {
            const unsigned key = keys_in[g];
            const unsigned d = (key >> shift) & 255u;
            const unsigned peers = __match_any_sync(mask, d);
            const unsigned pos = wh[warp][d] + __popc(peers & lt);
            __syncwarp(mask);
            if ((peers & lt) == 0) wh[warp][d] += __popc(peers);  // lowest lane of each digit group
            __syncwarp(mask);
            keys_out[pos] = key;
            vals_out[pos] = vals_in[g];
        }
    }
}

// keys[i] = global point id (batch*n + idx[i]) or the sentinel b*n for an out-of-range index; vals[i] = i
__global__ void scatter_keys_kernel(long long total, long long slots_per_batch, int n, unsigned sentinel,
                                    const int *__restrict__ idx, unsigned *__restrict__ keys,
                                    unsigned *__restrict__ vals) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long bb = i / slots_per_batch;
    const int a = idx[i];
    keys[i] = (a >= 0 && a < n) ? static_cast<unsigned>(bb * n + a) : sentinel;
    vals[i] = static_cast<unsigned>(i);
}

// One thread per (sorted position p, channel l).  The thread at the HEAD of a run of equal keys sums the run in ascending
// slot order (the sort is stable and the values started ascending) and writes the point's output; points that no slot
// refers to keep the zero of the preceding memset.  No search, no atomics.
__global__ void segmented_sum_kernel(long long total, int c, unsigned sentinel, const unsigned *__restrict__ keys,
                                     const unsigned *__restrict__ vals, const float *__restrict__ grad, float *__restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total * c) return;
    const long long p = i / c;
    const int l = static_cast<int>(i - p * c);
    const unsigned q = keys[p];
    if (q >= sentinel || (p > 0 && keys[p - 1] == q)) return;  // out-of-range index, or not the head of its run
    float acc = 0.0f;
    for (long long r = p; r < total && keys[r] == q; ++r) acc += grad[static_cast<size_t>(vals[r]) * c + l];
    out[static_cast<size_t>(q) * c + l] = acc;
}

// ---- one CTA per cloud: deterministic scatter-add without a global sort --------------------------------------------------
// For n <= 16384 points and L <= 65535 slots per cloud everything but the slot list fits one SM's shared memory:
//   1. count the slots of every point, separately for kScSeg contiguous SEGMENTS of the slot range (shared-memory atomics on
//      16-bit counters packed two per word; the counts do not depend on the order of the atomics),
//   2. block scan -> for every (segment, point) the first position of its slots in the cloud's CSR list,
//   3. placement: warp q walks segment q 32 slots at a time IN ORDER; match_any groups the lanes that hit the same point, the
//      group leader advances the (segment, point) cursor, lane order gives the rank: the list of every point is in ascending
//      slot order by construction (a stable counting sort whose only serial part is 4 warps x L/128 steps),
//   4. (second kernel, whole GPU) one thread per (point, channel) adds its rows in list order = ascending slot order: no atomics on
//      floats, bit-reproducible, equal to the reference's CPU statement (test/query_ball_point.cpp:68-84); points without slots get
//      their zero here.
// Two launches instead of eleven (keys, 3 x {histogram, scan, scatter}, memset, segmented sum).
constexpr int kScThreads = 512;

// P CTAs per cloud (P = 1, 2 or 4): CTA `part` owns the points [part * np, (part + 1) * np) and 4 P slot segments, so the shared
// memory stays 8 n + 2 n / P bytes while the serial part of the placement shrinks to L / (128 P) steps per warp.  Every CTA reads
// all the indices of its cloud (L2-resident) and ignores the ones outside its point range; the number of valid slots that refer
// to LOWER points is the base of its slice of the cloud's list.
__global__ void __launch_bounds__(kScThreads, 1)
scatter_cloud_kernel(int n, int L, int P, const int *__restrict__ idx, unsigned short *__restrict__ lists, unsigned short *__restrict__ offs) {
    extern __shared__ unsigned sc_smem[];
    const int seg = 4 * P;                                   // slot segments = placement warps; two 16-bit counters per word
    const int np = ((n + P - 1) / P + 31) / 32 * 32;         // points per CTA
    unsigned *cnt = sc_smem;                                 // [seg / 2][np]: (segment 2h | segment 2h+1 << 16)
    unsigned short *off = reinterpret_cast<unsigned short *>(cnt + (seg / 2) * np);  // [np]
    __shared__ unsigned warp_tot[kScThreads / 32];
    __shared__ unsigned below_tot[kScThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / P, part = blockIdx.x - cloud * P;
    const int pb = part * np, pe = min(pb + np, n);          // this CTA's points
    const int *ix = idx + static_cast<size_t>(cloud) * L;
    unsigned short *list = lists + static_cast<size_t>(cloud) * L;
    const int seg_len = ((L + seg - 1) / seg + 31) / 32 * 32;  // slots per segment, a multiple of 32

    for (int i = tid; i < (seg / 2) * np; i += kScThreads) cnt[i] = 0u;
    __syncthreads();
    // 1. counts of this CTA's points per segment; valid slots of lower points
    unsigned below = 0;
    for (int s = tid; s < L; s += kScThreads) {
        const int p = ix[s];
        if (p >= pb && p < pe) {
            const int q = s / seg_len;
            atomicAdd(&cnt[(q >> 1) * np + (p - pb)], 1u << (16 * (q & 1)));
        } else if (p >= 0 && p < pb) {
            ++below;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(kFull, below, o);
    if (lane == 0) below_tot[warp] = below;
    __syncthreads();
    // 2. exclusive scan over the points, counts -> first positions.  Warp w owns a contiguous block of points and reads it 32
    //    consecutive points at a time (conflict-free); pass A gives the warp totals, pass B the positions.
    constexpr int kWarps = kScThreads / 32;
    const int rows = ((np + kWarps - 1) / kWarps + 31) / 32;  // 32-point rows per warp
    const int wbase = warp * rows * 32;
    auto total_of = [&](int pl) -> unsigned {  // pl: point index local to this CTA
        unsigned t = 0;
        if (pl < pe - pb) {
            for (int h = 0; h < seg / 2; ++h) {
                const unsigned w = cnt[h * np + pl];
                t += (w & 0xffffu) + (w >> 16);
            }
        }
        return t;
    };
    unsigned wsum = 0;
    for (int r = 0; r < rows; ++r) wsum += total_of(wbase + r * 32 + lane);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(kFull, wsum, o);
    if (lane == 0) warp_tot[warp] = wsum;
    __syncthreads();
    unsigned base0 = 0;  // valid slots of lower points = first list position of this CTA
#pragma unroll
    for (int w = 0; w < kWarps; ++w) base0 += below_tot[w];
    __syncthreads();
    if (warp == 0) {
        const unsigned w = lane < kWarps ? warp_tot[lane] : 0u;
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(kFull, winc, o); if (lane >= o) winc += v; }
        if (lane < kWarps) warp_tot[lane] = winc - w;
        if (lane == kWarps - 1 && part == P - 1)  // end of the last point's list = number of valid slots of the cloud
            offs[static_cast<size_t>(cloud) * (n + 1) + n] = static_cast<unsigned short>(base0 + winc);
    }
    __syncthreads();
    unsigned carry = base0 + warp_tot[warp];
    for (int r = 0; r < rows; ++r) {
        const int pl = wbase + r * 32 + lane;
        const unsigned t = total_of(pl);
        unsigned inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
        unsigned run = carry + inc - t;
        carry += __shfl_sync(kFull, inc, 31);
        if (pl < pe - pb) {
            off[pl] = static_cast<unsigned short>(run);
            for (int h = 0; h < seg / 2; ++h) {
                const unsigned w = cnt[h * np + pl];
                const unsigned a = run, b2 = run + (w & 0xffffu);
                run = b2 + (w >> 16);
                cnt[h * np + pl] = a | (b2 << 16);
            }
        }
    }
    __syncthreads();
    // 3. stable placement: warp q walks segment q in slot order
    if (warp < seg) {
        const int s_begin = warp * seg_len, s_end = min(s_begin + seg_len, L);
        const unsigned lt = lanemask_lt();
        const int sh = 16 * (warp & 1);
        unsigned *cw = cnt + (warp >> 1) * np;
        int pnext = s_begin + lane < s_end ? ix[s_begin + lane] : -1;
        for (int s0 = s_begin; s0 < s_end; s0 += 32) {
            const int s = s0 + lane;
            int p = pnext;
            pnext = s + 32 < s_end ? ix[s + 32] : -1;  // the next chunk's indices are in flight during this chunk's ranking
            const bool valid = p >= pb && p < pe;
            if (!valid) p = -1 - lane;  // a private value: no peers
            const unsigned peers = __match_any_sync(kFull, p);
            const int leader = __ffs(peers) - 1;
            const int rank = __popc(peers & lt);
            unsigned base = 0;
            if (valid && lane == leader) base = (atomicAdd(&cw[p - pb], static_cast<unsigned>(__popc(peers)) << sh) >> sh) & 0xffffu;
            base = __shfl_sync(kFull, base, leader);
            if (valid) list[base + rank] = static_cast<unsigned short>(s);
        }
    }
    // the first positions (off) go to global memory for the summation kernel; `off` was final before the placement started
    unsigned short *og = offs + static_cast<size_t>(cloud) * (n + 1) + pb;
    for (int i = tid; i < pe - pb; i += kScThreads) og[i] = off[i];
}

// 4. one thread per (cloud, point, channel): adds the rows of the point's list in list order = ascending slot order
__global__ void __launch_bounds__(256)
scatter_sum_kernel(long long total, int n, int c, int L, const unsigned short *__restrict__ lists, const unsigned short *__restrict__ offs,
                   const float *__restrict__ grad, float *__restrict__ out) {
    const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long bp = t / c;  // cloud * n + point
    const int ch = static_cast<int>(t - bp * c);
    const long long cloud = bp / n;
    const int p = static_cast<int>(bp - cloud * n);
    const unsigned short *o = offs + cloud * (n + 1) + p;
    const int beg = o[0], end = o[1];
    const unsigned short *list = lists + cloud * L;
    const float *g = grad + cloud * L * c + ch;
    float acc = 0.0f;
    for (int r = beg; r < end; ++r) acc += __ldg(g + static_cast<size_t>(list[r]) * c);
    out[t] = acc;
}

static inline unsigned blocks_for(long long total, int per_block) {
    return static_cast<unsigned>((total + per_block - 1) / per_block);
}

// Sorts (keys, vals) by the low `bits` bits of the keys.  a = input buffers, b = scratch; returns which holds the result.
static int radix_sort_pairs(unsigned *&ka, unsigned *&va, unsigned *&kb, unsigned *&vb, unsigned *hist, long long n,
                            int bits, cudaStream_t st) {
    const int nblocks = static_cast<int>((n + kRsTile - 1) / kRsTile);
    for (int shift = 0; shift < bits; shift += 8) {
        rs_hist_kernel<<<nblocks, kRsThreads, 0, st>>>(ka, n, shift, hist, nblocks);
        int rc = check_launch("rs_hist_kernel");
        if (rc) return rc;
        rs_scan_kernel<<<1, 1024, 0, st>>>(hist, 256LL * nblocks);
        rc = check_launch("rs_scan_kernel");
        if (rc) return rc;
        rs_scatter_kernel<<<nblocks, kRsThreads, 0, st>>>(ka, va, kb, vb, n, shift, hist, nblocks);
        rc = check_launch("rs_scatter_kernel");
        if (rc) return rc;
        unsigned *t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    return 0;
}

// out (b,n,c) = scatter-add of grad (b,L,c) at idx (b,L)
static int scatter_add_sorted(int b, int n, int c, long long L, const float *grad, const int *idx, float *out,
                              void *workspace, size_t workspace_bytes, cudaStream_t st) {
    const long long total = static_cast<long long>(b) * L;
    const long long out_total = static_cast<long long>(b) * n * c;
    if (out_total == 0) return 0;
    if (static_cast<long long>(b) * n >= 0xffffffffLL || total >= 0xffffffffLL)
        return fail(F3D_ERR_UNSUPPORTED, "scatter-add: more than 2^32 points or slots");
    if (total == 0) {
        cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * out_total, st);
        return e == cudaSuccess ? 0 : fail(static_cast<int>(e), "scatter-add: memset");
    }
    if (!workspace || workspace_bytes < f3d_scatter_workspace_bytes(total))
        return fail(F3D_ERR_WORKSPACE_TOO_SMALL, "scatter-add: workspace too small");
    const size_t cloud_path_bytes = (static_cast<size_t>((total + 7) & ~7LL) + static_cast<size_t>(b) * (n + 1)) * sizeof(unsigned short);
    if (n <= 16384 && L <= 65535 && workspace_bytes >= cloud_path_bytes) {
        // one CTA per cloud builds the CSR lists, then a grid-wide ordered sum (see scatter_cloud_kernel); a workspace sized by
        // f3d_scatter_workspace_bytes always suffices unless the clouds have far more points than slots (then: the sort below)
        static int num_sms = 0;
        if (num_sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
            if (num_sms <= 0) num_sms = 148;
        }
        const int P = (4LL * b <= num_sms && n >= 4 * 1024) ? 4 : ((2LL * b <= num_sms && n >= 2 * 1024) ? 2 : 1);  // CTAs per cloud
        const int np = ((n + P - 1) / P + 31) / 32 * 32;
        const size_t smem = static_cast<size_t>(2 * P) * np * sizeof(unsigned) + (static_cast<size_t>(np) + 2) * sizeof(unsigned short);
        cudaError_t e = cudaFuncSetAttribute(scatter_cloud_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "scatter-add: cudaFuncSetAttribute");
        unsigned short *lists = static_cast<unsigned short *>(workspace);                 // b x L
        unsigned short *offs = lists + ((total + 7) & ~7LL);                              // b x (n + 1)
        scatter_cloud_kernel<<<b * P, kScThreads, smem, st>>>(n, static_cast<int>(L), P, idx, lists, offs);
        int rc = check_launch("scatter_cloud_kernel");
        if (rc) return rc;
        scatter_sum_kernel<<<blocks_for(out_total, 256), 256, 0, st>>>(out_total, n, c, static_cast<int>(L), lists, offs, grad, out);
        return check_launch("scatter_sum_kernel");
    }
    unsigned *ka = static_cast<unsigned *>(workspace);
    unsigned *va = ka + total;
    unsigned *kb = va + total;
    unsigned *vb = kb + total;
    unsigned *hist = vb + total;
    const unsigned sentinel = static_cast<unsigned>(static_cast<long long>(b) * n);
    scatter_keys_kernel<<<blocks_for(total, 256), 256, 0, st>>>(total, L, n, sentinel, idx, ka, va);
    int rc = check_launch("scatter_keys_kernel");
    if (rc) return rc;
    int bits = 1;
    while ((1ULL << bits) <= sentinel) ++bits;
    rc = radix_sort_pairs(ka, va, kb, vb, hist, total, bits, st);
    if (rc) return rc;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * out_total, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "scatter-add: memset");
    segmented_sum_kernel<<<blocks_for(total * c, 256), 256, 0, st>>>(total, c, sentinel, ka, va, grad, out);
    return check_launch("segmented_sum_kernel");
}

}  // namespace f3d

using namespace f3d;

// Workspace for a scatter-add of b clouds of n points and slots_per_cloud slots: the minimum (f3d_scatter_workspace_bytes) plus,
// when the clouds have more points than slots, the room the per-cloud path needs for its offset table.
F3D_API size_t f3d_scatter_add_workspace_bytes(int b, int n, long long slots_per_cloud) {
    if (b < 0) b = 0;
    if (n < 0) n = 0;
    if (slots_per_cloud < 0) slots_per_cloud = 0;
    const long long total = static_cast<long long>(b) * slots_per_cloud;
    const size_t base = f3d_scatter_workspace_bytes(total);
    const size_t cloud = (static_cast<size_t>((total + 7) & ~7LL) + static_cast<size_t>(b) * (static_cast<size_t>(n) + 1)) * sizeof(unsigned short) + 64;
    return base > cloud ? base : cloud;
}

F3D_API size_t f3d_scatter_workspace_bytes(long long num_slots) {
    if (num_slots < 0) num_slots = 0;
    const long long nblocks = (num_slots + kRsTile - 1) / kRsTile;
    return static_cast<size_t>(num_slots) * 16 + static_cast<size_t>(nblocks) * 256 * 4 + 1024;  // the per-cloud path needs 2 B per slot + 2 B per point
}

F3D_API int f3d_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || m < 0 || !out_g || !idx || !inp_g)
        return fail(F3D_ERR_INVALID_ARGUMENT, "gather_point_grad: bad arguments");
    return scatter_add_sorted(b, n, 3, m, out_g, idx, inp_g, workspace, workspace_bytes, as_stream(stream));
}

F3D_API int f3d_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                 float *grad_points, void *workspace, size_t workspace_bytes, void *stream) {
    if (b < 0 || n <= 0 || c <= 0 || m < 0 || nsample <= 0 || !grad_out || !idx || !grad_points)
        return fail(F3D_ERR_INVALID_ARGUMENT, "group_point_grad: bad arguments");
    return scatter_add_sorted(b, n, c, static_cast<long long>(m) * nsample, grad_out, idx, grad_points, workspace,
                              workspace_bytes, as_stream(stream));
}
