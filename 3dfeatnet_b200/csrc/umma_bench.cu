// umma_bench.cu -- measurement aid (not part of the reference surface): how long does the tensor pipe hold ONE tcgen05.mma
// of a given shape when instructions are issued back to back the way the fused row kernels issue them (A operand from
// tensor memory, B operand a bf16 no-swizzle image in shared memory, fp32 accumulation in TMEM)?
//
// det_rows_tc_kernel's timeline (profiles/r01_n_det_tc_timeline.txt) shows 60 instructions of 128x64x16 per 2 443 cycles,
// i.e. ~40.7 cycles per instruction against 32 cycles of math: the per-instruction overhead is what larger instructions
// (N = 128 / 256, or M = 256 over a CTA pair with cta_group::2) would amortise.  This kernel measures that directly, on
// every SM at once (so clocks and power are those of a full-chip tensor load), for
//     mode 0: cta_group::1, M = 128, N in {8 .. 256}
//     mode 1: cta_group::2, M = 256 over a pair of CTAs (cluster of 2), N in {16 .. 256}; each CTA supplies N/2 columns of B
// out[cta] = {cycles of the issue loop incl. the final commit wait, instructions issued}.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace f3d {
using namespace tc;

namespace {
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (count 1) on the barrier at the same shared-memory offset in BOTH CTAs of the pair once every MMA issued so far is done
__device__ __forceinline__ void umma2_commit_both(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}
}  // namespace

// B image: 8 K-steps of a K-major no-swizzle bf16 operand with `rows` rows (LBO = rows*16, SBO = 128), zero-filled.
template <int kMode>
__global__ void __launch_bounds__(128, 1) umma_bench_kernel(int N, int groups, int per_group, int b_mn_major, long long *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_group, bar_done;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    const int rows_b = kMode == 1 ? N / 2 : N;  // rows of B held by THIS CTA
    const uint32_t b_bytes = static_cast<uint32_t>(rows_b) * 16u * 2u * 8u;  // 8 K-steps of 16
    for (uint32_t i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar_group, 1);
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (kMode == 1) {
            tmem_alloc2(&tmem_base_s, 512);
            tmem_relinquish2();
        } else {
            tmem_alloc(&tmem_base_s, 512);
            tmem_relinquish();
        }
    }
    tcgen05_fence_before();
    if (kMode == 1) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const bool leader = kMode == 0 || cluster_ctarank() == 0;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        // D at columns 0..N-1 (and a second accumulator at N..2N-1 when it fits next to the A operand), A at columns 448..511
        const uint32_t idesc = make_idesc(1, kMode == 1 ? 256 : 128, static_cast<uint32_t>(N)) | (b_mn_major ? kIdescBMnMajor : 0u);
        const uint32_t sbase = smem_u32(smem);
        const uint32_t lbo = b_mn_major ? 128u : static_cast<uint32_t>(rows_b) * 16u;
        const uint32_t sbo = b_mn_major ? 16u * 128u : 128u;
        const uint32_t kstep = b_mn_major ? 2u * 128u : 2u * lbo;
        const bool two_acc = 2 * N <= 448;
        if (leader) {
            t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                if (elect_one()) {
                    const uint32_t d = tmem_base + ((g & 1) && two_acc ? static_cast<uint32_t>(N) : 0u);
                    for (int i = 0; i < per_group; ++i) {
                        const uint32_t k = static_cast<uint32_t>(i) & 7u;
                        const uint64_t db = make_smem_desc(sbase + k * kstep, lbo, sbo);
                        if (kMode == 1) umma2_f16_ts(d, tmem_base + 448 + k * 8, db, idesc, i > 0 ? 1u : 0u);
                        else umma_f16_ts(d, tmem_base + 448 + k * 8, db, idesc, i > 0 ? 1u : 0u);
                    }
                    if (kMode == 1) umma2_commit_both(&bar_group); else umma_commit(&bar_group);
                }
                __syncwarp();
            }
            if (elect_one()) {
                if (kMode == 1) umma2_commit_both(&bar_done); else umma_commit(&bar_done);
            }
            __syncwarp();
        }
        mbar_wait(&bar_done, 0);
        t1 = clock64();
        if (leader && threadIdx.x == 0) {
            out[blockIdx.x * 2 + 0] = t1 - t0;
            out[blockIdx.x * 2 + 1] = static_cast<long long>(groups) * per_group;
        }
    } else {
        mbar_wait(&bar_done, 0);
    }
    tcgen05_fence_before();
    if (kMode == 1) cluster_sync_all(); else __syncthreads();
    if (warp == 0) {
        if (kMode == 1) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace f3d

using namespace f3d;

// Measurement aid: `groups` x `per_group` back-to-back tcgen05.mma (M = 128 with cta_group 1, M = 256 with cta_group 2) x N x 16,
// one commit per group, on `ctas` CTAs (cta_group 2: an even number, launched as clusters of 2).  out: 2 int64 per CTA.
F3D_API int f3d_debug_umma_bench(int cta_group, int N, int groups, int per_group, int b_mn_major, int ctas, void *out, void *stream) {
    if ((cta_group != 1 && cta_group != 2) || N < 8 * cta_group || N > 256 || N % (8 * cta_group) || groups <= 0 || per_group <= 0 || ctas <= 0 || !out ||
        (cta_group == 2 && (ctas & 1)))
        return fail(F3D_ERR_INVALID_ARGUMENT, "umma_bench: bad arguments");
    const size_t smem = static_cast<size_t>(cta_group == 2 ? N / 2 : N) * 16 * 2 * 8 + 1024;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(ctas));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = as_stream(stream);
    cudaLaunchAttribute attr[1];
    cudaError_t e;
    if (cta_group == 2) {
        e = cudaFuncSetAttribute(umma_bench_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "umma_bench: cudaFuncSetAttribute");
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, umma_bench_kernel<1>, N, groups, per_group, b_mn_major, static_cast<long long *>(out));
    } else {
        e = cudaFuncSetAttribute(umma_bench_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "umma_bench: cudaFuncSetAttribute");
        e = cudaLaunchKernelEx(&cfg, umma_bench_kernel<0>, N, groups, per_group, b_mn_major, static_cast<long long *>(out));
    }
    if (e != cudaSuccess) return fail(static_cast<int>(e), "umma_bench: launch");
    return check_launch("umma_bench_kernel");
}
