// Pure tcgen05.mma issue/execute rate on B200 (sm_100a): operands resident in shared memory (no TMA in the
// loop, no epilogue), one thread issues `iters` x 4 back-to-back K=16 fp16 MMAs (SS mode, M = 128 per CTA,
// FP32 accumulate in TMEM), then one commit; cycles per MMA = (clock after the commit's barrier flips -
// clock before the first issue) / count.  Variants: N, cta_group 1 / 2, same vs alternating accumulators.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_microbench tools/mma_microbench.cu && /tmp/mma_microbench
//
// NOTE: only the "same-accumulator" columns (a straight-line loop of 8) measure the pipe.  The general loop
// behind the other columns and the gap / per-stage-cost rows carries a run-time modulo and tests, which makes the
// issuing thread itself the limit (~118 cycles per MMA at every N) -- the mistake the real kernel was making too.
// tools/mma_issue_bench.cu is the clean version of those experiments; this file is kept for its cta_group::2 rows.
//
// Bring-up instrument (run by hand on a B200); results are recorded in profiles/r1_mma_microbench.txt.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr >> 4) | (1u << 16); }
constexpr uint32_t DESC_HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B

template <bool TWO>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    if constexpr (TWO)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                     "setp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                     "setp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                     ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
}

template <bool TWO>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int alt, int shift_rows, long long* out, int gap_every = 0,
                                                  int gap_cycles = 0) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot, flag;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t smA = base, smB = base + 2 * 192 * 128;            // A: 2 planes x 192 rows x 128 B; B: 256 rows x 128 B x 2
    uint32_t rank = 0;
    if (TWO) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (uint32_t i = threadIdx.x * 4; i < 2 * 192 * 128 + 2 * 256 * 128; i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)) + i) = 0x3c003c00u;     // fp16 1.0 pairs
    if (threadIdx.x == 0) {
        flag = 1;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(&bar3)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        if (TWO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TWO) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((TWO ? 256 : 128) >> 4) << 24);
        const uint32_t a_lo = desc_lo(smA + shift_rows * 128), b_lo = desc_lo(smB);
        const int nacc = alt ? (512 / N) : 1;
        int since = 0; uint32_t lcg = 1;
        const long long t0 = clock64();
        if (!alt && !gap_every) {
            // clean path: straight-line groups of 8, no modulo, no gap test (the general loop below is itself
            // issue-bound at ~118 cycles per MMA -- see tools/mma_issue_bench.cu)
            for (int it = 0; it < iters / 2; ++it) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) mma<TWO>(tmem, a_lo + (ks & 3) * 2, b_lo + (ks & 3) * 2, idesc);
            }
        } else
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (uint32_t)((it % nacc) * N);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma<TWO>(d, a_lo + ks * 2, b_lo + ks * 2, idesc);
            if (gap_every && ++since == gap_every) {               // the issuing thread does something else for a while:
                since = 0;                                          // a dependent ALU chain (~4 cycles per step), no clock reads
                if (gap_cycles >= 0) {
                    for (int g = 0; g < gap_cycles; g += 4) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(lcg));
                } else {
                    // what a pipeline stage costs the issuing thread besides its MMAs: bit 0 = one try_wait on a
                    // barrier whose phase is already complete, bit 1 = one tcgen05.commit (to a dummy barrier)
                    const int kind = -gap_cycles;
                    if (kind & 1) {
                        uint32_t okk;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(okk) : "r"(smem_u32(&bar2)) : "memory");
                        lcg += okk;
                    }
                    if (kind & 8) {                                  // non-blocking test_wait instead of try_wait
                        uint32_t okk;
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                     : "=r"(okk) : "r"(smem_u32(&bar2)) : "memory");
                        lcg += okk;
                    }
                    if (kind & 16) {                                 // a plain volatile shared-memory load (a relayed flag)
                        uint32_t v;
                        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory");
                        lcg += v;
                    }
                    if (kind & 32) {                                 // try_wait whose result gates the next group (consumed)
                        uint32_t okk = 0;
                        while (!okk)
                            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                         : "=r"(okk) : "r"(smem_u32(&bar2)) : "memory");
                    }
                    if (kind & 64) {                                 // a volatile load whose result gates the next group
                        uint32_t v = 0;
                        while (!v) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory");
                    }
                    if (kind & 4) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (kind & 2) {
                        if (TWO)
                            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                         ::"r"(smem_u32(&bar3)), "h"((uint16_t)1) : "memory");
                        else
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar3)) : "memory");
                    }
                }
            }
        }
        if (TWO)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long t_issue = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        out[2 * blockIdx.x] = t1 - t0;
        out[2 * blockIdx.x + 1] = t_issue - t0 + (lcg == 12345u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TWO) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (threadIdx.x < 32) {
        if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <bool TWO>
static double run(int N, int iters, int alt, int shift, int grid, long long* d_out, double* issue_out, int gap_every = 0,
                  int gap_cycles = 0) {
    const size_t smem = 2 * 192 * 128 + 2 * 256 * 128 + 1024;
    cudaFuncSetAttribute(bench<TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_out, 0, sizeof(long long) * 2 * grid);
    for (int rep = 0; rep < 2; ++rep) {
        if (TWO) {
            cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, bench<TWO>, N, iters, alt, shift, d_out, gap_every, gap_cycles);
        } else {
            bench<TWO><<<grid, 128, smem>>>(N, iters, alt, shift, d_out, gap_every, gap_cycles);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
    }
    std::vector<long long> h(2 * grid);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double tot = 0, iss = 0; int n = 0;
    for (int i = 0; i < grid; ++i) if (h[2 * i] > 0) { tot += (double)h[2 * i]; iss += (double)h[2 * i + 1]; ++n; }
    *issue_out = iss / n / (iters * 4.0);
    return tot / n / (iters * 4.0);
}

int main() {
    long long* d_out = nullptr;
    cudaMalloc(&d_out, sizeof(long long) * 2 * 148);
    const int iters = 2000;
    printf("# cycles per tcgen05.mma (M 128 per CTA, K 16, fp16 -> fp32, SS mode, SWIZZLE_128B K-major), %d x 4 MMAs per CTA\n", iters);
    printf("# columns: mode grid N | same-accumulator: total, issue-only | alternating accumulators: total | A start shifted by 3 rows: total | math floor\n");
    for (int grid : {2, 148}) {
        for (int N : {32, 64, 96, 128, 192, 256}) {
            double is1, is2, is3;
            const double a = run<false>(N, iters, 0, 0, grid, d_out, &is1);
            const double b = run<false>(N, iters, 1, 0, grid, d_out, &is2);
            const double c = run<false>(N, iters, 0, 3, grid, d_out, &is3);
            printf("single grid %3d N %3d | %7.1f %7.1f | %7.1f | %7.1f | floor %5.1f\n", grid, N, a, is1, b, c, N / 2.0);
        }
        for (int N : {32, 64, 96, 128, 192, 256}) {
            double is1, is2, is3;
            const double a = run<true>(N, iters, 0, 0, grid, d_out, &is1);
            const double b = run<true>(N, iters, 1, 0, grid, d_out, &is2);
            const double c = run<true>(N, iters, 0, 3, grid, d_out, &is3);
            printf("pair   grid %3d N %3d | %7.1f %7.1f | %7.1f | %7.1f | floor %5.1f (M 256 over two SMs)\n", grid, N, a, is1, b, c, N / 2.0);
        }
    }
    // how much of a gap in the issue stream does the pipe's queue absorb?  12 MMAs (3 x 4) then a gap
    printf("# issue gaps: N 192, 12 MMAs then the issuing thread idles `gap` cycles; cycles per MMA (12 x 96 = 1152 per group if hidden)\n");
    for (int gap : {0, 100, 200, 400, 800, 1600}) {
        double is;
        const double a = run<false>(192, iters, 0, 0, 148, d_out, &is, 3, gap);
        const double b = run<true>(192, iters, 0, 0, 148, d_out, &is, 3, gap);
        printf("gap %4d cycles: single %7.1f  pair %7.1f   (fully exposed would be %7.1f)\n", gap, a, b, 96.0 + gap / 12.0);
    }
    printf("# per-stage fixed costs: N 192, after every 12 MMAs: (1) one try_wait on a completed barrier, (2) one tcgen05.commit, (4) tcgen05.fence::after; extra cycles per group\n");
    {
        double is;
        const double base1 = run<false>(192, iters, 0, 0, 148, d_out, &is, 3, 0), base2 = run<true>(192, iters, 0, 0, 148, d_out, &is, 3, 0);
        for (int kind : {1, 2, 3, 4, 7, 8, 16, 32, 64, 66, 34}) {
            const double a = run<false>(192, iters, 0, 0, 148, d_out, &is, 3, -kind);
            const double b = run<true>(192, iters, 0, 0, 148, d_out, &is, 3, -kind);
            printf("kind %d: single +%6.1f  pair +%6.1f cycles per group of 12 MMAs\n", kind, (a - base1) * 12, (b - base2) * 12);
        }
        for (int every : {1, 2}) {
            const double a = run<false>(192, iters, 0, 0, 148, d_out, &is, every, -7);
            printf("kind 7 after every %d MMAs x4: single %6.1f cycles per MMA\n", every, a);
        }
    }
    return 0;
}
