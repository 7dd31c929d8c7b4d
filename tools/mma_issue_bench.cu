// How fast can ONE thread feed the B200 tensor pipe, and what does anything else in its instruction stream cost?
// Straight-line code: groups of G back-to-back tcgen05.mma (SS mode, M 128, K 16, fp16 -> fp32, operands resident
// in shared memory), each group followed by a compile-time filler: F dependent integer multiply-adds, and/or one
// mbarrier.try_wait on a completed barrier (W), one tcgen05.commit (C).  No run-time loop control inside a group,
// 8 groups per loop iteration, so loop overhead is amortised away.  Reports cycles per MMA next to the math floor
// N/2 (8192 dense fp16 FLOP per cycle per SM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_issue_bench tools/mma_issue_bench.cu && build/mma_issue_bench
//
// Bring-up instrument (run by hand on a B200); results are recorded in profiles/r1_mma_microbench.txt.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr >> 4) | (1u << 16); }
constexpr uint32_t DESC_HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B

__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "setp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
}

// G MMAs per group, F dependent mads after each group, W try_waits, C commits, NACC accumulators used round-robin per group
template <int G, int F, int W, int C, int NACC>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t smA = base, smB = base + 128 * 128;                 // A: 128 rows x 128 B; B: 256 rows x 128 B
    for (uint32_t i = threadIdx.x * 4; i < 128 * 128 + 256 * 128; i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)) + i) = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 100000000;" ::"r"(smem_u32(&bar3)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        // N < 0: the concatenated form's pattern -- an N = 192 MMA into columns [0, 192) followed by an N = 96 MMA into
        // columns [96 + off, 192 + off) with off = 0 (what tc_gemm.cu does: the correction block is the upper half of
        // the wide MMA's output) or off = 96 (disjoint columns), per k-step
        const bool cat = N < 0;
        const uint32_t cat_off = N == -2 ? 96u : 0u;
        if (cat) N = 192;
        const uint32_t idesc96 = (1u << 4) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_lo = desc_lo(smA), b_lo = desc_lo(smB);
        const uint32_t bar2_a = smem_u32(&bar2), bar3_a = smem_u32(&bar3);
        uint32_t lcg = 1;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int grp = 0; grp < 8; ++grp) {
                const uint32_t d = tmem + (uint32_t)((grp % NACC) * 256);
#pragma unroll
                for (int m = 0; m < G; ++m) {
                    if (cat && (m & 1)) mma(d + 96 + cat_off, a_lo + (m & 3) * 2, b_lo + (m & 3) * 2, idesc96);
                    else mma(d, a_lo + (m & 3) * 2, b_lo + (m & 3) * 2, idesc);
                }
#pragma unroll
                for (int w = 0; w < W; ++w) {
                    uint32_t ok;
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(bar2_a) : "memory");
                    lcg += ok;
                }
#pragma unroll
                for (int f = 0; f < F; ++f) asm volatile("mad.lo.u32 %0, %0, 1664525, 1013904223;" : "+r"(lcg));
#pragma unroll
                for (int c = 0; c < C; ++c)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar3_a) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0 + (lcg == 12345u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int G, int F, int W, int C, int NACC>
static double run(int N, long long* d_out, int grid = 148) {
    const int iters = 400;
    const size_t smem = 128 * 128 + 256 * 128 + 1024;
    cudaFuncSetAttribute(bench<G, F, W, C, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_out, 0, sizeof(long long) * grid);
    for (int rep = 0; rep < 2; ++rep) {
        bench<G, F, W, C, NACC><<<grid, 128, smem>>>(N, iters, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < grid; ++i) tot += (double)h[i];
    return tot / grid / (iters * 8.0 * G);
}

template <int G, int F, int W, int C, int NACC>
static void row(long long* d_out) {
    const double c96 = run<G, F, W, C, NACC>(96, d_out), c192 = run<G, F, W, C, NACC>(192, d_out), c256 = run<G, F, W, C, NACC>(256, d_out);
    printf("G %2d  filler: %3d mads %d try_wait %d commit, %d acc | N 96: %6.1f (+%5.0f/group)  N 192: %6.1f (+%5.0f/group)  N 256: %6.1f (+%5.0f/group)\n",
           G, F, W, C, NACC, c96, (c96 - 64.0) * G, c192, (c192 - 96.0) * G, c256, (c256 - 128.0) * G);
}

int main() {
    long long* d_out = nullptr;
    cudaMalloc(&d_out, sizeof(long long) * 148);
    printf("# straight-line issue: cycles per tcgen05.mma (cta_group::1, M 128, K 16), no filler, 32 per group\n");
    for (int N : {16, 32, 64, 96, 128, 192, 256})
        printf("N %3d: %6.1f cycles per MMA (math floor %5.1f)\n", N, run<32, 0, 0, 0, 1>(N, d_out), N / 2.0);
    printf("# concatenated-form pattern, cycles per (N 192 + N 96) pair: correction block inside the wide MMA's columns %6.1f, disjoint columns %6.1f  (96 + 56 = 152 if independent)\n",
           2 * run<32, 0, 0, 0, 1>(-1, d_out), 2 * run<32, 0, 0, 0, 1>(-2, d_out));
    printf("# cost of instructions between groups; (+x/group) = cycles per group above G x max(64?, N/2) -- see the first table for the true floor\n");
    row<12, 0, 0, 0, 1>(d_out);
    row<12, 8, 0, 0, 1>(d_out);
    row<12, 16, 0, 0, 1>(d_out);
    row<12, 32, 0, 0, 1>(d_out);
    row<12, 64, 0, 0, 1>(d_out);
    row<12, 128, 0, 0, 1>(d_out);
    row<12, 0, 1, 0, 1>(d_out);
    row<12, 0, 0, 1, 1>(d_out);
    row<12, 0, 1, 1, 1>(d_out);
    row<12, 0, 2, 2, 1>(d_out);
    row<12, 32, 1, 1, 1>(d_out);
    row<12, 32, 1, 1, 2>(d_out);
    row<4, 0, 0, 0, 1>(d_out);
    row<4, 16, 0, 0, 1>(d_out);
    row<4, 0, 1, 1, 1>(d_out);
    row<4, 0, 1, 1, 2>(d_out);
    row<24, 0, 1, 1, 1>(d_out);
    row<24, 64, 1, 1, 1>(d_out);
    return 0;
}
