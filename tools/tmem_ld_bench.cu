// TMEM read bandwidth on B200 (tcgen05.ld), alone and under a running tensor pipe.
// The tap-GEMM's epilogue drains every accumulation segment out of TMEM (the tensor core's FP32 accumulate
// truncates, so long sums are finished in registers); this measures what a drain costs.
//   8 (or 4) warps each read COLS columns of their 32-lane quadrant, REPS times, with tcgen05.ld.32x32b.xN;
//   optionally one thread keeps issuing N=192 MMAs into other TMEM columns meanwhile.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_ld_bench tools/tmem_ld_bench.cu && build/tmem_ld_bench
//
// Bring-up instrument (run by hand on a B200); results are recorded in profiles/r1_mma_microbench.txt.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return (addr >> 4) | (1u << 16); }
constexpr uint32_t DESC_HI = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "setp.eq.u32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
}

template <int X>
__device__ __forceinline__ void tld(uint32_t taddr, uint32_t* r);
template <> __device__ __forceinline__ void tld<8>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tld<16>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tld<32>(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}

// WARPS readers (4: one per quadrant, COLS columns each; 8: two per quadrant, COLS columns each at different offsets)
template <int X, int COLS, int WARPS, bool WITH_MMA>
__global__ void __launch_bounds__(32 * WARPS + 32, 1) bench(int reps, int mma_count, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x * 4; i < 128 * 128 + 256 * 128; i += blockDim.x * 4)
        *reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)) + i) = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp < WARPS) {
        const int q = warp & 3, h = warp >> 2;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * COLS);
        uint32_t sum = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint32_t v[COLS];
#pragma unroll
            for (int c = 0; c < COLS; c += X) tld<X>(taddr + c, v + c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < COLS; ++c) sum += v[c];
        }
        const long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0 + (sum == 12345u);
    } else if (WITH_MMA && lane == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_lo = desc_lo(base), b_lo = desc_lo(base + 128 * 128);
        const long long t0 = clock64();
        for (int it = 0; it < mma_count / 8; ++it) {
#pragma unroll
            for (int m = 0; m < 8; ++m) mma(tmem + 256, a_lo + (m & 3) * 2, b_lo + (m & 3) * 2, idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        out[blockIdx.x * 16 + 15] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// WITH_MMA: the MMA stream is sized to outlast the reads when `mma_long` (-> read slowdown under MMA), or to be
// well inside them otherwise (-> MMA slowdown under reads)
template <int X, int COLS, int WARPS, bool WITH_MMA>
static void run(long long* d_out, bool mma_long = true) {
    const int grid = 148, reps = 2000;
    const size_t smem = 128 * 128 + 256 * 128 + 1024;
    // MMAs sized to outlast the reads: reps drains of WARPS*32*COLS*4 bytes at >= 64 B/clk
    const int mma_count = !WITH_MMA ? 0 : mma_long ? 8 * (int)(1.5 * reps * WARPS * 32 * COLS * 4 / 64 / 96 / 8 + 1) : 2000;
    cudaFuncSetAttribute(bench<X, COLS, WARPS, WITH_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemset(d_out, 0, sizeof(long long) * 16 * grid);
    for (int rep = 0; rep < 2; ++rep) {
        bench<X, COLS, WARPS, WITH_MMA><<<grid, 32 * WARPS + 32, smem>>>(reps, mma_count, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return; }
    }
    std::vector<long long> h(16 * grid);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost);
    double rd = 0, mm = 0;
    for (int i = 0; i < grid; ++i) {
        long long mx = 0;
        for (int w = 0; w < WARPS; ++w) mx = h[i * 16 + w] > mx ? h[i * 16 + w] : mx;
        rd += (double)mx; mm += (double)h[i * 16 + 15];
    }
    rd /= grid; mm /= grid;
    const double bytes = (double)reps * WARPS * 32 * COLS * 4;
    printf("x%-2d %3d cols x %d warps%s: %7.1f cycles per drain of %5.1f KB = %6.1f B/clk/SM", X, COLS, WARPS,
           WITH_MMA ? " + MMA" : "      ", rd / reps, bytes / reps / 1024, bytes / rd);
    if (WITH_MMA) printf("   | %d MMAs N=192: %6.1f cycles each (96 alone)%s", mma_count, mm / mma_count, mma_long ? " [outlast the reads]" : " [inside the reads]");
    printf("\n");
}

int main() {
    long long* d_out = nullptr;
    cudaMalloc(&d_out, sizeof(long long) * 16 * 148);
    run<8, 96, 8, false>(d_out);
    run<16, 96, 8, false>(d_out);
    run<32, 96, 8, false>(d_out);
    run<8, 96, 4, false>(d_out);
    run<32, 96, 4, false>(d_out);
    run<32, 192, 4, false>(d_out);
    run<32, 32, 8, false>(d_out);
    run<32, 64, 8, false>(d_out);
    run<8, 96, 8, true>(d_out, true);
    run<8, 96, 8, true>(d_out, false);
    run<32, 96, 8, true>(d_out, true);
    run<32, 96, 8, true>(d_out, false);
    run<32, 96, 4, true>(d_out, true);
    run<32, 96, 4, true>(d_out, false);
    return 0;
}
