// How should an epilogue whose data arrives one ROW PER LANE (tcgen05.ld 32x32b) write a row-major [rows, 192]
// float32 tensor?  (a) directly: each lane stores 32-byte sectors of its own row (STG.E.ENL2.256: 32 different
// 128-byte lines per instruction -- what tc_gemm.cu does; ncu: ~45 L1 data-pipe wavefronts per request, the LSU
// data pipe at 64 % in the 1x1 convolutions), or (b) staged: each warp writes its 32 x COLS block into a swizzled
// shared-memory buffer (STS.128) and one lane hands it to the TMA unit (cp.async.bulk.tensor store).
// 148 CTAs x 8 warps; warp (q, h) owns rows q*32.. and column half h of each 128 x 192 tile.  Reports TB/s.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_store_bench tools/tma_store_bench.cu && build/tma_store_bench
//
// Bring-up instrument (run by hand on a B200); results are recorded in profiles/r1_mma_microbench.txt.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int N = 192, BM = 128, WARPS = 8;

// ---- (a) direct
__global__ void __launch_bounds__(256, 1) direct_kernel(float* Y, int tiles_per_cta, long long* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, h = warp >> 2;
    const long long t0 = clock64();
    for (int t = 0; t < tiles_per_cta; ++t) {
        const long long row = ((long long)t * gridDim.x + blockIdx.x) * BM + q * 32 + lane;
        float* p = Y + row * N + h * 96;
#pragma unroll
        for (int c = 0; c < 96; c += 8) {
            const float v = (float)(c + t);
            asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p + c), "f"(v) : "memory");
        }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

// ---- (b) staged through shared memory + TMA store; COLS float32 columns per chunk (16 -> 64-byte rows, SWIZZLE_64B;
//      32 -> 128-byte rows, SWIZZLE_128B); NBUF staging buffers per warp
template <int COLS, int NBUF>
__global__ void __launch_bounds__(256, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int tiles_per_cta, long long* out) {
    extern __shared__ uint8_t raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, h = warp >> 2;
    constexpr int ROWB = COLS * 4, BUFB = 32 * ROWB;
    const uint32_t base = ((smem_u32(raw) + 1023u) & ~1023u) + (uint32_t)warp * NBUF * BUFB;
    constexpr int SW = ROWB == 128 ? 7 : 3;                         // 16-byte-chunk XOR mask: (row >> sh) & SW
    constexpr int SH = ROWB == 128 ? 0 : 1;
    const long long t0 = clock64();
    int buf = 0;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const int row0 = (int)(((long long)t * gridDim.x + blockIdx.x) * BM + q * 32);
#pragma unroll 1
        for (int c0 = 0; c0 < 96; c0 += COLS) {
            // the buffer's previous store must have finished READING shared memory
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncwarp();
            const uint32_t b = base + (uint32_t)buf * BUFB + (uint32_t)lane * ROWB;
#pragma unroll
            for (int k = 0; k < ROWB / 16; ++k) {
                const uint32_t kk = (uint32_t)k ^ (((uint32_t)lane >> SH) & SW);
                const float v = (float)(c0 + 4 * k + t);                      // element (row, c) = c + t
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(b + kk * 16), "f"(v), "f"(v + 1.f), "f"(v + 2.f), "f"(v + 3.f) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             ::"l"(&tm), "r"(h * 96 + c0), "r"(row0), "r"(base + (uint32_t)buf * BUFB) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf = buf + 1 == NBUF ? 0 : buf + 1;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int COLS, int NBUF>
static void run_tma(EncodeFn enc, float* Y, long long rows, int tiles, long long* d_out) {
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {(cuuint32_t)COLS, 32};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     COLS == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const size_t smem = (size_t)WARPS * NBUF * 32 * COLS * 4 + 1024;
    cudaFuncSetAttribute(tma_kernel<COLS, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        tma_kernel<COLS, NBUF><<<148, 256, smem>>>(tm, tiles, d_out);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d_out, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
    double cyc = 0; for (long long c : h) cyc += (double)c; cyc /= 148;
    const double bytes = (double)rows * N * 4;
    // spot check
    std::vector<float> chk(N);
    cudaMemcpy(chk.data(), Y + (long long)(148 + 3) * BM * N + 5 * N, N * 4, cudaMemcpyDeviceToHost);   // tile t=1 of CTA 3, row 5
    bool ok = true;
    for (int c = 0; c < N; ++c) ok &= chk[c] == (float)((c % 96) + 1);
    printf("staged + TMA store, %2d columns per chunk (%3d-byte rows), %d buffers per warp: %6.3f ms  %5.2f TB/s  %7.0f cycles per 128 x 192 tile  [%s]\n",
           COLS, COLS * 4, NBUF, best, bytes / best / 1e9, cyc / tiles, ok ? "data ok" : "DATA WRONG");
}

int main() {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qr;
    void* fp = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { printf("no encoder\n"); return 1; }
    enc = (EncodeFn)fp;
    const int tiles = 64;                                   // per CTA
    const long long rows = (long long)tiles * 148 * BM;     // 1.2 M rows x 192 x 4 B = 931 MB
    float* Y; cudaMalloc(&Y, (size_t)rows * N * 4);
    long long* d_out; cudaMalloc(&d_out, sizeof(long long) * 148);
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            direct_kernel<<<148, 256>>>(Y, tiles, d_out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        std::vector<long long> h(148);
        cudaMemcpy(h.data(), d_out, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
        double cyc = 0; for (long long c : h) cyc += (double)c; cyc /= 148;
        printf("direct, one 32-byte sector per lane per store:                              %6.3f ms  %5.2f TB/s  %7.0f cycles per 128 x 192 tile\n",
               best, (double)rows * N * 4 / best / 1e9, cyc / tiles);
    }
    run_tma<16, 2>(enc, Y, rows, tiles, d_out);
    run_tma<16, 3>(enc, Y, rows, tiles, d_out);
    run_tma<32, 2>(enc, Y, rows, tiles, d_out);
    run_tma<32, 3>(enc, Y, rows, tiles, d_out);
    return 0;
}
