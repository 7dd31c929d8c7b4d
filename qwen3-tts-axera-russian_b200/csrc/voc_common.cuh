// Shared declarations for the B200 vocoder library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#define VOC_MAX_TAPS 8

// ---- error codes of the C ABI (include/voc_b200.h) ----
#define VOC_OK 0
#define VOC_E_INVALID (-1)   // bad argument / code out of range
#define VOC_E_CUDA (-2)      // CUDA runtime or driver failure
#define VOC_E_STATE (-3)     // call order (weights missing, not finalized ...)
#define VOC_E_NOMEM (-4)

// ------------------------------------------------------------------------------------
// "tap GEMM": the one contraction shape every dense layer of the vocoder maps onto.
//
//   acc[b, m, n] = sum_{tap} sum_{k<K}  A[b, m + a_row0 + tap_off[tap], k] * W[tap*K + k, n]
//
// with rows of A outside [0, a_rows) reading as zero (causal left padding / transposed-conv
// edges).  Activations are channels-last ([batch][time][channel], channel contiguous), so
//   * a causal Conv1d(k, dilation d) is ntaps = k, tap_off[j] = -(k-1-j)*d;
//   * a ConvTranspose1d(k = 2s, stride s) is ntaps = 2, tap_off = {0, -1}, N = s*C_out
//     (column p*C_out+co = phase p), and its [M][s*C_out] output *is* the channels-last
//     [M*s][C_out] signal;
//   * Linear / 1x1 conv is ntaps = 1.
// Epilogue:  v = acc + bias[n];  act;  v *= scale[n];  v += R[b,m,n];
//            Y[b,m,n] = v (optional);   S[b,m,n] = snake(v; sn_a[n], sn_invb[n]) (optional).
// ------------------------------------------------------------------------------------
struct TapGemmParams {
    const float* A;  long long a_bstride;  int a_rows;  int lda;  int K;
    int a_row0;  int ntaps;  int tap_off[VOC_MAX_TAPS];
    const float* W;      // SIMT path: [ntaps*K][N], n contiguous
    int N;  int M;  int B;
    const float* bias;   // [N] or nullptr
    int act;             // 0 none, 1 exact GELU
    const float* scale;  // [N] or nullptr
    const float* R;  long long r_bstride;  int ldr;
    float* Y;        long long y_bstride;  int ldy;
    float* S;        long long s_bstride;  int lds;
    const float* sn_a;  const float* sn_invb;   // [N]: exp(alpha), 1/(exp(beta)+eps)
};

enum { VOC_ACT_NONE = 0, VOC_ACT_GELU = 1 };

// sin^2(t), absolute error <= 3e-7 for |t| < 1e4.  sin^2 has period pi, so one Cody-Waite
// reduction to r in [-pi/2, pi/2] and an even polynomial in r^2 -- no quadrant logic.
__device__ __forceinline__ float voc_sin2(float t) {
    const float k = rintf(t * 0.31830988618379067f);
    float r = fmaf(-k, 3.140625f, t);
    r = fmaf(-k, 0.0009670257568359375f, r);
    r = fmaf(-k, 6.278329465203569e-07f, r);
    const float u = r * r;
    float p = 8.086376368510173e-08f;
    p = fmaf(p, u, -4.2304154703742824e-06f);
    p = fmaf(p, u, 0.00014101542183198035f);
    p = fmaf(p, u, -0.0031745336018502712f);
    p = fmaf(p, u, 0.04444441571831703f);
    p = fmaf(p, u, -0.3333333432674408f);
    p = fmaf(p, u, 1.0f);
    return p * u;
}

// SnakeBeta with pre-exponentiated parameters: x + invb * sin^2(a*x)
__device__ __forceinline__ float voc_snake(float x, float a, float invb) {
    return fmaf(invb, voc_sin2(x * a), x);
}

__device__ __forceinline__ float voc_gelu(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// host-side launch wrappers (simt_kernels.cu)
cudaError_t voc_launch_tapgemm_simt(const TapGemmParams& p, cudaStream_t st);
cudaError_t voc_launch_rvq_gather(const long long* codes, int n_frames, int frames_per_win, int win_step,
                                  int n_windows, int n_q, int codebook_size, const float* tables,
                                  int dim, float* out, int* err_flag, cudaStream_t st);
cudaError_t voc_launch_rmsnorm(const float* x, const float* w, float* y, int rows, int C, float eps,
                               cudaStream_t st);
cudaError_t voc_launch_dwconv_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w,
                                 const float* ln_b, float* y, int B, int L, int C, int ksz, float eps,
                                 cudaStream_t st);
cudaError_t voc_launch_attention(const float* qkv, float* out, int B, int T, int heads, int head_dim,
                                 const float* rope_cos, const float* rope_sin, int window,
                                 cudaStream_t st);
cudaError_t voc_launch_swiglu(const float* gu, float* out, long long rows, int inter, cudaStream_t st);
cudaError_t voc_launch_head(const float* S, long long s_bstride, int L, int C, int ksz, const float* w,
                            float bias, float* out, long long o_bstride, int B, cudaStream_t st);
cudaError_t voc_launch_stitch(const float* chunks, long long chunk_stride, const int* win_meta,
                              int n_windows, int ov, const float* fade_out, const float* fade_in,
                              float* out_f32, short* out_i16, int max_a_len, cudaStream_t st);
cudaError_t voc_launch_append_window(float* res, long long res_len, const float* chunk, int a_len, int ov,
                                     int blend, const float* fade_out, const float* fade_in,
                                     cudaStream_t st);
cudaError_t voc_launch_pcm16(const float* in, short* out, long long n, cudaStream_t st);
