// Shared declarations for the B200 vocoder library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <string>

struct CUtensorMap_st;

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
//
// Operand formats.  The FP32 CUDA-core path keeps every operand in float32.  The tensor-core
// path keeps every *GEMM operand* as two float16 planes (hi = fp16(x), lo = fp16(x - hi)), which
// occupy exactly the bytes of the float32 tensor they replace; three tcgen05 MMAs
// (hi*hi + hi*lo + lo*hi) with FP32 accumulation in TMEM then reproduce the FP32 product to
// ~2^-22 relative (oracle/precision_study.py: 91 dB / 2.5e-5 on the full window, against the
// 60 dB / 1e-4 gate; a single bf16 or tf32 pass fails it).  The residual stream (Y, R) is
// always float32.  A_hi/A_lo (when non-null) replace A; S_hi/S_lo replace S.
struct TapGemmParams {
    const float* A;  long long a_bstride;  int a_rows;  int lda;  int K;
    const __half* A_hi;  const __half* A_lo;
    int a_row0;  int ntaps;  int tap_off[VOC_MAX_TAPS];
    const float* W;      // SIMT path: [ntaps*K][N], n contiguous
    int N;  int M;  int B;
    const float* bias;   // [N] or nullptr
    int act;             // 0 none, 1 exact GELU
    const float* scale;  // [N] or nullptr
    const float* R;  long long r_bstride;  int ldr;
    float* Y;        long long y_bstride;  int ldy;
    float* S;        long long s_bstride;  int lds;
    __half* S_hi;  __half* S_lo;                // split-fp16 form of S (same strides)
    const float* sn_a;  const float* sn_invb;   // [N]: 2*exp(alpha), 1/(exp(beta)+eps); null = S is v itself
    // tensor-core form of the weights (tc_gemm.cu): two fp16 planes [2][ntaps][N][K] of W * 2^wexp
    const __half* Wtc;  long long wtc_plane;  float wscale;   // wscale = 2^-wexp
};

enum { VOC_ACT_NONE = 0, VOC_ACT_GELU = 1 };

// SnakeBeta: x + 1/(e^beta + eps) * sin^2(x * e^alpha)  (sibling :3645-3683).
// Device convention: the per-channel arrays hold a2 = 2 * e^alpha and invb = 1/(e^beta + eps) (make_snake in
// voc_engine.cu), because sin^2(t) = 0.5 - 0.5 cos(2t): the doubled argument is reduced to [-pi, pi] by a two-term
// Cody-Waite step (round-to-nearest by the 1.5 * 2^23 trick, two FMAs) and the cosine is one SFU evaluation
// (cos.approx: max abs error 2^-21.19 on [-pi, pi], CUDA C Programming Guide; i.e. <= 2.2e-7 on sin^2, measured in
// tests/test_gpu_tapgemm.py::test_snake_accuracy).  9 instructions per element against 16 for the round-1
// polynomial: the epilogue warps of the tensor-core kernels are issue-bound.
__device__ __forceinline__ float voc_sin2_half(float t2) {
    const float k = fmaf(t2, 0.15915494309189535f, 12582912.f) - 12582912.f;     // rint(t2 / 2 pi), |t2| < 2^22
    float r = fmaf(-k, 6.2831854820251465f, t2);
    r = fmaf(-k, -1.7484555e-07f, r);
    return fmaf(-0.5f, __cosf(r), 0.5f);
}
__device__ __forceinline__ float voc_snake(float x, float a2, float invb) {
    return fmaf(invb, voc_sin2_half(x * a2), x);
}

__device__ __forceinline__ float voc_gelu(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}

// fp32 -> (hi, lo) float16 pair.  The conversions saturate to the largest finite fp16 (F2FP.SATFINITE), so a stray
// huge activation cannot become inf; no separate clamp instructions.
__device__ __forceinline__ uint32_t voc_cvt_f16x2_sat(float a, float b) {
    uint32_t h;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(h) : "f"(a), "f"(b));      // low half = a
    return h;
}
__device__ __forceinline__ void voc_split2(float a, float b, __half2& hi, __half2& lo) {
    const uint32_t h = voc_cvt_f16x2_sat(a, b);
    hi = *reinterpret_cast<const __half2*>(&h);
    const float2 f = __half22float2(hi);
    const uint32_t l = voc_cvt_f16x2_sat(a - f.x, b - f.y);
    lo = *reinterpret_cast<const __half2*>(&l);
}
__device__ __forceinline__ void voc_split1(float a, __half& hi, __half& lo) {
    __half2 h2, l2;
    voc_split2(a, 0.f, h2, l2);
    hi = __low2half(h2);
    lo = __low2half(l2);
}

// An activation tensor as the kernels see it: float32 (f) or split float16 (hi, lo).
struct VocAct {
    float* f;  __half* hi;  __half* lo;
};

// host-side launch wrappers (tc_gemm.cu): returns cudaErrorNotSupported when the shape is not
// eligible for the tensor-core kernel (the caller then uses the CUDA-core kernel)
bool voc_tc_eligible(const TapGemmParams& p);
// cached 4-D fp16 tensor map {d0 (contiguous), d1, d2, 2 planes} with box {bk, box_rows, 1, box_planes}; strides in bytes
bool voc_tc_get_map(const void* base, long long d0, long long d1, long long d2, long long s1, long long s2, long long s3,
                    int bk, int box_rows, int box_planes, CUtensorMap_st* out);
cudaError_t voc_launch_tapgemm_tc(const TapGemmParams& p, cudaStream_t st, int num_sms, int flags);
// the tile plan voc_launch_tapgemm_tc would use (pure host arithmetic): column tile, K chunk, cta_group::2 pairing, the
// 3-pass form on a 96-column tile, and the layer's MMA form (3-pass or concatenated); BN = 0: N is not a multiple of 32
struct TcTilePlan { int BN, BK; bool pair, p3, three_pass; };
TcTilePlan voc_tc_plan_tile(int N, int K, int ntaps, int M, int B, int sms, int flags);
void voc_tc_clear_cache();

// ------------------------------------------------------------------------------------
// Fused residual unit (ru_fused.cu):  one kernel per unit
//   T  = Snake2( conv7_dil(A) + b7 )            A = Snake1(x), the split-fp16 operand the previous layer wrote
//   x' = x + conv1(T) + b1                      T never leaves the SM (split-fp16 operand in shared memory)
//   Y  = x' (float32, optional)                 S = Snake_next(x') (split-fp16 operand of the next layer)
// All tensors channels-last and dense: [B][L][C].  Bit-identical to the two tap-GEMM launches it replaces.
// ------------------------------------------------------------------------------------
struct RuFusedParams {
    const __half* A_hi;  const __half* A_lo;  int L;  int C;  int B;  int dil;  int ksz;
    int a_halo;          // rows of carried history stored before row 0 of A (streaming decode, B = 1); 0 = causal zeros
    const __half* W7tc;  long long w7_plane;  float w7scale;  const float* bias7;
    const float* sn2_a;  const float* sn2_invb;
    const __half* W1tc;  long long w1_plane;  float w1scale;  const float* bias1;
    const float* R;  float* Y;
    __half* S_hi;  __half* S_lo;  const float* snn_a;  const float* snn_invb;
    // the output head folded into the last unit (C = 96): head_w [head_taps][C]; head_part [B][2 column parts][head_taps][L]
    // receives per-tap partial sums of w * Snake_next(x') instead of S (S_hi / S_lo / Y null)
    const float* head_w;  float* head_part;  int head_taps;
};
bool voc_ru_fused_eligible(const RuFusedParams& p);
// shared-memory plan of the fused unit: {halo box rows, halo stages, weight stages, T overlays the halo ring, bytes}
bool voc_ru_fused_plan(int C, int ksz, int dil, int* out5);
cudaError_t voc_launch_ru_fused(const RuFusedParams& p, cudaStream_t st, int num_sms, int flags);
// flags: bit 0 no tap reuse, bit 1 / bit 2 force 64- / 32-wide K chunks, bit 3 run-time (generic) epilogue only,
// bit 4 no double-length head segments, bit 6 / bit 7 always / never cta_group::2 pairs, bit 8 / bit 9 always the widest /
// the narrowest column tile of the layer's family (default: chosen per launch from the number of tiles, tc_gemm.cu),
// bits 16.. MMAs into the main accumulator per round-to-nearest flush (default 24)
enum { VOC_TC_NO_REUSE = 1, VOC_TC_BK64 = 2, VOC_TC_BK32 = 4, VOC_TC_GENERIC_EPI = 8, VOC_TC_NO_SEG_HEAD = 16,
       VOC_TC_NO_PIPE = 32 /* fused residual unit: simple order of work */, VOC_TC_FORCE_PAIR = 64, VOC_TC_NO_PAIR = 128,
       VOC_TC_FIXED_TILE = 256, VOC_TC_SMALL_TILE = 512 };

// host-side launch wrappers (simt_kernels.cu)
cudaError_t voc_launch_tapgemm_simt(const TapGemmParams& p, cudaStream_t st);
cudaError_t voc_launch_rvq_gather(const long long* codes, int n_frames, int frames_per_win, int win_step,
                                  int n_windows, int n_q, int codebook_size, const float* tables,
                                  int dim, VocAct out, int* err_flag, cudaStream_t st,
                                  const int* win_meta = nullptr);
cudaError_t voc_launch_rmsnorm(const float* x, const float* w, VocAct y, int rows, int C, float eps,
                               cudaStream_t st);
cudaError_t voc_launch_dwconv_ln(const float* x, const float* dw_w, const float* dw_b, const float* ln_w,
                                 const float* ln_b, VocAct y, int B, int L, int C, int ksz, float eps,
                                 cudaStream_t st, int halo = 0);
cudaError_t voc_launch_attention(const float* qkv, VocAct out, int B, int T, int heads, int head_dim,
                                 const float* rope_cos, const float* rope_sin, int window,
                                 cudaStream_t st);
cudaError_t voc_launch_attention_stream(const float* qkv, VocAct out, int T, int heads, int head_dim,
                                        const float* rope_cos, const float* rope_sin, int window, int kv_max,
                                        const int* pos_ptr, cudaStream_t st);
cudaError_t voc_launch_swiglu(const float* gu, VocAct out, long long rows, int inter, cudaStream_t st);
cudaError_t voc_launch_head(VocAct S, long long s_bstride, int L, int C, int ksz, const float* w,
                            float bias, float* out, long long o_bstride, int B, cudaStream_t st, int halo = 0);
cudaError_t voc_launch_head_finish(const float* part, int parts, int taps, int L, float bias, float* out,
                                   long long o_bstride, int B, cudaStream_t st);
cudaError_t voc_launch_stitch(const float* chunks, long long chunk_stride, const int* win_meta,
                              int n_windows, int ov, const float* fade_out, const float* fade_in,
                              float* out_f32, short* out_i16, int max_a_len, cudaStream_t st);
cudaError_t voc_launch_append_window(float* res, long long res_len, const float* chunk, int a_len, int ov,
                                     int blend, const float* fade_out, const float* fade_in,
                                     cudaStream_t st);
cudaError_t voc_launch_pcm16(const float* in, short* out, long long n, cudaStream_t st);
// split-fp16 -> float32 (debug captures of operand tensors)
cudaError_t voc_launch_unsplit(const __half* hi, const __half* lo, float* out, long long n, cudaStream_t st);
// two equally long device-to-device copies in one launch (d1 = nullptr: one)
cudaError_t voc_launch_copy_pair(void* d0, const void* s0, void* d1, const void* s1, size_t bytes, cudaStream_t st);
// range statistics of a split-fp16 operand tensor [B][rows][cols] (row stride ld, window stride bstride), added into out[6]
cudaError_t voc_launch_operand_stats(const __half* hi, const __half* lo, int B, long long rows, int cols, int ld,
                                     long long bstride, unsigned long long* out, cudaStream_t st);
