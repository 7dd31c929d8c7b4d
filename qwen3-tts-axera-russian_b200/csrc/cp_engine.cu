// Code predictor on the GPU: the 5-layer transformer that the reference runs as
// code_predictor_decode_step.onnx on ONNX Runtime, 17 sequential calls per codec frame
// (/root/reference/dual_npu/code_predictor_server.py:77-140; "86 % of per-token time", docs/ARCHITECTURE.md:95-107).
//
// This is the other roofline of the product: batch 1, sequential, every step a chain of matrix-VECTOR products that
// streams ~315 MB of float32 weights -- HBM/L2 bandwidth and, above all, the latency of 432 dependent phases per frame; no
// tensor cores.  So:
//   * float32 weights and arithmetic throughout (the reference is ORT FP32; logits feed a temperature-0.1 sampler);
//   * one warp per (output row, K segment), 128-bit evict-first weight loads, the input vector(s) staged in shared
//     memory, the surrounding element-wise work fused into the GEMV's prologue (RMSNorm) or epilogue (residual add,
//     SwiGLU); q/k-norm + rotary + cache append + attention over <= 17 positions in one small kernel;
//   * kernels chained by programmatic dependent launch, weight loads issued before the dependency wait;
//   * the whole predict() -- 2 prefill positions, 15 x (lm_head, top-k sample, embedding lookup, decode step) -- is ONE
//     CUDA graph of 430 launches replayed per frame, sampling included: a frame costs one launch and one 60-byte D2H;
//     cp_predict_batch carries B <= 8 independent streams through the same kernels (weights streamed once for all);
//   * opt-in: the frame as one persistent cooperative kernel (cp_frame_kernel), with its own phase profiler.
// Level 1 (cp_step / cp_logits) is the reference's _ort_step interface with the KV cache kept on the device.
#include "../../include/cp_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace {

struct CpCfg {
    int hidden = 1024, layers = 5, heads = 16, kv_heads = 8, head_dim = 128, inter = 3072, vocab = 2048, groups = 15;
    double rms_eps = 1e-6, rope_theta = 10000.0;
    int max_positions = 32;
    int qdim() const { return heads * head_dim; }
    int kvdim() const { return kv_heads * head_dim; }
};

// ---- tiny flat-JSON number reader (keys of CPConfig.to_json) ----
static bool json_num(const std::string& js, const char* key, double& v) {
    const std::string k = std::string("\"") + key + "\"";
    size_t p = js.find(k);
    if (p == std::string::npos) return false;
    p = js.find(':', p + k.size());
    if (p == std::string::npos) return false;
    char* e = nullptr;
    const double d = strtod(js.c_str() + p + 1, &e);
    if (e == js.c_str() + p + 1) return false;
    v = d;
    return true;
}

// Programmatic dependent launch: every kernel of a step lets its successor start while it is still running
// (griddepcontrol.launch_dependents at the top) and touches activations only after griddepcontrol.wait, which returns
// once the predecessor grid has completed and its writes are visible.  Weights are read-only for the life of the
// handle, so a GEMV issues its first batch of weight loads BEFORE the wait: the HBM stream of kernel N + 1 starts under
// the tail of kernel N instead of after a launch gap.
__device__ __forceinline__ void cp_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void cp_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int CP_MAX_S = 2;          // tokens per step call: 1, or the reference's 2-token batch prefill
constexpr int CP_MAX_B = 8;          // independent streams per cp_predict_batch

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------
// y[s] = w * x[s] * rsqrt(mean(x[s]^2) + eps)            (sibling Qwen3OmniMoeRMSNorm)
// Only the level-1 interface launches this (hidden_out of cp_step); inside a step the norms are GEMV prologues.
__global__ void cp_rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H, float eps) {
    const int s = blockIdx.x;
    const float* xr = x + (size_t)s * H;
    cp_pdl_launch_dependents();
    cp_pdl_wait();
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) ss += xr[i] * xr[i];
    __shared__ float red[32];
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const float inv = rsqrtf(tot / (float)H + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) y[(size_t)s * H + i] = w[i] * (xr[i] * inv);
}

// Matrix-vector product(s), the op this whole path is made of.  Each launch streams one weight matrix (8 - 25 MB)
// exactly once; everything else is arranged so that enough 128-bit loads are in flight to cover HBM latency:
//   * the input vector(s) are staged in shared memory once per block -- with the preceding RMSNorm folded into the
//     staging when NORM (sum of squares of the S x K inputs is a 256-thread reduction, cheaper than a launch);
//   * KSPLIT warps share one output row (contiguous K segments, combined through shared memory), so that even the
//     1024-row projections put >= 4096 warps on the machine and every lane issues its loads back to back.
//   MODE 0: out[s][n] = W[n] . x[s]
//   MODE 1: out[s][n] = res[s][n] + W[n] . x[s]                  (residual add; out may be res)
//   MODE 2: out[s][n] = silu(W[n] . x[s]) * (W[N + n] . x[s])    (SwiGLU: gate rows, then up rows)
enum { CP_PLAIN = 0, CP_RESIDUAL = 1, CP_SWIGLU = 2 };
constexpr int CP_GEMV_WARPS = 16;
// SM = the most input vectors a launch may carry: 2 (a decode step or the 2-token prefill) or CP_MAX_B (one token of each of
// up to 8 independent streams: the weights are streamed once for all of them)
template <int MODE, int KSPLIT, bool NORM, int SM>
__global__ void __launch_bounds__(CP_GEMV_WARPS * 32, SM > 2 ? 1 : 2)
cp_gemv_kernel(const float* __restrict__ W, const float* __restrict__ x, int S, int N, int K, float* __restrict__ out,
               const float* res, const float* __restrict__ ln_w, float eps) {
    extern __shared__ float xs[];                      // [S][K]
    constexpr int NT = CP_GEMV_WARPS * 32;
    __shared__ float red[SM][CP_GEMV_WARPS];
    __shared__ float part[CP_GEMV_WARPS][2 * SM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cp_pdl_launch_dependents();
    constexpr int RPB = CP_GEMV_WARPS / KSPLIT;        // output rows per block
    const int n = blockIdx.x * RPB + warp / KSPLIT, ks = warp % KSPLIT;
    const int K4 = K >> 2, seg = (K4 + KSPLIT - 1) / KSPLIT;
    const int k0 = ks * seg, k1 = n < N ? min(K4, k0 + seg) : 0;
    const float4* w0 = reinterpret_cast<const float4*>(W + (size_t)min(n, N - 1) * K);
    const float4* w1 = reinterpret_cast<const float4*>(W + (size_t)(MODE == CP_SWIGLU ? N + min(n, N - 1) : 0) * K);
    // four independent 128-bit loads per lane per batch; slots past the segment load nothing and
    // multiply zeros (branch-free).  Streamed once: evict-first keeps the L2 for activations.
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = MODE == CP_SWIGLU ? 2 : 4;       // float4 slots per matrix per batch: four loads in flight per lane
    float4 wv[U], uv[U];
    auto load_batch = [&](int kb) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k4 = kb + 32 * u;
            wv[u] = k4 < k1 ? __ldcs(w0 + k4) : z4;
            uv[u] = (MODE == CP_SWIGLU && k4 < k1) ? __ldcs(w1 + k4) : z4;
        }
    };
    load_batch(k0 + lane);                             // in flight across the wait and the staging below
    cp_pdl_wait();
    if (NORM) {
        float ss[SM] = {};
        for (int i = tid; i < (K >> 2); i += NT)
            for (int s = 0; s < S; ++s) {
                const float4 v = reinterpret_cast<const float4*>(x + (size_t)s * K)[i];
                reinterpret_cast<float4*>(xs + (size_t)s * K)[i] = v;
                ss[s] = fmaf(v.x, v.x, ss[s]); ss[s] = fmaf(v.y, v.y, ss[s]); ss[s] = fmaf(v.z, v.z, ss[s]); ss[s] = fmaf(v.w, v.w, ss[s]);
            }
#pragma unroll
        for (int s = 0; s < SM; ++s) {
            for (int o = 16; o; o >>= 1) ss[s] += __shfl_xor_sync(0xffffffffu, ss[s], o);
            if (lane == 0) red[s][warp] = ss[s];
        }
        __syncthreads();
        for (int s = 0; s < S; ++s) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < CP_GEMV_WARPS; ++w) tot += red[s][w];
            const float inv = rsqrtf(tot / (float)K + eps);
            for (int i = tid; i < (K >> 2); i += NT) {
                const float4 g = reinterpret_cast<const float4*>(ln_w)[i];
                float4 v = reinterpret_cast<float4*>(xs + (size_t)s * K)[i];
                v = make_float4(g.x * (v.x * inv), g.y * (v.y * inv), g.z * (v.z * inv), g.w * (v.w * inv));
                reinterpret_cast<float4*>(xs + (size_t)s * K)[i] = v;
            }
        }
    } else {
        for (int i = tid; i < ((S * K) >> 2); i += NT) reinterpret_cast<float4*>(xs)[i] = reinterpret_cast<const float4*>(x)[i];
    }
    __syncthreads();
    float a0[SM] = {}, a1[SM] = {};
    for (int kb = k0 + lane; kb < k1; kb += 32 * U) {
        float4 cw[U], cu[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { cw[u] = wv[u]; cu[u] = uv[u]; }
        if (kb + 32 * U < k1) load_batch(kb + 32 * U);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int k4 = min(kb + 32 * u, K4 - 1);
#pragma unroll
            for (int s = 0; s < SM; ++s) {
                if (s < S) {
                    const float4 xv = *reinterpret_cast<const float4*>(xs + (size_t)s * K + 4 * k4);
                    a0[s] = fmaf(cw[u].x, xv.x, a0[s]); a0[s] = fmaf(cw[u].y, xv.y, a0[s]);
                    a0[s] = fmaf(cw[u].z, xv.z, a0[s]); a0[s] = fmaf(cw[u].w, xv.w, a0[s]);
                    if (MODE == CP_SWIGLU) {
                        a1[s] = fmaf(cu[u].x, xv.x, a1[s]); a1[s] = fmaf(cu[u].y, xv.y, a1[s]);
                        a1[s] = fmaf(cu[u].z, xv.z, a1[s]); a1[s] = fmaf(cu[u].w, xv.w, a1[s]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < SM; ++s) {
        for (int o = 16; o; o >>= 1) {
            a0[s] += __shfl_xor_sync(0xffffffffu, a0[s], o);
            if (MODE == CP_SWIGLU) a1[s] += __shfl_xor_sync(0xffffffffu, a1[s], o);
        }
    }
    if (KSPLIT > 1) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < SM; ++s) { part[warp][2 * s] = a0[s]; part[warp][2 * s + 1] = a1[s]; }
        }
        __syncthreads();
        if (ks == 0 && lane == 0) {
#pragma unroll
            for (int s = 0; s < SM; ++s) {
                a0[s] = 0.f; a1[s] = 0.f;
                for (int q = 0; q < KSPLIT; ++q) { a0[s] += part[warp + q][2 * s]; a1[s] += part[warp + q][2 * s + 1]; }
            }
        }
    }
    if (ks == 0 && lane == 0 && n < N) {
        for (int s = 0; s < S; ++s) {
            float v = a0[s];
            if (MODE == CP_RESIDUAL) v += res[(size_t)s * N + n];
            if (MODE == CP_SWIGLU) v = (v / (1.f + expf(-v))) * a1[s];
            out[(size_t)s * N + n] = v;
        }
    }
}

// q/k RMSNorm over head_dim, rotary embedding, KV-cache append and causal attention for the S tokens of one call
// (sibling :2352-2424).  One block of four warps per (query head, token of this call); a lane owns four consecutive
// head dimensions (head_dim = 4 * live lanes, a power of two <= 128), so the norms, the rotation partner
// (lane ^ live/2) and the dot products are warp shuffles -- no block barrier until the final merge.  Warp w takes key
// positions w, w + 4, ... with its own online softmax.  Positions < pos0 come from the cache; this call's keys are
// normalised and rotated in place by whichever warp meets them (the writer block of that KV head may not have run
// yet), and the block of the group's first query head appends them to the cache.
// qkv [S][q + 2 kv]; cache K/V [kv_heads][max_pos][hd].
__global__ void __launch_bounds__(128)
cp_attn_kernel(const float* __restrict__ qkv, int S, int pos0, int heads, int kv_heads, int hd, int max_pos,
               const float* __restrict__ qn, const float* __restrict__ kn, const float* __restrict__ rope_cos,
               const float* __restrict__ rope_sin, float* __restrict__ kc, float* __restrict__ vc, float eps,
               float* __restrict__ att, long long stream_stride) {
    __shared__ float sm_m[4], sm_l[4];
    __shared__ float4 sm_o[4][32];
    const int h = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int s = blockIdx.y;
    const int live_lanes = hd >> 2, half = live_lanes >> 1, H2 = hd >> 1;
    const bool live = lane < live_lanes;
    const int rep = heads / kv_heads, g = h / rep;
    const int qd = heads * hd, kvd = kv_heads * hd, ld = qd + 2 * kvd;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (stream_stride) {
        // batch mode (cp_predict_batch): row s is the one new token of independent stream s, with its own cache
        qkv += (size_t)s * ld; att += (size_t)s * qd; kc += (size_t)s * stream_stride; vc += (size_t)s * stream_stride;
        s = 0;
    }
    cp_pdl_launch_dependents();
    cp_pdl_wait();
    auto wsum = [](float v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
    auto ld4 = [&](const float* p) { return live ? *reinterpret_cast<const float4*>(p + 4 * lane) : zero4; };
    auto norm_rope = [&](float4 v, const float* w, int pos) -> float4 {
        const float inv = rsqrtf(wsum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) / (float)hd + eps);
        const float4 wv = ld4(w);
        v = make_float4(wv.x * (v.x * inv), wv.y * (v.y * inv), wv.z * (v.z * inv), wv.w * (v.w * inv));
        float4 r;                                       // rotate_half (sibling :816-820): (-x2, x1)
        r.x = __shfl_xor_sync(0xffffffffu, v.x, half); r.y = __shfl_xor_sync(0xffffffffu, v.y, half);
        r.z = __shfl_xor_sync(0xffffffffu, v.z, half); r.w = __shfl_xor_sync(0xffffffffu, v.w, half);
        if (!live) return zero4;
        const float sg = lane < half ? -1.f : 1.f;
        const int f = 4 * (lane < half ? lane : lane - half);
        const float4 c = *reinterpret_cast<const float4*>(rope_cos + (size_t)pos * H2 + f);
        const float4 sn = *reinterpret_cast<const float4*>(rope_sin + (size_t)pos * H2 + f);
        return make_float4(v.x * c.x + sg * r.x * sn.x, v.y * c.y + sg * r.y * sn.y, v.z * c.z + sg * r.z * sn.z,
                           v.w * c.w + sg * r.w * sn.w);
    };
    const float4 q = norm_rope(ld4(qkv + (size_t)s * ld + h * hd), qn, pos0 + s);
    const int P = pos0 + s + 1;                                // keys 0 .. pos0 + s
    const float scaling = rsqrtf((float)hd);
    float m = -INFINITY, l = 0.f;
    float4 o = zero4;
    for (int j = warp; j < P; j += 4) {
        float4 k4, v4;
        if (j < pos0) {
            k4 = ld4(kc + ((size_t)g * max_pos + j) * hd);
            v4 = ld4(vc + ((size_t)g * max_pos + j) * hd);
        } else {
            const int t = j - pos0;
            k4 = norm_rope(ld4(qkv + (size_t)t * ld + qd + g * hd), kn, j);
            v4 = ld4(qkv + (size_t)t * ld + qd + kvd + g * hd);
            if (t == s && h % rep == 0 && live) {
                *reinterpret_cast<float4*>(kc + ((size_t)g * max_pos + j) * hd + 4 * lane) = k4;
                *reinterpret_cast<float4*>(vc + ((size_t)g * max_pos + j) * hd + 4 * lane) = v4;
            }
        }
        const float sc = wsum(q.x * k4.x + q.y * k4.y + q.z * k4.z + q.w * k4.w) * scaling;
        const float mn = fmaxf(m, sc), a = expf(m - mn), p = expf(sc - mn);
        l = l * a + p;
        o = make_float4(o.x * a + p * v4.x, o.y * a + p * v4.y, o.z * a + p * v4.z, o.w * a + p * v4.w);
        m = mn;
    }
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    sm_o[warp][lane] = o;
    __syncthreads();
    if (warp == 0 && live) {
        float M = sm_m[0];
        for (int w = 1; w < 4; ++w) M = fmaxf(M, sm_m[w]);
        float L = 0.f;
        float4 acc = zero4;
        for (int w = 0; w < 4; ++w) {
            const float a = expf(sm_m[w] - M);             // warps that met no key: exp(-inf) = 0
            const float4 ow = sm_o[w][lane];
            L += sm_l[w] * a;
            acc = make_float4(acc.x + ow.x * a, acc.y + ow.y * a, acc.z + ow.z * a, acc.w + ow.w * a);
        }
        const float il = 1.f / L;
        *reinterpret_cast<float4*>(att + (size_t)s * qd + h * hd + 4 * lane) = make_float4(acc.x * il, acc.y * il, acc.z * il, acc.w * il);
    }
}

// Device parameters of one predict() (updated by a small H2D copy before the graph is launched)
struct CpSampleParams { float temperature; int top_k; unsigned long long seed; };

__device__ __forceinline__ unsigned long long cp_splitmix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// code_predictor_server.py:87-92 on the device: the top_k largest logits, softmax of (l - max) / max(T, 1e-6), one
// categorical draw.  One block of 256 threads, 16 logits per thread in registers (vocab <= 4096):
//   1. the k-th largest logit by bisection on the order-preserving integer image of the floats: 32 counting passes,
//      one barrier each (no k rounds of arg-max: k = 50 of those cost 60 us);
//   2. the candidates (> threshold, plus the lowest-indexed ties) compacted in index order by a block scan;
//   3. warp 0: softmax over the <= 64 candidates, inclusive scan, inverse-CDF pick with a counter-based generator
//      keyed by (seed, group).  The reference draws from NumPy's global generator, which no device code can
//      reproduce: same distribution, different stream.  top_k = 1 is arg-max with NumPy's lowest-index tie-break.
// Also writes the next step's input: out_embed = emb_table[code].
constexpr int CP_VPT = 16;
__device__ __forceinline__ unsigned cp_ordered(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
cp_sample_kernel(const float* __restrict__ logits, int vocab, const CpSampleParams* __restrict__ sp, int group, int groups,
                 const float* __restrict__ emb_table, int H, int* __restrict__ codes, float* __restrict__ out_embed) {
    __shared__ int cnt[34];
    __shared__ int wtot[2][8];
    __shared__ float cand_v[64];
    __shared__ int cand_i[64];
    __shared__ int chosen;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {                                                  // block b = stream b
        const int b = blockIdx.x;
        logits += (size_t)b * vocab; sp += b; codes += (size_t)b * groups;
        if (out_embed) out_embed += (size_t)b * H;
    }
    cp_pdl_launch_dependents();
    cp_pdl_wait();
    unsigned key[CP_VPT];
#pragma unroll
    for (int r = 0; r < CP_VPT; ++r) {
        const int i = tid * CP_VPT + r;
        key[r] = i < vocab ? cp_ordered(logits[i]) : 0u;          // 0 sorts below every float
    }
    if (tid < 34) cnt[tid] = 0;
    int K = sp->top_k; if (K > 64) K = 64; if (K > vocab) K = vocab; if (K < 1) K = 1;
    __syncthreads();
    unsigned thr = 0;                                             // largest t with |{key >= t}| >= K
    for (int bit = 31; bit >= 0; --bit) {
        const unsigned cand = thr | (1u << bit);
        int c = 0;
#pragma unroll
        for (int r = 0; r < CP_VPT; ++r) c += key[r] >= cand;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0 && c) atomicAdd(&cnt[bit], c);
        __syncthreads();
        if (cnt[bit] >= K) thr = cand;
    }
    // compaction: all keys > thr, then ties == thr in index order until K candidates
    int ngt = 0, neq = 0;
#pragma unroll
    for (int r = 0; r < CP_VPT; ++r) { ngt += key[r] > thr; neq += key[r] == thr; }
    int sgt = ngt, seq = neq;                                     // inclusive scans over the block, thread order
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, sgt, o), b = __shfl_up_sync(0xffffffffu, seq, o);
        if (lane >= o) { sgt += a; seq += b; }
    }
    if (lane == 31) { wtot[0][warp] = sgt; wtot[1][warp] = seq; }
    __syncthreads();
    int bgt = 0, beq = 0, tgt = 0;
    for (int w = 0; w < 8; ++w) { if (w < warp) { bgt += wtot[0][w]; beq += wtot[1][w]; } tgt += wtot[0][w]; }
    int pgt = bgt + sgt - ngt, peq = beq + seq - neq;             // exclusive prefixes of this thread
    const int take_eq = K - tgt;                                  // >= 1 by construction of thr
#pragma unroll
    for (int r = 0; r < CP_VPT; ++r) {
        const int i = tid * CP_VPT + r;
        if (key[r] > thr) { cand_v[pgt] = logits[i]; cand_i[pgt] = i; ++pgt; }
        else if (key[r] == thr) { if (peq < take_eq) { cand_v[tgt + peq] = logits[i]; cand_i[tgt + peq] = i; } ++peq; }
    }
    __syncthreads();
    if (warp == 0) {
        const float T = fmaxf(sp->temperature, 1e-6f);
        const float v0 = lane < K ? cand_v[lane] : -INFINITY, v1 = lane + 32 < K ? cand_v[lane + 32] : -INFINITY;
        float mx = fmaxf(v0, v1);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = lane < K ? expf((v0 - mx) / T) : 0.f, e1 = lane + 32 < K ? expf((v1 - mx) / T) : 0.f;
        float c0 = e0, c1 = e1;
        for (int o = 1; o < 32; o <<= 1) {
            const float a = __shfl_up_sync(0xffffffffu, c0, o), b = __shfl_up_sync(0xffffffffu, c1, o);
            if (lane >= o) { c0 += a; c1 += b; }
        }
        const float tot0 = __shfl_sync(0xffffffffu, c0, 31), tot1 = __shfl_sync(0xffffffffu, c1, 31);
        c1 += tot0;
        const unsigned long long bits = cp_splitmix(sp->seed * 0x100000001B3ull + (unsigned long long)group);
        const float u = (float)((bits >> 40) * (1.0 / 16777216.0)) * (tot0 + tot1);      // uniform in [0, sum)
        const unsigned b0 = __ballot_sync(0xffffffffu, lane < K && u < c0);
        const unsigned b1 = __ballot_sync(0xffffffffu, lane + 32 < K && u < c1);
        if (lane == 0) {
            const int pick = b0 ? __ffs(b0) - 1 : b1 ? 32 + __ffs(b1) - 1 : K - 1;
            chosen = cand_i[pick];
            codes[group] = chosen;
        }
    }
    __syncthreads();
    if (out_embed) for (int i = tid; i < H; i += blockDim.x) out_embed[i] = emb_table[(size_t)chosen * H + i];
}

// ------------------------------------------------------------------------------------------------------------
// The whole frame as ONE persistent kernel.
//
// A frame is 432 strictly dependent phases (16 steps x 5 layers x {qkv, attention, o_proj, gate/up, down}, 15 x {lm_head,
// sample}), each shorter than a kernel boundary.  One cooperative launch, one 512-thread block per SM (so a thread may
// hold 128 registers), phases separated by a split grid barrier (arrive ... wait), and
//   * every warp owns up to two (row, K-segment) units of each GEMV phase and loads them -- 16 x 128-bit per lane --
//     into registers right after it has ARRIVED at the barrier that ends the phase before: the HBM stream of phase
//     p + 1 runs under the barrier wait and the staging, not after them;
//   * one bulk L2 prefetch per warp and unit is issued two phases ahead, so units that do not fit the registers (a third
//     of the gate/up rows) and the register loads themselves come from L2;
//   * activations (<= 12 KB per phase) are re-read from L2 by every block after the barrier (ld.global.cg: other SMs
//     wrote them) and staged in shared memory; the RMSNorm is folded in (norm weight staged before the barrier, sum of
//     squares reduced by every warp for itself -- one block barrier per phase);
//   * attention runs on the last `heads` blocks (which own no o_proj rows), the sampler on warp 0 of block 0 with all
//     logits in registers (no block barrier inside the 32 bisection passes).
constexpr int CP_MAX_LAYERS = 8;
constexpr int CP_FRAME_THREADS = 512;
constexpr int CP_FRAME_WARPS = CP_FRAME_THREADS / 32;
constexpr int CP_FRAME_MAX_VOCAB = 2048;             // 4 logits per thread of the sampling block

struct CpFrameArgs {
    int H, Q, KV, I, hd, heads, kv_heads, vocab, groups, layers, max_pos;
    float eps;
    const float *ln1[CP_MAX_LAYERS], *wqkv[CP_MAX_LAYERS], *qn[CP_MAX_LAYERS], *kn[CP_MAX_LAYERS], *wo[CP_MAX_LAYERS], *ln2[CP_MAX_LAYERS],
        *wgu[CP_MAX_LAYERS], *wd[CP_MAX_LAYERS];
    const float* fnorm;
    const float* const* emb;           // device arrays of `groups` table pointers
    const float* const* head;
    const float *rope_cos, *rope_sin;
    float *kc, *vc;
    float *x, *qkv, *att, *act, *logits;
    const float *in_hidden, *in_embed;
    int* codes;
    const CpSampleParams* sp;
    unsigned* bar;                     // grid-barrier counter, zeroed before every launch
    int l2_prefetch;                   // CP_FRAME_L2PF=0 turns the two-phases-ahead L2 prefetch off (experiment switch)
    unsigned long long* prof;          // CP_FRAME_PROF=1: block 0's globaltimer before / after every barrier (else NULL)
};

__device__ __forceinline__ unsigned long long cp_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Split grid barrier over the (co-resident, cooperative) grid; `epoch` = arrivals expected so far.  The release is
// done while thread 0 has no loads outstanding (a fence waits for them), the weight preload goes between the halves.
__device__ __forceinline__ void cp_bar_arrive(unsigned* bar, unsigned& epoch, unsigned long long* prof) {
    __syncthreads();
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) prof[2 * (epoch / gridDim.x)] = cp_globaltimer();
    epoch += gridDim.x;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
}
__device__ __forceinline__ void cp_bar_wait(const unsigned* bar, unsigned epoch, unsigned long long* prof) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v < epoch && clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a lost block must not hang the box
        } while (v < epoch);
    }
    __syncthreads();
    if (prof && blockIdx.x == 0 && threadIdx.x == 0) prof[2 * (epoch / gridDim.x) - 1] = cp_globaltimer();
}

struct CpPhase {
    const float* W; const float* x; float* out; const float* ln; int N, K, ksplit;
};

// K-segments per row: the smallest power of two that leaves <= 256 float4 (8 per lane, one register batch) per segment
__device__ __forceinline__ int cp_pick_ksplit(int K) {
    int k = 1;
    while (k < CP_FRAME_WARPS && ((K >> 2) + k - 1) / k > 256) k *= 2;
    return k;
}

// one register batch (8 x 128-bit per lane) of unit (row `n` of matrix `mat`, K-segment [.., k1)) starting at float4 kb
__device__ __forceinline__ void cp_unit_load(const CpPhase& p, int row, int kb, int k1, float4 (&w)[8]) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* w0 = reinterpret_cast<const float4*>(p.W + (size_t)row * p.K);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int k4 = kb + 32 * u;
        w[u] = k4 < k1 ? __ldcs(w0 + k4) : z4;           // streamed once: evict-first
    }
}
struct CpItem { int n, ks, k0, k1; };
__device__ __forceinline__ CpItem cp_item(const CpPhase& p, int item) {
    CpItem it;
    it.n = item / p.ksplit; it.ks = item - it.n * p.ksplit;
    const int K4 = p.K >> 2, seg = (K4 + p.ksplit - 1) / p.ksplit;
    it.k0 = it.ks * seg;
    it.k1 = it.n < p.N ? min(K4, it.k0 + seg) : 0;
    return it;
}
// registers for the phase about to start: plain phases -- items gw (wa) and gw + total_warps (wb);
// SwiGLU -- the gate (wa) and up (wb) rows of item gw
template <bool SWIGLU>
__device__ __forceinline__ void cp_phase_preload(const CpPhase& p, float4 (&wa)[8], float4 (&wb)[8]) {
    const int gw = blockIdx.x * CP_FRAME_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const CpItem a = cp_item(p, gw);
    cp_unit_load(p, min(a.n, p.N - 1), a.k0 + lane, a.k1, wa);
    if (SWIGLU) {
        cp_unit_load(p, p.N + min(a.n, p.N - 1), a.k0 + lane, a.k1, wb);
        // the rows this warp meets after its first (a third of the warps have a second gate/up pair): into L2 now, one
        // prefetch instruction per 4 KB, so that the in-phase loads are L2 hits
        for (int item = gw + gridDim.x * CP_FRAME_WARPS; item < p.N * p.ksplit; item += gridDim.x * CP_FRAME_WARPS) {
            const CpItem b = cp_item(p, item);
            for (int k4 = b.k0 + 8 * lane; k4 < b.k1; k4 += 256) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + (size_t)b.n * p.K + 4 * k4));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + (size_t)(p.N + b.n) * p.K + 4 * k4));
            }
        }
    } else {
        const CpItem b = cp_item(p, gw + gridDim.x * CP_FRAME_WARPS);
        cp_unit_load(p, min(b.n, p.N - 1), b.k0 + lane, b.k1, wb);
    }
}

// L2 prefetch of this warp's units of a LATER phase: one bulk-prefetch instruction per unit, no registers held
__device__ __forceinline__ void cp_phase_l2_prefetch(const CpPhase& p, bool swiglu) {
    if ((threadIdx.x & 31) != 0) return;
    const int total_warps = gridDim.x * CP_FRAME_WARPS;
    for (int item = blockIdx.x * CP_FRAME_WARPS + (threadIdx.x >> 5); item < p.N * p.ksplit; item += total_warps) {
        const CpItem it = cp_item(p, item);
        if (it.k1 <= it.k0) continue;
        const unsigned bytes = (unsigned)(it.k1 - it.k0) * 16u;
        const float* a = p.W + (size_t)it.n * p.K + 4 * it.k0;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
        if (swiglu) {
            const float* b = p.W + (size_t)(p.N + it.n) * p.K + 4 * it.k0;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b), "r"(bytes) : "memory");
        }
    }
}

// the RMSNorm weight of the NEXT normalising phase, staged before the barrier (it is constant data)
__device__ __forceinline__ void cp_stage_ln(const float* ln, int K, float* lns) {
    for (int i = threadIdx.x; i < K; i += CP_FRAME_THREADS) lns[i] = ln[i];
}

// dot products of one unit against the staged input: acc += sum_k w[k] * (ln[k] *) x[k] over the unit's segment;
// `w` holds the first batch, later batches (segments > 256 float4: not in the production shape) are loaded here
template <bool NORM>
__device__ __forceinline__ float cp_unit_dot(const CpPhase& p, int row, const CpItem& it, float4 (&w)[8], const float* xs, const float* lns) {
    const int lane = threadIdx.x & 31, K4 = p.K >> 2;
    float acc = 0.f;
    for (int kb = it.k0 + lane; kb < it.k1; kb += 256) {
        if (kb != it.k0 + lane) cp_unit_load(p, row, kb, it.k1, w);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k4 = min(kb + 32 * u, K4 - 1);
            float4 xv = reinterpret_cast<const float4*>(xs)[k4];
            if (NORM) {
                const float4 g = reinterpret_cast<const float4*>(lns)[k4];
                xv = make_float4(g.x * xv.x, g.y * xv.y, g.z * xv.z, g.w * xv.w);
            }
            acc = fmaf(w[u].x, xv.x, acc); acc = fmaf(w[u].y, xv.y, acc); acc = fmaf(w[u].z, xv.z, acc); acc = fmaf(w[u].w, xv.w, acc);
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// one GEMV phase for S = 1 (see cp_gemv_kernel for the modes); wa / wb hold the preloaded units (cp_phase_preload).
//   y[n] = rsqrt(mean(x^2) + eps) * sum_k W[n][k] * (ln[k] * x[k])        (NORM)
template <int MODE, bool NORM>
__device__ __forceinline__ void cp_phase_run(const CpPhase& p, float eps, float4 (&wa)[8], float4 (&wb)[8], float* xs, const float* lns,
                                             float (*part)[2], unsigned long long* fine = nullptr, bool have_regs = true) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, K4 = K >> 2;
    const bool stamp = fine && blockIdx.x == 0 && tid == 0;
    if (stamp) fine[0] = cp_globaltimer();
    for (int i = tid; i < K4; i += CP_FRAME_THREADS)
        reinterpret_cast<float4*>(xs)[i] = __ldcg(reinterpret_cast<const float4*>(p.x) + i);      // other SMs wrote it: L2
    __syncthreads();
    if (stamp) fine[1] = cp_globaltimer();
    float inv = 1.f;
    if (NORM) {
        float ss = 0.f;
        for (int k4 = lane; k4 < K4; k4 += 32) {
            const float4 v = reinterpret_cast<const float4*>(xs)[k4];
            ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        inv = rsqrtf(ss / (float)K + eps);
    }
    const int total_warps = gridDim.x * CP_FRAME_WARPS, items = p.N * p.ksplit;
    int j = have_regs ? 0 : 2;                          // without preloaded registers every item is loaded here
    for (int base = blockIdx.x * CP_FRAME_WARPS; base < items; base += total_warps, ++j) {      // block-uniform trip count
        const CpItem it = cp_item(p, base + warp);
        const int row = min(it.n, p.N - 1);
        float a0, a1 = 0.f;
        if (MODE == CP_SWIGLU) {
            if (j > 0) { cp_unit_load(p, row, it.k0 + lane, it.k1, wa); cp_unit_load(p, p.N + row, it.k0 + lane, it.k1, wb); }
            a0 = cp_unit_dot<NORM>(p, row, it, wa, xs, lns);
            a1 = cp_unit_dot<NORM>(p, p.N + row, it, wb, xs, lns);
        } else if (j == 1) {
            a0 = cp_unit_dot<NORM>(p, row, it, wb, xs, lns);
        } else {
            if (j > 1) cp_unit_load(p, row, it.k0 + lane, it.k1, wa);
            a0 = cp_unit_dot<NORM>(p, row, it, wa, xs, lns);
        }
        if (stamp && base == blockIdx.x * CP_FRAME_WARPS) fine[2] = cp_globaltimer();
        if (p.ksplit > 1) {                                  // block-uniform
            if (lane == 0) { part[warp][0] = a0; part[warp][1] = a1; }
            __syncthreads();
            if (it.ks == 0 && lane == 0) {
                a0 = 0.f; a1 = 0.f;
                for (int q = 0; q < p.ksplit; ++q) { a0 += part[warp + q][0]; a1 += part[warp + q][1]; }
            }
            __syncthreads();
        }
        if (it.ks == 0 && lane == 0 && it.n < p.N) {
            float v = a0 * inv;
            if (MODE == CP_RESIDUAL) v += __ldcg(p.out + it.n);
            if (MODE == CP_SWIGLU) v = (v / (1.f + expf(-v))) * (a1 * inv);
            p.out[it.n] = v;
        }
    }
}

// The cached keys / values of earlier positions do not depend on this step's qkv phase: the attention blocks load them
// while they wait at the barrier that ends it.  Warp w owns positions w and w + 16 (a frame has <= 32 positions).
__device__ __forceinline__ void cp_attn_prefetch(const CpFrameArgs& A, int h, int layer, int pos, float4 (&kpre)[2], float4 (&vpre)[2]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = A.hd, g = h / (A.heads / A.kv_heads);
    const bool live = lane < (hd >> 2);
    const size_t lc = (size_t)layer * A.kv_heads * A.max_pos * hd;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = warp + CP_FRAME_WARPS * t;
        const bool ok = live && j < pos;
        kpre[t] = ok ? __ldcg(reinterpret_cast<const float4*>(A.kc + lc + ((size_t)g * A.max_pos + j) * hd) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        vpre[t] = ok ? __ldcg(reinterpret_cast<const float4*>(A.vc + lc + ((size_t)g * A.max_pos + j) * hd) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// attention of one query head for the single token at position `pos` (cp_attn_kernel with S = 1, 16 warps over the keys)
__device__ __forceinline__ void cp_phase_attn(const CpFrameArgs& A, int h, int layer, int pos, const float4 (&kpre)[2], const float4 (&vpre)[2],
                                              float* sm_m, float* sm_l, float4 (*sm_o)[32]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = A.hd, live_lanes = hd >> 2, half = live_lanes >> 1, H2 = hd >> 1;
    const bool live = lane < live_lanes;
    const int rep = A.heads / A.kv_heads, g = h / rep;
    const size_t lc = (size_t)layer * A.kv_heads * A.max_pos * hd;
    float* kc = A.kc + lc;
    float* vc = A.vc + lc;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto wsum = [](float v) { for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; };
    auto ld4 = [&](const float* p) { return live ? __ldcg(reinterpret_cast<const float4*>(p) + lane) : zero4; };
    auto norm_rope = [&](float4 v, const float* w) -> float4 {
        const float inv = rsqrtf(wsum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) / (float)hd + A.eps);
        const float4 wv = live ? *reinterpret_cast<const float4*>(w + 4 * lane) : zero4;
        v = make_float4(wv.x * (v.x * inv), wv.y * (v.y * inv), wv.z * (v.z * inv), wv.w * (v.w * inv));
        float4 r;
        r.x = __shfl_xor_sync(0xffffffffu, v.x, half); r.y = __shfl_xor_sync(0xffffffffu, v.y, half);
        r.z = __shfl_xor_sync(0xffffffffu, v.z, half); r.w = __shfl_xor_sync(0xffffffffu, v.w, half);
        if (!live) return zero4;
        const float sg = lane < half ? -1.f : 1.f;
        const int f = 4 * (lane < half ? lane : lane - half);
        const float4 c = *reinterpret_cast<const float4*>(A.rope_cos + (size_t)pos * H2 + f);
        const float4 sn = *reinterpret_cast<const float4*>(A.rope_sin + (size_t)pos * H2 + f);
        return make_float4(v.x * c.x + sg * r.x * sn.x, v.y * c.y + sg * r.y * sn.y, v.z * c.z + sg * r.z * sn.z,
                           v.w * c.w + sg * r.w * sn.w);
    };
    const float4 q = norm_rope(ld4(A.qkv + h * hd), A.qn[layer]);
    const float scaling = rsqrtf((float)hd);
    float m = -INFINITY, l = 0.f;
    float4 o = zero4;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int j = warp + CP_FRAME_WARPS * t;
        if (j > pos) break;                                  // warp-uniform
        float4 k4, v4;
        if (j < pos) {
            k4 = kpre[t];
            v4 = vpre[t];
        } else {
            k4 = norm_rope(ld4(A.qkv + A.Q + g * hd), A.kn[layer]);
            v4 = ld4(A.qkv + A.Q + A.KV + g * hd);
            if (h % rep == 0 && live) {
                *reinterpret_cast<float4*>(kc + ((size_t)g * A.max_pos + j) * hd + 4 * lane) = k4;
                *reinterpret_cast<float4*>(vc + ((size_t)g * A.max_pos + j) * hd + 4 * lane) = v4;
            }
        }
        const float sc = wsum(q.x * k4.x + q.y * k4.y + q.z * k4.z + q.w * k4.w) * scaling;
        const float mn = fmaxf(m, sc), a = expf(m - mn), pe = expf(sc - mn);
        l = l * a + pe;
        o = make_float4(o.x * a + pe * v4.x, o.y * a + pe * v4.y, o.z * a + pe * v4.z, o.w * a + pe * v4.w);
        m = mn;
    }
    if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
    sm_o[warp][lane] = o;
    __syncthreads();
    if (warp == 0 && live) {
        float M = sm_m[0];
        for (int w = 1; w < CP_FRAME_WARPS; ++w) M = fmaxf(M, sm_m[w]);
        float L = 0.f;
        float4 acc = zero4;
        for (int w = 0; w < CP_FRAME_WARPS; ++w) {
            const float a = expf(sm_m[w] - M);
            const float4 ow = sm_o[w][lane];
            L += sm_l[w] * a;
            acc = make_float4(acc.x + ow.x * a, acc.y + ow.y * a, acc.z + ow.z * a, acc.w + ow.w * a);
        }
        const float il = 1.f / L;
        *reinterpret_cast<float4*>(A.att + h * hd + 4 * lane) = make_float4(acc.x * il, acc.y * il, acc.z * il, acc.w * il);
    }
}

// cp_sample_kernel's algorithm on block 0 of the frame kernel: 512 threads x 4 logits, two key bits per counting pass
// (three thresholds, three redux + at most three shared atomics per warp, one block barrier): 16 passes.
// Slot r of thread t holds logit 512 r + t: (slot, thread) order is index order.
__device__ __forceinline__ void cp_phase_sample(const CpFrameArgs& A, int group, bool more, int (*cnt)[4], unsigned long long (*wtot)[CP_FRAME_WARPS],
                                                float* cand_v, int* cand_i, int* chosen) {
    constexpr int VPT = CP_FRAME_MAX_VOCAB / CP_FRAME_THREADS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned key[VPT];
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
        const int i = r * CP_FRAME_THREADS + tid;
        key[r] = i < A.vocab ? cp_ordered(__ldcg(A.logits + i)) : 0u;
    }
    if (tid < 64) cnt[tid >> 2][tid & 3] = 0;
    int K = A.sp->top_k; if (K > 64) K = 64; if (K > A.vocab) K = A.vocab; if (K < 1) K = 1;
    __syncthreads();
    unsigned thr = 0;                                             // largest t with |{key >= t}| >= K
    for (int pass = 0; pass < 16; ++pass) {
        const int sh = 30 - 2 * pass;
        int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int r = 0; r < VPT; ++r) {
            c1 += key[r] >= (thr | (1u << sh)); c2 += key[r] >= (thr | (2u << sh)); c3 += key[r] >= (thr | (3u << sh));
        }
        c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2); c3 = __reduce_add_sync(0xffffffffu, c3);
        if (lane == 0) { if (c1) atomicAdd(&cnt[pass][1], c1); if (c2) atomicAdd(&cnt[pass][2], c2); if (c3) atomicAdd(&cnt[pass][3], c3); }
        __syncthreads();
        const unsigned d = cnt[pass][3] >= K ? 3u : cnt[pass][2] >= K ? 2u : cnt[pass][1] >= K ? 1u : 0u;
        thr |= d << sh;
    }
    // compaction in index order: per-slot counts ride in 16-bit fields of one 64-bit block scan
    unsigned long long g = 0, e = 0;
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
        g |= (unsigned long long)(key[r] > thr) << (16 * r);
        e |= (unsigned long long)(key[r] == thr) << (16 * r);
    }
    unsigned long long sg = g, se = e;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, sg, o), b = __shfl_up_sync(0xffffffffu, se, o);
        if (lane >= o) { sg += a; se += b; }
    }
    if (lane == 31) { wtot[0][warp] = sg; wtot[1][warp] = se; }
    __syncthreads();
    unsigned long long bg = 0, be = 0, tg = 0, te = 0;
    for (int w = 0; w < CP_FRAME_WARPS; ++w) {
        if (w < warp) { bg += wtot[0][w]; be += wtot[1][w]; }
        tg += wtot[0][w]; te += wtot[1][w];
    }
    const unsigned long long xg = bg + sg - g, xe = be + se - e;      // exclusive prefixes of this thread, per slot
    int tgt = 0;
#pragma unroll
    for (int r = 0; r < VPT; ++r) tgt += (int)((tg >> (16 * r)) & 0xffff);
    const int take_eq = K - tgt;
    int og = 0, oe = 0;                                               // candidates in earlier slots
#pragma unroll
    for (int r = 0; r < VPT; ++r) {
        const int i = r * CP_FRAME_THREADS + tid;
        const int pg = og + (int)((xg >> (16 * r)) & 0xffff), pe = oe + (int)((xe >> (16 * r)) & 0xffff);
        if (key[r] > thr) { cand_v[pg] = __ldcg(A.logits + i); cand_i[pg] = i; }
        else if (key[r] == thr && pe < take_eq) { cand_v[tgt + pe] = __ldcg(A.logits + i); cand_i[tgt + pe] = i; }
        og += (int)((tg >> (16 * r)) & 0xffff);
        oe += (int)((te >> (16 * r)) & 0xffff);
    }
    __syncthreads();
    if (warp == 0) {
        const float T = fmaxf(A.sp->temperature, 1e-6f);
        const float v0 = lane < K ? cand_v[lane] : -INFINITY, v1 = lane + 32 < K ? cand_v[lane + 32] : -INFINITY;
        float mx = fmaxf(v0, v1);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e0 = lane < K ? expf((v0 - mx) / T) : 0.f, e1 = lane + 32 < K ? expf((v1 - mx) / T) : 0.f;
        float c0 = e0, c1 = e1;
        for (int o = 1; o < 32; o <<= 1) {
            const float a = __shfl_up_sync(0xffffffffu, c0, o), b = __shfl_up_sync(0xffffffffu, c1, o);
            if (lane >= o) { c0 += a; c1 += b; }
        }
        const float tot0 = __shfl_sync(0xffffffffu, c0, 31), tot1 = __shfl_sync(0xffffffffu, c1, 31);
        c1 += tot0;
        const unsigned long long bits = cp_splitmix(A.sp->seed * 0x100000001B3ull + (unsigned long long)group);
        const float u = (float)((bits >> 40) * (1.0 / 16777216.0)) * (tot0 + tot1);
        const unsigned b0 = __ballot_sync(0xffffffffu, lane < K && u < c0);
        const unsigned b1 = __ballot_sync(0xffffffffu, lane + 32 < K && u < c1);
        if (lane == 0) {
            const int pick = b0 ? __ffs(b0) - 1 : b1 ? 32 + __ffs(b1) - 1 : K - 1;
            *chosen = cand_i[pick];
            A.codes[group] = cand_i[pick];
        }
    }
    __syncthreads();
    if (more) {
        const float* table = A.emb[group];
        const int code = *chosen;
        for (int i = tid; i < A.H; i += CP_FRAME_THREADS) A.x[i] = table[(size_t)code * A.H + i];
    }
}

// GEMV phases of a frame in execution order: (step, layer, kind) with kind 0..3 = qkv, o_proj, gate/up, down and
// kind 4 = the lm_head of group step - 1
struct CpCursor { int step, l, kind; };
__device__ __forceinline__ bool cp_cursor_next(const CpFrameArgs& A, CpCursor& c) {
    if (c.kind < 3) { ++c.kind; return true; }
    if (c.kind == 3) {
        if (c.l + 1 < A.layers) { ++c.l; c.kind = 0; return true; }
        if (c.step == 0) { c.step = 1; c.l = 0; c.kind = 0; return true; }
        c.kind = 4;
        return true;
    }
    if (c.step < A.groups) { ++c.step; c.l = 0; c.kind = 0; return true; }
    return false;
}
__device__ __forceinline__ CpPhase cp_phase_at(const CpFrameArgs& A, const CpCursor& c) {
    switch (c.kind) {
        case 0: return CpPhase{A.wqkv[c.l], A.x, A.qkv, A.ln1[c.l], A.Q + 2 * A.KV, A.H, cp_pick_ksplit(A.H)};
        case 1: return CpPhase{A.wo[c.l], A.att, A.x, nullptr, A.H, A.Q, cp_pick_ksplit(A.Q)};
        case 2: return CpPhase{A.wgu[c.l], A.x, A.act, A.ln2[c.l], A.I, A.H, cp_pick_ksplit(A.H)};
        case 3: return CpPhase{A.wd[c.l], A.act, A.x, nullptr, A.H, A.I, cp_pick_ksplit(A.I)};
        default: return CpPhase{A.head[c.step - 1], A.x, A.logits, A.fnorm, A.vocab, A.H, cp_pick_ksplit(A.H)};
    }
}

__global__ void __launch_bounds__(CP_FRAME_THREADS, 1)
cp_frame_kernel(const __grid_constant__ CpFrameArgs A) {
    extern __shared__ float dyn[];                       // xs [kmax] input vector | lns [H] norm weight of the next NORM phase
    __shared__ float part[CP_FRAME_WARPS][2];
    __shared__ float sm_m[CP_FRAME_WARPS], sm_l[CP_FRAME_WARPS];
    __shared__ float4 sm_o[CP_FRAME_WARPS][32];
    __shared__ int cnt[16][4];
    __shared__ unsigned long long wtot[2][CP_FRAME_WARPS];
    __shared__ float cand_v[64];
    __shared__ int cand_i[64];
    __shared__ int chosen;
    float* xs = dyn;
    float* lns = dyn + max(A.H, max(A.Q, A.I));
    unsigned epoch = 0;
    float4 wa[8], wb[8];
    CpCursor cur{0, 0, 0};                               // the GEMV phase about to run
    auto ahead = [&](int n, CpCursor& c) { c = cur; bool ok = true; for (int i = 0; i < n && ok; ++i) ok = cp_cursor_next(A, c); return ok; };
    auto prefetch2 = [&]() { CpCursor c; if (A.l2_prefetch && ahead(2, c)) cp_phase_l2_prefetch(cp_phase_at(A, c), c.kind == 2); };
    auto preload_at = [&](const CpCursor& c) {           // registers and norm weight of phase `c`
        const CpPhase p = cp_phase_at(A, c);
        if (c.kind == 2) cp_phase_preload<true>(p, wa, wb); else cp_phase_preload<false>(p, wa, wb);
        if (p.ln) cp_stage_ln(p.ln, p.K, lns);          // after an arrive: every warp is past its reads of the old weight
    };
    auto fine = [&]() { return A.prof ? A.prof + 1024 + 4 * (epoch / gridDim.x) : nullptr; };
    // end of GEMV phase `cur`: arrive, load the next phase's registers under the wait, advance
    auto finish = [&](bool preload) {
        cp_bar_arrive(A.bar, epoch, A.prof);
        CpCursor c;
        if (preload && ahead(1, c)) preload_at(c);
        cp_cursor_next(A, cur);
        cp_bar_wait(A.bar, epoch, A.prof);
    };

    if (blockIdx.x == 0) for (int i = threadIdx.x; i < A.H; i += CP_FRAME_THREADS) A.x[i] = A.in_hidden[i];
    cp_bar_arrive(A.bar, epoch, A.prof);
    preload_at(cur);
    { CpCursor c1; if (A.l2_prefetch && ahead(1, c1)) cp_phase_l2_prefetch(cp_phase_at(A, c1), false); }
    cp_bar_wait(A.bar, epoch, A.prof);
    for (int step = 0; step <= A.groups; ++step) {          // position `step`: 0 = hidden state, 1 = code_0, g + 1 = group g - 1's code
        for (int l = 0; l < A.layers; ++l) {
            prefetch2();
            cp_phase_run<CP_PLAIN, true>(cp_phase_at(A, cur), A.eps, wa, wb, xs, lns, part, fine());
            // o_proj registers are independent of the attention: every block but the attention blocks takes them now
            const int hb = (int)gridDim.x - 1 - (int)blockIdx.x;
            const bool attn_block = hb < A.heads;
            if (attn_block) {
                float4 kpre[2], vpre[2];
                cp_bar_arrive(A.bar, epoch, A.prof);
                cp_attn_prefetch(A, hb, l, step, kpre, vpre);
                cp_cursor_next(A, cur);
                cp_bar_wait(A.bar, epoch, A.prof);
                cp_phase_attn(A, hb, l, step, kpre, vpre, sm_m, sm_l, sm_o);
            } else {
                finish(true);
            }
            cp_bar_arrive(A.bar, epoch, A.prof);
            cp_bar_wait(A.bar, epoch, A.prof);
            prefetch2();
            cp_phase_run<CP_RESIDUAL, false>(cp_phase_at(A, cur), A.eps, wa, wb, xs, lns, part, fine(), !attn_block);
            finish(true);
            prefetch2();
            cp_phase_run<CP_SWIGLU, true>(cp_phase_at(A, cur), A.eps, wa, wb, xs, lns, part, fine());
            finish(true);
            prefetch2();
            cp_phase_run<CP_RESIDUAL, false>(cp_phase_at(A, cur), A.eps, wa, wb, xs, lns, part, fine());
            finish(true);
        }
        if (step == 0) {                                   // position 1 takes the embedding of code_0
            if (blockIdx.x == 0) for (int i = threadIdx.x; i < A.H; i += CP_FRAME_THREADS) A.x[i] = A.in_embed[i];
            cp_bar_arrive(A.bar, epoch, A.prof);
            cp_bar_wait(A.bar, epoch, A.prof);
            continue;
        }
        const int g = step - 1;
        const bool more = g + 1 < A.groups;
        prefetch2();
        cp_phase_run<CP_PLAIN, true>(cp_phase_at(A, cur), A.eps, wa, wb, xs, lns, part, fine());
        if (blockIdx.x == 0) {                             // the sampling block takes its registers after the draw
            finish(false);
            cp_phase_sample(A, g, more, cnt, wtot, cand_v, cand_i, &chosen);
            if (more) preload_at(cur);
        } else {
            finish(more);
        }
        if (more) { cp_bar_arrive(A.bar, epoch, A.prof); cp_bar_wait(A.bar, epoch, A.prof); }
    }
}

// ------------------------------------------------------------------------------------------------------------
struct CpEngine {
    CpCfg cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    bool finalized = false;
    std::map<std::string, std::vector<float>> raw;
    std::vector<void*> owned;
    struct Layer { float *ln1, *wqkv, *qn, *kn, *wo, *ln2, *wgu, *wd; };
    std::vector<Layer> L;
    float* fnorm = nullptr;
    std::vector<float*> emb, head;
    float *rope_cos = nullptr, *rope_sin = nullptr;
    float *kc = nullptr, *vc = nullptr;              // [stream][layers][kv_heads][max_pos][hd]; stream 0 is level 1's
    int cache_len = 0;                               // positions filled
    int last_S = 1;                                  // tokens of the last step (cp_logits reads the last one)
    long long graph_kernels = 0;
    // scratch
    float *d_x = nullptr, *d_qkv = nullptr, *d_att = nullptr, *d_act = nullptr, *d_out = nullptr, *d_logits = nullptr;
    float *d_in_hidden = nullptr, *d_in_embed = nullptr;
    int* d_codes = nullptr;
    CpSampleParams* d_sp = nullptr;
    cudaGraphExec_t graph = nullptr;                 // the frame for one stream
    std::map<int, cudaGraphExec_t> batch_graphs;     // ... and for B streams at once (built on first use)
    std::map<int, long long> batch_graph_kernels;
    long long launches = 0;
    // persistent frame kernel
    bool persistent = false;                         // cp_predict path: one cooperative kernel (default when it fits) or the graph
    int frame_grid = 0;
    size_t frame_smem = 0;
    CpFrameArgs fargs{};
    unsigned* d_bar = nullptr;

    ~CpEngine() {
        if (graph) cudaGraphExecDestroy(graph);
        for (auto& kv : batch_graphs) cudaGraphExecDestroy(kv.second);
        for (void* p : owned) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
    }
};

thread_local std::string g_cp_create_error;

#define CPK(expr)                                                                        \
    do { cudaError_t _e = (expr);                                                        \
         if (_e != cudaSuccess) {                                                        \
             E->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                \
             fprintf(stderr, "cp_b200: %s\n", E->err.c_str());                           \
             return CP_E_CUDA; } } while (0)

int cp_fail(CpEngine* E, int code, const std::string& m) {
    E->err = m;
    fprintf(stderr, "cp_b200: %s\n", m.c_str());
    return code;
}

template <class T>
T* cp_alloc(CpEngine* E, size_t n) {
    T* p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return nullptr;
    E->owned.push_back(p);
    return p;
}

float* cp_upload(CpEngine* E, const float* h, size_t n) {
    float* d = cp_alloc<float>(E, n);
    if (!d) return nullptr;
    if (cudaMemcpyAsync(d, h, n * sizeof(float), cudaMemcpyHostToDevice, E->stream) != cudaSuccess) return nullptr;
    return d;
}

const std::vector<float>* cp_raw(CpEngine* E, const std::string& name, size_t n) {
    auto it = E->raw.find(name);
    if (it == E->raw.end()) { E->err = "missing tensor " + name; return nullptr; }
    if (it->second.size() != n) { E->err = "tensor " + name + " has wrong size"; return nullptr; }
    return &it->second;
}

// every kernel of the path is launched with the programmatic-stream-serialization attribute (see cp_pdl_wait)
template <typename... Exp, typename... Act>
void cp_launch(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    static const bool pdl = !(getenv("CP_NO_PDL") && atoi(getenv("CP_NO_PDL")));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Act>(args)...);      // errors surface through cudaGetLastError
}

template <int MODE, int KSPLIT, bool NORM, int SM>
void cp_gemv_inst(cudaStream_t st, const float* W, const float* x, int S, int N, int K, float* out, const float* res,
                  const float* ln_w, float eps) {
    constexpr int RPB = CP_GEMV_WARPS / KSPLIT;
    const size_t smem = (size_t)S * K * sizeof(float);
    // dynamic + ~2 KB static must stay under the 48 KB default; only the batched form stages more (8 x 3072 floats = 96 KB)
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(cp_gemv_kernel<MODE, KSPLIT, NORM, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cp_launch(cp_gemv_kernel<MODE, KSPLIT, NORM, SM>, dim3((N + RPB - 1) / RPB), dim3(CP_GEMV_WARPS * 32), smem, st, W, x, S, N, K, out, res,
              ln_w, eps);
}
template <int MODE, bool NORM>
void cp_gemv(cudaStream_t st, const float* W, const float* x, int S, int N, int K, float* out, const float* res,
             const float* ln_w, float eps) {
    if (S <= CP_MAX_S) {
        if (N >= 3072) cp_gemv_inst<MODE, 2, NORM, CP_MAX_S>(st, W, x, S, N, K, out, res, ln_w, eps);
        else cp_gemv_inst<MODE, 4, NORM, CP_MAX_S>(st, W, x, S, N, K, out, res, ln_w, eps);
    } else {
        // batched streams: every block stages S x K inputs, so blocks are taller (16 or 8 rows) -- ~128-256 of them
        if (N >= 2048) cp_gemv_inst<MODE, 1, NORM, CP_MAX_B>(st, W, x, S, N, K, out, res, ln_w, eps);
        else cp_gemv_inst<MODE, 2, NORM, CP_MAX_B>(st, W, x, S, N, K, out, res, ln_w, eps);
    }
}

// one transformer pass over the S tokens in d_x (positions pos0 ..): the residual stream stays in d_x (un-normed; the
// final norm is the prologue of the lm_head GEMV, or cp_rmsnorm_kernel for level 1's hidden_out), caches grown.
// Five launches per layer.
int cp_forward(CpEngine* E, int S, int pos0, cudaStream_t st, bool batch = false) {
    const CpCfg& c = E->cfg;
    const int H = c.hidden, Q = c.qdim(), KV = c.kvdim(), I = c.inter, hd = c.head_dim;
    const float eps = (float)c.rms_eps;
    const size_t layer_cache = (size_t)c.kv_heads * c.max_positions * hd;
    for (int l = 0; l < c.layers; ++l) {
        auto& Ly = E->L[l];
        cp_gemv<CP_PLAIN, true>(st, Ly.wqkv, E->d_x, S, Q + 2 * KV, H, E->d_qkv, nullptr, Ly.ln1, eps);
        cp_launch(cp_attn_kernel, dim3(c.heads, S), dim3(128), 0, st, E->d_qkv, S, pos0, c.heads, c.kv_heads, hd, c.max_positions, Ly.qn,
                  Ly.kn, E->rope_cos, E->rope_sin, E->kc + l * layer_cache, E->vc + l * layer_cache, eps, E->d_att,
                  batch ? (long long)(c.layers * layer_cache) : 0ll);
        cp_gemv<CP_RESIDUAL, false>(st, Ly.wo, E->d_att, S, H, Q, E->d_x, E->d_x, nullptr, 0.f);
        cp_gemv<CP_SWIGLU, true>(st, Ly.wgu, E->d_x, S, I, H, E->d_act, nullptr, Ly.ln2, eps);
        cp_gemv<CP_RESIDUAL, false>(st, Ly.wd, E->d_act, S, H, I, E->d_x, E->d_x, nullptr, 0.f);
        E->launches += 5;
    }
    CPK(cudaGetLastError());
    return CP_OK;
}

// logits of `group` from row `row` of the residual stream: final RMSNorm folded into the lm_head GEMV
void cp_head(CpEngine* E, int group, int row, cudaStream_t st) {
    const CpCfg& c = E->cfg;
    cp_gemv<CP_PLAIN, true>(st, E->head[group], E->d_x + (size_t)row * c.hidden, 1, c.vocab, c.hidden, E->d_logits, nullptr, E->fnorm,
                            (float)c.rms_eps);
    E->launches += 1;
}

// the frame of B independent streams as one graph: every kernel carries B input vectors, so the weights are streamed
// once for all of them
int cp_build_graph(CpEngine* E, int B, cudaGraphExec_t* exec, long long* kernels) {
    // predict(): position 0 = hidden state, position 1 = embedding of code_0, then 15 x (lm_head, sample + embed, step)
    const CpCfg& c = E->cfg;
    const int H = c.hidden;
    cudaStream_t st = E->stream;
    CPK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    int rc = CP_OK;
    const long long before = E->launches;
    do {
        if (cudaMemcpyAsync(E->d_x, E->d_in_hidden, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = CP_E_CUDA; break; }
        if ((rc = cp_forward(E, B, 0, st, true))) break;
        if (cudaMemcpyAsync(E->d_x, E->d_in_embed, (size_t)B * H * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = CP_E_CUDA; break; }
        if ((rc = cp_forward(E, B, 1, st, true))) break;
        for (int g = 0; g < c.groups; ++g) {
            cp_gemv<CP_PLAIN, true>(st, E->head[g], E->d_x, B, c.vocab, H, E->d_logits, nullptr, E->fnorm, (float)c.rms_eps);
            const bool more = g + 1 < c.groups;
            cp_launch(cp_sample_kernel, dim3(B), dim3(256), 0, st, E->d_logits, c.vocab, E->d_sp, g, c.groups, more ? E->emb[g] : nullptr, H,
                      E->d_codes, more ? E->d_x : nullptr);
            E->launches += 2;
            if (more && (rc = cp_forward(E, B, g + 2, st, true))) break;
        }
    } while (0);
    *kernels = E->launches - before;
    E->launches = before;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }     // leave no stale error behind
    if (ce != cudaSuccess || !graph) return cp_fail(E, CP_E_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    CPK(cudaGraphInstantiate(exec, graph, 0));
    cudaGraphDestroy(graph);
    return CP_OK;
}

// CP_FRAME_PROF=1: where block 0 spends a frame -- work (barrier exit -> next barrier entry) and barrier time, by phase kind
void cp_print_prof(CpEngine* E) {
    static int printed = 0;
    if (++printed != 8) return;                       // one warmed-up frame
    const CpCfg& c = E->cfg;
    std::vector<unsigned long long> t(4096);
    if (cudaMemcpy(t.data(), E->fargs.prof, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    // barrier b: t[2b] = entry, t[2b+1] = exit.  Barrier 0 ends the copy-in; then per step 5 per layer (+1 copy / +2 head, sample)
    const char* names[] = {"qkv", "attn", "o_proj", "gate_up", "down", "copy_embed", "lm_head", "sample"};
    double work[8] = {0}, wait[8] = {0}, f_stage[8] = {0}, f_fma[8] = {0}, f_tail[8] = {0};
    int cnt[8] = {0};
    int b = 1;
    auto acc = [&](int kind) {
        if (2 * b + 1 >= 1024 || !t[2 * b + 1]) { ++b; return; }
        work[kind] += (double)(t[2 * b] - t[2 * b - 1]);
        wait[kind] += (double)(t[2 * b + 1] - t[2 * b]);
        const unsigned long long* f = &t[1024 + 4 * b];          // stamps of the GEMV that ran before barrier b
        if (f[0] && f[1] && f[2]) {
            f_stage[kind] += (double)(f[1] - f[0]); f_fma[kind] += (double)(f[2] - f[1]); f_tail[kind] += (double)(t[2 * b] - f[2]);
        }
        ++cnt[kind]; ++b;
    };
    for (int step = 0; step <= c.groups; ++step) {
        for (int l = 0; l < c.layers; ++l) for (int k = 0; k < 5; ++k) acc(k);
        if (step == 0) { acc(5); continue; }
        acc(6);
        if (step < c.groups) acc(7);
    }
    fprintf(stderr, "cp_frame_kernel, block 0, one frame: %d barriers, %.1f us total\n", b, (double)(t[2 * (b - 1)] - t[1]) / 1e3);
    for (int k = 0; k < 8; ++k)
        if (cnt[k]) fprintf(stderr, "  %-10s x%3d  work %6.2f us (stage %5.2f, fma+reduce %5.2f, tail %5.2f)  barrier %6.2f us\n", names[k], cnt[k],
                            work[k] / cnt[k] / 1e3, f_stage[k] / cnt[k] / 1e3, f_fma[k] / cnt[k] / 1e3, f_tail[k] / cnt[k] / 1e3, wait[k] / cnt[k] / 1e3);
}

// the persistent frame kernel: argument block, grid = one block per SM, cooperative launch required
int cp_setup_frame(CpEngine* E) {
    const CpCfg& c = E->cfg;
    E->persistent = false;
    if (c.layers > CP_MAX_LAYERS || c.vocab > CP_FRAME_MAX_VOCAB || c.groups + 1 > 2 * CP_FRAME_WARPS) return CP_OK;
    cudaDeviceProp prop{};
    CPK(cudaGetDeviceProperties(&prop, E->device));
    E->frame_smem = (size_t)(std::max(c.hidden, std::max(c.qdim(), c.inter)) + c.hidden) * sizeof(float);
    int per_sm = 0;
    CPK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cp_frame_kernel, CP_FRAME_THREADS, E->frame_smem));
    if (!prop.cooperativeLaunch || per_sm < 1 || prop.multiProcessorCount < c.heads) return CP_OK;
    E->frame_grid = prop.multiProcessorCount;
    std::vector<const float*> tabs(E->emb.begin(), E->emb.end()), heads(E->head.begin(), E->head.end());
    const float** d_emb = cp_alloc<const float*>(E, c.groups);
    const float** d_head = cp_alloc<const float*>(E, c.groups);
    E->d_bar = cp_alloc<unsigned>(E, 1);
    if (!d_emb || !d_head || !E->d_bar) return cp_fail(E, CP_E_NOMEM, "out of device memory");
    CPK(cudaMemcpy(d_emb, tabs.data(), c.groups * sizeof(float*), cudaMemcpyHostToDevice));
    CPK(cudaMemcpy(d_head, heads.data(), c.groups * sizeof(float*), cudaMemcpyHostToDevice));
    CpFrameArgs& A = E->fargs;
    A.H = c.hidden; A.Q = c.qdim(); A.KV = c.kvdim(); A.I = c.inter; A.hd = c.head_dim; A.heads = c.heads; A.kv_heads = c.kv_heads;
    A.vocab = c.vocab; A.groups = c.groups; A.layers = c.layers; A.max_pos = c.max_positions; A.eps = (float)c.rms_eps;
    for (int l = 0; l < c.layers; ++l) {
        const auto& Ly = E->L[l];
        A.ln1[l] = Ly.ln1; A.wqkv[l] = Ly.wqkv; A.qn[l] = Ly.qn; A.kn[l] = Ly.kn; A.wo[l] = Ly.wo; A.ln2[l] = Ly.ln2; A.wgu[l] = Ly.wgu; A.wd[l] = Ly.wd;
    }
    A.fnorm = E->fnorm; A.emb = d_emb; A.head = d_head; A.rope_cos = E->rope_cos; A.rope_sin = E->rope_sin; A.kc = E->kc; A.vc = E->vc;
    A.x = E->d_x; A.qkv = E->d_qkv; A.att = E->d_att; A.act = E->d_act; A.logits = E->d_logits;
    A.in_hidden = E->d_in_hidden; A.in_embed = E->d_in_embed; A.codes = E->d_codes; A.sp = E->d_sp; A.bar = E->d_bar;
    A.l2_prefetch = getenv("CP_FRAME_L2PF") && atoi(getenv("CP_FRAME_L2PF"));      // measured slower (UBLKPF issue cost): off
    A.prof = nullptr;
    if (getenv("CP_FRAME_PROF") && atoi(getenv("CP_FRAME_PROF"))) {
        A.prof = cp_alloc<unsigned long long>(E, 4096);
        if (A.prof) CPK(cudaMemset(A.prof, 0, 4096 * sizeof(unsigned long long)));
    }
    // Default: the graph.  Both paths measure 2.45 - 2.48 ms per frame on a B200 (profiles/r2_cp_*.txt); the graph's
    // small kernels interleave with a vocoder on the same GPU, the cooperative kernel holds every SM for the frame.
    const char* env = getenv("CP_PREDICT");
    E->persistent = env && !strcmp(env, "persistent");
    return CP_OK;
}

}  // namespace

extern "C" {

void* cp_create(const char* cfg_json, int device) {
    try {
        auto E = std::make_unique<CpEngine>();
        if (cfg_json && *cfg_json) {
            const std::string js = cfg_json;
            double v;
            auto geti = [&](const char* k, int& dst) { if (json_num(js, k, v)) dst = (int)v; };
            geti("hidden", E->cfg.hidden); geti("layers", E->cfg.layers); geti("heads", E->cfg.heads); geti("kv_heads", E->cfg.kv_heads);
            geti("head_dim", E->cfg.head_dim); geti("inter", E->cfg.inter); geti("vocab", E->cfg.vocab); geti("groups", E->cfg.groups);
            geti("max_positions", E->cfg.max_positions);
            if (json_num(js, "rms_eps", v)) E->cfg.rms_eps = v;
            if (json_num(js, "rope_theta", v)) E->cfg.rope_theta = v;
        }
        const CpCfg& c = E->cfg;
        auto bad = [&](const char* m) -> void* { g_cp_create_error = m; fprintf(stderr, "cp_create: %s\n", m); return nullptr; };
        if (c.hidden % 4 || c.inter % 4) return bad("hidden and inter must be multiples of 4");
        if (c.head_dim < 8 || c.head_dim > 128 || (c.head_dim & (c.head_dim - 1))) return bad("head_dim must be a power of two in [8, 128]");
        if (c.heads % c.kv_heads || c.layers < 1 || c.groups < 1 || c.vocab < 1 || c.vocab > 4096) return bad("bad head / layer / vocabulary configuration");
        if (c.max_positions < c.groups + 2) return bad("max_positions must cover groups + 2 positions");
        if ((size_t)CP_MAX_S * std::max(c.inter, std::max(c.hidden, c.qdim())) * 4 > 48 * 1024) return bad("layer too wide for the shared-memory staged GEMV");
        static_assert(CP_MAX_B >= CP_MAX_S, "activation buffers are sized for CP_MAX_B rows");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return bad("no usable CUDA device (there is no CPU fallback)");
        E->device = device;
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking) != cudaSuccess)
            return bad("cudaSetDevice / stream creation failed");
        return E.release();
    } catch (...) { g_cp_create_error = "cp_create: internal error"; return nullptr; }
}

void cp_destroy(void* h) {
    if (!h) return;
    CpEngine* E = (CpEngine*)h;
    cudaSetDevice(E->device);
    cudaDeviceSynchronize();
    delete E;
}

int cp_set_tensor(void* h, const char* name, const float* data, long long n) {
    CpEngine* E = (CpEngine*)h;
    if (!E || !name || !data || n <= 0) return CP_E_INVALID;
    if (E->finalized) return cp_fail(E, CP_E_STATE, "cp_set_tensor after cp_finalize");
    try { E->raw[name].assign(data, data + n); } catch (...) { return CP_E_NOMEM; }
    return CP_OK;
}

int cp_finalize(void* h) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    if (E->finalized) return cp_fail(E, CP_E_STATE, "already finalized");
    try {
        const CpCfg& c = E->cfg;
        CPK(cudaSetDevice(E->device));
        const int H = c.hidden, Q = c.qdim(), KV = c.kvdim(), I = c.inter, hd = c.head_dim;
#define CPREQ(x) do { if (!(x)) return cp_fail(E, E->err.rfind("missing", 0) == 0 || E->err.rfind("tensor", 0) == 0 ? CP_E_STATE : CP_E_CUDA, E->err.empty() ? std::string("cp_finalize failed: " #x) : E->err); } while (0)
        E->L.resize(c.layers);
        for (int l = 0; l < c.layers; ++l) {
            const std::string p = "layer_" + std::to_string(l) + "_";
            auto& Ly = E->L[l];
            auto up = [&](const char* nm, size_t n) -> float* { auto* v = cp_raw(E, p + nm, n); return v ? cp_upload(E, v->data(), n) : nullptr; };
            CPREQ(Ly.ln1 = up("input_ln", H)); CPREQ(Ly.ln2 = up("post_ln", H));
            CPREQ(Ly.qn = up("q_norm", hd)); CPREQ(Ly.kn = up("k_norm", hd));
            CPREQ(Ly.wo = up("o_proj", (size_t)H * Q)); CPREQ(Ly.wd = up("down_proj", (size_t)H * I));
            // q | k | v rows stacked: one GEMV; gate rows then up rows: one GEMV with the SwiGLU epilogue
            auto *wq = cp_raw(E, p + "q_proj", (size_t)Q * H), *wk = cp_raw(E, p + "k_proj", (size_t)KV * H), *wv = cp_raw(E, p + "v_proj", (size_t)KV * H);
            CPREQ(wq && wk && wv);
            std::vector<float> cat; cat.reserve((size_t)(Q + 2 * KV) * H);
            cat.insert(cat.end(), wq->begin(), wq->end()); cat.insert(cat.end(), wk->begin(), wk->end()); cat.insert(cat.end(), wv->begin(), wv->end());
            CPREQ(Ly.wqkv = cp_upload(E, cat.data(), cat.size()));
            CPK(cudaStreamSynchronize(E->stream));
            auto *wg = cp_raw(E, p + "gate_proj", (size_t)I * H), *wu = cp_raw(E, p + "up_proj", (size_t)I * H);
            CPREQ(wg && wu);
            cat.clear(); cat.insert(cat.end(), wg->begin(), wg->end()); cat.insert(cat.end(), wu->begin(), wu->end());
            CPREQ(Ly.wgu = cp_upload(E, cat.data(), cat.size()));
            CPK(cudaStreamSynchronize(E->stream));
        }
        { auto* v = cp_raw(E, "final_norm", H); CPREQ(v); CPREQ(E->fnorm = cp_upload(E, v->data(), H)); }
        E->emb.resize(c.groups); E->head.resize(c.groups);
        for (int g = 0; g < c.groups; ++g) {
            auto* e = cp_raw(E, "codec_emb_" + std::to_string(g), (size_t)c.vocab * H); CPREQ(e);
            auto* l = cp_raw(E, "lm_head_" + std::to_string(g), (size_t)c.vocab * H); CPREQ(l);
            CPREQ(E->emb[g] = cp_upload(E, e->data(), e->size()));
            CPREQ(E->head[g] = cp_upload(E, l->data(), l->size()));
        }
        // rotary table in float64, cast to float32: angle = pos * theta^(-2 i / hd)
        {
            const int H2 = hd / 2;
            std::vector<float> cs((size_t)c.max_positions * H2), sn((size_t)c.max_positions * H2);
            for (int p = 0; p < c.max_positions; ++p) for (int i = 0; i < H2; ++i) {
                const double inv = 1.0 / std::pow(c.rope_theta, (2.0 * i) / hd);
                cs[(size_t)p * H2 + i] = (float)std::cos(p * inv);
                sn[(size_t)p * H2 + i] = (float)std::sin(p * inv);
            }
            CPREQ(E->rope_cos = cp_upload(E, cs.data(), cs.size())); CPREQ(E->rope_sin = cp_upload(E, sn.data(), sn.size()));
            CPK(cudaStreamSynchronize(E->stream));
        }
        const size_t cache = (size_t)CP_MAX_B * c.layers * c.kv_heads * c.max_positions * hd;
        CPREQ(E->kc = cp_alloc<float>(E, cache)); CPREQ(E->vc = cp_alloc<float>(E, cache));
        constexpr int T = CP_MAX_B;                      // rows of every activation buffer (>= CP_MAX_S)
        CPREQ(E->d_x = cp_alloc<float>(E, (size_t)T * H));
        CPREQ(E->d_qkv = cp_alloc<float>(E, (size_t)T * (Q + 2 * KV))); CPREQ(E->d_att = cp_alloc<float>(E, (size_t)T * Q));
        CPREQ(E->d_act = cp_alloc<float>(E, (size_t)T * I)); CPREQ(E->d_out = cp_alloc<float>(E, (size_t)T * H));
        CPREQ(E->d_logits = cp_alloc<float>(E, (size_t)T * c.vocab)); CPREQ(E->d_in_hidden = cp_alloc<float>(E, (size_t)T * H));
        CPREQ(E->d_in_embed = cp_alloc<float>(E, (size_t)T * H));
        CPREQ(E->d_codes = cp_alloc<int>(E, (size_t)T * c.groups)); CPREQ(E->d_sp = cp_alloc<CpSampleParams>(E, T));
#undef CPREQ
        CPK(cudaStreamSynchronize(E->stream));
        E->raw.clear();
        E->finalized = true;
        if (int r = cp_build_graph(E, 1, &E->graph, &E->graph_kernels)) return r;
        CPK(cudaStreamSynchronize(E->stream));
        if (int r = cp_setup_frame(E)) return r;
        E->cache_len = 0;
        return CP_OK;
    } catch (const std::bad_alloc&) { return cp_fail(E, CP_E_NOMEM, "out of host memory"); }
    catch (...) { return cp_fail(E, CP_E_INVALID, "internal error"); }
}

int cp_reset(void* h) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    E->cache_len = 0;
    return CP_OK;
}

int cp_cache_len(void* h) { return h ? ((CpEngine*)h)->cache_len : CP_E_INVALID; }

int cp_step(void* h, const float* hidden_in, int S, int position, float* hidden_out) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    if (!E->finalized) return cp_fail(E, CP_E_STATE, "cp_finalize has not been called");
    if (!hidden_in || !hidden_out || S < 1 || S > CP_MAX_S) return cp_fail(E, CP_E_INVALID, "bad argument (1 or 2 tokens per step)");
    if (position != E->cache_len) return cp_fail(E, CP_E_INVALID, "position must equal the number of cached positions (cp_reset starts a frame)");
    if (position + S > E->cfg.max_positions) return cp_fail(E, CP_E_INVALID, "position beyond max_positions");
    CPK(cudaSetDevice(E->device));
    const int H = E->cfg.hidden;
    CPK(cudaMemcpyAsync(E->d_x, hidden_in, (size_t)S * H * 4, cudaMemcpyHostToDevice, E->stream));
    if (int r = cp_forward(E, S, position, E->stream)) return r;
    cp_launch(cp_rmsnorm_kernel, dim3(S), dim3(256), 0, E->stream, E->d_x, E->fnorm, E->d_out, H, (float)E->cfg.rms_eps);
    E->launches += 1;
    CPK(cudaGetLastError());
    CPK(cudaMemcpyAsync(hidden_out, E->d_out, (size_t)S * H * 4, cudaMemcpyDeviceToHost, E->stream));
    CPK(cudaStreamSynchronize(E->stream));
    E->cache_len = position + S;
    E->last_S = S;
    return CP_OK;
}

int cp_logits(void* h, int group, float* logits_out) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    if (!E->finalized) return cp_fail(E, CP_E_STATE, "cp_finalize has not been called");
    if (!logits_out || group < 0 || group >= E->cfg.groups) return cp_fail(E, CP_E_INVALID, "bad argument");
    if (E->cache_len < 1) return cp_fail(E, CP_E_STATE, "no step has run");
    CPK(cudaSetDevice(E->device));
    const CpCfg& c = E->cfg;
    cp_head(E, group, E->last_S - 1, E->stream);          // the LAST token of the last step
    CPK(cudaGetLastError());
    CPK(cudaMemcpyAsync(logits_out, E->d_logits, (size_t)c.vocab * 4, cudaMemcpyDeviceToHost, E->stream));
    CPK(cudaStreamSynchronize(E->stream));
    return CP_OK;
}

int cp_predict(void* h, const float* hidden_state, const float* code0_embed, float temperature, int top_k,
               unsigned long long seed, int* codes_out) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    if (!E->finalized || !E->graph) return cp_fail(E, CP_E_STATE, "cp_finalize has not been called");
    if (!hidden_state || !code0_embed || !codes_out || top_k < 1) return cp_fail(E, CP_E_INVALID, "bad argument");
    CPK(cudaSetDevice(E->device));
    const CpCfg& c = E->cfg;
    const CpSampleParams sp{temperature, top_k, seed};
    CPK(cudaMemcpyAsync(E->d_in_hidden, hidden_state, (size_t)c.hidden * 4, cudaMemcpyHostToDevice, E->stream));
    CPK(cudaMemcpyAsync(E->d_in_embed, code0_embed, (size_t)c.hidden * 4, cudaMemcpyHostToDevice, E->stream));
    CPK(cudaMemcpyAsync(E->d_sp, &sp, sizeof sp, cudaMemcpyHostToDevice, E->stream));
    if (E->persistent) {
        CPK(cudaMemsetAsync(E->d_bar, 0, sizeof(unsigned), E->stream));
        void* params[] = {(void*)&E->fargs};
        CPK(cudaLaunchCooperativeKernel((const void*)cp_frame_kernel, dim3(E->frame_grid), dim3(CP_FRAME_THREADS), params, E->frame_smem,
                                        E->stream));
        E->launches += 1;
    } else {
        CPK(cudaGraphLaunch(E->graph, E->stream));
        E->launches += E->graph_kernels;
    }
    CPK(cudaMemcpyAsync(codes_out, E->d_codes, (size_t)c.groups * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CPK(cudaStreamSynchronize(E->stream));
    if (E->persistent && E->fargs.prof) cp_print_prof(E);
    E->cache_len = c.groups + 1;
    E->last_S = 1;
    return CP_OK;
}

int cp_max_batch(void* h) { return h ? CP_MAX_B : CP_E_INVALID; }

int cp_predict_batch(void* h, int B, const float* hidden_states, const float* code0_embeds, float temperature, int top_k,
                     const unsigned long long* seeds, int* codes_out) {
    CpEngine* E = (CpEngine*)h;
    if (!E) return CP_E_INVALID;
    if (!E->finalized || !E->graph) return cp_fail(E, CP_E_STATE, "cp_finalize has not been called");
    if (!hidden_states || !code0_embeds || !codes_out || !seeds || top_k < 1 || B < 1 || B > CP_MAX_B)
        return cp_fail(E, CP_E_INVALID, "bad argument (1 .. cp_max_batch() streams)");
    CPK(cudaSetDevice(E->device));
    const CpCfg& c = E->cfg;
    if ((size_t)B * std::max(c.inter, std::max(c.hidden, c.qdim())) * 4 > 200 * 1024)
        return cp_fail(E, CP_E_INVALID, "batch too large for the shared-memory staged GEMV at this width");
    cudaGraphExec_t exec = B == 1 ? E->graph : nullptr;
    long long kernels = E->graph_kernels;
    if (B > 1) {
        auto it = E->batch_graphs.find(B);
        if (it == E->batch_graphs.end()) {
            cudaGraphExec_t g = nullptr;
            long long k = 0;
            if (int r = cp_build_graph(E, B, &g, &k)) return r;
            E->batch_graphs[B] = g;
            E->batch_graph_kernels[B] = k;
            it = E->batch_graphs.find(B);
        }
        exec = it->second;
        kernels = E->batch_graph_kernels[B];
    }
    CpSampleParams sp[CP_MAX_B];
    for (int b = 0; b < B; ++b) sp[b] = CpSampleParams{temperature, top_k, seeds[b]};
    CPK(cudaMemcpyAsync(E->d_in_hidden, hidden_states, (size_t)B * c.hidden * 4, cudaMemcpyHostToDevice, E->stream));
    CPK(cudaMemcpyAsync(E->d_in_embed, code0_embeds, (size_t)B * c.hidden * 4, cudaMemcpyHostToDevice, E->stream));
    CPK(cudaMemcpyAsync(E->d_sp, sp, (size_t)B * sizeof(CpSampleParams), cudaMemcpyHostToDevice, E->stream));
    CPK(cudaGraphLaunch(exec, E->stream));
    E->launches += kernels;
    CPK(cudaMemcpyAsync(codes_out, E->d_codes, (size_t)B * c.groups * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CPK(cudaStreamSynchronize(E->stream));
    E->cache_len = c.groups + 1;
    E->last_S = 1;
    return CP_OK;
}

int cp_hidden_size(void* h) { return h ? ((CpEngine*)h)->cfg.hidden : CP_E_INVALID; }
int cp_num_groups(void* h) { return h ? ((CpEngine*)h)->cfg.groups : CP_E_INVALID; }
int cp_vocab_size(void* h) { return h ? ((CpEngine*)h)->cfg.vocab : CP_E_INVALID; }
long long cp_launches(void* h) { return h ? ((CpEngine*)h)->launches : CP_E_INVALID; }
const char* cp_last_error(void* h) { return h ? ((CpEngine*)h)->err.c_str() : g_cp_create_error.c_str(); }
void* cp_stream(void* h) { return h ? (void*)((CpEngine*)h)->stream : nullptr; }

int cp_set_option(void* h, const char* key, const char* value) {
    CpEngine* E = (CpEngine*)h;
    if (!E || !key || !value) return CP_E_INVALID;
    if (!strcmp(key, "predict")) {
        if (!strcmp(value, "graph")) { E->persistent = false; return CP_OK; }
        if (!strcmp(value, "persistent")) {
            if (!E->frame_grid) return cp_fail(E, CP_E_STATE, "the persistent frame kernel is not available for this shape / device");
            E->persistent = true;
            return CP_OK;
        }
    }
    return cp_fail(E, CP_E_INVALID, std::string("unknown option ") + key + "=" + value);
}
const char* cp_predict_path(void* h) { return !h ? "" : ((CpEngine*)h)->persistent ? "persistent" : "graph"; }

}  // extern "C"
