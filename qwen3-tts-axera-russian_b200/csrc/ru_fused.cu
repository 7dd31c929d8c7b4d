// Fused residual unit for sm_100a (north_star subsystem 3): one kernel computes
//
//     T  = Snake2( conv7_dilated(A) + b7 )        A = Snake1(x), split-fp16 operand written by the previous layer
//     x' = x + conv1x1(T) + b1
//     Y  = x' (float32 residual stream, optional)  S = Snake_next(x') (split-fp16 operand of the next layer)
//
// where the two launches it replaces (tap-GEMM 7 taps, then tap-GEMM 1 tap + residual) moved T through HBM:
// per element of the unit 4 B written + 4 B read of the 24 B the pair moved.  Here T goes from the conv7
// accumulator (TMEM) through the epilogue warps' registers (bias, Snake, hi/lo split) into shared memory in the
// swizzled K-major layout the tensor core reads, and the 1x1 convolution is a second tcgen05 MMA chain on it.
//
// Structure (a specialisation of tapgemm_tc_kernel<BN, 64, TWO = true>, tc_gemm.cu, whose conventions it keeps):
// a cluster of two CTAs (one TPC) works on two consecutive 128-row time tiles with M = 256 cta_group::2 MMAs issued
// by the leader; BN = C (96 or 192) output channels = the whole channel dimension, so a tile's T rows are complete.
//   warp 0      TMA producer: per K chunk the 128-row A tile plus its halo of 6*dilation rows (once, all 7 taps
//               read it through shifted descriptors), the conv7 weights per (tap, chunk), then the 1x1 weights
//   warp 1      MMA issuer (leader CTA): conv7 in accumulation segments exactly as the unfused kernel, then waits
//               for both CTAs' T tiles and issues the 1x1 chain into the next TMEM buffer of the same ring
//   warps 2-3   idle (they fill warpgroup 0, the unit of setmaxnreg)
//   group A     4 CP warps, the conv7 side (4 TMEM lane quadrants x CP column parts): drain conv7 segments -> registers;
//               bias, Snake2, split -> the T tile in shared memory (st.shared, 128B swizzle by hand), fence.proxy.async,
//               arrive
//   group B     4 CP warps, the conv1 side: prefetch the residual rows into L2 a tile ahead; the 1x1 accumulator 16
//               columns at a time straight from TMEM; + bias + residual -> Y; Snake_next, split -> S
// Two groups because the two halves of the epilogue work are of equal weight and independent across tiles: A emits
// T(i+1) while B finishes tile i, each thread holding ONE tile's values (with a single group the conv7 sums of tile
// i+1 had to wait in registers through the final phase of tile i: no room at C = 192).
// The arithmetic -- segment schedule, pass order, k-step order, epilogue formulas -- is that of the two unfused
// launches, so the results are bit-identical to them (tests/test_gpu_ru_fused.py).
//
// Order of work (PIPE, C = 96): the issuing warp runs conv7 of tile i+1 BEFORE the 1x1 chain of tile i, so the tensor
// pipe works on tile i+1 while group A turns tile i's sums into T; a commit after each 1x1 chain (t_free) tells group A
// when T may be overwritten.
// C = 192: T (96 KB as two fp16 planes) does not fit next to a double-buffered 2 x 47 KB halo ring and a weight ring
// deep enough to cover the TMA latency (measured: with two weight stages the issuing warp waits for weights 27 % of
// the time), so there (ALIAS) T overlays the halo ring and the order is conv7(i), T(i), conv1(i): the issuing warp
// holds back the ring's last hand-backs of a tile until the 1x1 chain is issued, i.e. the producer's prefetch of the
// next tile's halo waits for it.  Group B's final phase still overlaps conv7(i+1).  Group A's 96 accumulators per
// thread need more than the 96 registers a 640-thread CTA starts with: warpgroup 0 gives registers up (setmaxnreg).
#include "voc_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <type_traits>

namespace {

constexpr int BM = 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 232448 - 1024 - 5120;   // opt-in maximum minus alignment slack and static smem
constexpr int FU_MAX_DEVICES = 64;

struct FuArgs {
    int M, B;
    int ntaps, a_min_off, a_box_rows, tap_row0, tap_step;
    int seg_iters, seg_head;
    int ring_k;                              // simple order: the first ring_k conv7 segments of a tile share one TMEM buffer
    int m_tiles, total_tiles;                // m_tiles counts tile PAIRS
    int k_chunks, kc_steps, kc_last;
    int SA, SB;
    float wscale7, wscale1;
    const float* bias7;  const float* sn2_a;  const float* sn2_invb;
    const float* bias1;  const float* snn_a;  const float* snn_invb;
    const float* R;  long long r_bstride;  int ldr;
    float* Y;        long long y_bstride;  int ldy;
    __half* S_hi;  __half* S_lo;  long long s_bstride;  int lds;
    const float* head_w;  float* head_part;   // HEAD: the output head's weights [taps][C]; per-tap partial sums out
};

// Bring-up profiling (-DVOC_TC_PROF, tools/ab_build.sh): cycles each role spends waiting, summed over CTAs
#ifdef VOC_TC_PROF
enum { RF_MMA_TOTAL, RF_MMA_W_A, RF_MMA_W_B, RF_MMA_W_ACC, RF_MMA_W_T, RF_EPI_TOTAL, RF_EPI_W_ACC7, RF_EPI_W_ACC1, RF_EPI_EMIT,
       RF_EPI_FINAL, RF_PROD_TOTAL, RF_PROD_W_A, RF_PROD_W_B, RF_CTAS, RF_TILES, RF_N };
__device__ unsigned long long g_ru_prof[RF_N];
#define PF_DECL(name) long long name = 0
#define PF_T0(t) const long long t = clock64()
#define PF_ACC(name, t) name += clock64() - (t)
#define PF_FLUSH(idx, name) atomicAdd(&g_ru_prof[idx], (unsigned long long)(name))
#else
#define PF_DECL(name)
#define PF_T0(t)
#define PF_ACC(name, t)
#define PF_FLUSH(idx, name)
#endif

// The hand-over of a T tile: every epilogue thread has written its row (st.shared) and issued fence.proxy.async, which
// makes those generic-proxy writes visible to the async proxy (the tensor core) of ITS OWN SM; the arrive on the
// leader's barrier then only has to be ordered after them in program order, for which the default semantics of
// mbarrier.arrive (.release at .cta scope) are enough -- each SM's tensor core reads its own CTA's shared memory.
// (A .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR and waits for the warp's outstanding global stores:
// 10 % of all stall samples, profiles/r2_ncu_ru_fused.txt.)  CUTLASS's ClusterBarrier::arrive(cta_id) is the same form.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Simple order of work (C = 192): which of the two TMEM accumulator buffers a use gets.  The 1x1 chain's accumulator is
// read by group B 16 columns at a time through its whole final phase (~20 k cycles: no registers to park it in), so a
// strictly alternating ring leaves the next tile's conv7 ONE buffer after its first segment and the issuing warp waits
// (14.5 k of 51 k cycles, profiles/r2_ru_prof.txt).  Instead the first K segments of a tile all use the buffer the
// previous chain is NOT in (group A drains a segment in a few hundred cycles), the rest alternate, and the chain takes
// the buffer the last segment is not in.  A static function of (previous chain's buffer, segment index): the issuing
// warp and both groups evaluate it independently.
__device__ __forceinline__ int fu_seg_buf(int cprev, int k, int K) {
    const int y = cprev ^ 1;
    return k < K ? y : (((k - K) & 1) ? y : cprev);
}

template <int BN, int CP, bool ALIAS, bool PIPE, bool HEAD = false>
__global__ void __launch_bounds__(128 + 256 * CP, 1)
ru_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW1b, const __grid_constant__ FuArgs a) {
    static_assert(BN == 96 || BN == 192, "the decoder blocks with C <= 192");
    static_assert((BN / CP) % 16 == 0, "the epilogue handles 16 operand columns (one 32-byte sector) at a time");
    static_assert(!(ALIAS && PIPE), "the pipelined order keeps the halo ring busy while T is live");
    constexpr int BK = 64;
    constexpr int EPI_WARPS = 4 * CP;                         // warps per epilogue group (A: conv7 side, B: conv1 side)
    constexpr int HN = BN / CP;                               // columns per epilogue thread
    // register budgets (setmaxnreg, multiples of 8): a 640-thread CTA starts at 96 per thread; group A at C = 192 holds
    // 96 accumulators per thread
#ifndef VOC_RU_REGS96
#define VOC_RU_REGS96 0          /* experiment: 0 = no rebalancing at C = 96; 1 = A 112 / B 96; 2 = A 120 / B 88 */
#endif
    constexpr bool REBALANCE = HN > 48 || VOC_RU_REGS96 != 0;
    // setmaxnreg.inc can only take what setmaxnreg.dec has released inside the CTA (the SM's never-allocated registers
    // are not in that pool: an inc that asks for more blocks for ever)
    constexpr int REGS_START = 96, REGS_WG0 = 56, REGS_A = HN > 48 ? 128 : (VOC_RU_REGS96 == 2 ? 120 : 112),
                  REGS_B = HN > 48 ? 80 : (VOC_RU_REGS96 == 2 ? 88 : 96);
    static_assert((REGS_START - REGS_WG0) * 128 + (REGS_START - REGS_B) * 32 * EPI_WARPS >= (REGS_A - REGS_START) * 32 * EPI_WARPS,
                  "register pool: released < requested");
    static_assert((128 + 64 * EPI_WARPS) * REGS_START <= 65536, "launch register budget");
    constexpr uint32_t ROWB = BK * 2;
    constexpr bool CAT = BN <= 128;                           // see tc_gemm.cu: A_hi x [B_hi; B_lo] as one N = 2 BN MMA
    constexpr int BROWS = BN / 2;
    constexpr uint32_t B_PLANE = CAT ? BN * ROWB : BROWS * ROWB;
    constexpr uint32_t B_STAGE = CAT ? (BN + BN / 2) * ROWB : 2 * B_PLANE;
    constexpr uint32_t ACC_COLS = CAT ? 2 * BN : BN;
    constexpr int NBUF = 2;
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
    constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
    // T tile: K chunks of 64 channels (128-byte rows, 128 rows, SWIZZLE_128B), two planes
    constexpr int NKC2 = (BN + 63) / 64;
    constexpr int KS2_LAST = (BN - (NKC2 - 1) * 64) / 16;     // k-steps of the last chunk (2 at C = 96)
    constexpr uint32_t A2_CHUNK = BM * ROWB;
    // C = 96: the second chunk holds 32 channels only -- 64-byte rows in the SWIZZLE_64B layout (16 KB less per tile,
    // which buys a fourth weight stage)
    constexpr bool T64 = (BN % 64) == 32;
    constexpr uint32_t A2_PLANE = T64 ? (NKC2 - 1) * A2_CHUNK + BM * 64 : NKC2 * A2_CHUNK;

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a_full[MAX_STAGES], bar_a_empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_b_full[MAX_STAGES], bar_b_empty[MAX_STAGES];
    // accumulator ring: one "empty" barrier per buffer (the issuing warp sees every use), but two "full" barriers --
    // conv7 segments complete on full7 (group A waits there), 1x1 chains on full1 (group B): a group that merely
    // stepped over the other group's uses of a buffer could not tell the phases of a shared barrier apart (one parity bit)
    __shared__ __align__(8) uint64_t bar_acc_full7[NBUF], bar_acc_full1[NBUF], bar_acc_empty[NBUF];
    __shared__ __align__(8) uint64_t bar_a2_full, bar_t_free;
    __shared__ uint32_t tmem_slot;
    // per-channel epilogue parameters: conv7 bias, Snake2 a / 1/b, conv1 bias, next Snake a / 1/b
    __shared__ __align__(16) float epi_par[6][BN];
    // HEAD (the last unit of the last block): instead of storing S = Snake_head(x') for the head kernel to read back
    // (4 B per element out, 4 B in), group B multiplies its columns of S with the head's 7 weight rows on the spot and
    // stores 7 partial sums per (row, column part): the output sample is their shifted sum (head_finish_kernel)
    constexpr int HEAD_TAPS = 7;
    __shared__ __align__(16) float head_par[HEAD ? HEAD_TAPS : 1][HEAD ? BN : 4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_plane = (uint32_t)a.a_box_rows * ROWB, a_stage = 2 * a_plane;
    const uint32_t a_ring = (uint32_t)a.SA * a_stage;
    const uint32_t smA = smem_base;
    const uint32_t smA2 = ALIAS ? smem_base : smem_base + a_ring;
    const uint32_t smB = ALIAS ? smem_base + (a_ring > 2 * A2_PLANE ? a_ring : 2 * A2_PLANE) : smem_base + a_ring + 2 * A2_PLANE;
    const int iters_per_tile = a.k_chunks * a.ntaps;
    const uint32_t rank = cluster_ctarank();
    const int walker = (int)(blockIdx.x >> 1), walkers = (int)(gridDim.x >> 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmB2);
        tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW1b);
        for (int i = 0; i < a.SA; ++i) { mbar_init(&bar_a_full[i], 1); mbar_init(&bar_a_empty[i], 1); }
        for (int i = 0; i < a.SB; ++i) { mbar_init(&bar_b_full[i], 1); mbar_init(&bar_b_empty[i], 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(&bar_acc_full7[i], 1); mbar_init(&bar_acc_full1[i], 1); mbar_init(&bar_acc_empty[i], 2 * EPI_WARPS); }
        mbar_init(&bar_a2_full, 2 * EPI_WARPS);
        mbar_init(&bar_t_free, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc2(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp < 4) {
        if constexpr (REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_WG0));
    }

    if (warp == 0) {
        // ================================ TMA producer ================================
        int sa = 0, pa = 0, sb = 0, pb = 0;
        PF_DECL(pf_w_a); PF_DECL(pf_w_b); PF_T0(pf_t0);
        const int n_my = walker < a.total_tiles ? (a.total_tiles - walker + walkers - 1) / walkers : 0;
        for (int s = 0; s < n_my + (PIPE ? 1 : 0); ++s) {
          if (!PIPE || s < n_my) {
            const int tile = walker + s * walkers;
            const int m_tile = 2 * (tile % a.m_tiles) + (int)rank, b = tile / a.m_tiles;
            const int row0 = m_tile * BM + a.a_min_off;
            const int n0 = CAT ? 0 : (int)rank * BROWS;
            for (int kc = 0; kc < a.k_chunks; ++kc) {
                for (int tap = 0; tap < a.ntaps; ++tap) {
                    if (tap == 0) {
                        { PF_T0(tw); mbar_wait(&bar_a_empty[sa], pa ^ 1); PF_ACC(pf_w_a, tw); }
                        if (elect_one()) {
                            if (rank == 0) mbar_expect_tx(&bar_a_full[sa], 2 * a_stage);
                            tma_load_4d_2sm(smA + sa * a_stage, &tmA, mapa_u32(&bar_a_full[sa], 0), kc * a.kc_steps * 16, row0, b, 0);
                        }
                        if (++sa == a.SA) { sa = 0; pa ^= 1; }
                    }
                    { PF_T0(tw); mbar_wait(&bar_b_empty[sb], pb ^ 1); PF_ACC(pf_w_b, tw); }
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&bar_b_full[sb], 2 * B_STAGE);
                        const uint32_t lb = mapa_u32(&bar_b_full[sb], 0);
                        if constexpr (CAT) {
                            tma_load_4d_2sm(smB + sb * B_STAGE, &tmB, lb, kc * a.kc_steps * 16, 0, tap, (int)rank);     // a whole plane
                            tma_load_4d_2sm(smB + sb * B_STAGE + B_PLANE, &tmB2, lb, kc * a.kc_steps * 16, (int)rank * (BN / 2), tap, 0);
                        } else {
                            tma_load_4d_2sm(smB + sb * B_STAGE, &tmB, lb, kc * a.kc_steps * 16, n0, tap, 0);
                        }
                    }
                    if (++sb == a.SB) { sb = 0; pb ^= 1; }
                }
            }
          }
          if (!PIPE || s > 0) {
            // the 1x1 weights, chunk by chunk, through the same ring
            const int n0 = CAT ? 0 : (int)rank * BROWS;
            for (int kc = 0; kc < NKC2; ++kc) {
                mbar_wait(&bar_b_empty[sb], pb ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&bar_b_full[sb], 2 * B_STAGE);
                    const uint32_t lb = mapa_u32(&bar_b_full[sb], 0);
                    if constexpr (CAT) {
                        tma_load_4d_2sm(smB + sb * B_STAGE, &tmW1, lb, kc * 64, 0, 0, (int)rank);
                        tma_load_4d_2sm(smB + sb * B_STAGE + B_PLANE, &tmW1b, lb, kc * 64, (int)rank * (BN / 2), 0, 0);
                    } else {
                        tma_load_4d_2sm(smB + sb * B_STAGE, &tmW1, lb, kc * 64, n0, 0, 0);
                    }
                }
                if (++sb == a.SB) { sb = 0; pb ^= 1; }
            }
          }
        }
#ifdef VOC_TC_PROF
        if (lane == 0) { PF_FLUSH(RF_PROD_TOTAL, clock64() - pf_t0); PF_FLUSH(RF_PROD_W_A, pf_w_a); PF_FLUSH(RF_PROD_W_B, pf_w_b); }
#endif
    } else if (warp == 1) {
        // ================================ MMA issuer (leader) =========================
        if (rank == 0) {
            const uint32_t a_full0 = opaque_u32(smem_u32(&bar_a_full[0])), a_empty0 = opaque_u32(smem_u32(&bar_a_empty[0]));
            const uint32_t b_full0 = opaque_u32(smem_u32(&bar_b_full[0])), b_empty0 = opaque_u32(smem_u32(&bar_b_empty[0]));
            const uint32_t acc_full0 = opaque_u32(smem_u32(&bar_acc_full7[0])), acc_empty0 = opaque_u32(smem_u32(&bar_acc_empty[0]));
            const uint32_t acc_full1_0 = opaque_u32(smem_u32(&bar_acc_full1[0]));
            const uint32_t a2_full = opaque_u32(smem_u32(&bar_a2_full)), t_free = opaque_u32(smem_u32(&bar_t_free));
            const uint32_t rt0 = tmem_base >> 24;
            auto reg = [&](uint32_t x) { return opaque_u32(x + rt0); };
            const int ks_regular = (int)reg((uint32_t)a.kc_steps), ks_last = (int)reg((uint32_t)a.kc_last);
            const int seg_iters = (int)reg((uint32_t)a.seg_iters), seg_head = (int)reg((uint32_t)a.seg_head);
            const int SA = (int)reg((uint32_t)a.SA), SB = (int)reg((uint32_t)a.SB);
            const int ipt = (int)reg((uint32_t)iters_per_tile);
            const int n_fills = (int)reg((uint32_t)a.k_chunks), n_inner = (int)reg((uint32_t)a.ntaps);
            const int first_short = (int)reg((uint32_t)(a.k_chunks - 1));
            const uint32_t a_desc0 = reg(smem_desc_lo(smA) + (uint32_t)a.tap_row0 * (ROWB >> 4));
            const uint32_t a_stage16 = reg(a_stage >> 4);
            const uint32_t tap_step16 = reg((uint32_t)(a.tap_step * (int)(ROWB >> 4)));
            const uint32_t b_desc0 = reg(smem_desc_lo(smB));
            const uint32_t a2_desc0 = reg(smem_desc_lo(smA2));
            int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pas = 0, p_a2 = 0;
            uint32_t ep = 0;                       // !PIPE: bit b = uses of accumulator buffer b so far (mod 2)
            int cprev = 1, nseg_m = 0;             // !PIPE: buffer of the previous tile's chain; segments per tile
            for (int rem = iters_per_tile; rem > 0; ++nseg_m) rem -= (nseg_m < a.seg_head ? 2 : 1) * a.seg_iters;
            const int ringK = a.ring_k < nseg_m ? a.ring_k : nseg_m;
            uint32_t b_lo = b_desc0;
            PF_DECL(pf_w_a); PF_DECL(pf_w_b); PF_DECL(pf_w_acc); PF_DECL(pf_w_t); PF_T0(pf_t0);
            const int n_my = walker < a.total_tiles ? (a.total_tiles - walker + walkers - 1) / walkers : 0;
            for (int s = 0; s < n_my + (PIPE ? 1 : 0); ++s) {
              uint32_t tmem_acc = 0;
              if (!PIPE || s < n_my) {
                uint32_t accum = 0;
                int seg_left = 0, iters_left = ipt, seg_idx = 0;
                for (int fill = 0; fill < n_fills; ++fill) {
                    { PF_T0(tw); mbar_spin_a(a_full0 + 8 * sa, pa); PF_ACC(pf_w_a, tw); }
                    uint32_t a_lo = a_desc0 + (uint32_t)sa * a_stage16;
                    const int ks_this = fill < first_short ? ks_regular : ks_last;
                    auto stages = [&](auto nks) {
                    constexpr int NKS = decltype(nks)::value;
                    for (int t = 0; t < n_inner; ++t) {
                        if (seg_left == 0) {
                            if constexpr (PIPE) {
                                PF_T0(tw); mbar_spin_a(acc_empty0 + 8 * as, pas ^ 1); PF_ACC(pf_w_acc, tw);
                            } else {
                                as = fu_seg_buf(cprev, seg_idx, ringK);
                                PF_T0(tw); mbar_spin_a(acc_empty0 + 8 * as, ((ep >> as) & 1u) ^ 1u); PF_ACC(pf_w_acc, tw);
                                ep ^= 1u << as;
                            }
                            tmem_acc = tmem_base + (uint32_t)as * ACC_COLS;
                            accum = 0;
                            const int want = seg_idx < seg_head ? 2 * seg_iters : seg_iters;
                            seg_left = iters_left < want ? iters_left : want;
                            ++seg_idx;
                        }
                        { PF_T0(tw); mbar_spin_a(b_full0 + 8 * sb, pb); PF_ACC(pf_w_b, tw); }
                        tc_fence_after();
                        --seg_left; --iters_left;
                        const bool last_of_seg = seg_left == 0;
                        if (elect_one()) {
                            if constexpr (CAT) {
#pragma unroll
                                for (int ks = 0; ks < NKS; ++ks) {
                                    if (ks == 0) mma2_f16_ss(tmem_acc, a_lo, b_lo, smem_desc_hi<BK>(), IDESC2, accum);
                                    else mma2_f16_ss_acc(tmem_acc, a_lo + ks * 2, b_lo + ks * 2, smem_desc_hi<BK>(), IDESC2);
                                    mma2_f16_ss_acc(tmem_acc + BN, a_lo + (a_plane >> 4) + ks * 2,
                                                    b_lo + (B_PLANE >> 4) + ks * 2, smem_desc_hi<BK>(), IDESC);
                                }
                            } else {
#pragma unroll
                                for (int pass = 0; pass < 3; ++pass) {
                                    const uint32_t ap = a_lo + (pass == 1 ? (a_plane >> 4) : 0u);
                                    const uint32_t bp = b_lo + (pass == 0 ? (B_PLANE >> 4) : 0u);
#pragma unroll
                                    for (int ks = 0; ks < NKS; ++ks) {
                                        if (pass == 0 && ks == 0) mma2_f16_ss(tmem_acc, ap, bp, smem_desc_hi<BK>(), IDESC, accum);
                                        else mma2_f16_ss_acc(tmem_acc, ap + ks * 2, bp + ks * 2, smem_desc_hi<BK>(), IDESC);
                                    }
                                }
                            }
                            mma2_commit_both_a(b_empty0 + 8 * sb);
                            if (last_of_seg) mma2_commit_both_a(acc_full0 + 8 * as);
                        }
                        __syncwarp();
                        accum = 1;
                        a_lo += tap_step16;
                        b_lo += B_STAGE >> 4;
                        if (++sb == SB) { sb = 0; pb ^= 1; b_lo = b_desc0; }
                        if constexpr (PIPE) { if (last_of_seg) { if (++as == NBUF) { as = 0; pas ^= 1; } } }
                    }
                    };
                    if (ks_this == 4) stages(std::integral_constant<int, 4>{});
                    else if (ks_this == 3) stages(std::integral_constant<int, 3>{});
                    else if (ks_this == 2) stages(std::integral_constant<int, 2>{});
                    else stages(std::integral_constant<int, 1>{});
                    // the A stage is free once everything issued so far has read it -- except, with T overlaying the
                    // ring, the last SA fills of a tile: their hand-back waits for the 1x1 chain below
                    if (!ALIAS || fill + SA < n_fills) {
                        if (elect_one()) mma2_commit_both_a(a_empty0 + 8 * sa);
                        __syncwarp();
                    }
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
              }
              if (!PIPE || s > 0) {
                // ---- the 1x1 convolution on the T tiles both CTAs' epilogue warps have written to shared memory
                { PF_T0(tw); mbar_spin_a(a2_full, (uint32_t)p_a2); PF_ACC(pf_w_t, tw); }
                p_a2 ^= 1;
                if constexpr (PIPE) {
                    PF_T0(tw); mbar_spin_a(acc_empty0 + 8 * as, pas ^ 1); PF_ACC(pf_w_acc, tw);
                } else {
                    as = fu_seg_buf(cprev, nseg_m - 1, ringK) ^ 1;      // not where the last segment is
                    cprev = as;
                    PF_T0(tw); mbar_spin_a(acc_empty0 + 8 * as, ((ep >> as) & 1u) ^ 1u); PF_ACC(pf_w_acc, tw);
                    ep ^= 1u << as;
                }
                tmem_acc = tmem_base + (uint32_t)as * ACC_COLS;
#pragma unroll
                for (int kc = 0; kc < NKC2; ++kc) {
                    constexpr int FULL = 4;
                    const int nks = kc == NKC2 - 1 ? KS2_LAST : FULL;
                    { PF_T0(tw); mbar_spin_a(b_full0 + 8 * sb, pb); PF_ACC(pf_w_b, tw); }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t t_lo = a2_desc0 + (uint32_t)kc * (A2_CHUNK >> 4);
                        if constexpr (CAT) {
                            // (the 32-channel last chunk of T is a SWIZZLE_64B tile: its own descriptor high word)
                            const uint32_t t_hi = (T64 && kc == NKC2 - 1) ? smem_desc_hi<32>() : smem_desc_hi<BK>();
#pragma unroll
                            for (int ks = 0; ks < FULL; ++ks) {
                                if (ks < nks) {
                                    if (kc == 0 && ks == 0) mma2_f16_ss(tmem_acc, t_lo, b_lo, smem_desc_hi<BK>(), IDESC2, 0u);
                                    else mma2_f16_ss_acc_hh(tmem_acc, t_lo + ks * 2, t_hi, b_lo + ks * 2, smem_desc_hi<BK>(), IDESC2);
                                    mma2_f16_ss_acc_hh(tmem_acc + BN, t_lo + (A2_PLANE >> 4) + ks * 2, t_hi,
                                                       b_lo + (B_PLANE >> 4) + ks * 2, smem_desc_hi<BK>(), IDESC);
                                }
                            }
                        } else {
#pragma unroll
                            for (int pass = 0; pass < 3; ++pass) {
                                const uint32_t ap = t_lo + (pass == 1 ? (A2_PLANE >> 4) : 0u);
                                const uint32_t bp = b_lo + (pass == 0 ? (B_PLANE >> 4) : 0u);
#pragma unroll
                                for (int ks = 0; ks < FULL; ++ks) {
                                    if (ks < nks) {
                                        if (kc == 0 && pass == 0 && ks == 0) mma2_f16_ss(tmem_acc, ap, bp, smem_desc_hi<BK>(), IDESC, 0u);
                                        else mma2_f16_ss_acc(tmem_acc, ap + ks * 2, bp + ks * 2, smem_desc_hi<BK>(), IDESC);
                                    }
                                }
                            }
                        }
                        mma2_commit_both_a(b_empty0 + 8 * sb);
                        if (kc == NKC2 - 1) { mma2_commit_both_a(acc_full1_0 + 8 * as); mma2_commit_both_a(t_free); }
                    }
                    __syncwarp();
                    b_lo += B_STAGE >> 4;
                    if (++sb == SB) { sb = 0; pb ^= 1; b_lo = b_desc0; }
                }
                if constexpr (PIPE) { if (++as == NBUF) { as = 0; pas ^= 1; } }
                if constexpr (ALIAS) {
                    // now the halo ring may be refilled: hand back the stages held since the conv7 loop
                    const int held = n_fills < SA ? n_fills : SA;
                    int hs = sa - held; if (hs < 0) hs += SA;
                    for (int i = 0; i < held; ++i) {
                        if (elect_one()) mma2_commit_both_a(a_empty0 + 8 * hs);
                        __syncwarp();
                        if (++hs == SA) hs = 0;
                    }
                }
              }
            }
#ifdef VOC_TC_PROF
            if (lane == 0) {
                PF_FLUSH(RF_MMA_TOTAL, clock64() - pf_t0); PF_FLUSH(RF_MMA_W_A, pf_w_a); PF_FLUSH(RF_MMA_W_B, pf_w_b);
                PF_FLUSH(RF_MMA_W_ACC, pf_w_acc); PF_FLUSH(RF_MMA_W_T, pf_w_t); PF_FLUSH(RF_CTAS, 1); PF_FLUSH(RF_TILES, n_my);
            }
#endif
        }
    } else if (warp >= 4) {
        // ================================ epilogue groups =============================
        const bool groupA = warp < 4 + EPI_WARPS;
        const int gw = groupA ? warp - 4 : warp - 4 - EPI_WARPS;     // warp within its group
        const int q = warp & 3;                       // the TMEM lane quadrant this warp can read
        const int h = gw >> 2;                        // which part of the channels
        int nseg = 0;
        for (int rem = iters_per_tile; rem > 0; ++nseg) rem -= (nseg < a.seg_head ? 2 : 1) * a.seg_iters;
        const int etid = threadIdx.x - 128;
        for (int i = etid; i < BN; i += 64 * EPI_WARPS) {
            epi_par[0][i] = __ldg(a.bias7 + i);
            epi_par[1][i] = __ldg(a.sn2_a + i);
            epi_par[2][i] = __ldg(a.sn2_invb + i);
            epi_par[3][i] = __ldg(a.bias1 + i);
            epi_par[4][i] = __ldg(a.snn_a + i);
            epi_par[5][i] = __ldg(a.snn_invb + i);
            if constexpr (HEAD) {
#pragma unroll
                for (int t = 0; t < HEAD_TAPS; ++t) head_par[t][i] = __ldg(a.head_w + t * BN + i);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(64 * EPI_WARPS) : "memory");
        // The accumulator ring is used in the issuing warp's order: per step the conv7 segments, then one 1x1 chain.
        // Each group waits for and hands back only its own buffers and steps over the other group's.
        int as = 0;                                   // PIPE: position in the alternating ring (all uses, both groups')
        uint32_t par = 0;                             // bit b: parity of this group's next wait on its full barrier of buffer b
        auto skip = [&](int n) { if constexpr (PIPE) { for (int i = 0; i < n; ++i) if (++as == NBUF) as = 0; } };
        int cprev = 1;                                // !PIPE: buffer of the previous tile's chain (fu_seg_buf)
        const int ringK = a.ring_k < nseg ? a.ring_k : nseg;
        uint32_t acc_empty_leader[NBUF];
#pragma unroll
        for (int i = 0; i < NBUF; ++i) acc_empty_leader[i] = mapa_u32(&bar_acc_empty[i], 0);
        const int n0 = h * HN;
        const uint32_t trow = (uint32_t)(q * 32 + lane);                        // row of the tile this thread owns
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0;
        const int n_my = walker < a.total_tiles ? (a.total_tiles - walker + walkers - 1) / walkers : 0;

        if (groupA) {
            // ---------------- group A: conv7 sums -> T tile ----------------
            if constexpr (REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_A));
            const uint32_t a2_full_leader = mapa_u32(&bar_a2_full, 0);
            const uint32_t t_row_addr = smA2 + trow * ROWB;
            const uint32_t sw = trow & 7u;                                      // 128B swizzle: 16-byte granule ^= row % 8
            int p_tf = 0;
            PF_DECL(pf_w7); PF_DECL(pf_wt); PF_DECL(pf_emit); PF_T0(pf_t0);
            for (int s = 0; s < n_my; ++s) {
                float acc[HN];
                // conv7 accumulation segments of tile s, added in FP32 with round-to-nearest as in the unfused kernel
                for (int seg = 0; seg < nseg; ++seg) {
                    if constexpr (!PIPE) as = fu_seg_buf(cprev, seg, ringK);
                    { PF_T0(tw); mbar_wait(&bar_acc_full7[as], (par >> as) & 1u); PF_ACC(pf_w7, tw); }
                    par ^= 1u << as;
                    tc_fence_after();
                    const uint32_t taddr = tlane + (uint32_t)as * ACC_COLS;
                    const bool first = seg == 0;
                    if constexpr (CAT) {
#pragma unroll
                        for (int c0 = 0; c0 < HN / 8; c0 += 2) {
                            uint32_t tm[2][8], tc[2][8];
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                if (c0 + c < HN / 8) { tmem_ld8(taddr + (c0 + c) * 8, tm[c]); tmem_ld8(taddr + BN + (c0 + c) * 8, tc[c]); }
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < 2; ++c)
                                if (c0 + c < HN / 8) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float v = __uint_as_float(tm[c][j]) + __uint_as_float(tc[c][j]);
                                        acc[(c0 + c) * 8 + j] = first ? v : acc[(c0 + c) * 8 + j] + v;
                                    }
                                }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc_empty_leader[as]);
                    } else {
                        constexpr int CH = 4;
#pragma unroll
                        for (int c0 = 0; c0 < HN / 8; c0 += CH) {
                            uint32_t tr[CH][8];
#pragma unroll
                            for (int c = 0; c < CH; ++c)
                                if (c0 + c < HN / 8) tmem_ld8(taddr + (c0 + c) * 8, tr[c]);
                            tmem_ld_wait();
                            if (c0 + CH >= HN / 8) {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive_cluster(acc_empty_leader[as]);
                            }
#pragma unroll
                            for (int c = 0; c < CH; ++c)
                                if (c0 + c < HN / 8) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float v = __uint_as_float(tr[c][j]);
                                        acc[(c0 + c) * 8 + j] = first ? v : acc[(c0 + c) * 8 + j] + v;
                                    }
                                }
                        }
                    }
                    skip(1);
                }
                // the 1x1 chain of the previous tile must have finished reading T.  Simple order: it was issued before
                // this tile's conv7, whose last segment has just been drained.  Pipelined order: it comes after.
                if constexpr (PIPE) {
                    if (s > 0) { PF_T0(tw); mbar_wait(&bar_t_free, (uint32_t)p_tf); p_tf ^= 1; PF_ACC(pf_wt, tw); }
                }
                PF_T0(te);
                // T = Snake2(conv7 + b7) -> split fp16 -> this thread's row of the T tile
#pragma unroll
                for (int g16 = 0; g16 < HN; g16 += 16) {
                    uint32_t hi16[8], lo16[8];
#pragma unroll
                    for (int g = g16; g < g16 + 16; g += 8) {
                        const int pc = n0 + g;
                        float v[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(&epi_par[0][pc]);
                        const float4 b1 = *reinterpret_cast<const float4*>(&epi_par[0][pc + 4]);
                        v[0] = acc[g + 0] * a.wscale7 + b0.x; v[1] = acc[g + 1] * a.wscale7 + b0.y;
                        v[2] = acc[g + 2] * a.wscale7 + b0.z; v[3] = acc[g + 3] * a.wscale7 + b0.w;
                        v[4] = acc[g + 4] * a.wscale7 + b1.x; v[5] = acc[g + 5] * a.wscale7 + b1.y;
                        v[6] = acc[g + 6] * a.wscale7 + b1.z; v[7] = acc[g + 7] * a.wscale7 + b1.w;
                        const float4 a0 = *reinterpret_cast<const float4*>(&epi_par[1][pc]);
                        const float4 a1 = *reinterpret_cast<const float4*>(&epi_par[1][pc + 4]);
                        const float4 i0 = *reinterpret_cast<const float4*>(&epi_par[2][pc]);
                        const float4 i1 = *reinterpret_cast<const float4*>(&epi_par[2][pc + 4]);
                        v[0] = voc_snake(v[0], a0.x, i0.x); v[1] = voc_snake(v[1], a0.y, i0.y);
                        v[2] = voc_snake(v[2], a0.z, i0.z); v[3] = voc_snake(v[3], a0.w, i0.w);
                        v[4] = voc_snake(v[4], a1.x, i1.x); v[5] = voc_snake(v[5], a1.y, i1.y);
                        v[6] = voc_snake(v[6], a1.z, i1.z); v[7] = voc_snake(v[7], a1.w, i1.w);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __half2 hh, ll;
                            voc_split2(v[2 * j], v[2 * j + 1], hh, ll);
                            hi16[(g - g16) / 2 + j] = *reinterpret_cast<uint32_t*>(&hh);
                            lo16[(g - g16) / 2 + j] = *reinterpret_cast<uint32_t*>(&ll);
                        }
                    }
                    // columns n0 + g16 .. + 15 of the tile: chunk (n0 + g16) / 64, 16-byte granules j0 and j0 + 1
                    const int col = n0 + g16;
                    const uint32_t j0 = (uint32_t)((col & 63) >> 3);
                    uint32_t p0, p1;
                    if (T64 && (col >> 6) == NKC2 - 1) {
                        // 64-byte rows, SWIZZLE_64B: granule ^= (row / 2) % 4
                        const uint32_t base = smA2 + (uint32_t)(NKC2 - 1) * A2_CHUNK + trow * 64u, sw64 = (trow >> 1) & 3u;
                        p0 = base + ((j0 ^ sw64) << 4); p1 = base + (((j0 + 1) ^ sw64) << 4);
                    } else {
                        const uint32_t base = t_row_addr + (uint32_t)(col >> 6) * A2_CHUNK;
                        p0 = base + (((j0) ^ sw) << 4); p1 = base + (((j0 + 1) ^ sw) << 4);
                    }
                    sts128(p0, hi16[0], hi16[1], hi16[2], hi16[3]);
                    sts128(p1, hi16[4], hi16[5], hi16[6], hi16[7]);
                    sts128(p0 + A2_PLANE, lo16[0], lo16[1], lo16[2], lo16[3]);
                    sts128(p1 + A2_PLANE, lo16[4], lo16[5], lo16[6], lo16[7]);
                }
                fence_proxy_async();                       // generic-proxy writes -> visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(a2_full_leader);
                PF_ACC(pf_emit, te);
                // step over the 1x1 chain's buffer that follows this tile's segments in the ring
                if constexpr (PIPE) { if (s > 0) skip(1); }
                else cprev = fu_seg_buf(cprev, nseg - 1, ringK) ^ 1;
            }
#ifdef VOC_TC_PROF
            if (warp == 4 && lane == 0) {
                PF_FLUSH(RF_EPI_TOTAL, clock64() - pf_t0); PF_FLUSH(RF_EPI_W_ACC7, pf_w7 + pf_wt); PF_FLUSH(RF_EPI_EMIT, pf_emit);
            }
#endif
        } else {
            // ---------------- group B: 1x1 accumulator -> the unit's outputs ----------------
            if constexpr (REBALANCE && REGS_B < REGS_START) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_B));
            PF_DECL(pf_w1); PF_DECL(pf_final); PF_T0(pf_t0);
            auto row_of = [&](int tile, int& m, int& b) { m = (2 * (tile % a.m_tiles) + (int)rank) * BM + (int)trow; b = tile / a.m_tiles; };
            auto prefetch_res = [&](int tile) {
                int m, b; row_of(tile, m, b);
                if (m < a.M) {
                    const float* r = a.R + (long long)b * a.r_bstride + (long long)m * a.ldr + n0;
#pragma unroll
                    for (int i = 0; i < HN; i += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(r + i));
                }
            };
            if (n_my > 0) prefetch_res(walker);
            if constexpr (PIPE) skip(nseg);                // step 0 holds only tile 0's conv7 segments
            for (int s = 0; s < n_my; ++s) {
                const int tile = walker + s * walkers;
                // ring order: simple -- segs(s), chain(s);  pipelined -- segs(s+1) (if any), chain(s)
                if constexpr (PIPE) { if (s + 1 < n_my) skip(nseg); }
                else { as = fu_seg_buf(cprev, nseg - 1, ringK) ^ 1; cprev = as; }
                int m, b; row_of(tile, m, b);
                const bool valid = m < a.M;
                const float* Rrow = a.R + (long long)b * a.r_bstride + (long long)(valid ? m : 0) * a.ldr + n0;
                float* Yrow = (a.Y && valid) ? a.Y + (long long)b * a.y_bstride + (long long)m * a.ldy + n0 : nullptr;
                const long long soff = (long long)b * a.s_bstride + (long long)m * a.lds + n0;
                float hp[HEAD ? HEAD_TAPS : 1];
#pragma unroll
                for (int t = 0; t < (HEAD ? HEAD_TAPS : 1); ++t) hp[t] = 0.f;
                float res[2][2][8];                       // two 16-column windows of the residual row, alternating
                if (valid) { ldg256(Rrow, res[0][0]); ldg256(Rrow + 8, res[0][1]); }
                if (s + 1 < n_my) prefetch_res(tile + walkers);   // the next tile's rows: HBM -> L2 while this one is finished
                { PF_T0(tw); mbar_wait(&bar_acc_full1[as], (par >> as) & 1u); PF_ACC(pf_w1, tw); }
                par ^= 1u << as;
                PF_T0(tf);
                tc_fence_after();
                const uint32_t taddr = tlane + (uint32_t)as * ACC_COLS;
#pragma unroll
                for (int g16 = 0; g16 < HN; g16 += 16) {
                    float v16[16];
                    if constexpr (CAT) {
                        uint32_t tm[2][8], tc[2][8];
                        tmem_ld8(taddr + g16, tm[0]); tmem_ld8(taddr + g16 + 8, tm[1]);
                        tmem_ld8(taddr + BN + g16, tc[0]); tmem_ld8(taddr + BN + g16 + 8, tc[1]);
                        if (valid && g16 + 16 < HN) { ldg256(Rrow + g16 + 16, res[((g16 >> 4) & 1) ^ 1][0]); ldg256(Rrow + g16 + 24, res[((g16 >> 4) & 1) ^ 1][1]); }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v16[j] = __uint_as_float(tm[0][j]) + __uint_as_float(tc[0][j]);
                            v16[8 + j] = __uint_as_float(tm[1][j]) + __uint_as_float(tc[1][j]);
                        }
                    } else {
                        uint32_t tm[2][8];
                        tmem_ld8(taddr + g16, tm[0]); tmem_ld8(taddr + g16 + 8, tm[1]);
                        if (valid && g16 + 16 < HN) { ldg256(Rrow + g16 + 16, res[((g16 >> 4) & 1) ^ 1][0]); ldg256(Rrow + g16 + 24, res[((g16 >> 4) & 1) ^ 1][1]); }
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) { v16[j] = __uint_as_float(tm[0][j]); v16[8 + j] = __uint_as_float(tm[1][j]); }
                    }
                    if (g16 + 16 >= HN) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc_empty_leader[as]);
                    }
                    if (valid) {
                        // 16 columns of the unit's output: x' = conv1 * 2^-e + b1 + x -> Y; Snake_next, split -> S
                        const float (&rs)[2][8] = res[(g16 >> 4) & 1];
                        uint32_t hi16[8], lo16[8];
#pragma unroll
                        for (int gg = 0; gg < 16; gg += 8) {
                            const int pc = n0 + g16 + gg;
                            float v[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = v16[gg + j] * a.wscale1;
                            const float4 b0 = *reinterpret_cast<const float4*>(&epi_par[3][pc]);
                            const float4 b1 = *reinterpret_cast<const float4*>(&epi_par[3][pc + 4]);
                            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += rs[gg / 8][j];
                            if (Yrow) stg256(Yrow + g16 + gg, v);
                            const float4 a0 = *reinterpret_cast<const float4*>(&epi_par[4][pc]);
                            const float4 a1 = *reinterpret_cast<const float4*>(&epi_par[4][pc + 4]);
                            const float4 i0 = *reinterpret_cast<const float4*>(&epi_par[5][pc]);
                            const float4 i1 = *reinterpret_cast<const float4*>(&epi_par[5][pc + 4]);
                            v[0] = voc_snake(v[0], a0.x, i0.x); v[1] = voc_snake(v[1], a0.y, i0.y);
                            v[2] = voc_snake(v[2], a0.z, i0.z); v[3] = voc_snake(v[3], a0.w, i0.w);
                            v[4] = voc_snake(v[4], a1.x, i1.x); v[5] = voc_snake(v[5], a1.y, i1.y);
                            v[6] = voc_snake(v[6], a1.z, i1.z); v[7] = voc_snake(v[7], a1.w, i1.w);
                            if constexpr (HEAD) {
#pragma unroll
                                for (int t = 0; t < HEAD_TAPS; ++t) {
                                    const float4 w0 = *reinterpret_cast<const float4*>(&head_par[t][pc]);
                                    const float4 w1 = *reinterpret_cast<const float4*>(&head_par[t][pc + 4]);
                                    float acc = hp[t];
                                    acc = fmaf(w0.x, v[0], acc); acc = fmaf(w0.y, v[1], acc); acc = fmaf(w0.z, v[2], acc); acc = fmaf(w0.w, v[3], acc);
                                    acc = fmaf(w1.x, v[4], acc); acc = fmaf(w1.y, v[5], acc); acc = fmaf(w1.z, v[6], acc); acc = fmaf(w1.w, v[7], acc);
                                    hp[t] = acc;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    __half2 hh, ll;
                                    voc_split2(v[2 * j], v[2 * j + 1], hh, ll);
                                    hi16[gg / 2 + j] = *reinterpret_cast<uint32_t*>(&hh);
                                    lo16[gg / 2 + j] = *reinterpret_cast<uint32_t*>(&ll);
                                }
                            }
                        }
                        if constexpr (!HEAD) {
                            stg256u(a.S_hi + soff + g16, hi16);
                            stg256u(a.S_lo + soff + g16, lo16);
                        }
                    }
                }
                if constexpr (HEAD) {
                    // planes [window][column part][tap][row]: consecutive lanes = consecutive rows, coalesced 4-byte stores
                    if (valid) {
                        float* pp = a.head_part + (((long long)b * CP + h) * HEAD_TAPS) * a.M + m;
#pragma unroll
                        for (int t = 0; t < HEAD_TAPS; ++t) pp[(long long)t * a.M] = hp[t];
                    }
                }
                skip(1);
                PF_ACC(pf_final, tf);
            }
#ifdef VOC_TC_PROF
            if (warp == 4 + EPI_WARPS && lane == 0) { PF_FLUSH(RF_EPI_W_ACC1, pf_w1); PF_FLUSH(RF_EPI_FINAL, pf_final); (void)pf_t0; }
#endif
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

template <int BN, int CP, bool ALIAS, bool PIPE, bool HEAD = false>
cudaError_t launch_fused(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, const CUtensorMap& tmW1,
                         const CUtensorMap& tmW1b, const FuArgs& a, int grid, size_t smem, cudaStream_t st) {
    static std::atomic<bool> attr_done[FU_MAX_DEVICES];
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FU_MAX_DEVICES) return cudaErrorInvalidDevice;
    if (!attr_done[dev].load(std::memory_order_acquire)) {
        // all of the 227 KB opt-in maximum that the kernel's static shared memory leaves
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, ru_fused_kernel<BN, CP, ALIAS, PIPE, HEAD>);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ru_fused_kernel<BN, CP, ALIAS, PIPE, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 232448 - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return e;
        attr_done[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128 + 256 * CP); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, ru_fused_kernel<BN, CP, ALIAS, PIPE, HEAD>, tmA, tmB, tmB2, tmW1, tmW1b, a);
}

struct FuPlan { int box_rows, SA, SB; bool alias; size_t smem; };

bool plan_fused(const RuFusedParams& p, FuPlan& pl) {
    const int BN = p.C, BK = 64;
    const bool cat = BN <= 128;
    const int span = (p.ksz - 1) * p.dil;
    pl.box_rows = ((BM + span + 7) / 8) * 8;                 // whole 8-row swizzle groups
    if (pl.box_rows > 256) return false;                     // one TMA box
    const int a_stage = 2 * pl.box_rows * BK * 2;
    const int b_stage = cat ? (BN + BN / 2) * BK * 2 : 2 * (BN / 2) * BK * 2;
    const int nkc2 = (BN + 63) / 64;
    const int a2 = (BN % 64) == 32 ? 2 * ((nkc2 - 1) * BM * BK * 2 + BM * 64) : 2 * nkc2 * BM * BK * 2;
    pl.SA = 2;
    // T next to the halo ring where that leaves at least two weight stages (a stage holds 450-1150 cycles of MMAs and
    // is refilled from L2; round 1 measured no difference between 2 and 8 stages), else overlaid on the ring
    int left = SMEM_BUDGET - pl.SA * a_stage - a2;
    pl.alias = left < 3 * b_stage;
    if (pl.alias) left = SMEM_BUDGET - std::max(pl.SA * a_stage, a2);
    pl.SB = std::min(MAX_STAGES, left / b_stage);
    if (pl.SB < 2) return false;
    pl.smem = (size_t)(pl.alias ? std::max(pl.SA * a_stage, a2) : pl.SA * a_stage + a2) + (size_t)pl.SB * b_stage + 1024;
    return true;
}

}  // namespace

#ifdef VOC_TC_PROF
extern "C" int voc_ru_prof_read(unsigned long long* out, int n, int reset) {
    unsigned long long h[RF_N];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, g_ru_prof, sizeof(h)) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < RF_N; ++i) out[i] = h[i];
    if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_ru_prof, h, sizeof(h)); }
    return RF_N;
}
#endif

// the shared-memory plan of a unit (host arithmetic; exported as voc_ru_plan for the CPU tests)
bool voc_ru_fused_plan(int C, int ksz, int dil, int* out5) {
    RuFusedParams p{};
    p.C = C; p.ksz = ksz; p.dil = dil;
    FuPlan pl;
    if ((C != 96 && C != 192) || ksz < 2 || ksz > VOC_MAX_TAPS || dil < 1 || !plan_fused(p, pl)) return false;
    out5[0] = pl.box_rows; out5[1] = pl.SA; out5[2] = pl.SB; out5[3] = pl.alias ? 1 : 0; out5[4] = (int)pl.smem;
    return true;
}

bool voc_ru_fused_eligible(const RuFusedParams& p) {
    if (p.C != 96 && p.C != 192) return false;
    if (p.ksz < 2 || p.ksz > VOC_MAX_TAPS || p.dil < 1) return false;
    if (p.L <= BM || p.B < 1) return false;                  // the pair form wants at least two M tiles per window
    if (p.a_halo < 0 || (p.a_halo && p.B != 1)) return false;
    if (!p.A_hi || !p.A_lo || !p.W7tc || !p.W1tc || !p.R) return false;
    if (p.head_w) {                                          // head folded into the unit: per-tap partial sums instead of S
        if (!p.head_part || p.head_taps != 7 || p.C != 96 || p.Y) return false;
    } else if (!p.S_hi || !p.S_lo) return false;
    if (!p.bias7 || !p.bias1 || !p.sn2_a || !p.sn2_invb || !p.snn_a || !p.snn_invb) return false;
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    if (!al32(p.A_hi) || !al32(p.A_lo) || !al32(p.R) || (p.S_hi && (!al32(p.S_hi) || !al32(p.S_lo))) || (p.Y && !al32(p.Y))) return false;
    if ((p.A_lo - p.A_hi) % 8 || p.A_lo <= p.A_hi || p.w7_plane % 8 || p.w1_plane % 8) return false;
    FuPlan pl;
    if (!plan_fused(p, pl)) return false;
    return !(p.head_w && pl.alias);
}

cudaError_t voc_launch_ru_fused(const RuFusedParams& p, cudaStream_t st, int num_sms, int flags) {
    if (!voc_ru_fused_eligible(p)) return cudaErrorNotSupported;
    FuPlan pl;
    if (!plan_fused(p, pl)) return cudaErrorNotSupported;
    const int BN = p.C, BK = 64;
    const bool cat = BN <= 128;
    FuArgs a;
    memset(&a, 0, sizeof(a));
    a.M = p.L; a.B = p.B; a.ntaps = p.ksz;
    a.a_min_off = p.a_halo - (p.ksz - 1) * p.dil;       // row 0 of the unit is row a_halo of the A tensor
    a.tap_row0 = 0; a.tap_step = p.dil;
    a.a_box_rows = pl.box_rows;
    // the same K chunking and segment schedule as voc_launch_tapgemm_tc gives these two layers (bit-identical results)
    const int seg_mmas = (flags >> 16) > 0 ? (flags >> 16) : 24;
    {
        const int ksteps = (p.C + 15) / 16;
        a.k_chunks = (p.C + BK - 1) / BK;
        a.kc_steps = (ksteps + a.k_chunks - 1) / a.k_chunks;
        a.k_chunks = (ksteps + a.kc_steps - 1) / a.kc_steps;
        a.kc_last = ksteps - (a.k_chunks - 1) * a.kc_steps;
    }
    a.seg_iters = std::max(1, seg_mmas / ((cat ? 1 : 3) * a.kc_steps));
    a.seg_head = (!cat && !(flags & VOC_TC_NO_SEG_HEAD)) ? 2 : 0;
    {
        static const int rk_env = []() { const char* e = getenv("VOC_RU_RINGK"); return e ? atoi(e) : -1; }();   // experiment hook
        a.ring_k = rk_env >= 1 ? rk_env : 3;      // measured: 1 (= plain alternation) 0.560 ms, 2 0.536, 3 0.528, 4 0.535, 5 0.546, 6 0.557 per unit on 8 windows
    }
    const int m_tiles = (p.L + BM - 1) / BM;
    a.m_tiles = (m_tiles + 1) / 2;
    a.total_tiles = a.m_tiles * p.B;
    a.SA = pl.SA; a.SB = pl.SB;
    a.wscale7 = p.w7scale; a.wscale1 = p.w1scale;
    a.bias7 = p.bias7; a.sn2_a = p.sn2_a; a.sn2_invb = p.sn2_invb;
    a.bias1 = p.bias1; a.snn_a = p.snn_a; a.snn_invb = p.snn_invb;
    const long long bs = (long long)p.L * p.C;
    a.R = p.R; a.r_bstride = bs; a.ldr = p.C;
    a.Y = p.Y; a.y_bstride = bs; a.ldy = p.C;
    a.S_hi = p.S_hi; a.S_lo = p.S_lo; a.s_bstride = bs; a.lds = p.C;
    a.head_w = p.head_w; a.head_part = p.head_part;

    CUtensorMap tmA, tmB, tmB2, tmW1, tmW1b;
    const long long a_pl = (long long)(p.A_lo - p.A_hi);
    if (!voc_tc_get_map(p.A_hi, p.C, p.L + p.a_halo, p.B, (long long)p.C * 2, ((long long)(p.L + p.a_halo) * p.C) * 2, a_pl * 2, BK,
                        a.a_box_rows, 2, &tmA))
        return cudaErrorInvalidValue;
    const long long wrow = (long long)p.C * 2, wtap = (long long)p.C * p.C * 2;
    if (!voc_tc_get_map(p.W7tc, p.C, p.C, p.ksz, wrow, wtap, p.w7_plane * 2, BK, cat ? BN : BN / 2, cat ? 1 : 2, &tmB))
        return cudaErrorInvalidValue;
    tmB2 = tmB;
    if (cat && !voc_tc_get_map(p.W7tc, p.C, p.C, p.ksz, wrow, wtap, p.w7_plane * 2, BK, BN / 2, 1, &tmB2))
        return cudaErrorInvalidValue;
    if (!voc_tc_get_map(p.W1tc, p.C, p.C, 1, wrow, wtap, p.w1_plane * 2, BK, cat ? BN : BN / 2, cat ? 1 : 2, &tmW1))
        return cudaErrorInvalidValue;
    tmW1b = tmW1;
    if (cat && !voc_tc_get_map(p.W1tc, p.C, p.C, 1, wrow, wtap, p.w1_plane * 2, BK, BN / 2, 1, &tmW1b))
        return cudaErrorInvalidValue;

    const int sms = num_sms > 0 ? num_sms : 148;
    const int grid = 2 * std::min(a.total_tiles, sms / 2);
    const bool pipe = !pl.alias && !(flags & VOC_TC_NO_PIPE);
    if (p.head_w) {
        if (BN != 96 || pl.alias) return cudaErrorNotSupported;
        return pipe ? launch_fused<96, 2, false, true, true>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st)
                    : launch_fused<96, 2, false, false, true>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st);
    }
    if (BN == 96) {
        if (pl.alias) return launch_fused<96, 2, true, false>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st);
        return pipe ? launch_fused<96, 2, false, true>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st)
                    : launch_fused<96, 2, false, false>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st);
    }
    return pl.alias ? launch_fused<192, 2, true, false>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st)
                    : launch_fused<192, 2, false, false>(tmA, tmB, tmB2, tmW1, tmW1b, a, grid, pl.smem, st);
}
