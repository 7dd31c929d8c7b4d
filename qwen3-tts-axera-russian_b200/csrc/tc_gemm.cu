// Tensor-core tap-GEMM for sm_100a: tcgen05.mma with the accumulator in TMEM, operands staged in
// shared memory by TMA, FP32-exact to ~2^-22 through a three-pass split-fp16 product.
//
//   acc[b, m, n] = sum_tap sum_k  A[b, m + a_row0 + tap_off[tap], k] * W[tap, n, k]
//
// (the contraction every dense layer of the vocoder maps onto; see voc_common.cuh).  M = time is
// the MMA M dimension (128 rows per tile), N = output channels, both operands K-major:
//   * A: activations, channels-last, two fp16 planes (hi, lo); 4-D tensor map {K, rows, window,
//     plane}.  Rows outside [0, rows) -- the causal left padding and the tile overhang -- are
//     zero-filled by the TMA unit, so there is no padding pass and no bounds code.
//   * B: weights, two fp16 planes [plane][tap][N][K] of W * 2^wexp (the power-of-two scale keeps the
//     lo plane out of the fp16 subnormals; the epilogue multiplies by 2^-wexp, which is exact).
//   * D += Ahi*Bhi + Ahi*Blo + Alo*Bhi, FP32 accumulation (north_star: "accumulation stays FP32"; CPU
//     emulation 91 dB / 2.5e-5 end to end, oracle/precision_study.py; measured on B200 92.6 dB / 2.5e-5).
//
// Tap reuse: for a k-tap causal conv the 128-row A tile *plus its halo of span = (k-1)*dilation
// rows* is loaded once per K-chunk ("the time-axis halo staged in shared memory", north_star (3));
// tap j then reads it through a shared-memory descriptor whose start address is shifted by
// (tap_off[j] - min_off) rows.  That start is not aligned to the 8-row swizzle period; measured on
// B200 the swizzle is applied to absolute shared-memory address bits, so the descriptor's base
// offset stays 0 (tests/test_gpu_tapgemm.py).  Weights are streamed per (tap, K-chunk).
//
// One persistent CTA per SM, 10 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane of a
// warp-uniform loop), warps 2-9 = epilogue (TMEM -> registers -> bias / GELU / LayerScale / residual /
// SnakeBeta -> float32 residual stream and split-fp16 operand for the next layer, all global accesses one
// 32-byte sector per lane).
//
// Accumulation is segmented: the tensor core's FP32 accumulator truncates, so a TMEM buffer holds at most
// ~24 MMAs into its main accumulator (48 for the first two segments of a 3-pass tile, while the epilogue warps
// are still busy with the previous tile) before the epilogue warps add it, round-to-nearest, into registers;
// 2 (or 4) TMEM buffers let drains overlap the MMAs of the next segments.
//
// The issuing warp is written for throughput of ONE thread: an MMA costs it ~45 cycles, the pipe queues only
// ~300-450 cycles of work, so everything it touches per stage sits in registers (raw barrier addresses, running
// descriptors, counters instead of modulos) and the k-step count is a compile-time constant chosen once per A
// fill.  The epilogue's shape is a template parameter for the layer kinds that matter (EPI_*).
// Measurements behind these choices: profiles/r1_mma_microbench.txt.
//
// Two MMA forms.  3-pass (BN = 192): (hi,lo), (lo,hi), (hi,hi) into one accumulator.  Concatenated
// (BN <= 128): A_hi x [B_hi; B_lo] as one N = 2*BN MMA plus A_lo x B_hi into a separate correction block.
// One MMA costs max(~44 + N/8, N/2) cycles (tools/mma_issue_bench.cu): N = 192 runs at the math floor, and the
// concatenated form trades three N = 96 MMAs (56 cycles each) for one of N = 192 and one of N = 96.
//
// TWO (cta_group::2): a cluster of two CTAs computes two consecutive M tiles with M = 256 MMAs issued by
// the leader; each CTA stages its own A tile and half of the weight rows (2-SM TMA credited to the leader's
// barrier, multicast commits, remote relaxed arrives for the TMEM hand-back).  Bit-identical to the
// single-CTA kernel; selected by layer shape only (never by batch size).
#include "voc_common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

namespace {

constexpr int BM = 128;
constexpr int EPI_WARPS = 8;                       // 4 TMEM lane quadrants x COL_PARTS column parts
constexpr int COL_PARTS = EPI_WARPS / 4;
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_STAGES = 8;

struct TcArgs {
    int M, N, K, B, ntaps, a_row0;
    int tap_off[VOC_MAX_TAPS];
    int a_reuse, a_min_off, a_box_rows, seg_iters;
    int seg_head;                            // the first seg_head segments of a tile are twice as long (0: none)
    int tap_row0, tap_step;                  // a_reuse: tap t starts tap_row0 + t * tap_step rows into the halo tile
    int m_tiles, n_tiles, k_chunks, total_tiles;
    int kc_steps, kc_last;                   // k-steps (of 16) per K chunk / in the last one; chunk c starts at c * kc_steps * 16
    int SA, SB;
    float wscale;
    const float* bias;  int act;  const float* scale;
    const float* R;  long long r_bstride;  int ldr;
    float* Y;        long long y_bstride;  int ldy;
    __half* S_hi;  __half* S_lo;  long long s_bstride;  int lds;
    const float* sn_a;  const float* sn_invb;
};

// ---------------------------------------------------------------------------------------------
// Bring-up profiling (compiled in only with -DVOC_TC_PROF, see tools/ab_build.sh): where each role of the
// kernel spends its cycles.  Counters are summed over CTAs; voc_tc_prof_read() fetches and clears them.
// ---------------------------------------------------------------------------------------------
#ifdef VOC_TC_PROF
enum { PF_MMA_TOTAL, PF_MMA_W_ACC, PF_MMA_W_A, PF_MMA_W_B, PF_EPI_TOTAL, PF_EPI_W_ACC, PF_EPI_DRAIN, PF_EPI_FINAL,
       PF_PROD_TOTAL, PF_PROD_W_A, PF_PROD_W_B, PF_CTAS, PF_TILES, PF_SEGS, PF_N };
__device__ unsigned long long g_tc_prof[PF_N];
#define PF_DECL(name) long long name = 0
#define PF_T0(t) const long long t = clock64()
#define PF_ACC(name, t) name += clock64() - (t)
#define PF_FLUSH(idx, name) atomicAdd(&g_tc_prof[idx], (unsigned long long)(name))
#else
#define PF_DECL(name)
#define PF_T0(t)
#define PF_ACC(name, t)
#define PF_FLUSH(idx, name)
#endif

// ---------------------------------------------------------------------------------------------
// TWO: cta_group::2.  A cluster of two CTAs computes two consecutive 128-row M tiles of the same column
// tile with M = 256 MMAs issued by the leader (rank 0): each CTA stages its own A tile and HALF of the
// weight rows: one issuing warp feeds two SMs' tensor pipes and each CTA streams half the weight bytes.
// EPI fixes the epilogue's shape at compile time for the four hot layer kinds (their run-time tests and the
// untaken variants' code -- erff for GELU alone is ~200 instructions per 8 columns -- otherwise sit in every
// 16-column group of a 5000-instruction straight-line epilogue that already stalls on instruction fetch):
enum { EPI_GENERIC = 0,      // everything decided at run time from TcArgs
       EPI_SNAKE_S = 1,      // bias, Snake -> split operand S                     (7-tap convs of a residual unit)
       EPI_RES_Y_S = 2,      // bias, + residual -> Y; Snake -> split operand S    (1x1 convs of a residual unit)
       EPI_Y_S = 3,          // bias -> Y; Snake -> split operand S                (transposed convs)
       EPI_RES_S = 4 };      // bias, + residual; Snake -> split operand S         (the last 1x1 conv of a block)

// P3: the 3-pass form on a 96-column tile -- bit-identical, column by column, to BN = 192 (same passes, k-step order
// and segment schedule); the launcher picks it when 192-column tiles would leave SMs idle (small batches).
template <int BN, int BK, bool TWO, int EPI = EPI_GENERIC, bool P3 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ TcArgs a) {
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 192, "UMMA N; BN/COL_PARTS columns per epilogue thread");
    static_assert((BN / COL_PARTS) % 16 == 0, "the epilogue stores 16 operand columns (one sector) at a time");
    static_assert(BK == 64 || BK == 32, "one swizzle atom per K-chunk");
    constexpr int HN = BN / COL_PARTS;                        // columns per epilogue thread
    constexpr int PB = HN <= 64 ? HN : 16;                    // residual prefetch window (columns)
    constexpr uint32_t ROWB = BK * 2;                         // bytes of one shared-memory row
    // CAT (BN <= 128): see below.  In pair mode the concatenated form stages, per CTA, one whole weight plane
    // (rank 0: B_hi, rank 1: B_lo -- so that the N = 2*BN MMA, which takes BN rows from each CTA, produces
    // [A_hi*B_hi | A_hi*B_lo]) followed by this CTA's half of B_hi for the second MMA (N = BN, BN/2 rows each).
    constexpr bool CAT = BN <= 128 && !P3;
    constexpr int BROWS = TWO ? BN / 2 : BN;                  // weight rows per plane staged by this CTA (3-pass form)
    constexpr uint32_t B_PLANE = (TWO && CAT) ? BN * ROWB : BROWS * ROWB;
    constexpr uint32_t B_STAGE = (TWO && CAT) ? (BN + BN / 2) * ROWB : 2 * B_PLANE;
    // CAT (BN <= 128): the two weight planes of a stage are contiguous rows, so A_hi x [B_hi; B_lo] is ONE
    // MMA of N = 2*BN whose right half lands in a separate "correction" block of the accumulator buffer;
    // A_lo x B_hi then accumulates into that block.  2 MMAs per k-step instead of 3 (an MMA costs
    // ~64 + N/2 cycles, so fewer and wider is cheaper), and the 2^-11-times-smaller cross terms no longer
    // share the main accumulator's truncation.  The epilogue adds the two blocks in FP32.
    constexpr uint32_t ACC_COLS = CAT ? 2 * BN : BN;          // TMEM columns of one accumulator buffer
    // accumulator buffers in flight: two, or four where 4 x ACC_COLS fit the 512 columns and the layer is
    // bound by the segment hand-off (the pair-mode C = 96 convs): the MMA warp then runs up to four
    // segments ahead of the drains
    constexpr int NBUF = (ACC_COLS <= 128) ? 4 : 2;
    constexpr uint32_t TMEM_COLS = (NBUF * ACC_COLS <= 64) ? 64 : (NBUF * ACC_COLS <= 128) ? 128
                                   : (NBUF * ACC_COLS <= 256) ? 256 : 512;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = F16, both K-major
    constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TWO ? 2 * BM : BM) >> 4) << 24);
    constexpr uint32_t IDESC2 = (1u << 4) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)((TWO ? 2 * BM : BM) >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a_full[MAX_STAGES], bar_a_empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_b_full[MAX_STAGES], bar_b_empty[MAX_STAGES];
    __shared__ __align__(8) uint64_t bar_acc_full[4], bar_acc_empty[4];
    __shared__ uint32_t tmem_slot;
    // per-channel epilogue parameters of the current n-tile: bias, scale, snake a, snake 1/b
    __shared__ __align__(16) float epi_par[4][BN];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_plane = (uint32_t)a.a_box_rows * ROWB, a_stage = 2 * a_plane;
    const uint32_t smA = smem_base, smB = smem_base + (uint32_t)a.SA * a_stage;
    const int iters_per_tile = a.k_chunks * a.ntaps;
    const uint32_t rank = TWO ? cluster_ctarank() : 0u;
    // tile walk: CTA pairs (TWO) or single CTAs stride over (n tile, M tile [pair], window)
    const int walker = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int walkers = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmB2);
        for (int i = 0; i < a.SA; ++i) { mbar_init(&bar_a_full[i], 1); mbar_init(&bar_a_empty[i], 1); }
        for (int i = 0; i < a.SB; ++i) { mbar_init(&bar_b_full[i], 1); mbar_init(&bar_b_empty[i], 1); }
        for (int i = 0; i < NBUF; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], TWO ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) { if constexpr (TWO) tmem_alloc2(&tmem_slot, TMEM_COLS); else tmem_alloc(&tmem_slot, TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    if constexpr (TWO) cluster_sync_all();       // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        // (the whole warp runs the loop so that addresses and phases stay in uniform registers; one
        // elected lane issues)
        {
            PF_DECL(pf_w_a); PF_DECL(pf_w_b); PF_T0(pf_t0);
            int sa = 0, pa = 0, sb = 0, pb = 0;
            for (int tile = walker; tile < a.total_tiles; tile += walkers) {
                const int n_tile = tile % a.n_tiles, ml = tile / a.n_tiles;
                const int m_tile = TWO ? 2 * (ml % a.m_tiles) + (int)rank : ml % a.m_tiles, b = ml / a.m_tiles;
                const int row0 = m_tile * BM + a.a_row0, n0 = n_tile * BN + ((TWO && !CAT) ? (int)rank * BROWS : 0);
                for (int kc = 0; kc < a.k_chunks; ++kc) {
                    for (int tap = 0; tap < a.ntaps; ++tap) {
                        if (tap == 0 || !a.a_reuse) {
                            { PF_T0(tw); mbar_wait(&bar_a_empty[sa], pa ^ 1); PF_ACC(pf_w_a, tw); }
                            if (elect_one()) {
                                const int arow = row0 + (a.a_reuse ? a.a_min_off : a.tap_off[tap]);
                                if constexpr (TWO) {
                                    // both CTAs' bytes are credited to the leader's barrier
                                    if (rank == 0) mbar_expect_tx(&bar_a_full[sa], 2 * a_stage);
                                    tma_load_4d_2sm(smA + sa * a_stage, &tmA, mapa_u32(&bar_a_full[sa], 0), kc * a.kc_steps * 16, arow, b, 0);
                                } else {
                                    mbar_expect_tx(&bar_a_full[sa], a_stage);
                                    tma_load_4d(smA + sa * a_stage, &tmA, &bar_a_full[sa], kc * a.kc_steps * 16, arow, b, 0);
                                }
                            }
                            if (++sa == a.SA) { sa = 0; pa ^= 1; }
                        }
                        { PF_T0(tw); mbar_wait(&bar_b_empty[sb], pb ^ 1); PF_ACC(pf_w_b, tw); }
                        if (elect_one()) {
                            if constexpr (TWO && CAT) {
                                if (rank == 0) mbar_expect_tx(&bar_b_full[sb], 2 * B_STAGE);
                                const uint32_t lb = mapa_u32(&bar_b_full[sb], 0);
                                tma_load_4d_2sm(smB + sb * B_STAGE, &tmB, lb, kc * a.kc_steps * 16, n0, tap, (int)rank);            // a whole plane
                                tma_load_4d_2sm(smB + sb * B_STAGE + B_PLANE, &tmB2, lb, kc * a.kc_steps * 16, n0 + (int)rank * (BN / 2), tap, 0);
                            } else if constexpr (TWO) {
                                if (rank == 0) mbar_expect_tx(&bar_b_full[sb], 2 * B_STAGE);
                                tma_load_4d_2sm(smB + sb * B_STAGE, &tmB, mapa_u32(&bar_b_full[sb], 0), kc * a.kc_steps * 16, n0, tap, 0);
                            } else {
                                mbar_expect_tx(&bar_b_full[sb], B_STAGE);
                                tma_load_4d(smB + sb * B_STAGE, &tmB, &bar_b_full[sb], kc * a.kc_steps * 16, n0, tap, 0);
                            }
                        }
                        if (++sb == a.SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
#ifdef VOC_TC_PROF
            if (lane == 0) { PF_FLUSH(PF_PROD_TOTAL, clock64() - pf_t0); PF_FLUSH(PF_PROD_W_A, pf_w_a); PF_FLUSH(PF_PROD_W_B, pf_w_b); }
#endif
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        // The tensor core's FP32 accumulator truncates (measured: a shrink of ~0.25 ulp per
        // accumulation, i.e. -23 dB end to end over chains of up to 1008 MMAs), so a TMEM buffer
        // only ever holds a *segment* of seg_iters stages; the epilogue warps add the segments in
        // registers with round-to-nearest.  Within a stage the two small cross terms go first.
        // The loop is warp-uniform with one elected lane issuing: inside a divergent region the compiler cannot
        // prove the descriptors uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall.
        if (!TWO || rank == 0) {
            // raw barrier addresses (element i at base + 8 i)
            // (opaque: otherwise the compiler rematerialises each address from SR_CgaCtaId at every use)
            const uint32_t a_full0 = opaque_u32(smem_u32(&bar_a_full[0])), a_empty0 = opaque_u32(smem_u32(&bar_a_empty[0]));
            const uint32_t b_full0 = opaque_u32(smem_u32(&bar_b_full[0])), b_empty0 = opaque_u32(smem_u32(&bar_b_empty[0]));
            const uint32_t acc_full0 = opaque_u32(smem_u32(&bar_acc_full[0])), acc_empty0 = opaque_u32(smem_u32(&bar_acc_empty[0]));
            // Everything the per-stage path touches lives in registers (opaque: the compiler would otherwise
            // re-read kernel parameters from the constant bank, a dependent ~30-cycle load each, every stage).
            // The tensor pipe queues only a few MMAs (tools/mma_issue_bench.cu: ~300-450 cycles of slack), so the
            // scalar work of a stage has to stay well below the stage's MMA time (300-1150 cycles).
            // rt0 is zero at run time (a TMEM address has no bits above the lane field's reach here) but not
            // at compile time: adding it pins a kernel parameter into a register for the whole loop.
            const uint32_t rt0 = tmem_base >> 24;     // lanes < 128 live in bits 16..22
            auto reg = [&](uint32_t x) { return opaque_u32(x + rt0); };
            const int ks_regular = (int)reg((uint32_t)a.kc_steps), ks_last = (int)reg((uint32_t)a.kc_last);
            const int seg_iters = (int)reg((uint32_t)a.seg_iters), seg_head = (int)reg((uint32_t)a.seg_head);
            const int SA = (int)reg((uint32_t)a.SA), SB = (int)reg((uint32_t)a.SB);
            const int ipt = (int)reg((uint32_t)iters_per_tile);
            // One "fill" of an A stage serves n_inner consecutive stages: all taps of a K chunk with tap reuse,
            // one stage otherwise; fills from first_short on belong to the (possibly short) last K chunk.
            const int n_fills = (int)reg((uint32_t)(a.a_reuse ? a.k_chunks : a.k_chunks * a.ntaps));
            const int n_inner = (int)reg((uint32_t)(a.a_reuse ? a.ntaps : 1));
            const int first_short = (int)reg((uint32_t)(a.a_reuse ? a.k_chunks - 1 : (a.k_chunks - 1) * a.ntaps));
            // descriptor low words advance by these (16-byte units): per A stage, per tap (rows into the halo tile)
            const uint32_t a_desc0 = reg(smem_desc_lo(smA) + (a.a_reuse ? (uint32_t)a.tap_row0 * (ROWB >> 4) : 0u));
            const uint32_t a_stage16 = reg(a_stage >> 4);
            const uint32_t tap_step16 = reg(a.a_reuse ? (uint32_t)(a.tap_step * (int)(ROWB >> 4)) : 0u);
            const uint32_t b_desc0 = reg(smem_desc_lo(smB));
            int sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pas = 0;
            uint32_t b_lo = b_desc0;
            PF_DECL(pf_w_acc); PF_DECL(pf_w_a); PF_DECL(pf_w_b); PF_DECL(pf_tiles); PF_T0(pf_t0);
            for (int tile = walker; tile < a.total_tiles; tile += walkers) {
                uint32_t tmem_acc = 0, accum = 0;
                int seg_left = 0, iters_left = ipt, seg_idx = 0;
                for (int fill = 0; fill < n_fills; ++fill) {
                    // tap reuse: the descriptor simply starts some rows into the halo tile and steps from tap to
                    // tap.  The swizzle is a function of the absolute shared-memory address, so no base-offset
                    // correction is applied (the documented (addr >> 7) & 7 value yields garbage).
                    { PF_T0(tw); mbar_spin_a(a_full0 + 8 * sa, pa); PF_ACC(pf_w_a, tw); }
                    uint32_t a_lo = a_desc0 + (uint32_t)sa * a_stage16;
                    const int ks_this = fill < first_short ? ks_regular : ks_last;      // k-steps of this fill's stages
                    // the stages of one fill, with the k-step count a compile-time constant: dispatched once per
                    // fill (a per-stage switch compiled to a constant-bank jump table: a dependent load + BRX)
                    auto stages = [&](auto nks) {
                    constexpr int NKS = decltype(nks)::value;
                    for (int t = 0; t < n_inner; ++t) {
                        if (seg_left == 0) {
                            { PF_T0(tw); mbar_spin_a(acc_empty0 + 8 * as, pas ^ 1); PF_ACC(pf_w_acc, tw); }
                            tmem_acc = tmem_base + (uint32_t)as * ACC_COLS;
                            accum = 0;
                            // the segments issued while the epilogue warps are still busy with the previous
                            // tile's final epilogue are twice as long, which doubles the issuing warp's run-ahead
                            const int want = seg_idx < seg_head ? 2 * seg_iters : seg_iters;
                            seg_left = iters_left < want ? iters_left : want;
                            ++seg_idx;
                        }
                        { PF_T0(tw); mbar_spin_a(b_full0 + 8 * sb, pb); PF_ACC(pf_w_b, tw); }
                        tc_fence_after();
                        --seg_left; --iters_left;
                        const bool last_of_seg = seg_left == 0;
                        if (elect_one()) {
                            // every path is branch-free and fully unrolled (a short last chunk picks one of the
                            // compile-time variants); only the first MMA of a stage takes a run-time accumulate flag
                            {
                                if constexpr (CAT && TWO) {
#pragma unroll
                                    for (int ks = 0; ks < NKS; ++ks) {
                                        {
                                            // [main | corr] = A_hi x [B_hi (rank 0's rows) ; B_lo (rank 1's rows)];
                                            // corr += A_lo x B_hi, whose halves sit after the plane in each CTA
                                            if (ks == 0) mma2_f16_ss(tmem_acc, a_lo, b_lo, smem_desc_hi<BK>(), IDESC2, accum);
                                            else mma2_f16_ss_acc(tmem_acc, a_lo + ks * 2, b_lo + ks * 2, smem_desc_hi<BK>(), IDESC2);
                                            mma2_f16_ss_acc(tmem_acc + BN, a_lo + (a_plane >> 4) + ks * 2,
                                                            b_lo + (B_PLANE >> 4) + ks * 2, smem_desc_hi<BK>(), IDESC);
                                        }
                                    }
                                } else if constexpr (CAT) {
#pragma unroll
                                    for (int ks = 0; ks < NKS; ++ks) {
                                        {
                                            // [main | corr] = A_hi x [B_hi; B_lo];  corr += A_lo x B_hi
                                            if (ks == 0) mma_f16_ss(tmem_acc, a_lo, b_lo, smem_desc_hi<BK>(), IDESC2, accum);
                                            else mma_f16_ss_acc(tmem_acc, a_lo + ks * 2, b_lo + ks * 2, smem_desc_hi<BK>(), IDESC2);
                                            mma_f16_ss_acc(tmem_acc + BN, a_lo + (a_plane >> 4) + ks * 2, b_lo + ks * 2,
                                                           smem_desc_hi<BK>(), IDESC);
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int pass = 0; pass < 3; ++pass) {
                                        // (A plane, B plane): (hi,lo), (lo,hi), (hi,hi); offsets in 16-byte units
                                        const uint32_t ap = a_lo + (pass == 1 ? (a_plane >> 4) : 0u);
                                        const uint32_t bp = b_lo + (pass == 0 ? (B_PLANE >> 4) : 0u);
#pragma unroll
                                        for (int ks = 0; ks < NKS; ++ks) {
                                            {
                                                if constexpr (TWO) {
                                                    if (pass == 0 && ks == 0) mma2_f16_ss(tmem_acc, ap, bp, smem_desc_hi<BK>(), IDESC, accum);
                                                    else mma2_f16_ss_acc(tmem_acc, ap + ks * 2, bp + ks * 2, smem_desc_hi<BK>(), IDESC);
                                                } else {
                                                    if (pass == 0 && ks == 0) mma_f16_ss(tmem_acc, ap, bp, smem_desc_hi<BK>(), IDESC, accum);
                                                    else mma_f16_ss_acc(tmem_acc, ap + ks * 2, bp + ks * 2, smem_desc_hi<BK>(), IDESC);
                                                }
                                            }
                                        }
                                    }
                                }
                            }
                            if constexpr (TWO) {
                                mma2_commit_both_a(b_empty0 + 8 * sb);
                                if (last_of_seg) mma2_commit_both_a(acc_full0 + 8 * as);
                            } else {
                                mma_commit_a(b_empty0 + 8 * sb);
                                if (last_of_seg) mma_commit_a(acc_full0 + 8 * as);
                            }
                        }
                        __syncwarp();
                        accum = 1;
                        a_lo += tap_step16;
                        b_lo += B_STAGE >> 4;
                        if (++sb == SB) { sb = 0; pb ^= 1; b_lo = b_desc0; }
                        if (last_of_seg) { if (++as == NBUF) { as = 0; pas ^= 1; } }
                    }
                    };
                    if (ks_this == BK / 16) stages(std::integral_constant<int, BK / 16>{});
                    else if (BK == 64 && ks_this == 3) stages(std::integral_constant<int, 3>{});
                    else if (BK == 64 && ks_this == 2) stages(std::integral_constant<int, 2>{});
                    else stages(std::integral_constant<int, 1>{});
                    // the A stage is free once everything issued so far has read it
                    if (elect_one()) {
                        if constexpr (TWO) mma2_commit_both_a(a_empty0 + 8 * sa);
                        else mma_commit_a(a_empty0 + 8 * sa);
                    }
                    __syncwarp();
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
#ifdef VOC_TC_PROF
                ++pf_tiles;
#endif
            }
#ifdef VOC_TC_PROF
            if (lane == 0) {
                PF_FLUSH(PF_MMA_TOTAL, clock64() - pf_t0); PF_FLUSH(PF_MMA_W_ACC, pf_w_acc); PF_FLUSH(PF_MMA_W_A, pf_w_a);
                PF_FLUSH(PF_MMA_W_B, pf_w_b); PF_FLUSH(PF_CTAS, 1); PF_FLUSH(PF_TILES, pf_tiles);
            }
#endif
        }
    } else {
        // ================================ epilogue ====================================
        // (16 epilogue warps and a shared-memory transpose for fully coalesced global traffic were both
        // measured and are not faster: profiles/r1_epilogue_experiments.txt)
        const int q = warp & 3;                       // the TMEM lane quadrant this warp can read
        const int h = (warp - 2) >> 2;                // which part of the tile's columns
        int nseg = 0;                                 // same segment schedule as the issuing warp
        for (int rem = iters_per_tile; rem > 0; ++nseg) rem -= (nseg < a.seg_head ? 2 : 1) * a.seg_iters;
        const int etid = threadIdx.x - 64;
        const bool has_bias = EPI != EPI_GENERIC || a.bias != nullptr;
        const bool has_gelu = EPI == EPI_GENERIC && a.act == VOC_ACT_GELU;
        const bool has_scale = EPI == EPI_GENERIC && a.scale != nullptr;
        const bool has_r = EPI == EPI_RES_Y_S || EPI == EPI_RES_S || (EPI == EPI_GENERIC && a.R != nullptr);
        const bool has_y = EPI == EPI_RES_Y_S || EPI == EPI_Y_S || (EPI == EPI_GENERIC && a.Y != nullptr);
        const bool has_s = EPI != EPI_GENERIC || a.S_hi != nullptr;
        const bool has_snake = EPI != EPI_GENERIC || a.sn_a != nullptr;
        int as = 0, pas = 0, par_tile = -1;
        PF_DECL(pf_w_acc); PF_DECL(pf_drain); PF_DECL(pf_final); PF_DECL(pf_segs); PF_T0(pf_t0);
        uint32_t acc_empty_leader[NBUF];
#pragma unroll
        for (int i = 0; i < NBUF; ++i) acc_empty_leader[i] = TWO ? mapa_u32(&bar_acc_empty[i], 0) : 0u;
        for (int tile = walker; tile < a.total_tiles; tile += walkers) {
            const int n_tile = tile % a.n_tiles, ml = tile / a.n_tiles;
            const int m_tile = TWO ? 2 * (ml % a.m_tiles) + (int)rank : ml % a.m_tiles, b = ml / a.m_tiles;
            const int m = m_tile * BM + q * 32 + lane, n0 = n_tile * BN + h * HN;
            const bool valid = m < a.M;
            if (n_tile != par_tile) {
                // (re)load the parameters of this column tile; loads of the per-column vectors at the
                // point of use would sit, at L2 latency, in every column group's dependent chain
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                for (int i = etid; i < BN; i += 32 * EPI_WARPS) {
                    const int n = n_tile * BN + i;
                    epi_par[0][i] = a.bias ? __ldg(a.bias + n) : 0.f;
                    epi_par[1][i] = a.scale ? __ldg(a.scale + n) : 1.f;
                    epi_par[2][i] = a.sn_a ? __ldg(a.sn_a + n) : 1.f;
                    epi_par[3][i] = a.sn_a ? __ldg(a.sn_invb + n) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                par_tile = n_tile;
            }
            // The residual is independent of the MMAs: its loads are issued before the accumulator
            // wait (whole row when it fits the register budget, else a rolling 16-column window).
            const float* Rrow = (has_r && valid) ? a.R + (long long)b * a.r_bstride + (long long)m * a.ldr + n0 : nullptr;
            float rpf[2][PB / 8][8];
            if (Rrow) {
#pragma unroll
                for (int i = 0; i < PB / 8; ++i) ldg256(Rrow + 8 * i, rpf[0][i]);
            }
            float acc[HN];
#pragma unroll
            for (int j = 0; j < HN; ++j) acc[j] = 0.f;
            for (int seg = 0; seg < nseg; ++seg) {
                { PF_T0(tw); mbar_wait(&bar_acc_full[as], pas); PF_ACC(pf_w_acc, tw); }
                PF_T0(pf_td);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * ACC_COLS + (uint32_t)(h * HN);
                if constexpr (CAT) {
                    // main and correction blocks, at most 24 columns of each in flight
#pragma unroll
                    for (int c0 = 0; c0 < HN / 8; c0 += 3) {
                        uint32_t tm[3][8], tc[3][8];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if (c0 + c < HN / 8) { tmem_ld8(taddr + (c0 + c) * 8, tm[c]); tmem_ld8(taddr + BN + (c0 + c) * 8, tc[c]); }
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            if (c0 + c < HN / 8) {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    acc[(c0 + c) * 8 + j] += __uint_as_float(tm[c][j]) + __uint_as_float(tc[c][j]);
                            }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (TWO) mbar_arrive_cluster(acc_empty_leader[as]);   // the leader's MMA warp owns the buffer hand-back
                        else mbar_arrive(&bar_acc_empty[as]);
                    }
                } else {
                    // in chunks of 32 columns: the kernel runs at 168 registers per thread (three warps share a
                    // 16 K-register partition), which a whole half row in flight next to the accumulators exceeds
                    constexpr int CH = 4;                               // 8-column loads per chunk
#pragma unroll
                    for (int c0 = 0; c0 < HN / 8; c0 += CH) {
                        uint32_t tr[CH][8];
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            if (c0 + c < HN / 8) tmem_ld8(taddr + (c0 + c) * 8, tr[c]);
                        tmem_ld_wait();
                        if (c0 + CH >= HN / 8) {
                            // the buffer is free as soon as its contents are in registers: hand it back before the adds
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) {
                                if constexpr (TWO) mbar_arrive_cluster(acc_empty_leader[as]);
                                else mbar_arrive(&bar_acc_empty[as]);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < CH; ++c)
                            if (c0 + c < HN / 8) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) acc[(c0 + c) * 8 + j] += __uint_as_float(tr[c][j]);
                            }
                    }
                }
                if (++as == NBUF) { as = 0; pas ^= 1; }
                PF_ACC(pf_drain, pf_td);
#ifdef VOC_TC_PROF
                ++pf_segs;
#endif
            }
            PF_T0(pf_tf);
#ifdef VOC_TC_PROF
            if (valid)
#else
            if (!valid) continue;
#endif
            {
            float* Yrow = has_y ? a.Y + (long long)b * a.y_bstride + (long long)m * a.ldy + n0 : nullptr;
            const long long soff = (long long)b * a.s_bstride + (long long)m * a.lds + n0;
#pragma unroll
            for (int g16 = 0; g16 < HN; g16 += 16) {
                uint32_t hi16[8], lo16[8];             // 16 columns of the operand planes = one 32-byte sector each
#pragma unroll
                for (int g = g16; g < g16 + 16; g += 8) {
                    const int pc = h * HN + g;         // column within the tile
                    const int pcur = (g / PB) & 1;
                    if (Rrow && (g % PB) == 0 && g + PB < HN) {
#pragma unroll
                        for (int i = 0; i < PB / 8; ++i) ldg256(Rrow + g + PB + 8 * i, rpf[pcur ^ 1][i]);
                    }
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = acc[g + j] * a.wscale;
                    if (has_bias) {
                        const float4 b0 = *reinterpret_cast<const float4*>(&epi_par[0][pc]);
                        const float4 b1 = *reinterpret_cast<const float4*>(&epi_par[0][pc + 4]);
                        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                    }
                    if (has_gelu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = voc_gelu(v[j]);
                    }
                    if (has_scale) {
                        const float4 s0 = *reinterpret_cast<const float4*>(&epi_par[1][pc]);
                        const float4 s1 = *reinterpret_cast<const float4*>(&epi_par[1][pc + 4]);
                        v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w;
                        v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
                    }
                    if (Rrow) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += rpf[pcur][(g % PB) / 8][j];
                    }
                    if (Yrow) stg256(Yrow + g, v);
                    if (has_s) {
                        if (has_snake) {
                            const float4 a0 = *reinterpret_cast<const float4*>(&epi_par[2][pc]);
                            const float4 a1 = *reinterpret_cast<const float4*>(&epi_par[2][pc + 4]);
                            const float4 i0 = *reinterpret_cast<const float4*>(&epi_par[3][pc]);
                            const float4 i1 = *reinterpret_cast<const float4*>(&epi_par[3][pc + 4]);
                            v[0] = voc_snake(v[0], a0.x, i0.x); v[1] = voc_snake(v[1], a0.y, i0.y);
                            v[2] = voc_snake(v[2], a0.z, i0.z); v[3] = voc_snake(v[3], a0.w, i0.w);
                            v[4] = voc_snake(v[4], a1.x, i1.x); v[5] = voc_snake(v[5], a1.y, i1.y);
                            v[6] = voc_snake(v[6], a1.z, i1.z); v[7] = voc_snake(v[7], a1.w, i1.w);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            __half2 hh, ll;
                            voc_split2(v[2 * j], v[2 * j + 1], hh, ll);
                            hi16[(g - g16) / 2 + j] = *reinterpret_cast<uint32_t*>(&hh);
                            lo16[(g - g16) / 2 + j] = *reinterpret_cast<uint32_t*>(&ll);
                        }
                    }
                }
                if (has_s) {
                    stg256u(a.S_hi + soff + g16, hi16);
                    stg256u(a.S_lo + soff + g16, lo16);
                }
            }
            }
            PF_ACC(pf_final, pf_tf);
        }
#ifdef VOC_TC_PROF
        if (warp == 2 && lane == 0) {
            PF_FLUSH(PF_EPI_TOTAL, clock64() - pf_t0); PF_FLUSH(PF_EPI_W_ACC, pf_w_acc); PF_FLUSH(PF_EPI_DRAIN, pf_drain);
            PF_FLUSH(PF_EPI_FINAL, pf_final); PF_FLUSH(PF_SEGS, pf_segs);
        }
#endif
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if constexpr (TWO) cluster_sync_all();       // neither CTA leaves while its pair may still read its memory
    if (warp == 1) { if constexpr (TWO) tmem_dealloc2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps (cuTensorMapEncodeTiled, resolved at run time so the library links
// against cudart only), stage planning, launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

using MapKey = std::tuple<const void*, long long, long long, long long, long long, long long, long long, int, int, int>;
std::mutex g_map_mu;
std::map<MapKey, CUtensorMap> g_maps;

// 4-D fp16 tensor {d0 (contiguous), d1, d2, 2 planes}, box {bk, box_rows, 1, 2}
bool get_map(const void* base, long long d0, long long d1, long long d2, long long s1, long long s2, long long s3,
             int bk, int box_rows, int box_planes, CUtensorMap* out) {
    const MapKey key{base, d0, d1, d2, s1, s2, s3, bk, box_rows, box_planes};
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2, 2};
    const cuuint64_t strides[3] = {(cuuint64_t)s1, (cuuint64_t)s2, (cuuint64_t)s3};   // bytes, dims 1..3
    const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 1, (cuuint32_t)box_planes};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "voc_b200: cuTensorMapEncodeTiled failed (%d) dims {%lld,%lld,%lld,2} strides {%lld,%lld,%lld} box {%d,%d}\n",
                (int)r, d0, d1, d2, s1, s2, s3, bk, box_rows);
        return false;
    }
    g_maps[key] = m;
    *out = m;
    return true;
}

int pick_bn(int N) {
    static const int cand[] = {192, 128, 96, 64, 32};
    for (int c : cand) if (N % c == 0) return c;
    return 0;
}

constexpr int SMEM_BUDGET = 232448 - 1024 - 5120;   // opt-in maximum minus alignment slack and static smem
constexpr int VOC_MAX_DEVICES = 64;
inline int current_device() {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= VOC_MAX_DEVICES) return -1;
    return d;
}

template <int BN, int BK, int EPI = EPI_GENERIC, bool P3 = false>
cudaError_t launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int grid, size_t smem,
                        cudaStream_t st) {
    // the shared-memory opt-in is a per-device attribute: one flag per device ordinal (a process may hold
    // handles on several GPUs)
    static std::atomic<bool> attr_done[VOC_MAX_DEVICES];
    const int dev = current_device();
    if (dev < 0) return cudaErrorInvalidDevice;
    if (!attr_done[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_kernel<BN, BK, false, EPI, P3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BUDGET + 1024);
        if (e != cudaSuccess) return e;
        attr_done[dev].store(true, std::memory_order_release);
    }
    tapgemm_tc_kernel<BN, BK, false, EPI, P3><<<grid, TC_THREADS, smem, st>>>(tmA, tmB, tmB, a);
    return cudaGetLastError();
}

// cta_group::2 launch: clusters of two CTAs
template <int BN, int BK, int EPI = EPI_GENERIC, bool P3 = false>
cudaError_t launch_inst2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, const TcArgs& a,
                         int grid, size_t smem, cudaStream_t st) {
    static std::atomic<bool> attr_done[VOC_MAX_DEVICES];
    const int dev = current_device();
    if (dev < 0) return cudaErrorInvalidDevice;
    if (!attr_done[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(tapgemm_tc_kernel<BN, BK, true, EPI, P3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BUDGET + 1024);
        if (e != cudaSuccess) return e;
        attr_done[dev].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tapgemm_tc_kernel<BN, BK, true, EPI, P3>, tmA, tmB, tmB2, a);
}

template <int BK>
cudaError_t launch_bn(int BN, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int grid, size_t smem,
                      cudaStream_t st) {
    switch (BN) {
        case 192: return launch_inst<192, BK>(tmA, tmB, a, grid, smem, st);
        case 128: return launch_inst<128, BK>(tmA, tmB, a, grid, smem, st);
        case 96:  return launch_inst<96, BK>(tmA, tmB, a, grid, smem, st);
        case 64:  return launch_inst<64, BK>(tmA, tmB, a, grid, smem, st);
        case 32:  return launch_inst<32, BK>(tmA, tmB, a, grid, smem, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

#ifdef VOC_TC_PROF
extern "C" int voc_tc_prof_read(unsigned long long* out, int n, int reset) {
    unsigned long long h[PF_N];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h)) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < PF_N; ++i) out[i] = h[i];
    if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_tc_prof, h, sizeof(h)); }
    return PF_N;
}
#endif

bool voc_tc_get_map(const void* base, long long d0, long long d1, long long d2, long long s1, long long s2, long long s3,
                    int bk, int box_rows, int box_planes, CUtensorMap* out) {
    return get_map(base, d0, d1, d2, s1, s2, s3, bk, box_rows, box_planes, out);
}

bool voc_tc_eligible(const TapGemmParams& p) {
    if (!p.A_hi || !p.A_lo || !p.Wtc || p.S) return false;
    if (p.K % 8 || p.lda % 8 || p.N % 32 || p.K < 16) return false;
    if (p.ntaps < 1 || p.ntaps > VOC_MAX_TAPS) return false;
    // the epilogue moves 32-byte sectors: 8 floats of Y / R, 16 halves of each operand plane
    auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
    if (p.Y && (p.ldy % 8 || p.y_bstride % 8 || !al32(p.Y))) return false;
    if (p.R && (p.ldr % 8 || p.r_bstride % 8 || !al32(p.R))) return false;
    if (p.S_hi && (p.lds % 16 || p.s_bstride % 16 || !al32(p.S_hi) || !al32(p.S_lo))) return false;
    if (p.B > 1 && (p.a_bstride % 8)) return false;
    if ((p.A_lo - p.A_hi) % 8 || p.A_lo <= p.A_hi || p.wtc_plane % 8) return false;
    if ((reinterpret_cast<uintptr_t>(p.A_hi) | reinterpret_cast<uintptr_t>(p.Wtc)) & 15) return false;
    int mn = 0, mx = 0;
    for (int i = 0; i < p.ntaps; ++i) { mn = std::min(mn, p.tap_off[i]); mx = std::max(mx, p.tap_off[i]); }
    if (mx - mn > 112) return false;           // halo tile: 128 + span rows <= 256 (one TMA box)
    return encode_fn() != nullptr;
}

void voc_tc_clear_cache() {
    std::lock_guard<std::mutex> lk(g_map_mu);
    g_maps.clear();
}

// The tile plan of one launch: pure host arithmetic (no device needed; tests/test_cabi.py checks the policy through
// voc_tc_plan).  What may depend on the batch and what may not:
//   * the MMA form (3-pass / concatenated) and cta_group::2 pairing are properties of the LAYER (N, K, taps, the
//     per-window length M): the two forms round differently, so a batch must never change them;
//   * the column tile WITHIN the family may follow the batch.  A narrower tile of the same form computes every output
//     column with the same passes, k-step order and segment schedule -- the same bits -- so when the widest tile
//     leaves SMs idle or strands a nearly empty last round (one window: 64 CTAs at C = 768, 160 tiles on 148 SMs at
//     C = 384, 4 CTAs for the transformer's 512-column projections) a narrower one spreads the same k-steps over
//     more SMs and each MMA is shorter.  Cost model: the issuing warp's cycles per k-step, one MMA =
//     max(44 + N/8, N/2) (tools/mma_issue_bench.cu), times k-steps per tile times rounds, plus the exposed final
//     epilogue of the last tile (~68 cycles per column, profiles/r1_mma_microbench.txt section 5); ties keep the widest.
TcTilePlan voc_tc_plan_tile(int N, int K, int ntaps, int M, int B, int sms, int flags) {
    TcTilePlan t{};
    const int BN0 = pick_bn(N);                   // the family's widest column tile: fixes the MMA form
    if (!BN0 || K < 1 || ntaps < 1 || M < 1 || B < 1) return t;
    if (sms <= 1) sms = 148;
    // One SS-mode MMA (M 128, K 16) costs ~64 + N/2 cycles from SWIZZLE_128B operands and ~100 + N/2
    // from SWIZZLE_64B ones (tools/mma_rate.py), so 64-wide K chunks are used whenever K > 32.
    int BK = K > 32 ? 64 : 32;
    if (flags & VOC_TC_BK32) BK = 32;
    if (flags & VOC_TC_BK64) BK = 64;
    // cta_group::2 pairs: measured per layer kind (see voc_launch_tapgemm_tc); M is the per-window length
    const bool two = !(flags & VOC_TC_NO_PAIR) && BK == 64 && M > BM && (BN0 == 192 || BN0 == 96) &&
                     ((flags & VOC_TC_FORCE_PAIR) || (BN0 == 192 && (long long)ntaps * K >= 768) ||
                      (BN0 == 96 && (long long)ntaps * K >= 384));
    int BN = BN0;
    bool p3 = false;                              // 3-pass form on a 96-column tile
    if (!(flags & VOC_TC_FIXED_TILE) && BK == 64) {
        const int m_tiles0 = (M + BM - 1) / BM;
        const long long mt = (long long)(two ? (m_tiles0 + 1) / 2 : m_tiles0) * B;
        const long long walkers = two ? sms / 2 : sms;
        const long long ksteps = (long long)((K + 15) / 16) * ntaps;
        auto mma = [](int n) { return std::max(44 + n / 8, n / 2); };
        auto est = [&](int bn, bool cat) {
            const long long tiles = mt * (N / bn), rounds = (tiles + walkers - 1) / walkers;
            return rounds * ksteps * (cat ? mma(2 * bn) + mma(bn) : 3 * mma(bn)) + 68LL * bn;
        };
        if (BN0 == 192) {
            if ((flags & VOC_TC_SMALL_TILE) || est(96, false) < est(192, false)) { BN = 96; p3 = true; }
        } else if (BN0 == 128 || BN0 == 64) {
            long long best = est(BN0, true);
            for (int bn = BN0 / 2; bn >= 32; bn /= 2) {
                const long long e = est(bn, true);
                if ((flags & VOC_TC_SMALL_TILE) || e < best) { best = e; BN = bn; }
            }
        }
    }
    t.BN = BN; t.BK = BK; t.pair = two; t.p3 = p3;
    t.three_pass = BN0 == 192;
    return t;
}

cudaError_t voc_launch_tapgemm_tc(const TapGemmParams& p, cudaStream_t st, int num_sms, int flags) {
    if (!voc_tc_eligible(p)) return cudaErrorNotSupported;
    if (p.M <= 0 || p.B <= 0) return cudaSuccess;
    const int sms = num_sms > 0 ? num_sms : 148;
    const TcTilePlan tp = voc_tc_plan_tile(p.N, p.K, p.ntaps, p.M, p.B, sms, flags);
    if (!tp.BN) return cudaErrorNotSupported;
    const int BN = tp.BN, BK = tp.BK;
    const bool two = tp.pair, p3 = tp.p3;
    const bool cat = BN <= 128 && !p3;            // the kernel's CAT

    TcArgs a;
    memset(&a, 0, sizeof(a));
    a.M = p.M; a.N = p.N; a.K = p.K; a.B = p.B; a.ntaps = p.ntaps; a.a_row0 = p.a_row0;
    int mn = p.tap_off[0], mx = p.tap_off[0];
    for (int i = 0; i < p.ntaps; ++i) { a.tap_off[i] = p.tap_off[i]; mn = std::min(mn, p.tap_off[i]); mx = std::max(mx, p.tap_off[i]); }
    a.a_reuse = (p.ntaps > 1 && !(flags & VOC_TC_NO_REUSE)) ? 1 : 0;
    // the issuing warp steps its A descriptor from tap to tap, so tap reuse wants equally spaced taps
    // (every convolution and transposed convolution of the decoder); anything else reloads A per tap
    for (int i = 2; i < p.ntaps; ++i)
        if (p.tap_off[i] - p.tap_off[i - 1] != p.tap_off[1] - p.tap_off[0]) a.a_reuse = 0;
    a.a_min_off = mn;
    a.tap_row0 = p.tap_off[0] - mn;
    a.tap_step = p.ntaps > 1 ? p.tap_off[1] - p.tap_off[0] : 0;
    a.a_box_rows = a.a_reuse ? ((BM + (mx - mn) + 15) / 16) * 16 : BM;
    // MMAs accumulated in the tensor core before a round-to-nearest flush (bits 8.. of flags).  Measured
    // on the full 64-frame window: 12 -> 98.8 dB / 1.2e-5, 24 -> 94.6 dB / 2.0e-5, 48 -> 88.6 dB / 3.9e-5,
    // 96 -> 82.6 dB / 6.9e-5, never -> 68.1 dB / 3.7e-4 (fails the 1e-4 gate).
    const int seg_mmas = (flags >> 16) > 0 ? (flags >> 16) : 24;
    // seg_mmas counts MMAs into the main accumulator per segment as in the 3-pass form (3 per k-step);
    // the concatenated form (BN <= 128) keeps the same number of k-steps per segment
    // (one MMA per k-step reaches its main accumulator, so a segment may span three times the k-steps)
    // cta_group::2 pairs: for the wide layers (column tile 192 or 128, several M tiles per window)
    // Measured (tools/probe_pair.py, 4 windows): conv7 C = 768 / 384 / 192: 0.202 -> 0.152, 0.260 -> 0.209,
    // 0.257 -> 0.225 ms; conv-in 0.131 -> 0.097; but the thin layers (1x1 convs, 2-tap transposed convs with
    // K <= 384) lose a few per cent to the pair's coupling, hence the taps * K thresholds (re-measured with the
    // final epilogue by forcing pairs on every layer: profiles/r1_pair_vs_single.txt).
    // Only for BN = 192, where the single-CTA kernel runs the same three passes: results are then bit-identical
    // whichever mode a batch size selects (the BN <= 128 single-CTA form concatenates two passes).
    // BN = 96 (C = 96, the 7-tap convs): the pair halves the weight bytes each CTA re-streams per tile, which is
    // what bounds that layer (258 KB per 128 rows).  Its single-CTA form is the concatenated one, so the choice
    // must not depend on the batch: M here is the per-window length, the same for any number of windows.
    // (`two` is computed above, from the family's widest tile)
    a.m_tiles = (p.M + BM - 1) / BM; a.n_tiles = p.N / BN; a.k_chunks = (p.K + BK - 1) / BK;
    // K chunks of equal depth: K = 96 runs as 48 + 48 (two 64-wide boxes, the second starting at column 48, three
    // k-steps used of each) instead of 64 + 32 -- a 2-k-step stage is shorter than the issuing warp's scalar path.
    {
        const int ksteps = (p.K + 15) / 16;
        a.kc_steps = (ksteps + a.k_chunks - 1) / a.k_chunks;
        a.k_chunks = (ksteps + a.kc_steps - 1) / a.kc_steps;
        a.kc_last = ksteps - (a.k_chunks - 1) * a.kc_steps;
    }
    a.seg_iters = std::max(1, seg_mmas / ((cat ? 1 : 3) * a.kc_steps));
    // 3-pass form (two 192-column TMEM buffers, a 13 k-cycle final epilogue per tile at C = 192): the first two
    // segments of a tile hold twice the MMAs.  Measured end to end: see profiles/r1_segment_sweep.txt.
    a.seg_head = (!cat && !(flags & VOC_TC_NO_SEG_HEAD)) ? 2 : 0;
    if (two) a.m_tiles = (a.m_tiles + 1) / 2;             // M-tile pairs
    a.total_tiles = a.m_tiles * a.n_tiles * p.B;
    a.wscale = p.wscale;
    a.bias = p.bias; a.act = p.act; a.scale = p.scale;
    a.R = p.R; a.r_bstride = p.r_bstride; a.ldr = p.ldr;
    a.Y = p.Y; a.y_bstride = p.y_bstride; a.ldy = p.ldy;
    a.S_hi = p.S_hi; a.S_lo = p.S_lo; a.s_bstride = p.s_bstride; a.lds = p.lds;
    a.sn_a = p.sn_a; a.sn_invb = p.sn_invb;

    // stage plan
    const bool two_cat = two && cat;                  // pair mode, concatenated form: a plane + a half of B_hi per CTA
    const int a_stage = 2 * a.a_box_rows * BK * 2;
    const int b_stage = two_cat ? (BN + BN / 2) * BK * 2 : 2 * (two ? BN / 2 : BN) * BK * 2;
    if (a.a_reuse) {
        a.SA = 2;
        a.SB = std::min(MAX_STAGES, (SMEM_BUDGET - a.SA * a_stage) / b_stage);
    } else {
        a.SA = a.SB = std::min(MAX_STAGES, SMEM_BUDGET / (a_stage + b_stage));
    }
    if (a.SB < 2 || a.SA < 1) return cudaErrorNotSupported;
    const size_t smem = (size_t)a.SA * a_stage + (size_t)a.SB * b_stage + 1024;

    // tensor maps
    CUtensorMap tmA, tmB;
    const long long a_bs = p.B > 1 ? p.a_bstride : (long long)p.a_rows * p.lda;
    const long long a_plane = (long long)(p.A_lo - p.A_hi);
    if (a_plane % 8 || a_plane <= 0) return cudaErrorInvalidValue;
    if (!get_map(p.A_hi, p.K, p.a_rows, p.B, (long long)p.lda * 2, a_bs * 2, a_plane * 2, BK, a.a_box_rows, 2, &tmA))
        return cudaErrorInvalidValue;
    CUtensorMap tmB2;
    if (!get_map(p.Wtc, p.K, p.N, p.ntaps, (long long)p.K * 2, (long long)p.N * p.K * 2, p.wtc_plane * 2, BK,
                 two_cat ? BN : (two ? BN / 2 : BN), two_cat ? 1 : 2, &tmB))
        return cudaErrorInvalidValue;
    tmB2 = tmB;
    if (two_cat && !get_map(p.Wtc, p.K, p.N, p.ntaps, (long long)p.K * 2, (long long)p.N * p.K * 2, p.wtc_plane * 2, BK,
                            BN / 2, 1, &tmB2))
        return cudaErrorInvalidValue;
    // the three hot epilogue shapes get their own instantiation on the tiles the decoder blocks use
    int epi = EPI_GENERIC;
    if (p.bias && p.act == VOC_ACT_NONE && !p.scale && p.S_hi && p.sn_a && !(flags & VOC_TC_GENERIC_EPI)) {
        if (!p.R && !p.Y) epi = EPI_SNAKE_S;
        else if (p.R && p.Y) epi = EPI_RES_Y_S;
        else if (p.Y) epi = EPI_Y_S;
        else epi = EPI_RES_S;
    }
    if (two) {
        const int grid2 = 2 * std::min(a.total_tiles, sms / 2);
#define VOC_TC_PAIR(BN_, P3_) \
        (epi == EPI_SNAKE_S ? launch_inst2<BN_, 64, EPI_SNAKE_S, P3_>(tmA, tmB, tmB2, a, grid2, smem, st) \
         : epi == EPI_RES_Y_S ? launch_inst2<BN_, 64, EPI_RES_Y_S, P3_>(tmA, tmB, tmB2, a, grid2, smem, st) \
         : epi == EPI_Y_S ? launch_inst2<BN_, 64, EPI_Y_S, P3_>(tmA, tmB, tmB2, a, grid2, smem, st) \
         : epi == EPI_RES_S ? launch_inst2<BN_, 64, EPI_RES_S, P3_>(tmA, tmB, tmB2, a, grid2, smem, st) \
                          : launch_inst2<BN_, 64, EPI_GENERIC, P3_>(tmA, tmB, tmB2, a, grid2, smem, st))
        return BN == 192 ? VOC_TC_PAIR(192, false) : p3 ? VOC_TC_PAIR(96, true) : VOC_TC_PAIR(96, false);
#undef VOC_TC_PAIR
    }

    const int grid = std::min(a.total_tiles, sms);
    if (BK == 64 && (BN == 192 || BN == 96) && (epi != EPI_GENERIC || p3)) {
#define VOC_TC_SINGLE(BN_, P3_) \
        (epi == EPI_SNAKE_S ? launch_inst<BN_, 64, EPI_SNAKE_S, P3_>(tmA, tmB, a, grid, smem, st) \
         : epi == EPI_RES_Y_S ? launch_inst<BN_, 64, EPI_RES_Y_S, P3_>(tmA, tmB, a, grid, smem, st) \
         : epi == EPI_RES_S ? launch_inst<BN_, 64, EPI_RES_S, P3_>(tmA, tmB, a, grid, smem, st) \
         : epi == EPI_Y_S ? launch_inst<BN_, 64, EPI_Y_S, P3_>(tmA, tmB, a, grid, smem, st) \
                          : launch_inst<BN_, 64, EPI_GENERIC, P3_>(tmA, tmB, a, grid, smem, st))
        return BN == 192 ? VOC_TC_SINGLE(192, false) : p3 ? VOC_TC_SINGLE(96, true) : VOC_TC_SINGLE(96, false);
#undef VOC_TC_SINGLE
    }
    if (BK == 64) return launch_bn<64>(BN, tmA, tmB, a, grid, smem, st);
    return launch_bn<32>(BN, tmA, tmB, a, grid, smem, st);
}
