"""Kernel-level GPU parity: the tcgen05 tap-GEMM (and the CUDA-core kernel on split operands)
against a float64 numpy restatement of the contraction, for every layer kind of the graph:
causal dilated Conv1d (k = 7, d = 1/3/9; k = 3), phase-decomposed ConvTranspose1d (two taps, row
offset 1 under transconv_trim = "both"), Linear / 1x1 conv, with the fused epilogues
(bias, exact GELU, per-channel scale, residual, SnakeBeta, split-fp16 operand output).

Tolerance: the tensor path multiplies split-fp16 operands (~22 mantissa bits) and accumulates in
FP32, so a dot product of O(1) terms is good to a few 1e-6 relative to the output scale.

Every call goes through voc_test_tapgemm, whose output buffers sit between guard bands of a byte pattern: a store outside
the rows / columns / planes the layer owns fails the call (the stand-in for compute-sanitizer, which is closed on the pool).

tc_flags: 0 = default (tap reuse through row-shifted descriptors, cta_group::2 pairs where the launcher
selects them; column tile chosen per launch within the layer's family), 1 = per-tap aligned loads, 8 = run-time epilogue
only, 128 = single-CTA kernel only, 256 / 512 = always the widest / the narrowest column tile of the family."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_tapgemm(A, W, tap_off, M, a_row0, bias=None, scale=None, act=0, R=None, sn_a=None, sn_invb=None):
    from math import erf
    B, a_rows, K = A.shape
    N = W.shape[1]
    acc = np.zeros((B, M, N), dtype=np.float64)
    A64, W64 = A.astype(np.float64), W.astype(np.float64)
    for t, off in enumerate(tap_off):
        rows = np.arange(M) + a_row0 + off
        ok = (rows >= 0) & (rows < a_rows)
        At = np.zeros((B, M, K))
        At[:, ok, :] = A64[:, rows[ok], :]
        acc += At @ W64[t * K:(t + 1) * K]
    v = acc + (0 if bias is None else bias.astype(np.float64))
    if act == 1:
        v = 0.5 * v * (1.0 + np.vectorize(erf)(v / np.sqrt(2.0)))
    if scale is not None:
        v = v * scale.astype(np.float64)
    if R is not None:
        v = v + R.astype(np.float64)
    s = v if sn_a is None else v + sn_invb.astype(np.float64) * np.sin(v * sn_a.astype(np.float64)) ** 2
    return v, s


CASES = [
    # name,              B, a_rows, K,   N,   M,  row0, tap_off
    ("linear",           1, 300,  128, 256, 300, 0, [0]),
    ("linear_k96",       2, 200,   96,  96, 200, 0, [0]),
    ("conv7_d1",         2, 333,   96,  96, 333, 0, [-6, -5, -4, -3, -2, -1, 0]),
    ("conv7_d3",         2, 333,  192, 192, 333, 0, [-18, -15, -12, -9, -6, -3, 0]),
    ("conv7_d9",         1, 400,  128, 384, 400, 0, [-54, -45, -36, -27, -18, -9, 0]),
    ("conv3",            3,  64,   64, 128,  64, 0, [-2, -1, 0]),
    ("convt_trim_both",  2, 100,  192, 288,  99, 1, [0, -1]),
    ("convt_trim_right", 2, 100,   64, 320, 100, 0, [0, -1]),
    ("tiny_k16",         1, 150,   16,  32, 150, 0, [-2, -1, 0]),
    # K that is not a multiple of the 64-wide chunk nor of the 16-wide k-step: chunks of equal depth (48 + 24 -> 3 + 2
    # k-steps, the second box starting at column 48), the last k-step half zero-filled by TMA
    ("linear_k72",       1, 130,   72,  64, 130, 0, [0]),
    ("conv3_k40",        2, 140,   40,  96, 140, 0, [-4, -2, 0]),
    ("conv7_k160",       1, 300,  160, 192, 300, 0, [-6, -5, -4, -3, -2, -1, 0]),
    # cta_group::2 pair mode (BN = 192 with taps*K >= 768; BN = 96 with taps*K >= 384, tc_gemm.cu): odd numbers of M tiles
    # (the pair's second CTA gets an all-padding tile), several windows, K tail (K = 96 in 64-wide chunks)
    ("pair_conv7_c192",  3, 650,  192, 192, 650, 0, [-54, -45, -36, -27, -18, -9, 0]),
    ("pair_conv7_c96",   2, 900,   96,  96, 900, 0, [-6, -5, -4, -3, -2, -1, 0]),
    ("pair_convt_k768",  2, 300,  768, 384, 299, 1, [0, -1]),
    ("pair_linear_2ntile", 1, 1000, 1024, 384, 1000, 0, [0]),
]


@pytest.mark.parametrize("flags", [0, 1, 128], ids=["default", "no_tap_reuse", "no_pairs"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_tapgemm_matches_float64(backend, case, flags):
    name, B, a_rows, K, N, M, row0, taps = case
    rng = np.random.default_rng(sum(map(ord, name)))
    A = rng.standard_normal((B, a_rows, K)).astype(np.float32)
    W = (rng.standard_normal((len(taps) * K, N)) / np.sqrt(len(taps) * K)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
    scale = (1.0 + 0.1 * rng.standard_normal(N)).astype(np.float32)
    R = rng.standard_normal((B, M, N)).astype(np.float32)
    sn_a = np.exp(0.1 * rng.standard_normal(N)).astype(np.float32)
    sn_invb = (1.0 / (np.exp(0.1 * rng.standard_normal(N)) + 1e-9)).astype(np.float32)
    v_ref, s_ref = ref_tapgemm(A, W, taps, M, row0, bias, scale, 0, R, sn_a, sn_invb)
    for mode in (2, 1):
        rc, Y, S, _ = backend.test_tapgemm(mode, A, W, taps, M, row0, bias=bias, scale=scale, R=R, sn_a=sn_a,
                                           sn_invb=sn_invb, want_y=True, want_s=True, tc_flags=flags)
        assert rc == 0, (name, mode, rc)
        ey = float(np.abs(Y - v_ref).max())
        es = float(np.abs(S - s_ref).max())
        print(f"{name} mode {mode} flags {flags}: max|dY| {ey:.2e}  max|dS| {es:.2e}")
        assert ey < 2e-5 and es < 2e-5, (name, mode, ey, es)


EPI_CASES = [c for c in CASES if c[0] in ("linear_k96", "conv7_d1", "conv7_d3", "convt_trim_both", "pair_conv7_c192",
                                           "pair_conv7_c96", "pair_convt_k768")]


@pytest.mark.parametrize("kind", ["snake_s", "res_y_s", "y_s", "res_s"])
@pytest.mark.parametrize("case", EPI_CASES, ids=[c[0] for c in EPI_CASES])
def test_tc_compile_time_epilogues(backend, case, kind):
    """The hot layer kinds (7-tap conv, 1x1 conv with residual with / without a float32 output, transposed conv) run an epilogue fixed at
    compile time on the 96- and 192-column tiles; it must equal the run-time one (tc_flags bit 3) bit for bit."""
    name, B, a_rows, K, N, M, row0, taps = case
    rng = np.random.default_rng(sum(map(ord, name + kind)))
    A = rng.standard_normal((B, a_rows, K)).astype(np.float32)
    W = (rng.standard_normal((len(taps) * K, N)) / np.sqrt(len(taps) * K)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
    R = rng.standard_normal((B, M, N)).astype(np.float32) if kind in ("res_y_s", "res_s") else None
    sn_a = np.exp(0.1 * rng.standard_normal(N)).astype(np.float32)
    sn_invb = (1.0 / (np.exp(0.1 * rng.standard_normal(N)) + 1e-9)).astype(np.float32)
    want_y = kind in ("res_y_s", "y_s")
    v_ref, s_ref = ref_tapgemm(A, W, taps, M, row0, bias, None, 0, R, sn_a, sn_invb)
    out = {}
    for flags in (0, 8):
        rc, Y, S, _ = backend.test_tapgemm(2, A, W, taps, M, row0, bias=bias, R=R, sn_a=sn_a, sn_invb=sn_invb,
                                           want_y=want_y, want_s=True, tc_flags=flags)
        assert rc == 0, (name, kind, flags, rc)
        assert float(np.abs(S - s_ref).max()) < 2e-5, (name, kind, flags)
        if want_y:
            assert float(np.abs(Y - v_ref).max()) < 2e-5, (name, kind, flags)
        out[flags] = (Y, S)
    assert np.array_equal(out[0][1], out[8][1]), (name, kind, "S differs between compile-time and run-time epilogue")
    if want_y:
        assert np.array_equal(out[0][0], out[8][0]), (name, kind, "Y differs")


TILE_CASES = [c for c in CASES if c[0] in ("linear", "conv7_d3", "conv7_d9", "conv3", "convt_trim_right", "linear_k72",
                                            "conv7_k160", "pair_conv7_c192", "pair_convt_k768", "pair_linear_2ntile")]


@pytest.mark.parametrize("kind", ["generic", "snake_s", "res_y_s", "y_s", "res_s"])
@pytest.mark.parametrize("case", TILE_CASES, ids=[c[0] for c in TILE_CASES])
def test_tc_column_tile_does_not_change_the_bits(backend, case, kind):
    """The launcher picks the column tile from the number of tiles (small batches get narrower tiles so that no SM
    idles).  That is only legitimate if every tile of a family produces the same bits: 192 and 96 columns in the
    3-pass form, 128 / 64 / 32 in the concatenated form, single CTAs and cta_group::2 pairs, every compile-time
    epilogue.  tc_flags 256 = widest, 512 = narrowest, 0 = the launcher's own choice."""
    name, B, a_rows, K, N, M, row0, taps = case
    rng = np.random.default_rng(sum(map(ord, name + kind)) + 1)
    A = rng.standard_normal((B, a_rows, K)).astype(np.float32)
    W = (rng.standard_normal((len(taps) * K, N)) / np.sqrt(len(taps) * K)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
    generic = kind == "generic"
    scale = (1.0 + 0.1 * rng.standard_normal(N)).astype(np.float32) if generic else None
    R = rng.standard_normal((B, M, N)).astype(np.float32) if kind in ("generic", "res_y_s", "res_s") else None
    sn_a = np.exp(0.1 * rng.standard_normal(N)).astype(np.float32)
    sn_invb = (1.0 / (np.exp(0.1 * rng.standard_normal(N)) + 1e-9)).astype(np.float32)
    want_y = kind in ("generic", "res_y_s", "y_s")
    v_ref, s_ref = ref_tapgemm(A, W, taps, M, row0, bias, scale, 0, R, sn_a, sn_invb)
    out = {}
    for flags in (256, 512, 0):
        rc, Y, S, _ = backend.test_tapgemm(2, A, W, taps, M, row0, bias=bias, scale=scale, R=R, sn_a=sn_a, sn_invb=sn_invb,
                                           want_y=want_y, want_s=True, tc_flags=flags)
        assert rc == 0, (name, kind, flags, rc)
        assert float(np.abs(S - s_ref).max()) < 2e-5, (name, kind, flags)
        out[flags] = (Y, S)
    for flags in (512, 0):
        assert np.array_equal(out[256][1], out[flags][1]), (name, kind, flags, "S differs between column tiles")
        if want_y:
            assert np.array_equal(out[256][0], out[flags][0]), (name, kind, flags, "Y differs between column tiles")


def test_tc_gelu_and_plain_operand_output(backend):
    """ConvNeXt pw1 form: GELU epilogue, operand output without Snake, no float32 output."""
    rng = np.random.default_rng(7)
    A = rng.standard_normal((1, 256, 128)).astype(np.float32)
    W = (rng.standard_normal((128, 512)) / np.sqrt(128)).astype(np.float32)
    bias = (0.1 * rng.standard_normal(512)).astype(np.float32)
    v_ref, s_ref = ref_tapgemm(A, W, [0], 256, 0, bias, None, 1)
    rc, Y, S, _ = backend.test_tapgemm(2, A, W, [0], 256, 0, bias=bias, act=1, want_y=False, want_s=True)
    assert rc == 0
    assert float(np.abs(S - s_ref).max()) < 2e-5


def test_ineligible_shape_is_reported(backend):
    rng = np.random.default_rng(3)
    A = rng.standard_normal((1, 40, 12)).astype(np.float32)
    W = rng.standard_normal((12, 8)).astype(np.float32)
    rc, *_ = backend.test_tapgemm(2, A, W, [0], 40)
    assert rc == 1
    rc, Y, _, _ = backend.test_tapgemm(1, A, W, [0], 40)
    assert rc == 0
    v_ref, _ = ref_tapgemm(A, W, [0], 40, 0)
    assert float(np.abs(Y - v_ref).max()) < 2e-5


def test_snake_accuracy():
    """SnakeBeta on the device (csrc/voc_common.cuh: two-term Cody-Waite reduction of the doubled argument + SFU cosine)
    against float64 x + invb * sin^2(a x), isolated from everything else: identity weights and inputs
    that are exact in fp16, through the FP32 CUDA-core kernel whose S output is plain float32.  Arguments up to
    |a x| = 40 exercise the range reduction."""
    import importlib
    backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
    rng = np.random.default_rng(12)
    K = 32
    A = (rng.integers(-256, 257, (2, 512, K)) / 64.0).astype(np.float32)           # multiples of 2^-6 in [-4, 4]
    W = np.eye(K, dtype=np.float32)
    sn_a = np.linspace(0.25, 10.0, K).astype(np.float32)
    sn_invb = np.ones(K, dtype=np.float32)
    rc, _, S, _ = backend.test_tapgemm(0, A, W, [0], 512, 0, sn_a=sn_a, sn_invb=sn_invb, want_y=False, want_s=True)
    assert rc == 0
    # the argument a*x is a float32 product on the device as in any FP32 implementation (the reference's ONNX graph
    # included): the reference value uses that same rounded argument, so that what is measured is the sin^2 evaluation
    t32 = (A * sn_a).astype(np.float32)
    x = A.astype(np.float64)
    ref = x + np.sin(t32.astype(np.float64)) ** 2
    err = float(np.abs(S - ref).max())
    exact = x + np.sin(x * sn_a.astype(np.float64)) ** 2
    print(f"snake max abs error {err:.3e} at the float32 argument, {float(np.abs(S - exact).max()):.3e} against the exact "
          f"product (|a x| up to {float(np.abs(x * sn_a).max()):.1f})")
    assert err < 7e-7          # 2.2e-7 from the SFU cosine + the reduction + float32 rounding of a result of magnitude <= 5
