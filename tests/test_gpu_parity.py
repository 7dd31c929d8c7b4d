"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Gate (BASELINE.json north_star): SNR >= 60 dB and max-abs <= 1e-4 on the float output of a
window; bit-exact for the integer/index work (stitch plan, crossfade, PCM16).
"""
import os

import numpy as np
import pytest

from oracle import stitch_oracle as SO
from oracle import vocoder_oracle as VO

pytestmark = pytest.mark.gpu

SNR_GATE_DB = 60.0
MAXABS_GATE = 1e-4


def _codes(cfg, shape, seed=1):
    return np.random.default_rng(seed).integers(0, cfg.codebook_size, shape, dtype=np.int64)


def _report(tag, ref, got):
    snr = VO.snr_db(ref, got)
    mx = float(np.abs(ref.astype(np.float64) - got).max())
    print(f"{tag}: SNR {snr:.1f} dB max-abs {mx:.3e} ref-rms {float(np.sqrt((ref.astype(np.float64) ** 2).mean())):.3f}")
    return snr, mx


@pytest.mark.parametrize("trim", ["both", "right"])
def test_tiny_stages(pkg, backend, trim):
    """Every stage of a small architecture, so a failure names the first wrong kernel."""
    cfg = pkg.VocoderConfig.tiny(transconv_trim=trim, chunk_frames=12)
    w = pkg.init_weights(cfg, 0)
    codes = _codes(cfg, (3, cfg.chunk_frames, 16))
    voc = backend.Vocoder(cfg, w, wave=2)          # 3 windows in waves of 2 -> ragged last wave
    voc.set_option("debug", "1")
    taps = {}
    ref, _ = VO.forward(codes, VO.Weights(w), cfg, taps)
    got = voc.infer_chunks(codes)
    # debug stages hold the last wave = window 2 only
    for name in ["rvq", "pre_conv", "xf", "up0", "up1", "conv_in", "dec0", "dec1", "dec2", "dec3"]:
        t = taps[name][2:3].permute(0, 2, 1).contiguous().numpy()
        g = voc.debug_stage(name).reshape(t.shape)
        snr, mx = _report(f"[{trim}] {name}", t, g)
        assert snr > 90.0, name
    snr, mx = _report(f"[{trim}] out", ref.numpy(), got)
    assert got.shape == (3, cfg.chunk_samples())
    assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE
    voc.close()


def test_default_arch_short_window(pkg, backend):
    """Default (full-width) architecture on 8-frame windows, stage by stage."""
    cfg = pkg.VocoderConfig(chunk_frames=8)
    w = pkg.init_weights(cfg, 0)
    codes = _codes(cfg, (2, 8, 16))
    voc = backend.Vocoder(cfg, w, wave=2)
    voc.set_option("debug", "1")
    taps = {}
    ref, _ = VO.forward(codes, VO.Weights(w), cfg, taps)
    got = voc.infer_chunks(codes)
    bad = []
    for name in ["rvq", "pre_conv", "xf", "up0", "up1", "conv_in", "dec0", "dec1", "dec2", "dec3"]:
        t = taps[name].permute(0, 2, 1).contiguous().numpy()
        g = voc.debug_stage(name).reshape(t.shape)
        snr, mx = _report(f"default/8 {name}", t, g)
        if snr < 90.0:
            bad.append(name)
    snr, mx = _report("default/8 out", ref.numpy(), got)
    assert not bad, bad
    assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE
    voc.close()


@pytest.fixture(scope="module")
def full_model(pkg, backend):
    cfg = pkg.VocoderConfig()
    w = pkg.init_weights(cfg, 0)
    voc = backend.Vocoder(cfg, w, wave=4)
    yield cfg, w, voc
    voc.close()


def test_full_chunk_parity(full_model):
    """BASELINE config: one 64-frame x 16-codebook window, default architecture."""
    cfg, w, voc = full_model
    codes = _codes(cfg, (1, 64, 16))
    got = voc.infer_chunks(codes)
    ref, lengths = VO.forward(codes, VO.Weights(w), cfg)
    ref = ref.numpy()
    assert got.shape == ref.shape == (1, cfg.chunk_samples())
    snr, mx = _report("full/64", ref, got)
    rms = float(np.sqrt((got.astype(np.float64) ** 2).mean()))
    clamp = float((np.abs(got) >= 1.0).mean())
    print(f"output rms {rms:.3f}, clamp fraction {clamp:.5f}")
    assert 0.05 < rms < 0.5 and clamp < 0.01          # non-degenerate signal
    assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE


def test_batch_invariance(full_model):
    """A window's result does not depend on its batch-mates or its wave."""
    cfg, w, voc = full_model
    codes = _codes(cfg, (6, 64, 16), seed=5)
    a = voc.infer_chunks(codes)
    b = voc.infer_chunks(codes[4:5])
    assert np.array_equal(a[4], b[0])


def test_operand_range_report(full_model, backend, pkg):
    """The diagnostic for checkpoints other than the random-init one: per layer, how the unscaled split-fp16 operands
    use their format.  On the random-init model no operand saturates and every layer is O(1); a Snake with a tiny beta
    (1 / (e^beta + 1e-9) ~ 1e9) must show up as saturated values in the layer that applies it, and the windowed result
    of the healthy model must not change by having looked."""
    cfg, w, voc = full_model
    codes = _codes(cfg, (2, 64, 16), seed=9)
    base = voc.infer_chunks(codes)
    voc.set_option("operand_stats", "1")
    try:
        got = voc.infer_chunks(codes)
        rep = {r["tag"]: r for r in voc.operand_report()}
    finally:
        voc.set_option("operand_stats", "0")
    assert np.array_equal(base, got)
    for tag in ("dec0.convt", "dec1.ru.conv7", "dec2.ru.fused", "dec3.ru.fused", "xf.gemm", "conv_in"):
        assert tag in rep, (tag, sorted(rep))
    for tag, r in rep.items():
        print(f"{tag:16s} elements {r['elements']:>12d} rms {r['rms']:.3f} max {r['max_abs']:.1f} "
              f"hi_subnormal {r['hi_subnormal'] / r['elements']:.2e} lo_subnormal {r['lo_subnormal'] / r['elements']:.2e}")
        assert r["elements"] > 0 and r["saturated"] == 0, (tag, r)
        assert 1e-3 < r["rms"] < 1e3, (tag, r)
    assert voc.operand_report() == []                      # the counters were cleared

    small = pkg.VocoderConfig(chunk_frames=8)
    w2 = dict(pkg.init_weights(small, seed=0))
    w2["dec.3.ru.1.snake1.beta"] = np.full_like(w2["dec.3.ru.1.snake1.beta"], -40.0)
    bad = backend.Vocoder(small, w2, wave=2)
    try:
        bad.set_option("operand_stats", "1")
        bad.infer_chunks(_codes(small, (1, 8, 16), seed=3))
        rep2 = {r["tag"]: r for r in bad.operand_report()}
    finally:
        bad.close()
    assert rep2["dec3.ru.fused"]["saturated"] > 0 and rep2["dec3.ru.fused"]["max_abs"] == 65504.0, rep2["dec3.ru.fused"]
    assert rep2["dec2.ru.fused"]["saturated"] == 0


def test_code_out_of_range_is_an_error(full_model, backend):
    cfg, w, voc = full_model
    codes = _codes(cfg, (1, 64, 16))
    codes[0, 3, 7] = cfg.codebook_size
    with pytest.raises(backend.VocoderError) as e:
        voc.infer_chunks(codes)
    assert e.value.code == backend.VOC_E_INVALID
    codes[0, 3, 7] = -1
    with pytest.raises(backend.VocoderError):
        voc.synthesize(codes[0])
    # and the handle keeps working afterwards
    codes[0, 3, 7] = 0
    voc.infer_chunks(codes)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 97, 111, 112, 113, 200])
def test_synthesize_equals_reference_stitching(full_model, n):
    """Level 2 == level 1 + the reference's Python loop (vocoder_server.py:73-121), bit for bit,
    including the short-last-window duplication quirk and the PCM16 truncation."""
    cfg, w, voc = full_model
    codes = _codes(cfg, (n, 16), seed=n)
    chunk_fn = lambda padded: voc.infer_chunks(padded)[0]
    ref = SO.synthesize(codes, chunk_fn, cfg.chunk_frames)
    got = voc.synthesize(codes)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    assert np.array_equal(voc.synthesize_pcm16(codes), SO.to_pcm16(ref))


def test_window_ranges_tile_the_output(full_model):
    import torch
    cfg, w, voc = full_model
    n = 500
    codes = _codes(cfg, (n, 16), seed=9)
    full = voc.synthesize_pcm16(codes)
    nw = voc.num_windows(n)
    d_codes = torch.from_numpy(codes).cuda()
    out = np.zeros_like(full)
    covered = 0
    for (a, b) in [(0, 3), (3, 4), (4, 9), (9, nw)]:
        buf = torch.zeros(len(full), dtype=torch.int16, device="cuda")
        off, cnt = voc.synthesize_range_dev(d_codes, n, a, b, d_out_i16=buf, cap=len(full))
        voc.check_dev()
        assert off == covered
        out[off:off + cnt] = buf[:cnt].cpu().numpy()
        covered += cnt
    assert covered == len(full)
    assert np.array_equal(out, full)


def test_batched_requests_equal_single_requests(full_model):
    """voc_synthesize_batch_pcm16: many requests share waves and one stitch; every request's PCM is
    bit-identical to synthesising it alone (single-window, multi-window, the short-last-window quirk)."""
    cfg, w, voc = full_model
    lens = [1, 64, 65, 97, 112, 200, 30, 48, 150]
    reqs = [_codes(cfg, (n, 16), seed=100 + n) for n in lens]
    got = voc.synthesize_batch_pcm16(reqs)
    assert len(got) == len(reqs)
    for r, g in zip(reqs, got):
        assert np.array_equal(g, voc.synthesize_pcm16(r)), len(r)


def test_batched_requests_reject_bad_input(full_model, backend):
    cfg, w, voc = full_model
    bad = _codes(cfg, (10, 16))
    bad[3, 2] = cfg.codebook_size
    with pytest.raises(backend.VocoderError):
        voc.synthesize_batch_pcm16([_codes(cfg, (5, 16)), bad])
    voc.synthesize_batch_pcm16([_codes(cfg, (5, 16))])


@pytest.mark.parametrize("n", [7500, 10000])
def test_full_size_requests_equal_reference_stitching(full_model, n):
    """BASELINE's 10-minute utterance (7 500 frames, 157 windows) and the protocol's maximum request
    (10 000 tokens, vocoder_server.py:149): level 2 and the batched entry point equal level 1 + the
    reference's Python loop bit for bit, and the length obeys the short-last-window rule
    (n + n mod 48 frames when 1 <= n mod 48 <= 15)."""
    cfg, w, voc = full_model
    codes = _codes(cfg, (n, 16), seed=n)
    chunk_fn = lambda padded: voc.infer_chunks(padded)[0]
    ref = SO.to_pcm16(SO.synthesize(codes, chunk_fn, cfg.chunk_frames))
    got = voc.synthesize_pcm16(codes)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)
    (batched,) = voc.synthesize_batch_pcm16([codes])
    assert np.array_equal(batched, ref)
    assert voc.out_samples(n) == len(ref)


def test_cuda_path_matches_the_sibling_forward_after_the_front_end(pkg, backend):
    """The CUDA path against executable third-party code, not only against the oracle: the whole chain after the
    pre-transformer (up-sampling stages, conv-in, four decoder blocks with the production strides, head, clamp) with
    the weights and the float32 output of ``transformers``' ``Qwen3OmniMoeCode2Wav.forward`` on a small random model
    (tests/golden/sibling_tail.npz).  The front end is made the identity -- codebook 0 holds the golden latent
    frames, unit out-projection, a pre-conv whose current-frame tap is the unit matrix -- so the engine's output IS
    the sibling's tail applied to the golden latent."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sibling_tail.npz"))
    cfg = pkg.VocoderConfig(codebook_size=8, codebook_dim=16, rvq_dim=16, latent_dim=16, num_quantizers=16,
                            num_semantic=1, decoder_dim=64, chunk_frames=3, pre_transformer=False)
    w = pkg.init_weights(cfg, 0)
    for k in G.files:
        if k in w:
            assert w[k].shape == G[k].shape, k
            w[k] = np.ascontiguousarray(G[k], dtype=np.float32)
    h = G["h"]                                              # [2, 16, 3]
    cb0 = np.zeros((8, 16), np.float32)
    codes = np.zeros((2, 3, 16), np.int64)
    for b in range(2):
        for t in range(3):
            cb0[3 * b + t] = h[b, :, t]
            codes[b, t, 0] = 3 * b + t
    w["rvq.codebook.0"] = cb0
    for q in range(1, 16):
        w[f"rvq.codebook.{q}"] = np.zeros((8, 16), np.float32)
    w["rvq.proj_sem.w"] = np.eye(16, dtype=np.float32)
    w["rvq.proj_ac.w"] = np.eye(16, dtype=np.float32)
    pc = np.zeros((16, 16, 3), np.float32)
    pc[:, :, 2] = np.eye(16)                                # tap k-1 of a causal conv is the current frame
    w["pre_conv.w"] = pc
    w["pre_conv.b"] = np.zeros(16, np.float32)
    ref = G["wav"][:, 0, :]                                 # [2, 5205]
    # the oracle on these weights reproduces the sibling exactly (the identity front is exact)
    out, _ = VO.forward(codes, VO.Weights(w), cfg)
    assert float(np.abs(out.numpy() - ref).max()) < 5e-5
    for gemm in ("auto", "simt"):
        voc = backend.Vocoder(cfg, w, wave=2)
        voc.set_option("gemm", gemm)
        got = voc.infer_chunks(codes)
        assert got.shape == ref.shape
        snr, mx = _report(f"sibling-tail/{gemm}", ref, got)
        # this random model's activations reach rms 15 inside, so the absolute gate is relative to that scale
        assert snr > 60.0 and mx < 2e-4, (gemm, snr, mx)
        assert float(np.abs(got).max()) == 1.0              # the clamp


def test_cuda_path_matches_the_sibling_model_after_the_code_embedding(pkg, backend):
    """As above, with the pre-transformer included: the CUDA path on the small random model of
    tests/golden/sibling_model.npz (2 transformer layers with RoPE, sliding-window causal attention, LayerScale and
    SwiGLU, then the whole decoder) against the float32 output of ``Qwen3OmniMoeCode2Wav`` itself -- M3 to M9 in one
    forward, 22 485 samples per window from 12 latent frames."""
    from helpers import sibling_model_case
    cfg, w, codes, ref, _ = sibling_model_case(pkg)
    for gemm in ("auto", "simt"):
        voc = backend.Vocoder(cfg, w, wave=2)
        voc.set_option("gemm", gemm)
        got = voc.infer_chunks(codes)
        assert got.shape == ref.shape
        snr, mx = _report(f"sibling-model/{gemm}", ref, got)
        assert snr > 60.0 and mx < 3e-4, (gemm, snr, mx)
        assert float(np.abs(got).max()) == 1.0              # the clamp


def test_production_size_decoder_matches_the_sibling_run_live(pkg, backend):
    """Production dimensions (latent 1024, decoder 1536 -> 96 channels, strides 8/5/4/3, 64 frames -> 122 325 samples;
    the sibling's own transformer shape: 8 layers, hidden 1024, 16 x 64 heads, SwiGLU 3072): the CUDA path against
    ``transformers``' ``Qwen3OmniMoeCode2Wav`` EXECUTED HERE on the CPU with the same variance-preserving random
    weights (copied into the torch module through tests/helpers.py::sibling_param_pairs).  Needs no golden file;
    skipped where ``transformers`` has no such model."""
    torch = pytest.importorskip("torch")
    try:
        import transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe as M
        from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeCode2WavConfig
    except Exception as e:                                   # pragma: no cover
        pytest.skip(f"no sibling implementation in this image: {e}")
    from helpers import identity_front, sibling_param_pairs
    H, T = 1024, 64
    cfg = pkg.VocoderConfig(codebook_size=128, codebook_dim=H, rvq_dim=H, latent_dim=H, xf_hidden=H, xf_inter=3072,
                            xf_layers=8, xf_heads=16, xf_head_dim=64)
    w = pkg.init_weights(cfg, 11)
    hcfg = Qwen3OmniMoeCode2WavConfig()                      # the published defaults = the production decoder
    assert (hcfg.hidden_size, hcfg.decoder_dim, tuple(hcfg.upsample_rates), tuple(hcfg.upsampling_ratios)) == \
           (H, cfg.decoder_dim, cfg.upsample_rates, cfg.upsampling_ratios)
    hcfg._attn_implementation = "eager"
    torch.manual_seed(0)
    with torch.no_grad():
        m = M.Qwen3OmniMoeCode2Wav(hcfg).eval()
        for name, p in sibling_param_pairs(m):
            assert tuple(p.shape) == w[name].shape, (name, tuple(p.shape), w[name].shape)
            p.copy_(torch.from_numpy(w[name]))
        rng = np.random.default_rng(5)
        hidden = rng.standard_normal((1, T, H)).astype(np.float32)
        x = m.pre_transformer(inputs_embeds=torch.from_numpy(hidden)).last_hidden_state.permute(0, 2, 1)
        for blocks in m.upsample:
            for blk in blocks:
                x = blk(x)
        for blk in m.decoder:
            x = blk(x)
        ref = x.clamp(min=-1, max=1)[:, 0, :].numpy()
    codes = identity_front(w, cfg, hidden)
    voc = backend.Vocoder(cfg, w, wave=1)
    got = voc.infer_chunks(codes)
    assert got.shape == ref.shape == (1, 122325)
    snr, mx = _report("sibling-live/production-size", ref, got)
    assert snr > 60.0 and mx < 1e-4, (snr, mx)


def test_full_chunk_parity_trim_right(pkg, backend):
    """The other setting of ambiguity A1 at full size: ``transconv_trim="right"`` (length-preserving,
    122 880 samples per 64-frame window) against the oracle, same gate."""
    cfg = pkg.VocoderConfig(transconv_trim="right")
    w = pkg.init_weights(cfg, 0)
    voc = backend.Vocoder(cfg, w, wave=2)
    voc.set_option("gemm", "tc")
    codes = _codes(cfg, (2, 64, 16), seed=3)
    got = voc.infer_chunks(codes)
    ref, _ = VO.forward(codes, VO.Weights(w), cfg)
    ref = ref.numpy()
    assert got.shape == ref.shape == (2, 122880)
    snr, mx = _report("full/64 trim=right", ref, got)
    assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE
    assert voc.simt_launches == 0
    voc.close()


def test_benchmarked_configuration_parity(pkg, backend):
    """The configuration bench.py times (BASELINE configs[2]: 256 windows per call, decoder waves of 32, front wave of
    256, gemm = "tc") is itself checked against the oracle on four of its windows -- the first and last of the batch
    and the two either side of a wave boundary -- and nothing on that path leaves the tcgen05 kernel family."""
    import torch
    cfg = pkg.VocoderConfig()
    w = pkg.init_weights(cfg, 0)
    voc = backend.Vocoder(cfg, w, wave=32)
    voc.set_option("gemm", "tc")
    B = 256
    codes = _codes(cfg, (B, 64, 16), seed=1)
    d_codes = torch.from_numpy(codes).cuda()
    d_out = torch.empty(B, voc.chunk_samples, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    voc.infer_chunks_dev(d_codes, B, d_out, st)
    voc.check_dev(st)
    assert voc.simt_launches == 0
    pick = [0, 31, 32, 255]
    got = d_out[pick].cpu().numpy()
    ref, _ = VO.forward(codes[pick], VO.Weights(w), cfg)
    ref = ref.numpy()
    for i, k in enumerate(pick):
        snr, mx = _report(f"bench-config window {k}", ref[i], got[i])
        assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE, (k, snr, mx)
    # and the same windows alone (batch 4, one wave) give the same bits: the result does not depend on the wave
    alone = voc.infer_chunks(codes[pick])
    assert np.array_equal(alone, got)
    voc.close()


def test_tc_mode_refuses_a_shape_the_tensor_kernel_does_not_take(pkg, backend):
    """north_star: no multi-backend dispatch.  With gemm = "tc" a layer the tcgen05 kernel cannot run is an error
    (VOC_E_INVALID naming the layer), not a silent switch to the CUDA-core kernel; gemm = "auto" runs it there and
    counts the launches."""
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)        # 64 >> 4 = 4 channels in the last block: not tensor-core shaped
    w = pkg.init_weights(cfg, 0)
    codes = _codes(cfg, (1, 8, 16))
    voc = backend.Vocoder(cfg, w, wave=1)
    voc.set_option("gemm", "tc")
    with pytest.raises(backend.VocoderError) as e:
        voc.infer_chunks(codes)
    assert e.value.code == backend.VOC_E_INVALID and "not eligible" in str(e.value)
    voc.set_option("gemm", "auto")
    n0 = voc.simt_launches
    got = voc.infer_chunks(codes)
    assert voc.simt_launches > n0
    ref, _ = VO.forward(codes, VO.Weights(w), cfg)
    snr, mx = _report("tiny/auto", ref.numpy(), got)
    assert snr >= SNR_GATE_DB and mx <= MAXABS_GATE
    voc.close()


def test_two_handles_on_two_devices_in_one_process(pkg, backend):
    """The ABI promises one handle per GPU, any number per process: the >48 KB shared-memory opt-in of the tcgen05
    kernels is a per-device attribute (round-1 bug: a process-wide flag).  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cfg = pkg.VocoderConfig(chunk_frames=8)
    w = pkg.init_weights(cfg, 0)
    codes = _codes(cfg, (2, 8, 16))
    a = backend.Vocoder(cfg, w, device=0, wave=2)
    b = backend.Vocoder(cfg, w, device=1, wave=2)
    for v in (a, b):
        v.set_option("gemm", "tc")
    ga = a.infer_chunks(codes)
    gb = b.infer_chunks(codes)
    ga2 = a.infer_chunks(codes)
    assert np.array_equal(ga, gb) and np.array_equal(ga, ga2)
    a.close(); b.close()
