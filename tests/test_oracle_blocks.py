"""The model oracle's building blocks against golden vectors from the executable sibling
implementation (tests/golden/sibling_blocks.npz; generator: make_sibling_golden.py), plus
internal consistency of the oracle (FP32 vs FP64, explicit-formula spot checks, I/O contract
of scripts/export_vocoder_traced.py:38-52)."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import vocoder_oracle as VO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sibling_blocks.npz"))
TOL = 2e-5


def T(k):
    return torch.from_numpy(G[k])


def close(a, b, tol=TOL):
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    err = float(np.abs(a - b).max())
    assert a.shape == b.shape and err <= tol * max(1.0, float(np.abs(b).max())), (a.shape, b.shape, err)


class Cfg:
    snake_eps = 1e-9
    ln_eps = 1e-6
    rms_eps = 1e-5
    rope_theta = 10000.0
    dilations = (1, 3, 9)


@pytest.mark.parametrize("d", [1, 3, 9])
def test_causal_conv(d):
    y = VO.causal_conv1d(T(f"conv_d{d}.x"), T(f"conv_d{d}.w"), T(f"conv_d{d}.b"), dilation=d)
    close(y, G[f"conv_d{d}.y"])


def test_causal_conv_explicit_formula():
    # y[co,t] = b[co] + sum_ci sum_j W[co,ci,j] x[ci, t-(k-1-j)d]     
    x, w, b = G["conv_d3.x"], G["conv_d3.w"], G["conv_d3.b"]
    co, t, d, k = 5, 17, 3, 7
    acc = float(b[co])
    for ci in range(x.shape[1]):
        for j in range(k):
            tt = t - (k - 1 - j) * d
            if tt >= 0:
                acc += float(w[co, ci, j]) * float(x[1, ci, tt])
    assert abs(acc - float(G["conv_d3.y"][1, co, t])) < 1e-4


@pytest.mark.parametrize("s", [8, 5, 4, 3, 2])
def test_transposed_conv_trim_both_is_the_sibling_behaviour(s):
    x, w, b = T(f"convt_s{s}.x"), T(f"convt_s{s}.w"), T(f"convt_s{s}.b")
    y = VO.causal_transconv1d(x, w, b, s, "both")
    close(y, G[f"convt_s{s}.y"])
    L = x.shape[-1]
    assert y.shape[-1] == ((L - 1) * s if s != 2 else L * s)
    yr = VO.causal_transconv1d(x, w, b, s, "right")
    assert yr.shape[-1] == L * s
    if s != 2:      # "both" is "right" shifted by one input step
        assert torch.equal(yr[..., s:], y[..., : (L - 1) * s])


def test_snake_beta():
    close(VO.snake_beta(T("snake.x"), T("snake.alpha"), T("snake.beta")), G["snake.y"])


def test_convnext_block():
    W = VO.Weights({k[4:]: G[k] for k in G.files if k.startswith("cnx.")})
    W._w = {"p." + k: v for k, v in W._w.items()}
    close(VO.convnext_block(T("cnx.x"), W, "p.", Cfg), G["cnx.y"])


def test_residual_unit():
    W = VO.Weights({"p." + k[3:]: G[k] for k in G.files if k.startswith("ru.")})
    close(VO.residual_unit(T("ru.x"), W, "p.", 3, Cfg), G["ru.y"])


def test_decoder_block():
    class C2(Cfg):
        upsample_rates = (4, 3)
        transconv_trim = "both"
    W = VO.Weights({"dec.1." + k[4:]: G[k] for k in G.files if k.startswith("blk.")})
    close(VO.decoder_block(T("blk.x"), W, 1, C2), G["blk.y"], tol=5e-5)


def test_transformer_stack_with_sliding_window():
    class C3(Cfg):
        xf_heads, xf_head_dim, sliding_window, xf_layers = 2, 16, 5, 2
    W = VO.Weights({k: G[k] for k in G.files if k.startswith("xf.") and k not in ("xf.x", "xf.y")})
    h = T("xf.x")
    for l in range(2):
        h = VO.transformer_layer(h, W, f"xf.{l}.", C3)
    h = VO.rms_norm(h, W["xf.norm.w"], C3.rms_eps)
    close(h, G["xf.y"], tol=5e-5)


# ---- whole-graph properties ------------------------------------------------------------

def test_forward_io_contract(pkg):
    cfg = pkg.VocoderConfig.tiny()
    w = pkg.init_weights(cfg, 0)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (2, cfg.chunk_frames, 16), dtype=np.int64)
    a, lengths = VO.forward(codes, VO.Weights(w), cfg)
    assert a.shape == (2, cfg.chunk_samples()) and a.dtype == torch.float32
    assert lengths.tolist() == [cfg.chunk_frames * 1920]      # export_vocoder_traced.py:50-51
    assert float(a.abs().max()) <= 1.0
    with pytest.raises(IndexError):
        bad = codes.copy(); bad[0, 0, 0] = cfg.codebook_size
        VO.forward(bad, VO.Weights(w), cfg)


def test_lengths_under_both_trim_settings(pkg):
    assert pkg.VocoderConfig().stage_lengths() == [64, 128, 256, 256, 2040, 10195, 40776, 122325]
    assert pkg.VocoderConfig(transconv_trim="right").stage_lengths() == [64, 128, 256, 256, 2048, 10240, 40960, 122880]
    assert abs(pkg.VocoderConfig().flops_per_chunk() / 1e9 - 317.49) < 0.01       # SURVEY 8d F_chunk
    for trim in ("both", "right"):
        cfg = pkg.VocoderConfig.tiny(transconv_trim=trim)
        w = pkg.init_weights(cfg, 0)
        codes = np.zeros((1, cfg.chunk_frames, 16), dtype=np.int64)
        a, _ = VO.forward(codes, VO.Weights(w), cfg)
        assert a.shape[1] == cfg.chunk_samples()


def test_fp32_oracle_tracks_fp64(pkg):
    cfg = pkg.VocoderConfig.tiny(decoder_dim=128, chunk_frames=16)
    w = pkg.init_weights(cfg, 0)
    codes = np.random.default_rng(3).integers(0, cfg.codebook_size, (1, 16, 16), dtype=np.int64)
    a32, _ = VO.forward(codes, VO.Weights(w, torch.float32), cfg)
    a64, _ = VO.forward(codes, VO.Weights(w, torch.float64), cfg)
    assert VO.snr_db(a64.numpy(), a32.numpy()) > 90.0


def test_causality_of_the_decoder_stack(pkg):
    """With trim='right' the whole graph is causal: frames after t do not change samples
    before t*1920 (what makes fixed windows + crossfade meaningful)."""
    cfg = pkg.VocoderConfig.tiny(transconv_trim="right", chunk_frames=10)
    w = pkg.init_weights(cfg, 0)
    rng = np.random.default_rng(4)
    c1 = rng.integers(0, cfg.codebook_size, (1, 10, 16), dtype=np.int64)
    c2 = c1.copy(); c2[0, 6:] = rng.integers(0, cfg.codebook_size, (4, 16))
    a1, _ = VO.forward(c1, VO.Weights(w), cfg)
    a2, _ = VO.forward(c2, VO.Weights(w), cfg)
    assert torch.equal(a1[0, : 6 * 1920], a2[0, : 6 * 1920])
    assert not torch.equal(a1[0, 6 * 1920:], a2[0, 6 * 1920:])


def test_weight_container_roundtrip(pkg, tmp_path):
    cfg = pkg.VocoderConfig.tiny()
    w = pkg.init_weights(cfg, 7)
    p = str(tmp_path / ("m" + pkg.MODEL_SUFFIX))
    pkg.save_model(p, cfg, w)
    cfg2, w2 = pkg.load_model(p)
    assert cfg2 == cfg and list(w2) == list(w)
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    w3 = pkg.init_weights(cfg, 7)
    assert all(np.array_equal(w[k], w3[k]) for k in w)           # deterministic in the seed


def test_sibling_whole_tail_composition():
    """Everything after the pre-transformer as ONE chain -- up-sampling stages, conv-in, the four decoder blocks
    with the production strides, head Snake + conv, clamp -- against the executable sibling's own forward
    (tests/golden/sibling_tail.npz; generator: make_sibling_tail_golden.py).  Pins the oracle's composition:
    block order, Snake placement, the transposed-conv trim carried from block to block, the clamp."""
    Gt = np.load(os.path.join(os.path.dirname(__file__), "golden", "sibling_tail.npz"))

    class TailCfg(Cfg):
        upsampling_ratios = (2, 2)
        upsample_rates = (8, 5, 4, 3)
        transconv_trim = "both"
        convnext = True

    names = [k for k in Gt.files if k not in ("h", "wav", "wav_unclamped")]
    W = VO.Weights({k: Gt[k] for k in names})
    taps = {}
    with torch.no_grad():
        wav = VO.decode_tail(torch.from_numpy(Gt["h"]), W, TailCfg, taps)
    ref = Gt["wav"]
    assert tuple(wav.shape) == ref.shape == (2, 1, 5205)            # ((((3*4 - 1)*8 - 1)*5 - 1)*4 - 1)*3
    err = float(np.abs(wav.numpy() - ref).max())
    assert err < 5e-5, err
    # the clamp is exercised (about 4 % of the samples) and placed last
    assert 0.01 < float((np.abs(Gt["wav_unclamped"]) > 1).mean()) < 0.2
    assert float(np.abs(wav.numpy()).max()) == 1.0
    # float64 evaluation of the same chain agrees with the sibling's float32 to its float32 round-off
    # (intermediate activations of this random model reach rms 15)
    W64 = VO.Weights({k: Gt[k] for k in names}, torch.float64)
    with torch.no_grad():
        wav64 = VO.decode_tail(torch.from_numpy(Gt["h"]).double(), W64, TailCfg)
    assert float(np.abs(wav64.numpy() - ref).max()) < 2e-4


def test_sibling_whole_model_after_the_code_embedding():
    """Pre-transformer + up-sampling + decoder + head + clamp as ONE forward against the executable sibling
    (tests/golden/sibling_model.npz; generator: make_sibling_model_golden.py), through the oracle's public
    ``forward`` with an identity front end (tests/helpers.py::sibling_model_case)."""
    import importlib

    from helpers import sibling_model_case
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    cfg, w, codes, ref, Gm = sibling_model_case(pkg)
    taps = {}
    out, lengths = VO.forward(codes, VO.Weights(w), cfg, taps)
    assert tuple(out.shape) == ref.shape == (2, 22485)
    # the transformer alone, then the whole chain
    close(taps["xf"].permute(0, 2, 1), Gm["xf_out"], 5e-5)
    err = float(np.abs(out.numpy() - ref).max())
    assert err < 1e-4, err
    assert 0.02 < float((np.abs(Gm["wav_unclamped"]) > 1).mean()) < 0.2 and float(np.abs(out.numpy()).max()) == 1.0


def test_production_size_oracle_matches_the_sibling_run_live():
    """Production dimensions (latent 1024, decoder 1536 -> 96 channels, 64 frames -> 122 325 samples; the sibling's
    transformer shape): the oracle's public ``forward`` with an identity front end against ``transformers``'
    ``Qwen3OmniMoeCode2Wav`` executed here with the same random weights.  No golden file; ~15 s of CPU."""
    import importlib
    try:
        import transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe as M
        from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeCode2WavConfig
    except Exception as e:                                   # pragma: no cover
        pytest.skip(f"no sibling implementation in this image: {e}")
    from helpers import identity_front, sibling_param_pairs
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    H, T_ = 1024, 64
    cfg = pkg.VocoderConfig(codebook_size=128, codebook_dim=H, rvq_dim=H, latent_dim=H, xf_hidden=H, xf_inter=3072,
                            xf_layers=8, xf_heads=16, xf_head_dim=64)
    w = pkg.init_weights(cfg, 11)
    hcfg = Qwen3OmniMoeCode2WavConfig()
    hcfg._attn_implementation = "eager"
    with torch.no_grad():
        m = M.Qwen3OmniMoeCode2Wav(hcfg).eval()
        for name, p in sibling_param_pairs(m):
            assert tuple(p.shape) == w[name].shape, name
            p.copy_(torch.from_numpy(w[name]))
        hidden = np.random.default_rng(5).standard_normal((1, T_, H)).astype(np.float32)
        x = m.pre_transformer(inputs_embeds=torch.from_numpy(hidden)).last_hidden_state.permute(0, 2, 1)
        for blocks in m.upsample:
            for blk in blocks:
                x = blk(x)
        for blk in m.decoder:
            x = blk(x)
        ref = x.clamp(min=-1, max=1)[:, 0, :].numpy()
    codes = identity_front(w, cfg, hidden)
    out, lengths = VO.forward(codes, VO.Weights(w), cfg)
    assert tuple(out.shape) == ref.shape == (1, 122325)
    err = float(np.abs(out.numpy() - ref).max())
    assert err < 5e-5, err
