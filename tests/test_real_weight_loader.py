"""SURVEY 8f N2 scaffolding, CPU only: a checkpoint fabricated in the upstream ``speech_tokenizer`` naming
(EMA codebooks, 1x1-conv output projections, sibling-style decoder keys) loads back into this repo's layout, the
architecture is inferred from tensor shapes, a missing key names what was tried, and the oracle runs on the result."""
import importlib

import numpy as np
import pytest


def _mods():
    pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
    W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
    return pkg, W


def test_round_trip_through_the_upstream_naming(tmp_path):
    pkg, W = _mods()
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)
    w = pkg.init_weights(cfg, 3)
    up = W.to_speech_tokenizer_names(cfg, w)
    assert "decoder.quantizer.rvq_first.vq.layers.0._codebook.embedding_sum" in up
    assert up["decoder.quantizer.rvq_rest.output_proj.weight"].shape == (cfg.rvq_dim, cfg.codebook_dim, 1)
    path = str(tmp_path / "model.safetensors")
    W.write_safetensors(path, up)
    cfg2, w2 = W.from_speech_tokenizer(path, xf_head_dim=cfg.xf_head_dim, chunk_frames=8)
    for f in ("codebook_size", "codebook_dim", "rvq_dim", "latent_dim", "xf_hidden", "xf_inter", "xf_layers", "xf_heads",
              "decoder_dim", "upsample_rates", "upsampling_ratios", "conv_kernel", "pre_conv_kernel"):
        assert getattr(cfg2, f) == getattr(cfg, f), f
    assert list(w2) == list(W.weight_shapes(cfg))
    for k in w:
        tol = 2e-6 if k.startswith("rvq.codebook") else 0.0          # embedding_sum / usage is one rounding away
        assert np.allclose(w2[k], w[k], rtol=tol, atol=0.0), k


def test_dead_codes_use_the_clamped_usage():
    pkg, W = _mods()
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)
    w = pkg.init_weights(cfg, 0)
    up = W.to_speech_tokenizer_names(cfg, w)
    k = "decoder.quantizer.rvq_rest.vq.layers.2._codebook."
    up[k + "cluster_usage"][5] = 0.0                                   # a code the EMA never saw
    up[k + "embedding_sum"][5] = 1e-6
    _, w2 = W.from_speech_tokenizer(up, cfg=cfg)
    assert np.allclose(w2["rvq.codebook.3"][5], 1e-6 / 1e-5)


def test_missing_key_names_what_was_tried_and_rename_fixes_it():
    pkg, W = _mods()
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)
    w = pkg.init_weights(cfg, 0)
    up = W.to_speech_tokenizer_names(cfg, w)
    up["decoder.pre_conv_renamed.weight"] = up.pop("decoder.pre_conv.conv.weight")
    with pytest.raises(KeyError) as e:
        W.from_speech_tokenizer(up, cfg=cfg)
    assert "pre_conv.conv.weight" in str(e.value) and "rename" in str(e.value)
    _, w2 = W.from_speech_tokenizer(up, cfg=cfg, rename={"pre_conv.w": "pre_conv_renamed.weight"})
    assert np.array_equal(w2["pre_conv.w"], w["pre_conv.w"])


def test_loaded_weights_drive_the_oracle():
    pkg, W = _mods()
    from oracle import vocoder_oracle as VO
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)
    w = pkg.init_weights(cfg, 1)
    _, w2 = W.from_speech_tokenizer(W.to_speech_tokenizer_names(cfg, w), cfg=cfg)
    codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, 8, 16), dtype=np.int64)
    a, _ = VO.forward(codes, VO.Weights(w), cfg)
    b, _ = VO.forward(codes, VO.Weights(w2), cfg)
    assert float(np.abs(a.numpy() - b.numpy()).max()) < 1e-5
