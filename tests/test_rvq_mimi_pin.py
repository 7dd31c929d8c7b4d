"""M1 (split residual vector quantizer decode) pinned against executable third-party code.

The reference's quantizer lives in the un-vendored ``qwen_tts`` package, and the sibling model this
repo pins M3-M9 against has a different front end.  The published design M1 restates -- EMA codebooks
``embed_sum / clamp(cluster_usage, 1e-5)``, one semantic + N acoustic residual quantizers, a bias-free 1x1-conv output
projection per group, the two groups added -- is the Mimi codec's, and ``transformers`` ships it in this image:
``MimiSplitResidualVectorQuantizer`` (modeling_mimi.py:1296-1350, codebook :1176-1219, group decode :1282-1293).
[non-reference evidence, same standing as the sibling pins]

CPU: Mimi's own ``state_dict()`` (its real key names, EMA buffers with dead codes, Conv1d projection kernels) goes through
``weights.from_speech_tokenizer`` and the oracle's ``rvq_decode`` must reproduce ``Mimi.decode`` on the same codes.
GPU: the CUDA path's codebook-sum stage at production dimensions (16 x 2048 x 256 -> 512) against Mimi executed live."""
import importlib

import numpy as np
import pytest
import torch

from oracle import vocoder_oracle as VO


def _mimi_quantizer(codebook_size, codebook_dim, out_dim, n_q, n_sem, seed):
    from transformers.models.mimi.configuration_mimi import MimiConfig
    from transformers.models.mimi.modeling_mimi import MimiSplitResidualVectorQuantizer
    mc = MimiConfig(codebook_size=codebook_size, codebook_dim=codebook_dim, vector_quantization_hidden_dimension=codebook_dim,
                    hidden_size=out_dim, num_quantizers=n_q, num_semantic_quantizers=n_sem)
    q = MimiSplitResidualVectorQuantizer(mc).eval()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for group in (q.semantic_residual_vector_quantizer, q.acoustic_residual_vector_quantizer):
            group.output_proj.weight.copy_(torch.randn(group.output_proj.weight.shape, generator=g) / codebook_dim ** 0.5)
            for layer in group.layers:
                cb = layer.codebook
                usage = torch.rand(codebook_size, generator=g) * 50.0 + 0.5
                usage[3] = 0.0                      # codes the EMA never saw: the clamp at 1e-5 decides their rows
                usage[7] = 1e-7
                cb.cluster_usage.copy_(usage)
                cb.embed_sum.copy_(torch.randn(codebook_size, codebook_dim, generator=g) * usage.clamp(min=1e-3)[:, None])
                cb._embed = None
    return q


def _our_weights(pkg, W, cfg, q):
    """Mimi's state_dict under the upstream prefix, completed with a fabricated rest of the checkpoint, through the
    loader; returns (loaded weights, the quantizer keys the loader consumed)."""
    w0 = pkg.init_weights(cfg, 1)
    up = W.to_speech_tokenizer_names(cfg, w0)
    up = {k: v for k, v in up.items() if ".quantizer." not in k}
    sd = {"decoder.quantizer." + k: v.detach().numpy() for k, v in q.state_dict().items()}
    up.update(sd)
    _, w = W.from_speech_tokenizer(up, cfg=cfg)
    return w, sorted(sd)


def test_oracle_and_loader_reproduce_mimi_decode(pkg):
    W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)
    q = _mimi_quantizer(cfg.codebook_size, cfg.codebook_dim, cfg.rvq_dim, cfg.num_quantizers, cfg.num_semantic, seed=5)
    w, keys = _our_weights(pkg, W, cfg, q)
    assert any(k.endswith("semantic_residual_vector_quantizer.layers.0.codebook.embed_sum") for k in keys)
    rng = np.random.default_rng(2)
    codes = rng.integers(0, cfg.codebook_size, (3, cfg.num_quantizers, 11), dtype=np.int64)
    codes[0, :, 0] = 3                                  # the dead codes are used
    codes[1, :, 1] = 7
    with torch.no_grad():
        want = q.decode(torch.from_numpy(codes)).numpy()                    # [B, rvq_dim, T]
    got = VO.rvq_decode(torch.from_numpy(codes), VO.Weights(w), cfg).numpy()
    assert got.shape == want.shape
    scale = float(np.abs(want).max())
    err = float(np.abs(got - want).max())
    print(f"rvq_decode vs Mimi: max-abs {err:.2e} at scale {scale:.2f}")
    assert err <= 2e-6 * max(1.0, scale)
    # the clamp, explicitly: a dead code's row is embed_sum / 1e-5
    cb0 = q.semantic_residual_vector_quantizer.layers[0].codebook
    assert np.allclose(w["rvq.codebook.0"][3], (cb0.embed_sum[3] / 1e-5).numpy(), rtol=1e-6)


@pytest.mark.gpu
def test_cuda_rvq_stage_matches_mimi_at_production_dimensions(pkg, backend):
    W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
    cfg = pkg.VocoderConfig(chunk_frames=8)
    q = _mimi_quantizer(cfg.codebook_size, cfg.codebook_dim, cfg.rvq_dim, cfg.num_quantizers, cfg.num_semantic, seed=6)
    w, _ = _our_weights(pkg, W, cfg, q)
    codes = np.random.default_rng(4).integers(0, cfg.codebook_size, (2, cfg.chunk_frames, cfg.num_quantizers), dtype=np.int64)
    codes[0, 0, :] = 3
    with torch.no_grad():
        want = q.decode(torch.from_numpy(codes).permute(0, 2, 1).contiguous()).permute(0, 2, 1).numpy()   # [B, T, rvq_dim]
    voc = backend.Vocoder(cfg, w, wave=2)
    try:
        voc.set_option("debug", "1")
        voc.infer_chunks(codes)
        got = voc.debug_stage("rvq").reshape(want.shape)
    finally:
        voc.close()
    # the device folds the output projections into the codebooks (16 tables [2048][512]) and stores the sum as a
    # split-fp16 operand: float32 rounding of a 16-term sum plus 2^-22 relative of the operand format
    scale = float(np.abs(want).max())
    err = float(np.abs(got - want).max())
    print(f"CUDA rvq stage vs Mimi: max-abs {err:.2e} at scale {scale:.1f}")
    assert err <= 4e-6 * max(1.0, scale)
