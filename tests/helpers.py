"""Shared test helpers (deterministic fake model, reference loader)."""
import importlib.util
import os
import zlib

import numpy as np

REFERENCE = "/root/reference"


def fake_chunk_fn(chunk_samples: int):
    """A stand-in for ``_inference_chunk``: a pure integer-hash function of the padded codes,
    reproducible on any platform (no RNG library involved)."""
    idx = np.arange(chunk_samples, dtype=np.uint64)

    def fn(padded: np.ndarray) -> np.ndarray:
        seed = np.uint64(zlib.crc32(np.ascontiguousarray(padded, dtype=np.int64).tobytes()))
        x = (idx * np.uint64(2654435761) + seed * np.uint64(40503)) % np.uint64(1 << 32)
        x = (x ^ (x >> np.uint64(15))) * np.uint64(2246822519) % np.uint64(1 << 32)
        return ((x.astype(np.float64) / float(1 << 32)) * 2.4 - 1.2).astype(np.float32)

    return fn


def load_reference_server():
    """Import /root/reference/dual_npu/vocoder_server.py as a module (numpy only; its
    onnxruntime import is lazy).  Returns None when the tree is absent (GPU box)."""
    path = os.path.join(REFERENCE, "dual_npu", "vocoder_server.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_vocoder_server", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_server_with(mod, chunk_fn, max_tokens=64, socket_path="/tmp/unused.sock"):
    """An instance of the reference's VocoderServer whose model call is `chunk_fn`."""
    class Fake(mod.VocoderServer):
        def __init__(self):
            self.socket_path = socket_path
            self.is_onnx = True
            self.max_tokens = max_tokens
            self._running = True

        def _inference_chunk(self, padded):
            return chunk_fn(padded)

    return Fake()


def sibling_model_case(pkg):
    """The small random model of tests/golden/sibling_model.npz (``transformers`` ``Qwen3OmniMoeCode2Wav`` after its
    code embedding: pre-transformer + up-sampling + decoder + head + clamp) as a VocoderConfig + weight dict whose front
    end is the identity: codebook 0 holds the golden latent frames, the other codebooks are zero, unit out-projections,
    a pre-conv whose current-frame tap is the unit matrix, unit transformer in/out projections.
    Returns (cfg, weights, codes [2, 12, 16], reference wav [2, L])."""
    import os

    import numpy as np
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sibling_model.npz"))
    H, T = 32, 12
    cfg = pkg.VocoderConfig(codebook_size=32, codebook_dim=H, rvq_dim=H, latent_dim=H, num_quantizers=16, num_semantic=1,
                            decoder_dim=64, chunk_frames=T, xf_hidden=H, xf_inter=48, xf_layers=2, xf_heads=2,
                            xf_head_dim=16, sliding_window=5)
    w = pkg.init_weights(cfg, 0)
    for k in G.files:
        if k in w:
            assert w[k].shape == G[k].shape, (k, w[k].shape, G[k].shape)
            w[k] = np.ascontiguousarray(G[k], dtype=np.float32)
    hidden = G["hidden"]                                    # [2, T, H]
    cb0 = np.zeros((32, H), np.float32)
    codes = np.zeros((2, T, 16), np.int64)
    for b in range(2):
        for t in range(T):
            cb0[T * b + t] = hidden[b, t]
            codes[b, t, 0] = T * b + t
    w["rvq.codebook.0"] = cb0
    for q in range(1, 16):
        w[f"rvq.codebook.{q}"] = np.zeros((32, H), np.float32)
    eye = np.eye(H, dtype=np.float32)
    w["rvq.proj_sem.w"] = eye.copy()
    w["rvq.proj_ac.w"] = eye.copy()
    pc = np.zeros((H, H, 3), np.float32)
    pc[:, :, 2] = eye                                       # tap k-1 of a causal conv is the current frame
    w["pre_conv.w"] = pc
    w["pre_conv.b"] = np.zeros(H, np.float32)
    w["xf.in_proj.w"] = eye.copy(); w["xf.in_proj.b"] = np.zeros(H, np.float32)
    w["xf.out_proj.w"] = eye.copy(); w["xf.out_proj.b"] = np.zeros(H, np.float32)
    return cfg, w, codes, G["wav"][:, 0, :], G


def sibling_param_pairs(m):
    """(oracle weight name, parameter) for everything a ``transformers`` ``Qwen3OmniMoeCode2Wav`` runs after its
    code embedding -- the name map of tests/golden/make_sibling_model_golden.py, usable in either direction."""
    t = m.pre_transformer
    for l, ly in enumerate(t.layers):
        for k, p in (("ln1.w", ly.input_layernorm.weight), ("q.w", ly.self_attn.q_proj.weight),
                     ("k.w", ly.self_attn.k_proj.weight), ("v.w", ly.self_attn.v_proj.weight),
                     ("o.w", ly.self_attn.o_proj.weight), ("ls_attn", ly.self_attn_layer_scale.scale),
                     ("ln2.w", ly.post_attention_layernorm.weight), ("gate.w", ly.mlp.gate_proj.weight),
                     ("up.w", ly.mlp.up_proj.weight), ("down.w", ly.mlp.down_proj.weight),
                     ("ls_mlp", ly.mlp_layer_scale.scale)):
            yield f"xf.{l}.{k}", p
    yield "xf.norm.w", t.norm.weight
    for u, blocks in enumerate(m.upsample):
        ct, cn = blocks[0], blocks[1]
        yield f"up.{u}.convt.w", ct.conv.weight
        yield f"up.{u}.convt.b", ct.conv.bias
        for k, p in (("dw.w", cn.dwconv.conv.weight), ("dw.b", cn.dwconv.conv.bias), ("ln.w", cn.norm.weight),
                     ("ln.b", cn.norm.bias), ("pw1.w", cn.pwconv1.weight), ("pw1.b", cn.pwconv1.bias),
                     ("pw2.w", cn.pwconv2.weight), ("pw2.b", cn.pwconv2.bias), ("gamma", cn.gamma)):
            yield f"up.{u}.{k}", p
    dec = m.decoder
    nb = len(m.config.upsample_rates)
    yield "dec.conv_in.w", dec[0].conv.weight
    yield "dec.conv_in.b", dec[0].conv.bias
    for b in range(nb):
        blk = dec[1 + b].block
        yield f"dec.{b}.snake.alpha", blk[0].alpha
        yield f"dec.{b}.snake.beta", blk[0].beta
        yield f"dec.{b}.convt.w", blk[1].conv.weight
        yield f"dec.{b}.convt.b", blk[1].conv.bias
        for j in range(3):
            r = blk[2 + j]
            for k, p in (("snake1.alpha", r.act1.alpha), ("snake1.beta", r.act1.beta), ("conv1.w", r.conv1.conv.weight),
                         ("conv1.b", r.conv1.conv.bias), ("snake2.alpha", r.act2.alpha), ("snake2.beta", r.act2.beta),
                         ("conv2.w", r.conv2.conv.weight), ("conv2.b", r.conv2.conv.bias)):
                yield f"dec.{b}.ru.{j}.{k}", p
    yield "head.snake.alpha", dec[1 + nb].alpha
    yield "head.snake.beta", dec[1 + nb].beta
    yield "head.conv.w", dec[2 + nb].conv.weight
    yield "head.conv.b", dec[2 + nb].conv.bias


def identity_front(w, cfg, hidden):
    """Make the front end of weight dict ``w`` the identity on ``hidden`` [B, T, H] (H = codebook_dim = rvq_dim =
    latent_dim = xf_hidden): returns the codes [B, T, 16] that select the latent frames from codebook 0."""
    import numpy as np
    B, T, H = hidden.shape
    assert cfg.codebook_dim == cfg.rvq_dim == cfg.latent_dim == H and cfg.codebook_size >= B * T
    cb0 = np.zeros((cfg.codebook_size, H), np.float32)
    codes = np.zeros((B, T, 16), np.int64)
    for b in range(B):
        for t in range(T):
            cb0[T * b + t] = hidden[b, t]
            codes[b, t, 0] = T * b + t
    w["rvq.codebook.0"] = cb0
    for q in range(1, 16):
        w[f"rvq.codebook.{q}"] = np.zeros((cfg.codebook_size, H), np.float32)
    eye = np.eye(H, dtype=np.float32)
    w["rvq.proj_sem.w"] = eye.copy()
    w["rvq.proj_ac.w"] = eye.copy()
    pc = np.zeros((H, H, cfg.pre_conv_kernel), np.float32)
    pc[:, :, cfg.pre_conv_kernel - 1] = eye                 # the last tap of a causal conv is the current frame
    w["pre_conv.w"] = pc
    w["pre_conv.b"] = np.zeros(H, np.float32)
    if cfg.pre_transformer:
        assert cfg.xf_hidden == H
        w["xf.in_proj.w"] = eye.copy(); w["xf.in_proj.b"] = np.zeros(H, np.float32)
        w["xf.out_proj.w"] = eye.copy(); w["xf.out_proj.b"] = np.zeros(H, np.float32)
    return codes
