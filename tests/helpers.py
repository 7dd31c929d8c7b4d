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
