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
