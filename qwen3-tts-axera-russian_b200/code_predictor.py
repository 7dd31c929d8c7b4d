"""ctypes binding of the code-predictor half of ``libvoc_b200.so`` (``include/cp_b200.h``).

``CodePredictor`` is the device-side replacement of what ``/root/reference/dual_npu/code_predictor_server.py`` keeps in
an ``ort.InferenceSession`` plus NumPy arrays: the 5-layer decode-step transformer, the 15 ``lm_head`` matrices and the
15 codec embedding tables, with the KV cache resident on the GPU.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import json
import zlib
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

from . import backend as _backend

CP_OK, CP_E_INVALID, CP_E_CUDA, CP_E_STATE, CP_E_NOMEM = 0, -1, -2, -3, -4

# name -> (restype, argtypes): checked against include/cp_b200.h by tests/test_cabi.py
SIGNATURES = {
    "cp_create": (C.c_void_p, [C.c_char_p, C.c_int]),
    "cp_destroy": (None, [C.c_void_p]),
    "cp_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "cp_finalize": (C.c_int, [C.c_void_p]),
    "cp_reset": (C.c_int, [C.c_void_p]),
    "cp_cache_len": (C.c_int, [C.c_void_p]),
    "cp_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cp_logits": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "cp_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_ulonglong, C.c_void_p]),
    "cp_max_batch": (C.c_int, [C.c_void_p]),
    "cp_predict_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "cp_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "cp_predict_path": (C.c_char_p, [C.c_void_p]),
    "cp_hidden_size": (C.c_int, [C.c_void_p]),
    "cp_num_groups": (C.c_int, [C.c_void_p]),
    "cp_vocab_size": (C.c_int, [C.c_void_p]),
    "cp_launches": (C.c_longlong, [C.c_void_p]),
    "cp_last_error": (C.c_char_p, [C.c_void_p]),
    "cp_stream": (C.c_void_p, [C.c_void_p]),
}


@dataclass(frozen=True)
class CPConfig:
    """Shape of the code predictor (docs/ARCHITECTURE.md:101; code_predictor_server.py:39,65-67)."""
    hidden: int = 1024
    layers: int = 5
    heads: int = 16
    kv_heads: int = 8
    head_dim: int = 128
    inter: int = 3072
    vocab: int = 2048
    groups: int = 15
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_positions: int = 32

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), sort_keys=True)

    @staticmethod
    def tiny(**kw) -> "CPConfig":
        base = dict(hidden=64, layers=2, heads=4, kv_heads=2, head_dim=16, inter=96, vocab=32, groups=4)
        base.update(kw)
        return CPConfig(**base)

    @staticmethod
    def from_weights(w) -> "CPConfig":
        """Read the shape off a ``code_predictor_weights.npz`` the way the reference does (``head_dim =
        q_proj.shape[0] // 16``, 8 KV heads; code_predictor_server.py:65-67)."""
        layers = 0
        while f"layer_{layers}_q_proj" in w:
            layers += 1
        groups = 0
        while f"lm_head_{groups}" in w:
            groups += 1
        q, hidden = w["layer_0_q_proj"].shape
        head_dim = q // 16
        return CPConfig(hidden=hidden, layers=layers, heads=16, kv_heads=w["layer_0_k_proj"].shape[0] // head_dim,
                        head_dim=head_dim, inter=w["layer_0_gate_proj"].shape[0], vocab=w["lm_head_0"].shape[0],
                        groups=groups)


def weight_shapes(cfg: CPConfig) -> Dict[str, Tuple[int, ...]]:
    """Arrays of code_predictor_weights.npz by the names scripts/export_code_predictor_weights.py:50-70 writes."""
    s: Dict[str, Tuple[int, ...]] = {}
    q, kv = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        p = f"layer_{i}_"
        s[p + "input_ln"] = (cfg.hidden,)
        s[p + "q_proj"] = (q, cfg.hidden)
        s[p + "k_proj"] = (kv, cfg.hidden)
        s[p + "v_proj"] = (kv, cfg.hidden)
        s[p + "o_proj"] = (cfg.hidden, q)
        s[p + "q_norm"] = (cfg.head_dim,)
        s[p + "k_norm"] = (cfg.head_dim,)
        s[p + "post_ln"] = (cfg.hidden,)
        s[p + "gate_proj"] = (cfg.inter, cfg.hidden)
        s[p + "up_proj"] = (cfg.inter, cfg.hidden)
        s[p + "down_proj"] = (cfg.hidden, cfg.inter)
    s["final_norm"] = (cfg.hidden,)
    for i in range(cfg.groups):
        s[f"codec_emb_{i}"] = (cfg.vocab, cfg.hidden)
        s[f"lm_head_{i}"] = (cfg.vocab, cfg.hidden)
    return s


def init_weights(cfg: CPConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Deterministic random weights of that shape (there is no checkpoint in this environment): residual stream O(1),
    logits with a spread of ~2 so that top-k sampling has something to choose from."""
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_shapes(cfg).items():
        g = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if name.endswith("_ln") or name.endswith("_norm") or name == "final_norm":
            w = 1.0 + 0.05 * g.standard_normal(shape)
        elif name.startswith("codec_emb_"):
            w = g.standard_normal(shape)
        elif name.startswith("lm_head_"):
            w = g.standard_normal(shape) * (2.0 / np.sqrt(shape[1]))
        else:
            gain = 0.5 if (name.endswith("o_proj") or name.endswith("down_proj")) else 1.0
            w = g.standard_normal(shape) * (gain / np.sqrt(shape[1]))
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


class CodePredictorError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cp_b200 error {code}: {msg}")
        self.code = code


_bound = False


def load_library(path: Optional[str] = None):
    global _bound
    lib = _backend.load_library(path)
    if not _bound or path is not None:
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _bound = True
    return lib


def _f32(a, n=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    return a if n is None else np.ascontiguousarray(a[:n])


class CodePredictor:
    """One handle = one GPU: weights, KV cache, scratch, a stream and the captured frame graph."""

    def __init__(self, cfg: Optional[CPConfig] = None, weights: Optional[Dict[str, np.ndarray]] = None, device: int = 0,
                 seed: int = 0, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        if cfg is None:
            cfg = CPConfig.from_weights(weights) if weights is not None else CPConfig()
        self.cfg = cfg
        self._h = self.lib.cp_create(cfg.to_json().encode(), device)
        if not self._h:
            raise CodePredictorError(CP_E_CUDA, (self.lib.cp_last_error(None) or b"cp_create failed").decode())
        if weights is None:
            weights = init_weights(cfg, seed)
        try:
            for name, shape in weight_shapes(cfg).items():
                w = np.ascontiguousarray(weights[name], dtype=np.float32)
                if tuple(w.shape) != tuple(shape):
                    raise ValueError(f"{name}: shape {w.shape}, expected {shape}")
                self._ck(self.lib.cp_set_tensor(self._h, name.encode(), w.ctypes.data, w.size))
            self._ck(self.lib.cp_finalize(self._h))
        except Exception:
            self.close()
            raise

    def _ck(self, rc: int):
        if rc != 0:
            raise CodePredictorError(rc, (self.lib.cp_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.cp_launches(self._h))

    @property
    def cache_len(self) -> int:
        return int(self.lib.cp_cache_len(self._h))

    @property
    def max_batch(self) -> int:
        return int(self.lib.cp_max_batch(self._h))

    def predict_batch(self, hidden_states, code0_embeds, temperature: float = 0.1, top_k: int = 50, seeds=None) -> np.ndarray:
        """The frames of B independent streams in one launch (``cp_predict_batch``): [B, H] x 2 -> codes [B, groups]."""
        H = self.cfg.hidden
        h = np.ascontiguousarray(np.asarray(hidden_states, dtype=np.float32).reshape(-1, H))
        e = np.ascontiguousarray(np.asarray(code0_embeds, dtype=np.float32).reshape(-1, H))
        B = h.shape[0]
        if e.shape[0] != B:
            raise CodePredictorError(CP_E_INVALID, "hidden_states and code0_embeds need the same number of rows")
        sd = np.ascontiguousarray(np.arange(B) if seeds is None else seeds, dtype=np.uint64)
        if sd.size != B:
            raise CodePredictorError(CP_E_INVALID, "one seed per stream")
        codes = np.empty((B, self.cfg.groups), dtype=np.int32)
        self._ck(self.lib.cp_predict_batch(self._h, B, h.ctypes.data, e.ctypes.data, float(temperature), int(top_k),
                                           sd.ctypes.data, codes.ctypes.data))
        return codes

    def set_option(self, key: str, value: str):
        self._ck(self.lib.cp_set_option(self._h, key.encode(), value.encode()))

    @property
    def predict_path(self) -> str:
        return self.lib.cp_predict_path(self._h).decode()

    def reset(self):
        self._ck(self.lib.cp_reset(self._h))

    def step(self, hidden, position: int) -> np.ndarray:
        """``_ort_step`` (code_predictor_server.py:77-85): hidden [S, H] (or [1, S, H]) at positions position..,
        returns the final-normed hidden states [S, H]; the caches grow on the device."""
        H = self.cfg.hidden
        x = np.ascontiguousarray(np.asarray(hidden, dtype=np.float32).reshape(-1, H))
        out = np.empty_like(x)
        self._ck(self.lib.cp_step(self._h, x.ctypes.data, x.shape[0], int(position), out.ctypes.data))
        return out

    def logits(self, group: int) -> np.ndarray:
        out = np.empty(self.cfg.vocab, dtype=np.float32)
        self._ck(self.lib.cp_logits(self._h, int(group), out.ctypes.data))
        return out

    def predict(self, hidden_state, code0_embed, temperature: float = 0.1, top_k: int = 50, seed: int = 0) -> np.ndarray:
        """A whole frame in one graph launch, sampler on the device (``cp_predict``)."""
        H = self.cfg.hidden
        h, e = _f32(hidden_state, H), _f32(code0_embed, H)
        if h.size != H or e.size != H:
            raise CodePredictorError(CP_E_INVALID, f"hidden_state and code0_embed need {H} floats")
        codes = np.empty(self.cfg.groups, dtype=np.int32)
        self._ck(self.lib.cp_predict(self._h, h.ctypes.data, e.ctypes.data, float(temperature), int(top_k),
                                     int(seed) & (2 ** 64 - 1), codes.ctypes.data))
        return codes
