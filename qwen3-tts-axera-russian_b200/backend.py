"""ctypes binding of ``libvoc_b200.so`` (the C ABI in ``include/voc_b200.h``).

Loaded the way the reference loads its own native engine -- ``ctypes.CDLL`` on a shared
library that sits next to the Python file (``/root/reference/dual_npu/llama_cpp_bindings.py:
41-81``).  There is no CPU fallback: if the library is missing, or no sm_100 device is
usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Tuple

import numpy as np

from .config import VocoderConfig
from .weights import init_weights, load_model, weight_shapes

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libvoc_b200.so"
LIB_PATH = os.path.join(_HERE, LIB_NAME)

VOC_OK, VOC_E_INVALID, VOC_E_CUDA, VOC_E_STATE, VOC_E_NOMEM = 0, -1, -2, -3, -4

# name -> (restype, argtypes); the single source the symbol-export test checks against the header
SIGNATURES = {
    "voc_abi_version": (C.c_int, []),
    "voc_create": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int]),
    "voc_create_from_file": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int]),
    "voc_destroy": (None, [C.c_void_p]),
    "voc_set_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "voc_finalize": (C.c_int, [C.c_void_p]),
    "voc_max_tokens": (C.c_int, [C.c_void_p]),
    "voc_chunk_samples": (C.c_longlong, [C.c_void_p]),
    "voc_out_samples": (C.c_longlong, [C.c_void_p, C.c_int]),
    "voc_num_windows": (C.c_int, [C.c_void_p, C.c_int]),
    "voc_infer_chunks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "voc_infer_chunks_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "voc_synthesize_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                     C.POINTER(C.c_longlong)]),
    "voc_synthesize_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                       C.POINTER(C.c_longlong)]),
    "voc_synthesize_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_longlong, C.POINTER(C.c_longlong), C.c_void_p]),
    "voc_synthesize_range_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_longlong,
                                           C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                           C.c_void_p]),
    "voc_synthesize_batch_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                             C.c_void_p]),
    "voc_check_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "voc_stream_reset": (C.c_int, [C.c_void_p]),
    "voc_stream_position": (C.c_longlong, [C.c_void_p]),
    "voc_stream_decode_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]),
    "voc_stream_decode_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong)]),
    "voc_plan": (C.c_int, [C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                           C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "voc_fade_tables": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p]),
    "voc_last_error": (C.c_char_p, [C.c_void_p]),
    "voc_kernel_launches": (C.c_longlong, [C.c_void_p]),
    "voc_simt_launches": (C.c_longlong, [C.c_void_p]),
    "voc_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "voc_stream": (C.c_void_p, [C.c_void_p]),
    "voc_profile_report": (C.c_longlong, [C.c_void_p, C.c_void_p, C.c_longlong]),
    "voc_operand_report": (C.c_longlong, [C.c_void_p, C.c_void_p, C.c_longlong]),
    "voc_tc_plan": (C.c_int, [C.c_int] * 7 + [C.c_void_p]),
    "voc_ru_plan": (C.c_int, [C.c_int] * 3 + [C.c_void_p]),
    "voc_debug_stage": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_longlong]),
    "voc_test_tapgemm": (C.c_int, [C.c_int] * 10 + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 5
                         + [C.c_int, C.c_void_p]),
    "voc_test_ru": (C.c_int, [C.c_int] * 8 + [C.c_void_p] * 12 + [C.c_int, C.c_void_p]),
}

_lib = None


class VocoderError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"voc_b200 error {code}: {msg}")
        self.code = code


def load_library(path: Optional[str] = None):
    """``ctypes.CDLL`` the backend; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The B200 vocoder backend has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def plan(max_tokens: int, chunk_samples: int, n_tokens: int):
    """Host-side window plan of ``synthesize`` (no GPU needed).

    Returns (meta int32 [n_windows, 6], total_samples, pairwise) with meta columns
    {dst, a_len, blended, next_blended, prev_a_len, start_frame}."""
    lib = load_library()
    total, pw = C.c_longlong(0), C.c_int(0)
    nw = lib.voc_plan(max_tokens, chunk_samples, n_tokens, 0, None, C.byref(total), C.byref(pw))
    if nw < 0:
        raise VocoderError(nw, "voc_plan: bad argument")
    meta = np.zeros((nw, 6), dtype=np.int32)
    lib.voc_plan(max_tokens, chunk_samples, n_tokens, meta.size, meta.ctypes.data, C.byref(total), C.byref(pw))
    return meta, total.value, bool(pw.value)


def fade_tables(ov: int):
    lib = load_library()
    fo = np.empty(ov, dtype=np.float32)
    fi = np.empty(ov, dtype=np.float32)
    rc = lib.voc_fade_tables(ov, fo.ctypes.data, fi.ctypes.data)
    if rc:
        raise VocoderError(rc, "voc_fade_tables: bad argument")
    return fo, fi


def tc_plan(N: int, K: int, ntaps: int, M: int, B: int, sms: int = 148, tc_flags: int = 0) -> dict:
    """The tcgen05 kernel's tile plan for a dense layer and a batch (host arithmetic; no GPU needed)."""
    lib = load_library()
    out = (C.c_int * 5)()
    rc = lib.voc_tc_plan(N, K, ntaps, M, B, sms, tc_flags, out)
    if rc:
        raise ValueError(f"no tensor-core tile for N = {N}")
    return {"BN": out[0], "BK": out[1], "pair": bool(out[2]), "p3": bool(out[3]), "three_pass": bool(out[4])}


def ru_plan(Cc: int, ksz: int, dil: int) -> dict:
    """Shared-memory plan of the fused residual-unit kernel (host arithmetic; no GPU needed)."""
    lib = load_library()
    out = (C.c_int * 5)()
    if lib.voc_ru_plan(Cc, ksz, dil, out):
        raise ValueError(f"the fused residual unit does not take C = {Cc}, k = {ksz}, dilation {dil}")
    return {"box_rows": out[0], "halo_stages": out[1], "weight_stages": out[2], "alias": bool(out[3]), "smem": out[4]}


def test_tapgemm(mode: int, A: np.ndarray, W: np.ndarray, tap_off, M: int, a_row0: int = 0, bias=None,
                 scale=None, act: int = 0, R=None, sn_a=None, sn_invb=None, want_y: bool = True,
                 want_s: bool = False, tc_flags: int = 0, iters: int = 0, device: int = 0):
    """One tap-GEMM through ``voc_test_tapgemm`` (see include/voc_b200.h).  A [B, a_rows, K] f32,
    W [ntaps*K, N] f32.  Returns (rc, Y or None, S or None, ms)."""
    lib = load_library()
    A = np.ascontiguousarray(A, dtype=np.float32)
    W = np.ascontiguousarray(W, dtype=np.float32)
    B, a_rows, K = A.shape
    ntaps = len(tap_off)
    N = W.shape[1]
    assert W.shape[0] == ntaps * K
    to = np.asarray(tap_off, dtype=np.int32)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    bias, scale, R, sn_a, sn_invb = f(bias), f(scale), f(R), f(sn_a), f(sn_invb)
    Y = np.zeros((B, M, N), dtype=np.float32) if want_y else None
    S = np.zeros((B, M, N), dtype=np.float32) if want_s else None
    ms = C.c_float(0.0)
    rc = lib.voc_test_tapgemm(device, mode, tc_flags, B, a_rows, K, N, M, a_row0, ntaps, to.ctypes.data,
                              A.ctypes.data, W.ctypes.data, _ptr(bias), _ptr(scale), act, _ptr(R), _ptr(sn_a),
                              _ptr(sn_invb), _ptr(Y), _ptr(S), iters, C.addressof(ms))
    return rc, Y, S, ms.value


def test_ru(fused: int, A: np.ndarray, W7: np.ndarray, b7, sn2_a, sn2_invb, W1: np.ndarray, b1, R: np.ndarray,
            snn_a, snn_invb, dil: int, ksz: int = 7, want_y: bool = True, tc_flags: int = 0, iters: int = 0,
            device: int = 0):
    """One residual unit through ``voc_test_ru``: A (= Snake1(x)) and R (= x) are [B, L, C] f32.
    Returns (rc, Y or None, S, ms)."""
    lib = load_library()
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    A, W7, W1, R = f(A), f(W7), f(W1), f(R)
    b7, sn2_a, sn2_invb, b1, snn_a, snn_invb = map(f, (b7, sn2_a, sn2_invb, b1, snn_a, snn_invb))
    B, L, Cc = A.shape
    assert W7.shape == (ksz * Cc, Cc) and W1.shape == (Cc, Cc) and R.shape == A.shape
    Y = np.zeros_like(A) if want_y else None
    S = np.zeros_like(A)
    ms = C.c_float(0.0)
    rc = lib.voc_test_ru(device, fused, tc_flags, B, L, Cc, ksz, dil, A.ctypes.data, W7.ctypes.data, b7.ctypes.data,
                         sn2_a.ctypes.data, sn2_invb.ctypes.data, W1.ctypes.data, b1.ctypes.data, R.ctypes.data,
                         snn_a.ctypes.data, snn_invb.ctypes.data, _ptr(Y), S.ctypes.data, iters, C.addressof(ms))
    return rc, Y, S, ms.value


def _ptr(a) -> int:
    """Address of a numpy array, a torch tensor (host or device) or a raw int."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Vocoder:
    """One handle = one GPU (SURVEY 8b "Ownership").  Owns weights, workspaces and a stream;
    the caller owns code and output buffers."""

    def __init__(self, cfg: Optional[VocoderConfig] = None, weights: Optional[Dict[str, np.ndarray]] = None,
                 device: int = 0, wave: int = 32, seed: int = 0, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        self.cfg = cfg or VocoderConfig()
        self.device = device
        self.wave = wave
        self._h = self.lib.voc_create(self.cfg.to_json().encode(), device, wave)
        if not self._h:
            raise VocoderError(VOC_E_CUDA, (self.lib.voc_last_error(None) or b"voc_create failed").decode())
        if weights is None:
            weights = init_weights(self.cfg, seed)
        try:
            for name, shape in weight_shapes(self.cfg).items():
                w = np.ascontiguousarray(weights[name], dtype=np.float32)
                if tuple(w.shape) != tuple(shape):
                    raise ValueError(f"{name}: shape {w.shape}, expected {shape}")
                self._ck(self.lib.voc_set_tensor(self._h, name.encode(), w.ctypes.data, w.size))
            self._ck(self.lib.voc_finalize(self._h))
        except Exception:
            self.close()
            raise
        self.max_tokens = self.lib.voc_max_tokens(self._h)
        self.chunk_samples = self.lib.voc_chunk_samples(self._h)

    @classmethod
    def from_file(cls, path: str, **kw) -> "Vocoder":
        cfg, w = load_model(path)
        return cls(cfg, w, **kw)

    # ---- plumbing ----
    def _ck(self, rc: int):
        if rc != 0:
            raise VocoderError(rc, (self.lib.voc_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.voc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: str):
        self._ck(self.lib.voc_set_option(self._h, key.encode(), value.encode()))

    @property
    def kernel_launches(self) -> int:
        return self.lib.voc_kernel_launches(self._h)

    @property
    def simt_launches(self) -> int:
        """Request-path dense-layer launches that left the tcgen05 kernel family (0 on the production model)."""
        return self.lib.voc_simt_launches(self._h)

    def out_samples(self, n_tokens: int) -> int:
        return self.lib.voc_out_samples(self._h, n_tokens)

    def num_windows(self, n_tokens: int) -> int:
        return self.lib.voc_num_windows(self._h, n_tokens)

    # ---- level 1: chunk interface (vocoder_server.py:67-71) ----
    def infer_chunks(self, codes: np.ndarray) -> np.ndarray:
        """int64 [B, max_tokens, 16] host array -> float32 [B, chunk_samples] host array."""
        codes = np.ascontiguousarray(codes, dtype=np.int64)
        if codes.ndim != 3 or codes.shape[1] != self.max_tokens or codes.shape[2] != 16:
            raise ValueError(f"codes must be [B, {self.max_tokens}, 16], got {codes.shape}")
        out = np.empty((codes.shape[0], self.chunk_samples), dtype=np.float32)
        self._ck(self.lib.voc_infer_chunks(self._h, codes.ctypes.data, codes.shape[0], out.ctypes.data))
        return out

    def infer_chunks_dev(self, d_codes, B: int, d_out, stream: int = 0):
        self._ck(self.lib.voc_infer_chunks_dev(self._h, _ptr(d_codes), B, _ptr(d_out), stream))

    # ---- level 2: whole request (vocoder_server.py:73-121,175) ----
    def synthesize(self, codes: np.ndarray) -> np.ndarray:
        codes = self._codes2d(codes)
        out = np.empty(self.out_samples(len(codes)), dtype=np.float32)
        n = C.c_longlong(0)
        self._ck(self.lib.voc_synthesize_f32(self._h, codes.ctypes.data, len(codes), out.ctypes.data,
                                             out.size, C.byref(n)))
        return out[: n.value]

    def synthesize_pcm16(self, codes: np.ndarray) -> np.ndarray:
        codes = self._codes2d(codes)
        out = np.empty(self.out_samples(len(codes)), dtype=np.int16)
        n = C.c_longlong(0)
        self._ck(self.lib.voc_synthesize_pcm16(self._h, codes.ctypes.data, len(codes), out.ctypes.data,
                                               out.size, C.byref(n)))
        return out[: n.value]

    def synthesize_batch_pcm16(self, requests):
        """Many requests ([n_i, 16] int64 arrays) in one call; returns the list of their int16 PCM arrays,
        each bit-identical to ``synthesize_pcm16`` on that request alone."""
        reqs = [self._codes2d(r) for r in requests]
        if not reqs:
            return []
        lens = np.asarray([len(r) for r in reqs], dtype=np.int32)
        codes = np.ascontiguousarray(np.concatenate(reqs, axis=0))
        cap = int(sum(self.out_samples(int(n)) for n in lens))
        out = np.empty(cap, dtype=np.int16)
        offs = np.zeros(len(reqs) + 1, dtype=np.int64)
        self._ck(self.lib.voc_synthesize_batch_pcm16(self._h, codes.ctypes.data, lens.ctypes.data, len(reqs),
                                                     out.ctypes.data, cap, offs.ctypes.data))
        return [out[offs[i]:offs[i + 1]] for i in range(len(reqs))]

    # ---- carried-state decode (opt-in) ----
    def stream_reset(self):
        self._ck(self.lib.voc_stream_reset(self._h))

    def stream_decode(self, codes: np.ndarray, pcm16: bool = False) -> np.ndarray:
        """The next `len(codes)` frames of the current sequence -> their len(codes) * 1920 samples (float32, or int16
        with the reference's truncating conversion).  Successive calls concatenate to the un-chunked decode."""
        codes = self._codes2d(codes)
        n = len(codes)
        out = np.empty(n * self.cfg.samples_per_frame, dtype=np.int16 if pcm16 else np.float32)
        cnt = C.c_longlong(0)
        fn = self.lib.voc_stream_decode_pcm16 if pcm16 else self.lib.voc_stream_decode_f32
        self._ck(fn(self._h, codes.ctypes.data, n, out.ctypes.data, out.size, C.byref(cnt)))
        return out[: cnt.value]

    def synthesize_dev(self, d_codes, n_tokens: int, d_out_f32=None, d_out_i16=None, cap: int = 0,
                       stream: int = 0) -> int:
        n = C.c_longlong(0)
        self._ck(self.lib.voc_synthesize_dev(self._h, _ptr(d_codes), n_tokens, _ptr(d_out_f32),
                                             _ptr(d_out_i16), cap, C.byref(n), stream))
        return n.value

    def synthesize_range_dev(self, d_codes, n_tokens: int, w0: int, w1: int, d_out_f32=None,
                             d_out_i16=None, cap: int = 0, stream: int = 0) -> Tuple[int, int]:
        """Windows [w0, w1) of one request -> (offset, count) of the output span they own."""
        off, n = C.c_longlong(0), C.c_longlong(0)
        self._ck(self.lib.voc_synthesize_range_dev(self._h, _ptr(d_codes), n_tokens, w0, w1,
                                                   _ptr(d_out_f32), _ptr(d_out_i16), cap,
                                                   C.byref(off), C.byref(n), stream))
        return off.value, n.value

    def synthesize_range_pcm16(self, codes: np.ndarray, w0: int, w1: int):
        """Windows [w0, w1) of one request: (offset, int16 CUDA tensor of the span they own).
        torch is used only to hold device memory."""
        import torch
        codes = self._codes2d(codes)
        dev = torch.device("cuda", self.device)
        d_codes = torch.from_numpy(codes).to(dev)
        cap = self.out_samples(len(codes))
        buf = torch.empty(cap, dtype=torch.int16, device=dev)
        off, cnt = self.synthesize_range_dev(d_codes, len(codes), w0, w1, d_out_i16=buf, cap=cap)
        self.check_dev()
        return off, buf[:cnt]

    def check_dev(self, stream: int = 0):
        """Synchronise `stream` and raise if a *_dev call met an out-of-range code."""
        self._ck(self.lib.voc_check_dev(self._h, stream))

    @property
    def stream(self) -> int:
        """The handle's cudaStream_t (as an int) for event bracketing of the host entry points."""
        return int(self.lib.voc_stream(self._h) or 0)

    def profile_report(self):
        """Aggregated per-layer CUDA-event timings since the last call (needs profile=1)."""
        import json
        n = self.lib.voc_profile_report(self._h, None, 0)
        if n < 0:
            self._ck(int(n))
        buf = C.create_string_buffer(int(n))
        n2 = self.lib.voc_profile_report(self._h, buf, n)
        if n2 < 0:
            self._ck(int(n2))
        return json.loads(buf.value.decode() or "[]")

    def operand_report(self):
        """Per-layer range statistics of the split-fp16 operands written since the last call (needs operand_stats=1):
        saturated / subnormal counts, rms and largest magnitude -- the check to run on a real checkpoint."""
        import json
        n = self.lib.voc_operand_report(self._h, None, 0)
        if n < 0:
            self._ck(int(n))
        buf = C.create_string_buffer(int(n))
        n2 = self.lib.voc_operand_report(self._h, buf, n)
        if n2 < 0:
            self._ck(int(n2))
        return json.loads(buf.value.decode() or "[]")

    def debug_stage(self, name: str) -> np.ndarray:
        n = self.lib.voc_debug_stage(self._h, name.encode(), None, 0)
        if n < 0:
            self._ck(int(n))
        out = np.empty(n, dtype=np.float32)
        n2 = self.lib.voc_debug_stage(self._h, name.encode(), out.ctypes.data, out.size)
        if n2 < 0:
            self._ck(int(n2))
        return out

    @staticmethod
    def _codes2d(codes) -> np.ndarray:
        codes = np.asarray(codes)
        if codes.ndim != 2 or codes.shape[1] < 16:
            raise ValueError(f"codes must be [n_tokens, 16], got {codes.shape}")
        # the reference keeps only the first 16 columns (vocoder_server.py:79,94)
        return np.ascontiguousarray(codes[:, :16], dtype=np.int64)
