#!/usr/bin/env python3
"""Latency of the carried-state decode for live pieces: k new frames per voc_stream_decode_pcm16 call (host codes -> host
PCM), p50 of 60 calls after 10, for k in 1, 2, 4, 8, 16, 64, with CUDA graphs on and off (option "graphs")."""
import ctypes
import dataclasses
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")

cfg = dataclasses.replace(pkg.VocoderConfig(), transconv_trim="right")
voc = backend.Vocoder(cfg, None, device=0, wave=32)
voc.set_option("gemm", "tc")
codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (8192, 16), dtype=np.int64)
out = np.empty(64 * 1920, dtype=np.int16)
cnt = ctypes.c_longlong(0)
for graphs in ("1", "0"):
    voc.set_option("graphs", graphs)
    for k in (1, 2, 4, 8, 16, 64):
        voc.stream_reset()
        ts = []
        for i in range(70):
            c = np.ascontiguousarray(codes[i * k:(i + 1) * k])
            t0 = time.perf_counter()
            rc = voc.lib.voc_stream_decode_pcm16(voc._h, c.ctypes.data, k, out.ctypes.data, out.size, ctypes.byref(cnt))
            assert rc == 0, voc.lib.voc_last_error(voc._h)
            if i >= 10:
                ts.append((time.perf_counter() - t0) * 1e3)
        print(f"graphs {graphs}  k = {k:2d} frames per call ({k * 80:4d} ms of audio): p50 {np.percentile(ts, 50):.3f} ms  "
              f"p95 {np.percentile(ts, 95):.3f} ms", flush=True)
