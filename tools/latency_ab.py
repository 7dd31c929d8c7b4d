#!/usr/bin/env python3
"""Batch-1 latency (BASELINE configs[1]: host codes -> host float32 window, p50 / p95 of 200 calls) for several
settings of tc_flags in ONE process on ONE box, and the bits of the window under each:
    python tools/latency_ab.py 256 0        # widest column tiles (the pre-adaptive choice) against the launcher's own"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")

flag_sets = [int(x) for x in sys.argv[1:]] or [256, 0]
cfg = pkg.VocoderConfig()
voc = backend.Vocoder(cfg, None, device=0, wave=32)
voc.set_option("gemm", "tc")
codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, cfg.chunk_frames, 16), dtype=np.int64)
outs = {}
for rep in range(2):
    for fl in flag_sets:
        voc.set_option("tc_flags", str(fl))
        for _ in range(20):
            out = voc.infer_chunks(codes)
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            out = voc.infer_chunks(codes)
            ts.append((time.perf_counter() - t0) * 1e3)
        outs[fl] = out.copy()
        print(f"tc_flags {fl:4d}: p50 {np.percentile(ts, 50):.3f} ms  p95 {np.percentile(ts, 95):.3f} ms  "
              f"min {min(ts):.3f} ms", flush=True)
ref = outs[flag_sets[0]]
for fl in flag_sets[1:]:
    print(f"tc_flags {fl} vs {flag_sets[0]}: bit-equal {np.array_equal(ref, outs[fl])}")
