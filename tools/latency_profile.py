#!/usr/bin/env python3
"""Per-layer CUDA-event times of ONE window at batch 1 (the latency configuration, BASELINE configs[1])."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")

cfg = pkg.VocoderConfig()
voc = backend.Vocoder(cfg, None, device=0, wave=32)
codes = np.random.default_rng(0).integers(0, cfg.codebook_size, (1, cfg.chunk_frames, 16), dtype=np.int64)
for _ in range(5):
    voc.infer_chunks(codes)
voc.set_option("profile", "1")
voc.profile_report()
N = 20
for _ in range(N):
    voc.infer_chunks(codes)
prof = voc.profile_report()
tot = sum(p["ms"] for p in prof) / N
print(f"one window, batch 1: {tot:.3f} ms of kernels, {sum(p['calls'] for p in prof) // N} launches")
for p in sorted(prof, key=lambda q: -q["ms"]):
    print(f"  {p['tag']:18s} {p['ms'] / N * 1e3:8.1f} us  {p['calls'] // N:3d} launches  {p['ms'] / p['calls'] * 1e3:6.1f} us each")
