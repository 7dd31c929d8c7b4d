#!/usr/bin/env python3
"""Does replaying small batches as CUDA graphs pay beyond 4 windows?  p50 of the host-to-host call for B windows with
option graph_max_wave = 4 (default) and 32, alternating in one process."""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")

cfg = pkg.VocoderConfig()
voc = backend.Vocoder(cfg, None, device=0, wave=32)
voc.set_option("gemm", "tc")
rng = np.random.default_rng(0)
for B in (2, 4, 8, 16, 32):
    codes = rng.integers(0, cfg.codebook_size, (B, cfg.chunk_frames, 16), dtype=np.int64)
    for rep in range(2):
        for gmw in ("4", "32"):
            voc.set_option("graph_max_wave", gmw)
            ts = []
            for i in range(30):
                t0 = time.perf_counter()
                voc.infer_chunks(codes)
                if i >= 6:
                    ts.append((time.perf_counter() - t0) * 1e3)
            print(f"B = {B:2d}  graph_max_wave {gmw:>2s}: p50 {np.percentile(ts, 50):.3f} ms", flush=True)
