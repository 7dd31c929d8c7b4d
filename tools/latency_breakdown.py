"""Per-layer CUDA-event breakdown of ONE 64-frame window at batch 1 (BASELINE configs[1]); run on a B200."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("qwen3-tts-axera-russian_b200")
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
cfg = pkg.VocoderConfig()
voc = backend.Vocoder(cfg, pkg.init_weights(cfg, 0), wave=1)
codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (1, 64, 16), dtype=np.int64)
for _ in range(3):
    voc.infer_chunks(codes)
voc.set_option("profile", "1"); voc.profile_report()
N = 10
for _ in range(N):
    voc.infer_chunks(codes)
rep = sorted(voc.profile_report(), key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rep)
print(f"sum of kernel times {tot / N:.3f} ms per window, {sum(r['calls'] for r in rep) // N} launches")
for r in rep:
    print(f"{r['tag']:16s} {r['ms'] / N:7.3f} ms  {r['calls'] // N:3d} launches")
