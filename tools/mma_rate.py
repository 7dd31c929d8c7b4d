"""Per-instruction cost of tcgen05.mma in the tap-GEMM kernel (run by hand on a B200): a compute-bound
Linear (M = 148*128 rows, K = 2048) for several N / K-chunk widths, accumulation segments disabled."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
rng = np.random.default_rng(0)
M, K = 148 * 128, 2048
A = rng.standard_normal((1, M, K), dtype=np.float32)
for N in (32, 64, 96, 128, 192, 384):
    W = (rng.standard_normal((K, N), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    for fl, tag in [((4000 << 8) | 2, "BK64 noflush"), ((4000 << 8) | 4, "BK32 noflush"), (2, "BK64 seg24")]:
        rc, _, _, ms = backend.test_tapgemm(2, A, W, [0], M, 0, want_y=False, want_s=False, tc_flags=fl, iters=5)
        bn = next(c for c in (192, 128, 96, 64, 32) if N % c == 0)
        tiles_per_cta = (N // bn)
        mmas = tiles_per_cta * (K // 16) * 3
        print(f"N {N:4d} BN {bn:3d} {tag:13s} rc {rc} {ms:.4f} ms  {ms * 1e6 / mmas:.1f} ns/MMA  "
              f"{2.0 * M * N * K * 3 / ms / 1e9:.0f} MMA-TFLOP/s", flush=True)
