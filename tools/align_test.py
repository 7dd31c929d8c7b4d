"""Does a row-shifted (not 8-row aligned) A descriptor cost extra operand-fetch time?  Same 7-tap conv with
dilation 8 (all tap shifts multiples of the swizzle atom) vs 9 vs 1; run by hand on a B200."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
rng = np.random.default_rng(0)
for (C, L) in [(96, 122325), (192, 40776)]:
    A = rng.standard_normal((4, L, C), dtype=np.float32)
    W = (rng.standard_normal((7 * C, C), dtype=np.float32) / np.sqrt(7 * C)).astype(np.float32)
    for d in (8, 9, 1, 16):
        taps = [-(6 - j) * d for j in range(7)]
        for fl in (0, 128):
            rc, _, _, ms = backend.test_tapgemm(2, A, W, taps, L, 0, want_y=False, want_s=False, tc_flags=fl, iters=5)
            print(f"C {C} dilation {d:2d} flags {fl:3d} rc {rc} {ms:.4f} ms", flush=True)
