"""Bring-up probe (run by hand on a B200): error and signed bias of the tensor-core tap-GEMM
against float64 for several accumulation-segment lengths (tc_flags bits 8..)."""
import importlib, sys, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
from test_gpu_tapgemm import ref_tapgemm, CASES
for name, B, a_rows, K, N, M, row0, taps in CASES:
    rng = np.random.default_rng(1)
    A = rng.standard_normal((B, a_rows, K)).astype(np.float32)
    W = (rng.standard_normal((len(taps) * K, N)) / np.sqrt(len(taps) * K)).astype(np.float32)
    v_ref, _ = ref_tapgemm(A, W, taps, M, row0)
    out = []
    for mode, fl in [(0, 0), (1, 0), (2, 0), (2, 1), (2, 3 << 8), (2, 24 << 8), (2, 48 << 8), (2, 4000 << 8)]:
        rc, Y, _, _ = backend.test_tapgemm(mode, A, W, taps, M, row0, tc_flags=fl)
        e = np.abs(Y - v_ref)
        out.append(f"m{mode}f{fl}:{e.max():.2e}/{e.mean():.1e}/b{((Y - v_ref) * np.sign(v_ref)).mean():+.1e}")
    print(name, " ".join(out), flush=True)
