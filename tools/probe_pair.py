"""cta_group::2 bring-up probe (run by hand on a B200): pair mode (default; tc_flags bit 7 disables it) against float64 and against
the single-CTA kernel, then timing on the wide decoder layers."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
from test_gpu_tapgemm import ref_tapgemm
CASES = [("linear", 1, 300, 128, 256, 300, 0, [0]), ("linear_odd", 2, 700, 192, 192, 700, 0, [0]),
         ("conv7_d3", 2, 333, 192, 192, 333, 0, [-18, -15, -12, -9, -6, -3, 0]),
         ("conv7_d9", 1, 400, 128, 384, 400, 0, [-54, -45, -36, -27, -18, -9, 0]),
         ("convt", 2, 300, 192, 384, 299, 1, [0, -1])]
for name, B, a_rows, K, N, M, row0, taps in CASES:
    rng = np.random.default_rng(1)
    A = rng.standard_normal((B, a_rows, K)).astype(np.float32)
    W = (rng.standard_normal((len(taps) * K, N)) / np.sqrt(len(taps) * K)).astype(np.float32)
    R = rng.standard_normal((B, M, N)).astype(np.float32)
    v_ref, _ = ref_tapgemm(A, W, taps, M, row0, R=R)
    for fl in (128, 0):
        rc, Y, _, _ = backend.test_tapgemm(2, A, W, taps, M, row0, R=R, tc_flags=fl)
        print(name, "flags", fl, "rc", rc, "max err", float(np.abs(Y - v_ref).max()), flush=True)
if "--time" in sys.argv:
    import subprocess
    subprocess.call([sys.executable, os.path.join(os.path.dirname(__file__), "gemm_bench.py"), "--windows", "4", "--flags", "128", "0",
                     "--layers", "dec0.c7d1", "dec1.c7d1", "dec2.c7d1", "dec0.convt", "dec1.convt", "dec2.convt", "conv_in", "dec0.c1", "dec1.c1", "dec2.c1", "--check"])
