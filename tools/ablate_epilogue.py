"""Ablation of the tensor-core kernel's epilogue on one layer shape (run by hand on a B200):
which of {residual load, float32 store, Snake, operand store} costs the time."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
rng = np.random.default_rng(0)
for (name, B, L, K, N, taps) in [("dec3.c1", 4, 122325, 96, 96, [0]), ("dec3.c7", 4, 122325, 96, 96, [-6, -5, -4, -3, -2, -1, 0]),
                                 ("dec2.c1", 4, 40776, 192, 192, [0])]:
    A = rng.standard_normal((B, L, K), dtype=np.float32)
    W = (rng.standard_normal((len(taps) * K, N), dtype=np.float32) / np.sqrt(len(taps) * K)).astype(np.float32)
    bias = np.zeros(N, np.float32)
    R = rng.standard_normal((B, L, N), dtype=np.float32)
    one = np.ones(N, np.float32)
    flags_list = [int(x) for x in sys.argv[1:]] or [0]
    for fl in flags_list:
      for tag, kw in [("R+Y+snake+S", dict(R=R, sn_a=one, sn_invb=one, want_y=True, want_s=True)),
                    ("Y+snake+S", dict(sn_a=one, sn_invb=one, want_y=True, want_s=True)),
                    ("R+Y", dict(R=R, want_y=True, want_s=False)),
                    ("Y", dict(want_y=True, want_s=False)),
                    ("snake+S", dict(sn_a=one, sn_invb=one, want_y=False, want_s=True)),
                    ("S plain", dict(want_y=False, want_s=True)),
                    ("nothing", dict(want_y=False, want_s=False))]:
        if len(flags_list) > 1 and tag not in ("R+Y+snake+S", "snake+S", "nothing"):
            continue
        rc, _, _, ms = backend.test_tapgemm(2, A, W, taps, L, 0, bias=bias, iters=5, tc_flags=fl, **kw)
        print(f"{name:8s} flags {fl:5d} {tag:12s} rc {rc} {ms:.4f} ms", flush=True)
