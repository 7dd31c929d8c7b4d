#!/usr/bin/env python3
"""One residual unit (production shape) for profiling: python tools/ru_one.py C dil windows fused [tc_flags]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
if os.environ.get("VOC_LIB"):                      # an A/B build (build/libvoc_*.so)
    backend._lib = backend.load_library(os.path.join(ROOT, os.environ["VOC_LIB"]))
from test_gpu_ru_fused import make_case
C, dil, win, fused = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
L = {96: 122325, 192: 40776}.get(C, 10195)
k = make_case(C, L, win, dil, seed=1)
rc, _, _, ms = backend.test_ru(fused, tc_flags=fl, iters=3, **k)
print(f"C={C} d={dil} x{win} fused={fused} flags={fl}: rc {rc}  {ms:.3f} ms")
