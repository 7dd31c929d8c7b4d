#!/usr/bin/env python3
"""Residual unit at the production shapes of decoder blocks 2 and 3 through voc_test_ru: the fused kernel (pipelined and
simple order of work) against the two tap-GEMM launches it replaces.  ms per launch on `--windows` windows, algorithmic
TFLOP/s (2 * L * C^2 * 8 per window) and the bytes the unit must move (read S and x, write x' and S': 16 B per element)."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
    from test_gpu_ru_fused import make_case
    for C, L in ((96, 122325), (192, 40776)):
        for dil in (1, 3, 9):
            k = make_case(C, L, args.windows, dil, seed=1)
            row = []
            for name, fused, fl in (("two launches", 0, 0), ("fused", 1, 0), ("fused, simple order", 1, 32)):
                rc, _, _, ms = backend.test_ru(fused, tc_flags=fl, iters=args.iters, **k)
                assert rc == 0, rc
                fl_ = 2.0 * args.windows * L * C * C * 8
                by = 16.0 * args.windows * L * C
                row.append(f"{name}: {ms:.3f} ms {fl_ / ms / 1e9:.0f} TFLOP/s {by / ms / 1e6:.0f} GB/s")
            print(f"C={C} d={dil} L={L} x{args.windows}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()
