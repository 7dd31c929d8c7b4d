#!/usr/bin/env python3
"""Where the roles of the fused residual-unit kernel spend their cycles (needs the -DVOC_TC_PROF build:
tools/build_prof.sh, then VOC_LIB=build/libvoc_prof.so python tools/ru_prof.py C dil windows [tc_flags])."""
import ctypes as C
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
lib = backend.load_library(os.environ.get("VOC_LIB") and os.path.join(ROOT, os.environ["VOC_LIB"]))
backend._lib = lib
from test_gpu_ru_fused import make_case

NAMES = ["mma_total", "mma_wait_a", "mma_wait_b", "mma_wait_acc", "mma_wait_T", "epi_total", "epi_wait_acc7", "epi_wait_acc1",
         "epi_emit", "epi_final", "prod_total", "prod_wait_a", "prod_wait_b", "ctas", "tiles"]


def read(reset=1):
    buf = (C.c_ulonglong * 16)()
    lib.voc_ru_prof_read.restype = C.c_int
    lib.voc_ru_prof_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
    n = lib.voc_ru_prof_read(buf, 16, reset)
    return {NAMES[i]: int(buf[i]) for i in range(min(n, len(NAMES)))}


def main():
    Cc, dil, win = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    fl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    L = {96: 122325, 192: 40776}[Cc]
    k = make_case(Cc, L, win, dil, seed=1)
    backend.test_ru(1, tc_flags=fl, iters=0, **k)
    read()
    rc, _, _, ms = backend.test_ru(1, tc_flags=fl, iters=4, **k)
    p = read()
    launches = 5
    ctas = p["ctas"] / launches                     # leader CTAs per launch
    tiles = p["tiles"] / launches
    per = lambda key, n: p[key] / launches / n
    print(f"C={Cc} d={dil} x{win} flags={fl}: {ms:.3f} ms/launch, {tiles:.0f} tile pairs on {ctas:.0f} clusters "
          f"({tiles / ctas:.1f} per cluster)")
    tp = tiles / ctas
    print(f"  cycles per tile pair (leader's MMA warp): total {per('mma_total', ctas) / tp:.0f}  wait A {per('mma_wait_a', ctas) / tp:.0f}  "
          f"wait B {per('mma_wait_b', ctas) / tp:.0f}  wait TMEM buffer {per('mma_wait_acc', ctas) / tp:.0f}  wait T tile {per('mma_wait_T', ctas) / tp:.0f}")
    e = 2 * ctas
    print(f"  epilogue warp: total {per('epi_total', e) / tp:.0f}  wait conv7 acc {per('epi_wait_acc7', e) / tp:.0f}  "
          f"wait conv1 acc {per('epi_wait_acc1', e) / tp:.0f}  emit T {per('epi_emit', e) / tp:.0f}  final (incl. its wait) {per('epi_final', e) / tp:.0f}")
    print(f"  producer: total {per('prod_total', e) / tp:.0f}  wait A slot {per('prod_wait_a', e) / tp:.0f}  wait B slot {per('prod_wait_b', e) / tp:.0f}")


if __name__ == "__main__":
    main()
