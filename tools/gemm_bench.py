"""Per-layer micro-benchmark of the tap-GEMM kernels on the real layer shapes of the decoder
(one wave of `--windows` 64-frame windows), through voc_test_tapgemm.  Prints algorithmic
TFLOP/s (2*M*N*K*taps) and, for the tensor path, the MMA rate (3 passes per product).

    python tools/gemm_bench.py [--windows 4] [--modes 2 0] [--flags 0] [--layers dec3.c7d1 ...]
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def conv_taps(k, d):
    return [-(k - 1 - j) * d for j in range(k)]


def layers(trim_both=True):
    L = [256]
    for s in (8, 5, 4, 3):
        L.append((L[-1] - 1) * s if trim_both else L[-1] * s)
    C = [1536, 768, 384, 192, 96]
    out = [("conv_in", 256, 1024, 1536, 256, 0, conv_taps(7, 1), "s")]
    for b, s in enumerate((8, 5, 4, 3)):
        M = L[b + 1] // s
        out.append((f"dec{b}.convt", L[b], C[b], s * C[b + 1], M, 1 if trim_both else 0, [0, -1], "ys"))
        for d in (1, 3, 9):
            out.append((f"dec{b}.c7d{d}", L[b + 1], C[b + 1], C[b + 1], L[b + 1], 0, conv_taps(7, d), "s"))
        out.append((f"dec{b}.c1", L[b + 1], C[b + 1], C[b + 1], L[b + 1], 0, [0], "rys"))
    out.append(("up.pw1", 256, 1024, 4096, 256, 0, [0], "s"))
    out.append(("up.pw2", 256, 4096, 1024, 256, 0, [0], "rs"))
    out.append(("xf.qkv", 64, 512, 3072, 64, 0, [0], "y"))
    return out


PROF_NAMES = ["mma_total", "mma_w_acc", "mma_w_a", "mma_w_b", "epi_total", "epi_w_acc", "epi_drain", "epi_final",
              "prod_total", "prod_w_a", "prod_w_b", "ctas", "tiles", "segs"]


def read_prof(backend, iters):
    """Cycle counters of a -DVOC_TC_PROF build (tools/ab_build.sh), per tile; None for a production build."""
    import ctypes as C
    lib = backend.load_library()
    if not hasattr(lib, "voc_tc_prof_read"):
        return None
    buf = (C.c_ulonglong * 16)()
    lib.voc_tc_prof_read.restype = C.c_int
    lib.voc_tc_prof_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int, C.c_int]
    if lib.voc_tc_prof_read(buf, 16, 1) < 0:
        return None
    v = dict(zip(PROF_NAMES, [int(x) for x in buf]))
    tiles = max(v["tiles"], 1)                    # tiles walked by the issuing warps (pairs count once)
    ctas = max(v["ctas"], 1)
    out = {"tiles_per_walker": round(tiles / ctas, 2), "segs_per_tile": round(v["segs"] / max(tiles, 1), 2)}
    for k in PROF_NAMES[:11]:
        out[k] = round(v[k] / tiles)              # cycles per tile (epilogue: one warp per CTA sampled; pairs: both CTAs summed)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=4)
    ap.add_argument("--modes", type=int, nargs="*", default=[2])
    ap.add_argument("--flags", type=int, nargs="*", default=[0])
    ap.add_argument("--layers", nargs="*", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="compare mode 2 against mode 0 on the same data")
    args = ap.parse_args()
    backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
    rng = np.random.default_rng(0)
    rows = []
    for name, a_rows, K, N, M, row0, taps, outs in layers():
        if args.layers and not any(name.startswith(p) for p in args.layers):
            continue
        B = args.windows
        A = rng.standard_normal((B, a_rows, K), dtype=np.float32)
        W = (rng.standard_normal((len(taps) * K, N), dtype=np.float32) / np.sqrt(len(taps) * K)).astype(np.float32)
        bias = (0.1 * rng.standard_normal(N)).astype(np.float32)
        R = rng.standard_normal((B, M, N), dtype=np.float32) if "r" in outs else None
        sn_a = np.ones(N, dtype=np.float32)
        sn_b = np.ones(N, dtype=np.float32)
        flops = 2.0 * B * M * N * K * len(taps)
        ref = None
        for mode in args.modes:
            for fl in (args.flags if mode == 2 else [0]):
                rc, Y, S, ms = backend.test_tapgemm(mode, A, W, taps, M, row0, bias=bias, R=R, sn_a=sn_a, sn_invb=sn_b,
                                                    want_y="y" in outs, want_s="s" in outs, tc_flags=fl,
                                                    iters=args.iters)
                rec = {"layer": name, "mode": mode, "flags": fl, "rc": rc, "B": B, "M": M, "N": N, "K": K,
                       "taps": len(taps), "ms": round(ms, 4),
                       "tflops": round(flops / ms / 1e9, 2) if ms > 0 else None}
                if mode == 2 and ms > 0:
                    rec["mma_tflops"] = round(3 * flops / ms / 1e9, 1)
                prof = read_prof(backend, args.iters)
                if prof:
                    rec["prof"] = prof
                out = S if S is not None else Y
                if args.check and out is not None:
                    if ref is None:
                        ref = out
                    else:
                        rec["max_abs_vs_first"] = float(np.abs(out - ref).max())
                print(json.dumps(rec), flush=True)
                rows.append(rec)
    return rows


if __name__ == "__main__":
    main()
