"""Operand-format study for the tensor-core path (TEST INFRASTRUCTURE; run by hand).

Emulates, on the CPU oracle, what the tcgen05 kernels do when every dense contraction's
operands are split into 16-bit pieces and multiplied in several passes with FP32
accumulation, and reports SNR / max-abs against the plain FP32 oracle and the FP64 oracle.
The gate (BASELINE.json north_star): SNR >= 60 dB and max-abs <= 1e-4.

    python -m oracle.precision_study [--frames 64] [--modes bf16x1 tf32x1 bf16x3 fp16x3]
"""
from __future__ import annotations

import argparse
import sys
import time
import types

import numpy as np
import torch
import torch.nn.functional as TF

from . import vocoder_oracle as O


def _split(x: torch.Tensor, kind: str):
    if kind == "bf16":
        hi = x.to(torch.bfloat16).to(torch.float32)
        lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    elif kind == "fp16":
        hi = x.to(torch.float16).to(torch.float32)
        lo = (x - hi).to(torch.float16).to(torch.float32)
    elif kind == "tf32":
        # round-to-nearest-even to 10 explicit mantissa bits
        i = x.contiguous().view(torch.int32)
        r = ((i >> 13) & 1) + 0x0FFF
        hi = ((i + r) & ~0x1FFF).view(torch.float32)
        lo = x - hi
        i2 = lo.contiguous().view(torch.int32)
        r2 = ((i2 >> 13) & 1) + 0x0FFF
        lo = ((i2 + r2) & ~0x1FFF).view(torch.float32)
    else:
        raise ValueError(kind)
    return hi, lo


def _round_mant(x: torch.Tensor, bits: int) -> torch.Tensor:
    """Round-to-nearest-even to `bits` explicit mantissa bits, unlimited exponent range: the BEST case of an 8-bit
    float operand (e4m3: bits = 3, e5m2: bits = 2) under ideal power-of-two scaling of the plane."""
    sh = 23 - bits
    i = x.contiguous().view(torch.int32)
    r = ((i >> sh) & 1) + ((1 << (sh - 1)) - 1)
    return ((i + r) & ~((1 << sh) - 1)).view(torch.float32)


class SplitF(types.SimpleNamespace):
    """Stand-in for torch.nn.functional inside the oracle: dense ops use split operands."""

    def __init__(self, kind: str, passes: int, cross_bits: int = 0):
        super().__init__()
        self.kind, self.passes, self.cross_bits = kind, passes, cross_bits
        for name in ("pad", "layer_norm", "gelu", "silu"):
            setattr(self, name, getattr(TF, name))

    def _mm(self, fn, x, w, b, **kw):
        xh, xl = _split(x, self.kind)
        wh, wl = _split(w, self.kind)
        y = fn(xh, wh, None, **kw)
        if self.passes == 2:            # activations split, weights rounded once: (xh + xl) * wh
            y = y + fn(xl, wh, None, **kw)
        if self.passes >= 3 and self.cross_bits:
            # the two cross terms as 8-bit-float MMAs (kind::f8f6f4): both factors of hi*lo and
            # lo*hi carry only `cross_bits` mantissa bits
            q = lambda t: _round_mant(t, self.cross_bits)
            y = y + fn(q(xh), q(wl), None, **kw) + fn(q(xl), q(wh), None, **kw)
        elif self.passes >= 3:
            y = y + fn(xh, wl, None, **kw) + fn(xl, wh, None, **kw)
        if self.passes >= 4:
            y = y + fn(xl, wl, None, **kw)
        if b is not None:
            y = y + (b.view(1, -1, 1) if fn is not TF.linear else b)
        return y

    def conv1d(self, x, w, b=None, dilation=1, groups=1):
        if groups != 1:                       # depth-wise conv stays on CUDA cores in FP32
            return TF.conv1d(x, w, b, dilation=dilation, groups=groups)
        return self._mm(TF.conv1d, x, w, b, dilation=dilation)

    def conv_transpose1d(self, x, w, b=None, stride=1):
        return self._mm(TF.conv_transpose1d, x, w, b, stride=stride)

    def linear(self, x, w, b=None):
        return self._mm(TF.linear, x, w, b)


def run(cfg, weights, codes, mode: str) -> np.ndarray:
    cross = 0
    if "c" in mode[4:]:                      # e.g. fp16x3c3: three passes, cross terms with 3 mantissa bits (e4m3)
        mode, cross = mode[:mode.rindex("c")], int(mode[mode.rindex("c") + 1:])
    kind, passes = mode[:-2], int(mode[-1])
    saved = O.F
    O.F = SplitF(kind, passes, cross)
    try:
        a, _ = O.forward(codes, O.Weights(weights, torch.float32), cfg)
    finally:
        O.F = saved
    return a.numpy()


def main(argv=None):
    sys.path.insert(0, ".")
    import voc_b200 as V
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--modes", nargs="*", default=["bf16x1", "tf32x1", "bf16x3", "fp16x3", "tf32x3"])
    ap.add_argument("--trim", default="both")
    args = ap.parse_args(argv)
    cfg = V.VocoderConfig(chunk_frames=args.frames, transconv_trim=args.trim)
    w = V.init_weights(cfg, 0)
    codes = np.random.default_rng(1).integers(0, cfg.codebook_size, (1, args.frames, 16), dtype=np.int64)
    t = time.time()
    ref32, _ = O.forward(codes, O.Weights(w, torch.float32), cfg)
    ref32 = ref32.numpy()
    print(f"fp32 oracle {time.time() - t:.2f}s  rms {np.sqrt((ref32 ** 2).mean()):.4f} "
          f"clamp {(np.abs(ref32) >= 1).mean():.5f}")
    ref64, _ = O.forward(codes, O.Weights(w, torch.float64), cfg)
    ref64 = ref64.numpy()
    print(f"fp32 vs fp64: snr {O.snr_db(ref64, ref32):.1f} dB  max-abs {np.abs(ref64 - ref32).max():.2e}")
    for m in args.modes:
        a = run(cfg, w, codes, m)
        print(f"{m:8s} vs fp32: snr {O.snr_db(ref32, a):6.1f} dB max-abs {np.abs(ref32 - a).max():.2e}"
              f"   vs fp64: snr {O.snr_db(ref64, a):6.1f} dB max-abs {np.abs(ref64 - a).max():.2e}")


if __name__ == "__main__":
    main()
