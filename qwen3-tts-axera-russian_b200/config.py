"""Architecture description of the Qwen3-TTS 12 Hz codec vocoder (decoder side).

Every dimension the kernels, the oracle and the stitcher use is derived from this one
dataclass; nothing else in the repo hard-codes ``T * 1920``.

What the reference itself pins (paths relative to /root/reference):
  * 16 codebooks per frame, int64 codes           dual_npu/vocoder_server.py:10,78
  * code range [0, 2048)                           scripts/export_vocoder_traced.py:85
  * 1920 samples per frame at 24 kHz               dual_npu/vocoder_server.py:29-30
  * 64-frame fixed window                          dual_npu/vocoder_server.py:45-46,57
  * Snake activation, dilation-9 Conv1d            README.md:58,61
The remaining dimensions are the upstream ``qwen_tts`` decoder_config values recorded in
SURVEY.md section 8a (not verifiable offline) and are therefore configurable.
"""
from __future__ import annotations

import dataclasses
import json
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass(frozen=True)
class VocoderConfig:
    # --- RVQ front end (M1) ---
    codebook_size: int = 2048
    codebook_dim: int = 256          # E_q row width
    num_quantizers: int = 16         # 1 semantic + 15 acoustic
    num_semantic: int = 1
    rvq_dim: int = 512               # output of the two 1x1 out-projections
    # --- pre-conv (M2) ---
    latent_dim: int = 1024
    pre_conv_kernel: int = 3
    # --- pre-transformer (M3) ---
    pre_transformer: bool = True
    xf_hidden: int = 512
    xf_inter: int = 1024
    xf_layers: int = 8
    xf_heads: int = 16
    xf_head_dim: int = 64
    rope_theta: float = 10000.0
    rms_eps: float = 1e-5
    sliding_window: int = 72
    # --- upsample stages (M4) ---
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    convnext: bool = True
    convnext_mult: int = 4
    ln_eps: float = 1e-6
    # --- decoder (M5-M9) ---
    decoder_dim: int = 1536
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    dilations: Tuple[int, ...] = (1, 3, 9)
    conv_kernel: int = 7
    snake_eps: float = 1e-9
    # --- ambiguity A1: how ConvTranspose1d(k=2s, stride s) is trimmed ---
    #   "both"  : drop k-s samples from each end  -> (L-1)*s   (behaviour of the executable sibling)
    #   "right" : drop k-s samples from the right -> L*s       (length preserving, causal)
    transconv_trim: str = "both"
    # --- chunk interface (H2) ---
    chunk_frames: int = 64
    sample_rate: int = 24000

    def __post_init__(self):
        if self.transconv_trim not in ("both", "right"):
            raise ValueError("transconv_trim must be 'both' or 'right'")
        if self.xf_head_dim % 2 or self.xf_heads <= 0:
            raise ValueError("xf_head_dim must be even (rotary halves)")
        if self.num_quantizers <= self.num_semantic:
            raise ValueError("need at least one acoustic codebook")

    # ---- derived quantities -------------------------------------------------
    @property
    def samples_per_frame(self) -> int:
        """Nominal upsampling factor (1920): product of all rates."""
        p = 1
        for r in tuple(self.upsampling_ratios) + tuple(self.upsample_rates):
            p *= r
        return p

    @property
    def attn_dim(self) -> int:
        return self.xf_heads * self.xf_head_dim

    def block_channels(self) -> List[Tuple[int, int]]:
        """(C_in, C_out) of each decoder block."""
        out = []
        for i in range(len(self.upsample_rates)):
            out.append((self.decoder_dim >> i, self.decoder_dim >> (i + 1)))
        return out

    @property
    def head_channels(self) -> int:
        return self.decoder_dim >> len(self.upsample_rates)

    def transconv_out_len(self, length: int, stride: int) -> int:
        """Length after ConvTranspose1d(k=2*stride, stride) + trim (A1)."""
        if self.transconv_trim == "both":
            return (length - 1) * stride
        return length * stride

    def stage_lengths(self, frames: int | None = None) -> List[int]:
        """Time length after: frames, each upsample stage, conv-in, each decoder block."""
        t = self.chunk_frames if frames is None else frames
        lens = [t]
        for r in self.upsampling_ratios:       # k == s: no trim
            t = t * r
            lens.append(t)
        lens.append(t)                          # conv-in keeps the length
        for s in self.upsample_rates:
            t = self.transconv_out_len(t, s)
            lens.append(t)
        return lens

    def chunk_samples(self, frames: int | None = None) -> int:
        """Number of samples the model emits for one window (L in SURVEY 8a H2)."""
        return self.stage_lengths(frames)[-1]

    def flops_per_chunk(self, frames: int | None = None, nominal: bool = True) -> float:
        """Algorithmic FLOPs (2*MAC of conv/linear/attention) for one window.

        ``nominal=True`` uses the length-preserving lengths so the figure equals
        SURVEY 8d's F_chunk (317.49 GFLOP at 64 frames) whatever the trim mode is.
        """
        t0 = self.chunk_frames if frames is None else frames
        cfg = self if not nominal else dataclasses.replace(self, transconv_trim="right")
        mac = 0.0
        mac += t0 * self.codebook_dim * self.rvq_dim * 2                    # two out-projections
        mac += t0 * self.rvq_dim * self.latent_dim * self.pre_conv_kernel   # pre_conv
        if self.pre_transformer:
            h, i, a = self.xf_hidden, self.xf_inter, self.attn_dim
            mac += t0 * 2 * self.latent_dim * h
            per = t0 * (3 * h * a + a * h) + 2 * self.xf_heads * t0 * t0 * self.xf_head_dim + t0 * 3 * h * i
            mac += self.xf_layers * per
        t = t0
        c = self.latent_dim
        for r in self.upsampling_ratios:
            t *= r
            mac += t * c * c
            if self.convnext:
                mac += t * (self.conv_kernel * c + 2 * self.convnext_mult * c * c)
        mac += t * c * self.decoder_dim * self.conv_kernel
        for (ci, co), s in zip(self.block_channels(), self.upsample_rates):
            t = cfg.transconv_out_len(t, s)
            mac += t * ci * co * 2
            mac += len(self.dilations) * t * co * co * (self.conv_kernel + 1)
        mac += t * self.head_channels * self.conv_kernel
        return 2.0 * mac

    # ---- (de)serialisation ----------------------------------------------------
    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), sort_keys=True)

    @staticmethod
    def from_json(s: str) -> "VocoderConfig":
        d = json.loads(s)
        for k in ("upsampling_ratios", "upsample_rates", "dilations"):
            if k in d:
                d[k] = tuple(d[k])
        return VocoderConfig(**d)

    @staticmethod
    def tiny(**kw) -> "VocoderConfig":
        """A small architecture with the same topology, for fast CPU/GPU tests."""
        base = dict(codebook_size=64, codebook_dim=16, rvq_dim=32, latent_dim=64,
                    xf_hidden=32, xf_inter=64, xf_layers=2, xf_heads=2, xf_head_dim=16,
                    decoder_dim=64, chunk_frames=8)
        base.update(kw)
        return VocoderConfig(**base)
