"""CPU oracle of the vocoder graph  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product path
(``qwen3-tts-axera-russian_b200``) never does and has no CPU fallback.

PARITY STATUS: **unpinned at the model boundary.**  The arithmetic of the reference's
``vocoder_traced_64.onnx`` lives in the un-vendored, un-pinned PyPI package ``qwen-tts``
(``/root/reference/scripts/export_vocoder_traced.py:74-79``, ``docs/SETUP.md:38``); the
reference holds no golden vector, hash or tolerance for it, and neither ``onnxruntime``
nor the weights exist in this image.  This file restates the published structure of that
decoder with explicit formulas.  What *is* pinned:
  * the graph I/O contract   -> ``scripts/export_vocoder_traced.py:38-52``  (``forward``)
  * each building block      -> golden vectors under ``tests/golden/sibling_*.npz`` produced
    by ``tests/golden/make_sibling_golden.py`` from the executable sibling implementation
    of the same lineage that ships in this image (``transformers`` ``Qwen3OmniMoeCode2Wav*``;
    non-reference evidence).
  * the composition of M4-M9 -> ``decode_tail`` against that sibling's whole forward after its front end
    (``tests/golden/sibling_tail.npz``, ``tests/golden/make_sibling_tail_golden.py``), and with the
    pre-transformer included, M3-M9 (``tests/golden/sibling_model.npz``, ``make_sibling_model_golden.py``);
    at production dimensions against the sibling executed live with the same random weights
    (``tests/test_oracle_blocks.py::test_production_size_oracle_matches_the_sibling_run_live``: 7e-6).

All functions take ``x`` as ``[B, C, L]`` (torch layout) and a dict of named weights
(layouts documented in ``weights.py``) and are dtype-generic (float32 = the stand-in for
ONNX Runtime FP32; float64 = the error yard-stick).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _t(w, dtype):
    if isinstance(w, torch.Tensor):
        return w.to(dtype)
    return torch.from_numpy(np.ascontiguousarray(w)).to(dtype)


class Weights:
    """Name -> torch tensor view of a numpy weight dict, converted once."""

    def __init__(self, weights: Dict[str, np.ndarray], dtype=torch.float32):
        self.dtype = dtype
        self._w = {k: _t(v, dtype) for k, v in weights.items()}

    def __getitem__(self, k):
        return self._w[k]

    def __contains__(self, k):
        return k in self._w


# ----------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------

def causal_conv1d(x, w, b, dilation: int = 1, groups: int = 1):
    """CausalConvNet, stride 1: left-pad (k-1)*d zeros, no right pad.

    y[co,t] = b[co] + sum_ci sum_j W[co,ci,j] * x[ci, t-(k-1-j)*d]
    (sibling modeling_qwen3_omni_moe.py:3283-3316; with stride 1 the 'extra padding' is 0).
    """
    k = w.shape[-1]
    x = F.pad(x, ((k - 1) * dilation, 0))
    return F.conv1d(x, w, b, dilation=dilation, groups=groups)


def causal_transconv1d(x, w, b, stride: int, trim: str):
    """CausalTransConvNet: ConvTranspose1d(k, stride) then trim k-stride samples.

    full[co, t*s+j] += x[ci,t] * W[ci,co,j]  (+ b[co]);  length (L-1)*s + k.
    trim="both" drops k-s from each end (sibling :3319-3331), "right" only from the right.
    For k == s nothing is trimmed in either mode.
    """
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    pad = k - stride
    if pad == 0:
        return y
    if trim == "both":
        return y[..., pad: y.shape[-1] - pad]
    return y[..., : y.shape[-1] - pad]


def snake_beta(x, alpha, beta, eps: float = 1e-9):
    """SnakeBeta: x + 1/(exp(beta)+eps) * sin(x*exp(alpha))**2   (sibling :3645-3683)."""
    a = torch.exp(alpha).view(1, -1, 1)
    b = torch.exp(beta).view(1, -1, 1)
    return x + (1.0 / (b + eps)) * torch.pow(torch.sin(x * a), 2)


def rms_norm(x, w, eps):
    """RMSNorm over the last dim (sibling :3458-3476)."""
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rotary_cos_sin(T: int, head_dim: int, theta: float, dtype):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
    ang = torch.arange(T, dtype=torch.float64)[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], dim=-1)
    # the sibling computes the table in float32 (default rope init); keep the oracle's
    # dtype for the values themselves
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def attention(x, W: Weights, p: str, heads: int, head_dim: int, theta: float, window: int):
    """Causal sliding-window MHA with RoPE, no biases (sibling :3370-3439)."""
    B, T, _ = x.shape
    q = F.linear(x, W[p + "q.w"]).view(B, T, heads, head_dim).transpose(1, 2)
    k = F.linear(x, W[p + "k.w"]).view(B, T, heads, head_dim).transpose(1, 2)
    v = F.linear(x, W[p + "v.w"]).view(B, T, heads, head_dim).transpose(1, 2)
    cos, sin = rotary_cos_sin(T, head_dim, theta, x.dtype)
    q = q * cos + _rotate_half(q) * sin
    k = k * cos + _rotate_half(k) * sin
    s = torch.matmul(q, k.transpose(-1, -2)) * (head_dim ** -0.5)
    i = torch.arange(T)[:, None]
    j = torch.arange(T)[None, :]
    allowed = (j <= i) & (j > i - window)
    s = s.masked_fill(~allowed, float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, T, heads * head_dim)
    return F.linear(o, W[p + "o.w"])


def transformer_layer(x, W: Weights, p: str, cfg):
    """x += ls_a * MHA(RMSNorm(x)); x += ls_m * SwiGLU(RMSNorm(x))  (sibling :3494-3553)."""
    h = rms_norm(x, W[p + "ln1.w"], cfg.rms_eps)
    x = x + W[p + "ls_attn"] * attention(h, W, p, cfg.xf_heads, cfg.xf_head_dim,
                                         cfg.rope_theta, cfg.sliding_window)
    h = rms_norm(x, W[p + "ln2.w"], cfg.rms_eps)
    m = F.linear(F.silu(F.linear(h, W[p + "gate.w"])) * F.linear(h, W[p + "up.w"]),
                 W[p + "down.w"])
    return x + W[p + "ls_mlp"] * m


def pre_transformer(x, W: Weights, cfg):
    """[B, latent, T] -> [B, latent, T]: in-proj, layers, final RMSNorm, out-proj (M3)."""
    h = F.linear(x.transpose(1, 2), W["xf.in_proj.w"], W["xf.in_proj.b"])
    for l in range(cfg.xf_layers):
        h = transformer_layer(h, W, f"xf.{l}.", cfg)
    h = rms_norm(h, W["xf.norm.w"], cfg.rms_eps)
    h = F.linear(h, W["xf.out_proj.w"], W["xf.out_proj.b"])
    return h.transpose(1, 2)


def convnext_block(x, W: Weights, p: str, cfg):
    """dw causal conv k7 -> LayerNorm(C) -> Linear -> exact GELU -> Linear -> gamma -> +x
    (sibling :3334-3366)."""
    c = x.shape[1]
    h = causal_conv1d(x, W[p + "dw.w"], W[p + "dw.b"], groups=c)
    h = h.transpose(1, 2)
    h = F.layer_norm(h, (c,), W[p + "ln.w"], W[p + "ln.b"], cfg.ln_eps)
    h = F.linear(h, W[p + "pw1.w"], W[p + "pw1.b"])
    h = F.gelu(h)                      # erf form
    h = F.linear(h, W[p + "pw2.w"], W[p + "pw2.b"])
    h = W[p + "gamma"] * h
    return x + h.transpose(1, 2)


def residual_unit(x, W: Weights, p: str, dilation: int, cfg):
    """y = x + Conv_k1(Snake2(CausalConv_k7,dil(Snake1(x))))   (sibling :3686-3702)."""
    h = snake_beta(x, W[p + "snake1.alpha"], W[p + "snake1.beta"], cfg.snake_eps)
    h = causal_conv1d(h, W[p + "conv1.w"], W[p + "conv1.b"], dilation=dilation)
    h = snake_beta(h, W[p + "snake2.alpha"], W[p + "snake2.beta"], cfg.snake_eps)
    h = causal_conv1d(h, W[p + "conv2.w"], W[p + "conv2.b"])
    return x + h


def decoder_block(x, W: Weights, b: int, cfg):
    """Snake -> transposed conv (stride s, k=2s) -> 3 residual units (sibling :3705-3727)."""
    p = f"dec.{b}."
    s = cfg.upsample_rates[b]
    h = snake_beta(x, W[p + "snake.alpha"], W[p + "snake.beta"], cfg.snake_eps)
    h = causal_transconv1d(h, W[p + "convt.w"], W[p + "convt.b"], s, cfg.transconv_trim)
    for j, d in enumerate(cfg.dilations):
        h = residual_unit(h, W, p + f"ru.{j}.", d, cfg)
    return h


def rvq_decode(codes, W: Weights, cfg):
    """codes int64 [B, n_q, T] -> [B, rvq_dim, T]:
    P_sem * E_0[c_0] + P_ac * sum_{q>=1} E_q[c_q]   (SURVEY 8a M1)."""
    B, nq, T = codes.shape
    sem = None
    ac = None
    for q in range(nq):
        e = W[f"rvq.codebook.{q}"][codes[:, q, :]]           # [B, T, D]
        if q < cfg.num_semantic:
            sem = e if sem is None else sem + e
        else:
            ac = e if ac is None else ac + e
    h = F.linear(sem, W["rvq.proj_sem.w"]) + F.linear(ac, W["rvq.proj_ac.w"])
    return h.transpose(1, 2)


# ----------------------------------------------------------------------------------
# whole graph
# ----------------------------------------------------------------------------------

def decode(codes_bqt, W: Weights, cfg, taps: Optional[dict] = None):
    """The upstream ``decoder(codes[B,16,T])`` -> wav ``[B, 1, L]`` clamped to [-1, 1]."""
    if codes_bqt.shape[1] != cfg.num_quantizers:
        raise ValueError(f"expected {cfg.num_quantizers} codebooks, got {codes_bqt.shape[1]}")
    if codes_bqt.min() < 0 or codes_bqt.max() >= cfg.codebook_size:
        raise IndexError("code out of range")           # ORT Gather would throw here

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()
        return v

    h = tap("rvq", rvq_decode(codes_bqt, W, cfg))
    h = tap("pre_conv", causal_conv1d(h, W["pre_conv.w"], W["pre_conv.b"]))
    if cfg.pre_transformer:
        h = tap("xf", pre_transformer(h, W, cfg))
    return decode_tail(h, W, cfg, taps)


def decode_tail(h, W: Weights, cfg, taps: Optional[dict] = None):
    """Latent ``[B, latent, T]`` (the pre-transformer's output) -> wav ``[B, 1, L]``: up-sampling stages,
    conv-in, decoder blocks, head, clamp (M4-M9).  The composition the executable sibling runs after
    its own front end (``Qwen3OmniMoeCode2Wav.forward``: ``for blocks in self.upsample ... for block in
    self.decoder ... clamp``); pinned as a whole by tests/golden/sibling_tail.npz."""

    def tap(name, v):
        if taps is not None:
            taps[name] = v.detach().clone()
        return v

    for u, r in enumerate(cfg.upsampling_ratios):
        p = f"up.{u}."
        h = causal_transconv1d(h, W[p + "convt.w"], W[p + "convt.b"], r, cfg.transconv_trim)
        tap(f"up{u}.convt", h)
        if cfg.convnext:
            h = convnext_block(h, W, p, cfg)
        tap(f"up{u}", h)
    h = tap("conv_in", causal_conv1d(h, W["dec.conv_in.w"], W["dec.conv_in.b"]))
    for b in range(len(cfg.upsample_rates)):
        h = tap(f"dec{b}", decoder_block(h, W, b, cfg))
    h = snake_beta(h, W["head.snake.alpha"], W["head.snake.beta"], cfg.snake_eps)
    h = causal_conv1d(h, W["head.conv.w"], W["head.conv.b"])
    return h.clamp(min=-1, max=1)


def forward(audio_codes, W: Weights, cfg, taps: Optional[dict] = None):
    """The ONNX graph contract: ``VocoderWrapper.forward``
    (/root/reference/scripts/export_vocoder_traced.py:46-52).

    audio_codes int64 [B, T, 16]  ->  (audio_values [B, L], lengths int64 [1] = T*total_upsample)
    """
    if isinstance(audio_codes, np.ndarray):
        audio_codes = torch.from_numpy(np.ascontiguousarray(audio_codes))
    codes = audio_codes.permute(0, 2, 1).long()
    with torch.no_grad():
        wav = decode(codes, W, cfg, taps)
    audio_values = wav.squeeze(1)
    lengths = torch.tensor([audio_codes.shape[1] * cfg.samples_per_frame], dtype=torch.int64)
    return audio_values, lengths


class OracleVocoder:
    """``_inference_chunk``-shaped callable for the reference's chunker
    (/root/reference/dual_npu/vocoder_server.py:67-71)."""

    def __init__(self, cfg, weights: Dict[str, np.ndarray], dtype=torch.float32,
                 threads: Optional[int] = None):
        self.cfg = cfg
        self.W = Weights(weights, dtype)
        self.max_tokens = cfg.chunk_frames
        if threads:
            torch.set_num_threads(threads)

    def infer_chunks(self, padded: np.ndarray) -> np.ndarray:
        a, _ = forward(padded, self.W, self.cfg)
        return a.to(torch.float32).numpy() if a.dtype != torch.float64 else a.numpy()

    def _inference_chunk(self, padded: np.ndarray) -> np.ndarray:
        return self.infer_chunks(padded)[0].astype(np.float32).flatten()


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    """SNR_dB = 10 log10( sum ref^2 / sum (ref-test)^2 )   (SURVEY 8c)."""
    ref = np.asarray(ref, dtype=np.float64)
    test = np.asarray(test, dtype=np.float64)
    num = float(np.sum(ref * ref))
    den = float(np.sum((ref - test) ** 2))
    if den == 0.0:
        return math.inf
    return 10.0 * math.log10(num / den)
