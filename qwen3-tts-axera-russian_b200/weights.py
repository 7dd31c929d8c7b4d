"""Weight inventory, deterministic synthetic initialisation and the on-disk container.

The reference ships no weights (``/root/reference/.gitignore:4``) and there is no network, so
benchmarks and parity tests run on random-init weights of the named architecture
(SURVEY.md 8d "Weights").  The initialisation is variance preserving so that the signal
neither vanishes nor saturates the final clamp -- a degenerate output would make the
SNR / max-abs gate meaningless.

Tensor layouts follow torch conventions so that real ``speech_tokenizer`` weights
(``scripts/export_vocoder_traced.py:28-35``) can be mapped one-to-one later:
  Conv1d            [C_out, C_in, K]
  ConvTranspose1d   [C_in, C_out, K]
  Linear            [out, in]

Container: the safetensors byte layout (8-byte LE header length, JSON header, raw
little-endian tensor data) with the architecture JSON under ``__metadata__["voc_config"]``.
Files carry the suffix ``.b200voc`` -- the suffix ``VocoderServer`` dispatches on, the same
way the reference dispatches on ``.onnx`` (``dual_npu/vocoder_server.py:36``).
"""
from __future__ import annotations

import json
import struct
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import VocoderConfig

MODEL_SUFFIX = ".b200voc"


def weight_shapes(cfg: VocoderConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    """Name -> shape for every parameter of the decoder graph."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    for q in range(cfg.num_quantizers):
        s[f"rvq.codebook.{q}"] = (cfg.codebook_size, cfg.codebook_dim)
    s["rvq.proj_sem.w"] = (cfg.rvq_dim, cfg.codebook_dim)
    s["rvq.proj_ac.w"] = (cfg.rvq_dim, cfg.codebook_dim)
    s["pre_conv.w"] = (cfg.latent_dim, cfg.rvq_dim, cfg.pre_conv_kernel)
    s["pre_conv.b"] = (cfg.latent_dim,)
    if cfg.pre_transformer:
        h, i, a = cfg.xf_hidden, cfg.xf_inter, cfg.attn_dim
        s["xf.in_proj.w"] = (h, cfg.latent_dim)
        s["xf.in_proj.b"] = (h,)
        for l in range(cfg.xf_layers):
            p = f"xf.{l}."
            s[p + "ln1.w"] = (h,)
            s[p + "q.w"] = (a, h)
            s[p + "k.w"] = (a, h)
            s[p + "v.w"] = (a, h)
            s[p + "o.w"] = (h, a)
            s[p + "ls_attn"] = (h,)
            s[p + "ln2.w"] = (h,)
            s[p + "gate.w"] = (i, h)
            s[p + "up.w"] = (i, h)
            s[p + "down.w"] = (h, i)
            s[p + "ls_mlp"] = (h,)
        s["xf.norm.w"] = (h,)
        s["xf.out_proj.w"] = (cfg.latent_dim, h)
        s["xf.out_proj.b"] = (cfg.latent_dim,)
    c = cfg.latent_dim
    for u, r in enumerate(cfg.upsampling_ratios):
        p = f"up.{u}."
        s[p + "convt.w"] = (c, c, r)
        s[p + "convt.b"] = (c,)
        if cfg.convnext:
            s[p + "dw.w"] = (c, 1, cfg.conv_kernel)
            s[p + "dw.b"] = (c,)
            s[p + "ln.w"] = (c,)
            s[p + "ln.b"] = (c,)
            s[p + "pw1.w"] = (cfg.convnext_mult * c, c)
            s[p + "pw1.b"] = (cfg.convnext_mult * c,)
            s[p + "pw2.w"] = (c, cfg.convnext_mult * c)
            s[p + "pw2.b"] = (c,)
            s[p + "gamma"] = (c,)
    s["dec.conv_in.w"] = (cfg.decoder_dim, c, cfg.conv_kernel)
    s["dec.conv_in.b"] = (cfg.decoder_dim,)
    for b, ((ci, co), st) in enumerate(zip(cfg.block_channels(), cfg.upsample_rates)):
        p = f"dec.{b}."
        s[p + "snake.alpha"] = (ci,)
        s[p + "snake.beta"] = (ci,)
        s[p + "convt.w"] = (ci, co, 2 * st)
        s[p + "convt.b"] = (co,)
        for j in range(len(cfg.dilations)):
            r = p + f"ru.{j}."
            s[r + "snake1.alpha"] = (co,)
            s[r + "snake1.beta"] = (co,)
            s[r + "conv1.w"] = (co, co, cfg.conv_kernel)
            s[r + "conv1.b"] = (co,)
            s[r + "snake2.alpha"] = (co,)
            s[r + "snake2.beta"] = (co,)
            s[r + "conv2.w"] = (co, co, 1)
            s[r + "conv2.b"] = (co,)
    ch = cfg.head_channels
    s["head.snake.alpha"] = (ch,)
    s["head.snake.beta"] = (ch,)
    s["head.conv.w"] = (1, ch, cfg.conv_kernel)
    s["head.conv.b"] = (1,)
    return s


def _rng(seed: int, name: str) -> np.random.Generator:
    # one independent stream per tensor: the value of a tensor does not depend on the
    # order or presence of the others.
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


# Gains chosen once from a CPU run of the graph so that,
# for the default architecture and seed 0, every stage stays O(1), the output RMS is
# ~0.1-0.3 and < 1 % of the samples reach the clamp.
_RESIDUAL_GAIN = 0.5      # conv2 of a residual unit / o_proj / down_proj / pwconv2
_LAYER_SCALE = 0.1        # transformer LayerScale (upstream initial value 0.01)
_CONVNEXT_GAMMA = 0.1     # ConvNeXt gamma (upstream initial value 1e-6)
_HEAD_GAIN = 0.035
_SNAKE_JITTER = 0.1
_BIAS_STD = 0.01


def init_weights(cfg: VocoderConfig, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Deterministic variance-preserving random weights (float32) for ``cfg``."""
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in weight_shapes(cfg).items():
        g = _rng(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if name.startswith("rvq.codebook."):
            w = g.standard_normal(shape, dtype=np.float32)
        elif leaf in ("alpha", "beta"):
            w = (_SNAKE_JITTER * g.standard_normal(shape, dtype=np.float32))
        elif leaf == "b":
            w = (_BIAS_STD * g.standard_normal(shape, dtype=np.float32))
        elif name.endswith("ln1.w") or name.endswith("ln2.w") or name.endswith("norm.w") \
                or name.endswith("ln.w"):
            w = (1.0 + 0.05 * g.standard_normal(shape, dtype=np.float32)).astype(np.float32)
        elif leaf in ("ls_attn", "ls_mlp"):
            w = (_LAYER_SCALE * (1.0 + 0.1 * g.standard_normal(shape, dtype=np.float32)))
        elif leaf == "gamma":
            w = (_CONVNEXT_GAMMA * (1.0 + 0.1 * g.standard_normal(shape, dtype=np.float32)))
        elif leaf == "w":
            if ".convt." in name:
                ci, co, k = shape
                # k = stride : one tap per output sample; k = 2*stride : two taps
                fan_in = ci * (2 if name.startswith("dec.") else 1)
            elif ".dw." in name:
                fan_in = shape[2]
            elif len(shape) == 3:
                fan_in = shape[1] * shape[2]
            else:
                fan_in = shape[1]
            std = 1.0 / np.sqrt(fan_in)
            if name.startswith("rvq.proj_"):
                # sum of 15 unit-variance acoustic codewords vs one semantic codeword
                n = cfg.num_semantic if name.startswith("rvq.proj_sem") \
                    else cfg.num_quantizers - cfg.num_semantic
                std = std / np.sqrt(2.0 * n)
            if ".conv2." in name or name.endswith("o.w") or name.endswith("down.w") \
                    or name.endswith("pw2.w"):
                std = std * _RESIDUAL_GAIN
            if name == "head.conv.w":
                std = std * _HEAD_GAIN
            w = (std * g.standard_normal(shape, dtype=np.float32))
        else:
            raise KeyError(name)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


# ------------------------------------------------------------------------------------
# container (safetensors byte layout)
# ------------------------------------------------------------------------------------

def save_model(path: str, cfg: VocoderConfig, weights: Dict[str, np.ndarray]) -> None:
    header = OrderedDict()
    header["__metadata__"] = {"voc_config": cfg.to_json(), "format": "b200voc-1"}
    off = 0
    for name, shape in weight_shapes(cfg).items():
        w = weights[name]
        if tuple(w.shape) != tuple(shape) or w.dtype != np.float32:
            raise ValueError(f"{name}: expected float32{shape}, got {w.dtype}{w.shape}")
        n = w.nbytes
        header[name] = {"dtype": "F32", "shape": list(shape), "data_offsets": [off, off + n]}
        off += n
    hb = json.dumps(header, separators=(",", ":")).encode()
    hb += b" " * ((8 - len(hb) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for name in weight_shapes(cfg):
            f.write(np.ascontiguousarray(weights[name]).tobytes())


def load_model(path: str) -> Tuple[VocoderConfig, "OrderedDict[str, np.ndarray]"]:
    with open(path, "rb") as f:
        (hl,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hl).decode())
        base = 8 + hl
        meta = header.pop("__metadata__", {})
        if "voc_config" not in meta:
            raise ValueError(f"{path}: no voc_config in metadata (not a {MODEL_SUFFIX} file)")
        cfg = VocoderConfig.from_json(meta["voc_config"])
        data = np.memmap(path, dtype=np.uint8, mode="r", offset=base)
        out: "OrderedDict[str, np.ndarray]" = OrderedDict()
        for name, shape in weight_shapes(cfg).items():
            if name not in header:
                raise ValueError(f"{path}: missing tensor {name}")
            e = header[name]
            if e["dtype"] != "F32" or tuple(e["shape"]) != tuple(shape):
                raise ValueError(f"{path}: {name} has {e['dtype']}{e['shape']}, want F32{shape}")
            a, b = e["data_offsets"]
            out[name] = np.frombuffer(data[a:b], dtype="<f4").reshape(shape)
    return cfg, out
