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


# ------------------------------------------------------------------------------------
# real weights: the upstream ``speech_tokenizer`` checkpoint  
# ------------------------------------------------------------------------------------
# The reference builds its graph from ``Qwen3TTSTokenizerV2Model.from_pretrained(<snapshot>/speech_tokenizer).decoder``
# (/root/reference/scripts/export_vocoder_traced.py:28-35,74-79).  Neither that package nor the checkpoint exists in
# this image, so the key names below are NOT verified against a real file: the decoder stack (pre_transformer,
# upsample, decoder) follows the state_dict of the executable sibling of the same lineage (``transformers``
# ``Qwen3OmniMoeCode2Wav``: tests/helpers.py::sibling_param_pairs is the same table), the front end (split RVQ with
# EMA codebooks ``embedding_sum / cluster_usage`` and 1x1 output projections, pre-conv, the projections around the
# transformer) follows the published Mimi-style quantizer layout.  Every entry lists alternatives; a key that cannot be
# found raises with the names that were tried, and ``rename`` lets a caller supply the real name without a code change.

_CODEBOOK_EPS = 1e-5          # EuclideanCodebook: embedding = embedding_sum / clamp(cluster_usage, min=eps)


def _upstream_candidates(cfg: VocoderConfig) -> "OrderedDict[str, Tuple[str, ...]]":
    """our tensor name -> candidate upstream keys (relative to the decoder module)."""
    c: "OrderedDict[str, Tuple[str, ...]]" = OrderedDict()
    c["rvq.proj_sem.w"] = ("quantizer.rvq_first.output_proj.weight",
                           "quantizer.semantic_residual_vector_quantizer.output_proj.weight")
    c["rvq.proj_ac.w"] = ("quantizer.rvq_rest.output_proj.weight",
                          "quantizer.acoustic_residual_vector_quantizer.output_proj.weight")
    c["pre_conv.w"] = ("pre_conv.conv.weight", "pre_conv.weight")
    c["pre_conv.b"] = ("pre_conv.conv.bias", "pre_conv.bias")
    if cfg.pre_transformer:
        c["xf.in_proj.w"] = ("pre_transformer.input_proj.weight",)
        c["xf.in_proj.b"] = ("pre_transformer.input_proj.bias",)
        c["xf.out_proj.w"] = ("pre_transformer.output_proj.weight",)
        c["xf.out_proj.b"] = ("pre_transformer.output_proj.bias",)
        for l in range(cfg.xf_layers):
            u = f"pre_transformer.layers.{l}."
            p = f"xf.{l}."
            for ours, theirs in (("ln1.w", "input_layernorm.weight"), ("q.w", "self_attn.q_proj.weight"),
                                 ("k.w", "self_attn.k_proj.weight"), ("v.w", "self_attn.v_proj.weight"),
                                 ("o.w", "self_attn.o_proj.weight"), ("ls_attn", "self_attn_layer_scale.scale"),
                                 ("ln2.w", "post_attention_layernorm.weight"), ("gate.w", "mlp.gate_proj.weight"),
                                 ("up.w", "mlp.up_proj.weight"), ("down.w", "mlp.down_proj.weight"),
                                 ("ls_mlp", "mlp_layer_scale.scale")):
                c[p + ours] = (u + theirs,)
        c["xf.norm.w"] = ("pre_transformer.norm.weight",)
    for u in range(len(cfg.upsampling_ratios)):
        p, t = f"up.{u}.", f"upsample.{u}."
        c[p + "convt.w"] = (t + "0.conv.weight",)
        c[p + "convt.b"] = (t + "0.conv.bias",)
        if cfg.convnext:
            for ours, theirs in (("dw.w", "dwconv.conv.weight"), ("dw.b", "dwconv.conv.bias"), ("ln.w", "norm.weight"),
                                 ("ln.b", "norm.bias"), ("pw1.w", "pwconv1.weight"), ("pw1.b", "pwconv1.bias"),
                                 ("pw2.w", "pwconv2.weight"), ("pw2.b", "pwconv2.bias"), ("gamma", "gamma")):
                c[p + ours] = (t + "1." + theirs,)
    c["dec.conv_in.w"] = ("decoder.0.conv.weight",)
    c["dec.conv_in.b"] = ("decoder.0.conv.bias",)
    nb = len(cfg.upsample_rates)
    for b in range(nb):
        p, t = f"dec.{b}.", f"decoder.{1 + b}.block."
        c[p + "snake.alpha"] = (t + "0.alpha",)
        c[p + "snake.beta"] = (t + "0.beta",)
        c[p + "convt.w"] = (t + "1.conv.weight",)
        c[p + "convt.b"] = (t + "1.conv.bias",)
        for j in range(len(cfg.dilations)):
            r, tr = p + f"ru.{j}.", t + f"{2 + j}."
            for ours, theirs in (("snake1.alpha", "act1.alpha"), ("snake1.beta", "act1.beta"),
                                 ("conv1.w", "conv1.conv.weight"), ("conv1.b", "conv1.conv.bias"),
                                 ("snake2.alpha", "act2.alpha"), ("snake2.beta", "act2.beta"),
                                 ("conv2.w", "conv2.conv.weight"), ("conv2.b", "conv2.conv.bias")):
                c[r + ours] = (tr + theirs,)
    c["head.snake.alpha"] = (f"decoder.{1 + nb}.alpha",)
    c["head.snake.beta"] = (f"decoder.{1 + nb}.beta",)
    c["head.conv.w"] = (f"decoder.{2 + nb}.conv.weight",)
    c["head.conv.b"] = (f"decoder.{2 + nb}.conv.bias",)
    return c


def _codebook_candidates(q: int, cfg: VocoderConfig):
    """(embedding_sum key, cluster_usage key) alternatives, then plain-embedding alternatives, for quantizer q."""
    if q < cfg.num_semantic:
        groups, i = ("quantizer.rvq_first", "quantizer.semantic_residual_vector_quantizer"), q
    else:
        groups, i = ("quantizer.rvq_rest", "quantizer.acoustic_residual_vector_quantizer"), q - cfg.num_semantic
    ema, plain = [], []
    for g in groups:
        for layer in (f"{g}.vq.layers.{i}._codebook", f"{g}.layers.{i}.codebook", f"{g}.vq.layers.{i}.codebook"):
            ema.append((layer + ".embedding_sum", layer + ".cluster_usage"))
            ema.append((layer + ".embed_sum", layer + ".cluster_usage"))
            plain.append(layer + ".embed")
            plain.append(layer + ".embedding")
    return ema, plain


def read_safetensors(path: str) -> Dict[str, np.ndarray]:
    """All tensors of a .safetensors file as float32 numpy arrays (F32 / F16 / BF16 / F64 inputs)."""
    with open(path, "rb") as f:
        (hl,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hl).decode())
        base = 8 + hl
    header.pop("__metadata__", None)
    data = np.memmap(path, dtype=np.uint8, mode="r", offset=base)
    out: Dict[str, np.ndarray] = {}
    for name, e in header.items():
        a, b = e["data_offsets"]
        raw = data[a:b]
        dt = e["dtype"]
        if dt == "F32":
            arr = np.frombuffer(raw, dtype="<f4")
        elif dt == "F16":
            arr = np.frombuffer(raw, dtype="<f2").astype(np.float32)
        elif dt == "F64":
            arr = np.frombuffer(raw, dtype="<f8").astype(np.float32)
        elif dt == "BF16":
            arr = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
        else:
            continue                                            # integer buffers (step counters ...) are not weights
        out[name] = np.ascontiguousarray(arr, dtype=np.float32).reshape(e["shape"])
    return out


def infer_config(tensors: Dict[str, np.ndarray], prefix: str = "decoder.", **overrides) -> VocoderConfig:
    """Read every dimension the checkpoint pins from tensor shapes; what it cannot pin (transconv_trim, the rotary base,
    epsilons, sliding window) keeps the VocoderConfig default unless overridden."""
    def get(*names):
        for n in names:
            if prefix + n in tensors:
                return tensors[prefix + n]
        return None
    kw = {}
    pc = get("pre_conv.conv.weight", "pre_conv.weight")
    if pc is not None:
        kw.update(latent_dim=int(pc.shape[0]), rvq_dim=int(pc.shape[1]), pre_conv_kernel=int(pc.shape[2]))
    q = get("pre_transformer.layers.0.self_attn.q_proj.weight")
    if q is not None:
        kw.update(xf_hidden=int(q.shape[1]))
        g = get("pre_transformer.layers.0.mlp.gate_proj.weight")
        if g is not None:
            kw.update(xf_inter=int(g.shape[0]))
        n = 0
        while prefix + f"pre_transformer.layers.{n}.self_attn.q_proj.weight" in tensors:
            n += 1
        kw.update(xf_layers=n, pre_transformer=True)
        hd = overrides.get("xf_head_dim", VocoderConfig.xf_head_dim)
        kw.update(xf_heads=int(q.shape[0]) // hd, xf_head_dim=hd)
    ci = get("decoder.0.conv.weight")
    if ci is not None:
        kw.update(decoder_dim=int(ci.shape[0]), conv_kernel=int(ci.shape[2]))
    rates, b = [], 0
    while prefix + f"decoder.{1 + b}.block.1.conv.weight" in tensors:
        rates.append(int(tensors[prefix + f"decoder.{1 + b}.block.1.conv.weight"].shape[2]) // 2)
        b += 1
    if rates:
        kw.update(upsample_rates=tuple(rates))
    ups, u = [], 0
    while prefix + f"upsample.{u}.0.conv.weight" in tensors:
        ups.append(int(tensors[prefix + f"upsample.{u}.0.conv.weight"].shape[2]))
        u += 1
    if ups:
        kw.update(upsampling_ratios=tuple(ups), convnext=prefix + "upsample.0.1.pwconv1.weight" in tensors)
    kw.update(overrides)
    cfg = VocoderConfig(**kw)
    ema, plain = _codebook_candidates(0, cfg)
    for es, _ in ema:
        if prefix + es in tensors:
            t = tensors[prefix + es]
            return dataclass_replace(cfg, codebook_size=int(t.shape[0]), codebook_dim=int(t.shape[1]))
    for p in plain:
        if prefix + p in tensors:
            t = tensors[prefix + p]
            return dataclass_replace(cfg, codebook_size=int(t.shape[0]), codebook_dim=int(t.shape[1]))
    return cfg


def dataclass_replace(cfg: VocoderConfig, **kw) -> VocoderConfig:
    import dataclasses
    return dataclasses.replace(cfg, **kw)


def from_speech_tokenizer(path_or_tensors, cfg: VocoderConfig = None, prefix: str = "decoder.", rename: Dict[str, str] = None,
                          **cfg_overrides) -> Tuple[VocoderConfig, "OrderedDict[str, np.ndarray]"]:
    """Map an upstream ``speech_tokenizer`` checkpoint (a .safetensors path or a name -> array dict) onto this
    repo's tensor names and layouts: codebooks folded from their EMA form (``embedding_sum / clamp(cluster_usage)``),
    the two RVQ output projections as [rvq_dim, codebook_dim] matrices (their 1x1-conv kernel axis dropped),
    everything else copied.  ``rename`` maps one of OUR names to the upstream key to use for it."""
    tensors = read_safetensors(path_or_tensors) if isinstance(path_or_tensors, str) else dict(path_or_tensors)
    if cfg is None:
        cfg = infer_config(tensors, prefix, **cfg_overrides)
    rename = dict(rename or {})
    shapes = weight_shapes(cfg)
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def fetch(ours, cands):
        if ours in rename:
            cands = (rename[ours],)
        for k in cands:
            if prefix + k in tensors:
                return np.asarray(tensors[prefix + k], dtype=np.float32)
            if k in tensors:
                return np.asarray(tensors[k], dtype=np.float32)
        raise KeyError(f"{ours}: none of {[prefix + k for k in cands]} is in the checkpoint "
                       f"(pass rename={{'{ours}': '<upstream key>'}})")

    for q in range(cfg.num_quantizers):
        ours = f"rvq.codebook.{q}"
        ema, plain = _codebook_candidates(q, cfg)
        cb = None
        if ours in rename:
            cb = fetch(ours, ())
        else:
            for es, cu in ema:
                if prefix + es in tensors and prefix + cu in tensors:
                    usage = np.maximum(np.asarray(tensors[prefix + cu], dtype=np.float32), _CODEBOOK_EPS)
                    cb = np.asarray(tensors[prefix + es], dtype=np.float32) / usage[:, None]
                    break
            if cb is None:
                for p in plain:
                    if prefix + p in tensors:
                        cb = np.asarray(tensors[prefix + p], dtype=np.float32)
                        break
        if cb is None:
            raise KeyError(f"{ours}: no codebook found; tried {[prefix + e for e, _ in ema[:3]]} ...")
        out[ours] = cb
    for ours, cands in _upstream_candidates(cfg).items():
        w = fetch(ours, cands)
        want = shapes[ours]
        if ours.startswith("rvq.proj_") and w.ndim == 3 and w.shape[2] == 1:
            w = w[:, :, 0]                                     # Conv1d(k=1) -> matrix
        if ours.endswith("ls_attn") or ours.endswith("ls_mlp") or ours.endswith("gamma"):
            w = w.reshape(-1)
        if tuple(w.shape) != tuple(want):
            raise ValueError(f"{ours}: checkpoint shape {tuple(w.shape)}, architecture wants {tuple(want)}")
        out[ours] = w
    ordered: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in shapes.items():
        if tuple(out[name].shape) != tuple(shape):
            raise ValueError(f"{name}: checkpoint shape {tuple(out[name].shape)}, architecture wants {tuple(shape)}")
        ordered[name] = np.ascontiguousarray(out[name], dtype=np.float32)
    return cfg, ordered


def to_speech_tokenizer_names(cfg: VocoderConfig, weights: Dict[str, np.ndarray], prefix: str = "decoder.",
                              seed: int = 0) -> Dict[str, np.ndarray]:
    """The inverse of ``from_speech_tokenizer`` (first candidate name of every entry; codebooks in EMA form with a
    random positive ``cluster_usage``): used by the tests to fabricate a checkpoint in the upstream naming."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for q in range(cfg.num_quantizers):
        ema, _ = _codebook_candidates(q, cfg)
        es, cu = ema[0]
        usage = rng.uniform(0.5, 50.0, cfg.codebook_size).astype(np.float32)
        out[prefix + cu] = usage
        out[prefix + es] = (weights[f"rvq.codebook.{q}"] * usage[:, None]).astype(np.float32)
    for ours, cands in _upstream_candidates(cfg).items():
        w = np.asarray(weights[ours], dtype=np.float32)
        if ours.startswith("rvq.proj_"):
            w = w[:, :, None]
        out[prefix + cands[0]] = w
    return out


def write_safetensors(path: str, tensors: Dict[str, np.ndarray]) -> None:
    header, off = OrderedDict(), 0
    for name, w in tensors.items():
        w = np.ascontiguousarray(w, dtype=np.float32)
        header[name] = {"dtype": "F32", "shape": list(w.shape), "data_offsets": [off, off + w.nbytes]}
        off += w.nbytes
    hb = json.dumps(header, separators=(",", ":")).encode()
    hb += b" " * ((8 - len(hb) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for w in tensors.values():
            f.write(np.ascontiguousarray(w, dtype=np.float32).tobytes())
