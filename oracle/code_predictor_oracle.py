"""CPU restatement of the reference's code-predictor path (TEST INFRASTRUCTURE ONLY).

The reference predicts codec groups 1-15 of a frame with a 5-layer transformer run as
``code_predictor_decode_step.onnx`` on ONNX Runtime (/root/reference/dual_npu/code_predictor_server.py:77-140): one
"decode step" takes ``hidden [1, S, H]``, ``position [S]`` and the per-layer ``past_k / past_v`` and returns the
final-normed hidden states plus the grown caches; ``predict()`` runs it on the talker's hidden state (position 0), the
embedding of code_0 (position 1) and then 14 times on the embedding of the group it has just sampled.

As with the vocoder, the graph's arithmetic is not in the reference: the ONNX file is exported from the un-vendored
``qwen_tts`` package (scripts/export_code_predictor_onnx.py:70-91) -- **parity unpinned** against the reference's own
model.  What the reference does pin: the weight inventory and names (scripts/export_code_predictor_weights.py:50-70),
"5-layer transformer, 1024-dim, GQA 16/8 heads" (docs/ARCHITECTURE.md:101), ``head_dim = q_proj.shape[0] // 16`` and 8 KV
heads (code_predictor_server.py:66-67), the step's I/O contract (:77-85), the sampler (:87-92) and the predict loop
(:94-140).  The layer arithmetic below follows the executable sibling of the same lineage in this image,
``transformers`` ``Qwen3OmniMoeTalkerCodePredictorModel`` (cited ``sib:line``; modeling_qwen3_omni_moe.py), and is
pinned against it by tests/golden/sibling_cp.npz (tests/test_cp_oracle.py); ``predict`` is pinned against the reference's
own ``predict()`` executing with a fake step function.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import dataclasses
import json
import zlib
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class CPConfig:
    hidden: int = 1024            # docs/ARCHITECTURE.md:101
    layers: int = 5
    heads: int = 16
    kv_heads: int = 8             # code_predictor_server.py:67
    head_dim: int = 128           # q_proj.shape[0] // 16 (:66); 128 in the sibling's defaults
    inter: int = 3072
    vocab: int = 2048
    groups: int = 15              # :39
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_positions: int = 32       # a frame uses positions 0 .. 16

    def to_json(self) -> str:
        return json.dumps(dataclasses.asdict(self), sort_keys=True)

    @staticmethod
    def tiny(**kw) -> "CPConfig":
        base = dict(hidden=64, layers=2, heads=4, kv_heads=2, head_dim=16, inter=96, vocab=32, groups=4)
        base.update(kw)
        return CPConfig(**base)


def weight_shapes(cfg: CPConfig) -> Dict[str, Tuple[int, ...]]:
    """The arrays of code_predictor_weights.npz, by the names scripts/export_code_predictor_weights.py:50-70 gives them."""
    s: Dict[str, Tuple[int, ...]] = {}
    q, kv = cfg.heads * cfg.head_dim, cfg.kv_heads * cfg.head_dim
    for i in range(cfg.layers):
        p = f"layer_{i}_"
        s[p + "input_ln"] = (cfg.hidden,)
        s[p + "q_proj"] = (q, cfg.hidden)
        s[p + "k_proj"] = (kv, cfg.hidden)
        s[p + "v_proj"] = (kv, cfg.hidden)
        s[p + "o_proj"] = (cfg.hidden, q)
        s[p + "q_norm"] = (cfg.head_dim,)
        s[p + "k_norm"] = (cfg.head_dim,)
        s[p + "post_ln"] = (cfg.hidden,)
        s[p + "gate_proj"] = (cfg.inter, cfg.hidden)
        s[p + "up_proj"] = (cfg.inter, cfg.hidden)
        s[p + "down_proj"] = (cfg.hidden, cfg.inter)
    s["final_norm"] = (cfg.hidden,)
    for i in range(cfg.groups):
        s[f"codec_emb_{i}"] = (cfg.vocab, cfg.hidden)
        s[f"lm_head_{i}"] = (cfg.vocab, cfg.hidden)
    return s


def init_weights(cfg: CPConfig, seed: int = 0) -> Dict[str, np.ndarray]:
    """Deterministic random weights that keep the residual stream O(1) and the logits spread out (std ~ 2)."""
    out: Dict[str, np.ndarray] = {}
    for name, shape in weight_shapes(cfg).items():
        g = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if name.endswith("_ln") or name.endswith("_norm") or name == "final_norm":
            w = 1.0 + 0.05 * g.standard_normal(shape)
        elif name.startswith("codec_emb_"):
            w = g.standard_normal(shape)
        elif name.startswith("lm_head_"):
            w = g.standard_normal(shape) * (2.0 / np.sqrt(shape[1]))
        else:
            gain = 0.5 if (name.endswith("o_proj") or name.endswith("down_proj")) else 1.0
            w = g.standard_normal(shape) * (gain / np.sqrt(shape[1]))
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


class Weights:
    def __init__(self, w: Dict[str, np.ndarray], dtype=torch.float32):
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in w.items()}
        self.dtype = dtype

    def __getitem__(self, k):
        return self.t[k]


def rms_norm(x, w, eps):
    """x * rsqrt(mean(x^2) + eps) * w, statistics in the input dtype promoted to float32+ (sib: Qwen3OmniMoeRMSNorm)."""
    v = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(v + eps))


def rope_tables(positions: Sequence[int], cfg: CPConfig, dtype):
    """cos / sin [S, head_dim] with the two halves equal (sib: Qwen3OmniMoeRotaryEmbedding.forward: emb = cat(freqs, freqs))."""
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.float64) / cfg.head_dim))
    ang = torch.tensor(list(positions), dtype=torch.float64)[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2                                        # sib:816-820
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def step(hidden, positions: Sequence[int], kv: Optional[List[Tuple[torch.Tensor, torch.Tensor]]], W: Weights, cfg: CPConfig):
    """One call of the decode-step graph (code_predictor_server.py:77-85): hidden [S, H] at `positions`, caches
    kv[l] = (K [kv_heads, P, head_dim], V [...]) or None  ->  (final-normed hidden [S, H], grown caches).

    Layer (sib:2442-2478): x += Attn(RMSNorm(x)); x += MLP(RMSNorm(x)).  Attention (sib:2352-2424): q / k / v projections
    without bias, RMSNorm over head_dim on q and k, rotary embedding, GQA with heads // kv_heads query heads per KV head,
    causal softmax in float32 scaled by head_dim^-0.5.  MLP (sib:2426-2439): down(silu(gate(x)) * up(x))."""
    x = hidden.to(W.dtype)
    S = x.shape[0]
    cos, sin = rope_tables(positions, cfg, W.dtype)
    new_kv = []
    rep = cfg.heads // cfg.kv_heads
    for l in range(cfg.layers):
        p = f"layer_{l}_"
        hn = rms_norm(x, W[p + "input_ln"], cfg.rms_eps)
        q = (hn @ W[p + "q_proj"].T).view(S, cfg.heads, cfg.head_dim)
        k = (hn @ W[p + "k_proj"].T).view(S, cfg.kv_heads, cfg.head_dim)
        v = (hn @ W[p + "v_proj"].T).view(S, cfg.kv_heads, cfg.head_dim)
        q = rms_norm(q, W[p + "q_norm"], cfg.rms_eps)
        k = rms_norm(k, W[p + "k_norm"], cfg.rms_eps)
        q = q * cos[:, None, :] + rotate_half(q) * sin[:, None, :]
        k = k * cos[:, None, :] + rotate_half(k) * sin[:, None, :]
        k = k.transpose(0, 1)                                    # [kv_heads, S, hd]
        v = v.transpose(0, 1)
        if kv is not None and kv[l] is not None and kv[l][0].shape[1] > 0:
            k = torch.cat([kv[l][0].to(W.dtype), k], dim=1)
            v = torch.cat([kv[l][1].to(W.dtype), v], dim=1)
        new_kv.append((k, v))
        P = k.shape[1]
        kr = k.repeat_interleave(rep, dim=0)                     # [heads, P, hd]
        vr = v.repeat_interleave(rep, dim=0)
        sc = torch.einsum("shd,hpd->hsp", q, kr) * (cfg.head_dim ** -0.5)
        # causal: query i of this call sits at cache index P - S + i
        qi = torch.arange(S)[:, None] + (P - S)
        mask = torch.arange(P)[None, :] > qi
        sc = sc.masked_fill(mask[None], float("-inf"))
        pr = torch.softmax(sc.float(), dim=-1).to(W.dtype)
        a = torch.einsum("hsp,hpd->shd", pr, vr).reshape(S, cfg.heads * cfg.head_dim)
        x = x + a @ W[p + "o_proj"].T
        hn = rms_norm(x, W[p + "post_ln"], cfg.rms_eps)
        g = hn @ W[p + "gate_proj"].T
        u = hn @ W[p + "up_proj"].T
        x = x + (torch.nn.functional.silu(g) * u) @ W[p + "down_proj"].T
    return rms_norm(x, W["final_norm"], cfg.rms_eps), new_kv


def sample_topk(logits: np.ndarray, temperature: float, top_k: int, rng) -> int:
    """code_predictor_server.py:87-92 with the random draw injected: argpartition top-k, softmax of
    (l - max) / max(T, 1e-6), one categorical draw over the partition's order."""
    top = np.argpartition(logits, -top_k)[-top_k:]
    tl = logits[top]
    pr = np.exp((tl - tl.max()) / max(temperature, 1e-6))
    pr /= pr.sum()
    return int(top[rng.choice(len(top), p=pr)])


def greedy(logits: np.ndarray) -> int:
    return int(np.argmax(logits))


def predict(hidden_state: np.ndarray, code0_embed: np.ndarray, W: Weights, cfg: CPConfig,
            sampler: Callable[[np.ndarray], int] = greedy, logits_out: Optional[list] = None) -> List[int]:
    """code_predictor_server.py:94-140, sequential prefill: position 0 = the talker's hidden state, position 1 = the
    embedding of code_0, group 0 sampled from lm_head_0; then group g from the embedding codec_emb_{g-1}[previous code]
    at position g + 1 through lm_head_g."""
    H = cfg.hidden
    h0 = torch.from_numpy(np.asarray(hidden_state, dtype=np.float32).reshape(-1)[:H].copy())[None]
    h1 = torch.from_numpy(np.asarray(code0_embed, dtype=np.float32).reshape(-1)[:H].copy())[None]
    with torch.no_grad():
        _, kv = step(h0, [0], None, W, cfg)
        out, kv = step(h1, [1], kv, W, cfg)
        codes: List[int] = []
        for g in range(cfg.groups):
            logits = (out[-1] @ W[f"lm_head_{g}"].T).to(torch.float32).numpy()
            if logits_out is not None:
                logits_out.append(logits.copy())
            tok = sampler(logits)
            codes.append(tok)
            if g + 1 < cfg.groups:
                emb = W[f"codec_emb_{g}"][tok][None]
                out, kv = step(emb, [g + 2], kv, W, cfg)
    return codes
