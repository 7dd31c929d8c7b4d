"""The code-predictor oracle (oracle/code_predictor_oracle.py) on CPU:
  * its decode step against golden vectors from the executable sibling (``transformers``
    ``Qwen3OmniMoeTalkerCodePredictorModel``; tests/golden/make_sibling_cp_golden.py), as a whole sequence, token by token
    with the cache, and as the reference's 2-token batch prefill;
  * its predict loop and sampler against the reference's own ``CodePredictorServer.predict`` / ``_sample`` EXECUTING
    (/root/reference/dual_npu/code_predictor_server.py:87-140) with the model call replaced by the oracle's step."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import REFERENCE
from oracle import code_predictor_oracle as CP

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sibling_cp.npz")


def _golden():
    G = np.load(GOLD)
    cfg = CP.CPConfig(hidden=64, layers=2, heads=4, kv_heads=2, head_dim=16, inter=96, vocab=32, groups=4,
                      rms_eps=float(G["rms_eps"]), rope_theta=float(G["rope_theta"]))
    w = {k: G[k] for k in G.files if k.startswith("layer_") or k == "final_norm"}
    for g in range(cfg.groups):
        w[f"codec_emb_{g}"] = np.zeros((cfg.vocab, cfg.hidden), np.float32)
        w[f"lm_head_{g}"] = np.zeros((cfg.vocab, cfg.hidden), np.float32)
    return G, cfg, CP.Weights(w)


def test_whole_sequence_matches_the_sibling():
    G, cfg, W = _golden()
    out, _ = CP.step(torch.from_numpy(G["x"]), list(range(6)), None, W, cfg)
    assert float(np.abs(out.numpy() - G["full"]).max()) < 5e-6


def test_token_by_token_with_the_cache_matches_the_sibling():
    G, cfg, W = _golden()
    kv = None
    for t in range(6):
        out, kv = CP.step(torch.from_numpy(G["x"][t:t + 1]), [t], kv, W, cfg)
        assert float(np.abs(out.numpy()[0] - G["steps"][t]).max()) < 5e-6, t
    assert kv[0][0].shape == (cfg.kv_heads, 6, cfg.head_dim)


def test_two_token_prefill_equals_two_single_steps():
    """The reference's --batch_prefill (code_predictor_server.py:107-118) feeds positions 0 and 1 in one call."""
    G, cfg, W = _golden()
    a, kva = CP.step(torch.from_numpy(G["x"][:2]), [0, 1], None, W, cfg)
    _, kv = CP.step(torch.from_numpy(G["x"][:1]), [0], None, W, cfg)
    b, kvb = CP.step(torch.from_numpy(G["x"][1:2]), [1], kv, W, cfg)
    assert float(np.abs(a.numpy()[1] - b.numpy()[0]).max()) < 2e-6
    assert float(np.abs(kva[1][0].numpy() - kvb[1][0].numpy()).max()) < 2e-6


def _reference_server():
    path = os.path.join(REFERENCE, "dual_npu", "code_predictor_server.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_cp_server", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_predict_loop_equals_the_reference_executing():
    """The reference's predict() with its ONNX session replaced by the oracle's step (same I/O contract, :77-85) and a
    seeded numpy RNG must give the codes of the oracle's predict() with the restated sampler on the same RNG stream."""
    mod = _reference_server()
    if mod is None:
        pytest.skip("/root/reference not present on this box")
    cfg = CP.CPConfig.tiny(hidden=mod.HIDDEN_SIZE, heads=4, kv_heads=2, head_dim=16, inter=64, layers=1, vocab=64, groups=15)
    w = CP.init_weights(cfg, 0)
    W = CP.Weights(w)
    talker_table = np.random.default_rng(5).standard_normal((32, cfg.hidden)).astype(np.float32)

    class Fake(mod.CodePredictorServer):
        def __init__(self):
            self.temperature, self.top_k, self.num_groups, self.batch_prefill = 0.7, 10, 15, False
            self.codec_embeddings = [w[f"codec_emb_{i}"] for i in range(15)]
            self.lm_heads = [w[f"lm_head_{i}"] for i in range(15)]
            self.codec_embedding = talker_table
            self.num_layers, self.head_dim, self.num_kv_heads = cfg.layers, cfg.head_dim, cfg.kv_heads

        def _ort_step(self, hidden, position, kv_caches):
            kv = None
            if kv_caches["past_k_0"].shape[2] > 0:
                kv = [(torch.from_numpy(kv_caches[f"past_k_{i}"][0]), torch.from_numpy(kv_caches[f"past_v_{i}"][0]))
                      for i in range(cfg.layers)]
            with torch.no_grad():
                out, nkv = CP.step(torch.from_numpy(np.ascontiguousarray(hidden[0])), [position], kv, W, cfg)
            new = {}
            for i in range(cfg.layers):
                new[f"past_k_{i}"] = nkv[i][0].numpy()[None]
                new[f"past_v_{i}"] = nkv[i][1].numpy()[None]
            return out.numpy()[None], new

    hidden_state = np.random.default_rng(1).standard_normal(cfg.hidden).astype(np.float32)
    code_0 = 7
    np.random.seed(123)
    ref_codes = Fake().predict(hidden_state, code_0)
    rs = np.random.RandomState(123)
    ours = CP.predict(hidden_state, talker_table[code_0], W, cfg, sampler=lambda l: CP.sample_topk(l, 0.7, 10, rs))
    assert ours == ref_codes and len(ours) == 15


def test_greedy_predict_is_reproducible_and_in_range():
    cfg = CP.CPConfig.tiny()
    W = CP.Weights(CP.init_weights(cfg, 1))
    hs = np.random.default_rng(2).standard_normal(cfg.hidden).astype(np.float32)
    e0 = np.random.default_rng(3).standard_normal(cfg.hidden).astype(np.float32)
    a = CP.predict(hs, e0, W, cfg)
    assert a == CP.predict(hs, e0, W, cfg) and len(a) == cfg.groups and all(0 <= c < cfg.vocab for c in a)


def test_production_size_oracle_matches_the_sibling_run_live():
    """The oracle at the PRODUCTION shape (1024 / 5 layers / 16-8 heads x 128 / 3072) against the sibling model built at
    that shape with the same random weights and run live, token by token with its own KV cache, over the 17 positions of
    a frame (the golden file pins a small model; this pins the dimensions the product runs)."""
    M = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeTalkerCodePredictorConfig
    cfg = CP.CPConfig()
    w = CP.init_weights(cfg, 2)
    W = CP.Weights(w)
    c = Qwen3OmniMoeTalkerCodePredictorConfig(hidden_size=cfg.hidden, intermediate_size=cfg.inter, num_hidden_layers=cfg.layers,
                                              num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, head_dim=cfg.head_dim,
                                              vocab_size=cfg.vocab, num_code_groups=cfg.groups + 1, max_position_embeddings=64,
                                              rms_norm_eps=cfg.rms_eps)
    c._attn_implementation = "eager"
    assert float(c.rope_parameters["rope_theta"]) == cfg.rope_theta
    m = M.Qwen3OmniMoeTalkerCodePredictorModel(c).eval()
    names = {"input_ln": "input_layernorm.weight", "q_proj": "self_attn.q_proj.weight", "k_proj": "self_attn.k_proj.weight",
             "v_proj": "self_attn.v_proj.weight", "o_proj": "self_attn.o_proj.weight", "q_norm": "self_attn.q_norm.weight",
             "k_norm": "self_attn.k_norm.weight", "post_ln": "post_attention_layernorm.weight",
             "gate_proj": "mlp.gate_proj.weight", "up_proj": "mlp.up_proj.weight", "down_proj": "mlp.down_proj.weight"}
    sd = m.state_dict()
    with torch.no_grad():
        for l in range(cfg.layers):
            for ours, theirs in names.items():
                sd[f"layers.{l}.{theirs}"].copy_(torch.from_numpy(w[f"layer_{l}_{ours}"]))
        sd["norm.weight"].copy_(torch.from_numpy(w["final_norm"]))
    xs = np.random.default_rng(21).standard_normal((cfg.groups + 2, cfg.hidden)).astype(np.float32)
    past = kv = None
    worst = 0.0
    with torch.no_grad():
        for t in range(len(xs)):
            r = m(inputs_embeds=torch.from_numpy(xs[None, t:t + 1]), past_key_values=past, use_cache=True,
                  cache_position=torch.tensor([t]), position_ids=torch.tensor([[t]]))
            past = r.past_key_values
            out, kv = CP.step(torch.from_numpy(xs[t:t + 1]), [t], kv, W, cfg)
            worst = max(worst, float(np.abs(out.numpy()[0] - r.last_hidden_state[0, 0].numpy()).max()))
    assert worst < 2e-5, worst
