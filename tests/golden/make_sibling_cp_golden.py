#!/usr/bin/env python3
"""Golden vectors for the code-predictor oracle from the executable sibling implementation in this image,
``transformers`` ``Qwen3OmniMoeTalkerCodePredictorModel`` (non-reference evidence; the reference's own model is the
un-vendored qwen_tts package, scripts/export_code_predictor_onnx.py:70-91): a small random model run (a) on a whole
6-token sequence and (b) token by token with its KV cache, which is the contract of the reference's decode-step graph
(dual_npu/code_predictor_server.py:77-85).  Writes tests/golden/sibling_cp.npz.

    python tests/golden/make_sibling_cp_golden.py
"""
import os

import numpy as np
import torch

import transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe as M
from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeTalkerCodePredictorConfig

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.manual_seed(0)
    c = Qwen3OmniMoeTalkerCodePredictorConfig(hidden_size=64, intermediate_size=96, num_hidden_layers=2, num_attention_heads=4,
                                              num_key_value_heads=2, head_dim=16, vocab_size=32, num_code_groups=5,
                                              max_position_embeddings=64)
    c._attn_implementation = "eager"
    m = M.Qwen3OmniMoeTalkerCodePredictorModel(c).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("norm.weight") or n.endswith("layernorm.weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif "proj" in n:
                p.copy_(torch.randn_like(p) / np.sqrt(p.shape[1]))
    x = torch.randn(1, 6, 64)
    out = {}
    with torch.no_grad():
        full = m(inputs_embeds=x, use_cache=False).last_hidden_state[0].numpy()
        past, steps = None, []
        for t in range(6):
            r = m(inputs_embeds=x[:, t:t + 1], past_key_values=past, use_cache=True,
                  cache_position=torch.tensor([t]), position_ids=torch.tensor([[t]]))
            past = r.past_key_values
            steps.append(r.last_hidden_state[0, 0].numpy())
    out["x"] = x[0].numpy()
    out["full"] = full
    out["steps"] = np.stack(steps)
    sd = m.state_dict()
    names = {"input_ln": "input_layernorm.weight", "q_proj": "self_attn.q_proj.weight", "k_proj": "self_attn.k_proj.weight",
             "v_proj": "self_attn.v_proj.weight", "o_proj": "self_attn.o_proj.weight", "q_norm": "self_attn.q_norm.weight",
             "k_norm": "self_attn.k_norm.weight", "post_ln": "post_attention_layernorm.weight",
             "gate_proj": "mlp.gate_proj.weight", "up_proj": "mlp.up_proj.weight", "down_proj": "mlp.down_proj.weight"}
    for l in range(2):                       # the npz names of scripts/export_code_predictor_weights.py:50-63
        for ours, theirs in names.items():
            out[f"layer_{l}_{ours}"] = sd[f"layers.{l}.{theirs}"].numpy()
    out["final_norm"] = sd["norm.weight"].numpy()
    out["rms_eps"] = np.float64(c.rms_norm_eps)
    out["rope_theta"] = np.float64(c.rope_parameters["rope_theta"])
    print("full vs stepwise max diff", float(np.abs(full - out["steps"]).max()))
    np.savez_compressed(os.path.join(HERE, "sibling_cp.npz"), **out)
    print("wrote sibling_cp.npz", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
