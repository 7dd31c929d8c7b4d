#!/usr/bin/env python3
"""Generates tests/golden/sibling_model.npz: everything the in-image sibling implementation (``transformers``
``Qwen3OmniMoeCode2Wav``) runs after its code embedding -- the pre-transformer (2 layers, RoPE, sliding-window
causal attention, LayerScale, SwiGLU), the up-sampling stages, conv-in, the four decoder blocks with the production
strides, the head and the clamp -- on a small random model, as ONE forward from a latent ``[2, 12, 32]``, with the
weights under the oracle's names.  Together with an identity front end (codebook 0 = the latent frames, unit
projections) this lets both the oracle and the CUDA path be run against executable third-party code for M3-M9.

NON-REFERENCE evidence, same status as make_sibling_golden.py.

    python tests/golden/make_sibling_model_golden.py
"""
import os

import numpy as np
import torch

import transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe as M
from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeCode2WavConfig

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


def put(k, t):
    out[k] = t.detach().to(torch.float32).numpy().copy()


def main():
    g = torch.Generator().manual_seed(97531)
    torch.set_grad_enabled(False)
    cfg = Qwen3OmniMoeCode2WavConfig(hidden_size=32, num_attention_heads=2, num_key_value_heads=2,
                                     intermediate_size=48, num_hidden_layers=2, decoder_dim=64,
                                     upsample_rates=(8, 5, 4, 3), upsampling_ratios=(2, 2), sliding_window=5,
                                     codebook_size=32, num_quantizers=16, layer_scale_initial_scale=0.5)
    cfg._attn_implementation = "eager"
    m = M.Qwen3OmniMoeCode2Wav(cfg)
    m.eval()
    for name, p in m.named_parameters():
        if name.endswith("alpha") or name.endswith("beta"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)              # log-scale Snake parameters
        elif name.endswith("gamma") or name.endswith("scale"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.3 + 0.5)        # ConvNeXt gamma, LayerScale
        elif "norm" in name and name.endswith("weight"):
            p.copy_(torch.randn(p.shape, generator=g) * 0.1 + 1.0)
        elif name.endswith("bias") or p.dim() == 1:
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        else:
            p.copy_(torch.randn(p.shape, generator=g) / p[0].numel() ** 0.5)
    m.decoder[-1].conv.weight.mul_(0.03)          # keep most of the output inside the clamp, some of it outside
    m.decoder[-1].conv.bias.mul_(0.1)
    T = 12
    hidden = torch.randn(2, T, cfg.hidden_size, generator=g)
    put("hidden", hidden)
    t = m.pre_transformer
    for l, ly in enumerate(t.layers):
        for k, p in (("ln1.w", ly.input_layernorm.weight), ("q.w", ly.self_attn.q_proj.weight),
                     ("k.w", ly.self_attn.k_proj.weight), ("v.w", ly.self_attn.v_proj.weight),
                     ("o.w", ly.self_attn.o_proj.weight), ("ls_attn", ly.self_attn_layer_scale.scale),
                     ("ln2.w", ly.post_attention_layernorm.weight), ("gate.w", ly.mlp.gate_proj.weight),
                     ("up.w", ly.mlp.up_proj.weight), ("down.w", ly.mlp.down_proj.weight),
                     ("ls_mlp", ly.mlp_layer_scale.scale)):
            put(f"xf.{l}.{k}", p)
    put("xf.norm.w", t.norm.weight)
    x = t(inputs_embeds=hidden).last_hidden_state
    put("xf_out", x)
    x = x.permute(0, 2, 1)
    for u, blocks in enumerate(m.upsample):
        ct, cn = blocks[0], blocks[1]
        put(f"up.{u}.convt.w", ct.conv.weight); put(f"up.{u}.convt.b", ct.conv.bias)
        for k, tt in (("dw.w", cn.dwconv.conv.weight), ("dw.b", cn.dwconv.conv.bias), ("ln.w", cn.norm.weight),
                      ("ln.b", cn.norm.bias), ("pw1.w", cn.pwconv1.weight), ("pw1.b", cn.pwconv1.bias),
                      ("pw2.w", cn.pwconv2.weight), ("pw2.b", cn.pwconv2.bias), ("gamma", cn.gamma)):
            put(f"up.{u}.{k}", tt)
        for blk in blocks:
            x = blk(x)
    dec = m.decoder
    put("dec.conv_in.w", dec[0].conv.weight); put("dec.conv_in.b", dec[0].conv.bias)
    nb = len(cfg.upsample_rates)
    for b in range(nb):
        blk = dec[1 + b].block
        put(f"dec.{b}.snake.alpha", blk[0].alpha); put(f"dec.{b}.snake.beta", blk[0].beta)
        put(f"dec.{b}.convt.w", blk[1].conv.weight); put(f"dec.{b}.convt.b", blk[1].conv.bias)
        for j in range(3):
            r = blk[2 + j]
            for k, tt in (("snake1.alpha", r.act1.alpha), ("snake1.beta", r.act1.beta), ("conv1.w", r.conv1.conv.weight),
                          ("conv1.b", r.conv1.conv.bias), ("snake2.alpha", r.act2.alpha), ("snake2.beta", r.act2.beta),
                          ("conv2.w", r.conv2.conv.weight), ("conv2.b", r.conv2.conv.bias)):
                put(f"dec.{b}.ru.{j}.{k}", tt)
    put("head.snake.alpha", dec[1 + nb].alpha); put("head.snake.beta", dec[1 + nb].beta)
    put("head.conv.w", dec[2 + nb].conv.weight); put("head.conv.b", dec[2 + nb].conv.bias)
    wav = x
    for block in dec:
        wav = block(wav)
    put("wav_unclamped", wav)
    put("wav", wav.clamp(min=-1, max=1))
    np.savez_compressed(os.path.join(HERE, "sibling_model.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "sibling_model.npz")), "bytes; wav", tuple(wav.shape),
          "rms", float(wav.pow(2).mean().sqrt()), "clamped fraction", float((wav.abs() > 1).float().mean()))


if __name__ == "__main__":
    main()
