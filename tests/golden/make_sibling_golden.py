#!/usr/bin/env python3
"""Generates tests/golden/sibling_blocks.npz: inputs, weights and outputs of the building
blocks of the in-image *sibling* implementation of the same decoder lineage
(``transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe``: CausalConvNet,
CausalTransConvNet, ConvNeXtBlock, SnakeBeta, DecoderResidualUnit, DecoderBlock,
Code2WavTransformerModel) at small sizes.

NON-REFERENCE evidence: the reference's own model code lives in the
un-vendored ``qwen-tts`` package; these vectors pin the oracle's block semantics (causal
padding, transposed-conv trim, SnakeBeta, ConvNeXt, RoPE/sliding-window attention,
LayerScale) against executable code, not against the reference itself.

    python tests/golden/make_sibling_golden.py
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


def randomize(mod, g):
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.3)


def main():
    g = torch.Generator().manual_seed(1234)
    torch.set_grad_enabled(False)
    # 1. causal convs, dilation 1/3/9
    for d in (1, 3, 9):
        m = M.Qwen3OmniMoeCausalConvNet(8, 12, 7, dilation=d)
        randomize(m, g)
        x = torch.randn(2, 8, 40, generator=g)
        put(f"conv_d{d}.x", x); put(f"conv_d{d}.w", m.conv.weight); put(f"conv_d{d}.b", m.conv.bias)
        put(f"conv_d{d}.y", m(x))
    # 2. transposed convs: k = 2s (decoder blocks) and k = s (upsample stages)
    for s, k in ((8, 16), (5, 10), (4, 8), (3, 6), (2, 2)):
        m = M.Qwen3OmniMoeCausalTransConvNet(8, 6, k, s)
        randomize(m, g)
        x = torch.randn(2, 8, 11, generator=g)
        put(f"convt_s{s}.x", x); put(f"convt_s{s}.w", m.conv.weight); put(f"convt_s{s}.b", m.conv.bias)
        put(f"convt_s{s}.y", m(x))
    # 3. SnakeBeta
    m = M.SnakeBeta(8)
    randomize(m, g)
    x = torch.randn(2, 8, 33, generator=g) * 3
    put("snake.x", x); put("snake.alpha", m.alpha); put("snake.beta", m.beta); put("snake.y", m(x))
    # 4. ConvNeXt block
    m = M.Qwen3OmniMoeConvNeXtBlock(16)
    randomize(m, g)
    x = torch.randn(2, 16, 21, generator=g)
    put("cnx.x", x)
    for k, t in (("dw.w", m.dwconv.conv.weight), ("dw.b", m.dwconv.conv.bias), ("ln.w", m.norm.weight),
                 ("ln.b", m.norm.bias), ("pw1.w", m.pwconv1.weight), ("pw1.b", m.pwconv1.bias),
                 ("pw2.w", m.pwconv2.weight), ("pw2.b", m.pwconv2.bias), ("gamma", m.gamma)):
        put("cnx." + k, t)
    put("cnx.y", m(x))
    # 5. residual unit
    m = M.Qwen3OmniMoeCode2WavDecoderResidualUnit(8, dilation=3)
    randomize(m, g)
    x = torch.randn(2, 8, 50, generator=g)
    put("ru.x", x)
    for k, t in (("snake1.alpha", m.act1.alpha), ("snake1.beta", m.act1.beta), ("conv1.w", m.conv1.conv.weight),
                 ("conv1.b", m.conv1.conv.bias), ("snake2.alpha", m.act2.alpha), ("snake2.beta", m.act2.beta),
                 ("conv2.w", m.conv2.conv.weight), ("conv2.b", m.conv2.conv.bias)):
        put("ru." + k, t)
    put("ru.y", m(x))
    # 6. decoder block (Snake -> transposed conv -> 3 residual units), and 7. the transformer
    cfg = Qwen3OmniMoeCode2WavConfig(hidden_size=32, num_attention_heads=2, num_key_value_heads=2,
                                     intermediate_size=48, num_hidden_layers=2, decoder_dim=32,
                                     upsample_rates=(4, 3), upsampling_ratios=(2,), sliding_window=5,
                                     layer_scale_initial_scale=0.5)
    cfg._attn_implementation = "eager"
    m = M.Qwen3OmniMoeCode2WavDecoderBlock(cfg, 1)          # 16 -> 8 channels, stride 3
    randomize(m, g)
    x = torch.randn(2, 16, 9, generator=g)
    put("blk.x", x)
    put("blk.snake.alpha", m.block[0].alpha); put("blk.snake.beta", m.block[0].beta)
    put("blk.convt.w", m.block[1].conv.weight); put("blk.convt.b", m.block[1].conv.bias)
    for j in range(3):
        r = m.block[2 + j]
        for k, t in (("snake1.alpha", r.act1.alpha), ("snake1.beta", r.act1.beta), ("conv1.w", r.conv1.conv.weight),
                     ("conv1.b", r.conv1.conv.bias), ("snake2.alpha", r.act2.alpha), ("snake2.beta", r.act2.beta),
                     ("conv2.w", r.conv2.conv.weight), ("conv2.b", r.conv2.conv.bias)):
            put(f"blk.ru.{j}.{k}", t)
    put("blk.y", m(x))
    t = M.Qwen3OmniMoeCode2WavTransformerModel(cfg)
    randomize(t, g)
    t.eval()
    x = torch.randn(2, 12, 32, generator=g)
    put("xf.x", x)
    for l, ly in enumerate(t.layers):
        for k, p in (("ln1.w", ly.input_layernorm.weight), ("q.w", ly.self_attn.q_proj.weight),
                     ("k.w", ly.self_attn.k_proj.weight), ("v.w", ly.self_attn.v_proj.weight),
                     ("o.w", ly.self_attn.o_proj.weight), ("ls_attn", ly.self_attn_layer_scale.scale),
                     ("ln2.w", ly.post_attention_layernorm.weight), ("gate.w", ly.mlp.gate_proj.weight),
                     ("up.w", ly.mlp.up_proj.weight), ("down.w", ly.mlp.down_proj.weight),
                     ("ls_mlp", ly.mlp_layer_scale.scale)):
            put(f"xf.{l}.{k}", p)
    put("xf.norm.w", t.norm.weight)
    put("xf.y", t(inputs_embeds=x).last_hidden_state)
    np.savez_compressed(os.path.join(HERE, "sibling_blocks.npz"), **out)
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "sibling_blocks.npz")), "bytes")


if __name__ == "__main__":
    main()
