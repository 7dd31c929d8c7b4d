#!/usr/bin/env python3
"""Parity of the B200 backend against the reference's own model, for a box that has it.

    python tools/onnx_parity.py --onnx vocoder_traced_64.onnx --weights <snapshot>/speech_tokenizer/model.safetensors

Needs what this image lacks: ``onnxruntime`` and the two files the reference builds / downloads
(/root/reference/scripts/export_vocoder_traced.py:28-35,74-99: HF snapshot c27fe8aa..., traced at T = 64, opset 17).
For each setting of ambiguity A1 (``transconv_trim``: "both" / "right") it loads the checkpoint through
``weights.from_speech_tokenizer``, runs the same random codes (``randint(0, 2048, (1, 64, 16))``, the reference's own
smoke input, :85,:137) through ONNX Runtime CPU FP32 -- session options as ``dual_npu/vocoder_server.py:40-44`` -- and
through ``voc_infer_chunks``, and prints the output lengths, SNR and max-abs error (gate: 60 dB / 1e-4).  The setting
whose length equals the ONNX output's is the real one; ``--rename ours=upstream`` fixes a checkpoint key name without a
code change.  Also reports the oracle (CPU restatement) against ONNX Runtime, which pins the oracle itself.
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def snr_db(ref, got):
    ref = ref.astype(np.float64); got = got.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - got) ** 2).sum(), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--onnx", required=True)
    ap.add_argument("--weights", required=True, help="speech_tokenizer safetensors file")
    ap.add_argument("--prefix", default="decoder.")
    ap.add_argument("--rename", action="append", default=[], metavar="OURS=UPSTREAM")
    ap.add_argument("--seeds", type=int, default=3)
    ap.add_argument("--no-gpu", action="store_true", help="oracle vs ONNX Runtime only")
    args = ap.parse_args()
    try:
        import onnxruntime as ort
    except ImportError:
        sys.exit("onnxruntime is not installed: this script is for a box that can run the reference's model")
    W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
    from oracle import vocoder_oracle as VO
    so = ort.SessionOptions()
    so.intra_op_num_threads = 4
    so.inter_op_num_threads = 1
    sess = ort.InferenceSession(args.onnx, so, providers=["CPUExecutionProvider"])
    T = sess.get_inputs()[0].shape[1]
    rename = dict(r.split("=", 1) for r in args.rename)
    report = {"onnx": args.onnx, "max_tokens": T, "settings": {}}
    for trim in ("both", "right"):
        cfg, w = W.from_speech_tokenizer(args.weights, prefix=args.prefix, rename=rename, transconv_trim=trim, chunk_frames=T)
        voc = None
        if not args.no_gpu:
            backend = importlib.import_module("qwen3-tts-axera-russian_b200.backend")
            voc = backend.Vocoder(cfg, w, wave=1)
            voc.set_option("gemm", "tc")
        rows = []
        for seed in range(args.seeds):
            codes = np.random.default_rng(seed).integers(0, cfg.codebook_size, (1, T, 16), dtype=np.int64)
            ref = sess.run(None, {"audio_codes": codes})[0].reshape(-1)
            orc, _ = VO.forward(codes, VO.Weights(w), cfg)
            orc = orc.numpy().reshape(-1)
            row = {"seed": seed, "onnx_len": int(ref.size), "our_len": int(orc.size)}
            n = min(ref.size, orc.size)
            row["oracle_vs_onnx"] = {"snr_db": float(snr_db(ref[:n], orc[:n])), "max_abs": float(np.abs(ref[:n] - orc[:n]).max())}
            if voc is not None:
                got = voc.infer_chunks(codes).reshape(-1)
                row["b200_vs_onnx"] = {"snr_db": float(snr_db(ref[:n], got[:n])), "max_abs": float(np.abs(ref[:n] - got[:n]).max())}
                row["gate_pass"] = bool(row["onnx_len"] == row["our_len"] and row["b200_vs_onnx"]["snr_db"] >= 60.0
                                        and row["b200_vs_onnx"]["max_abs"] <= 1e-4)
            rows.append(row)
        if voc is not None:
            # how the real checkpoint's activations sit in the unscaled split-fp16 operand format (voc_operand_report):
            # a saturated count > 0 or an rms far below 2^-3 names the layer that needs an activation scale
            voc.set_option("operand_stats", "1")
            voc.infer_chunks(codes)
            report.setdefault("operand_ranges", {})[trim] = [
                dict(r, flag=("saturated" if r["saturated"] else "small" if r["rms"] < 2.0 ** -6 else "ok"))
                for r in voc.operand_report()]
            voc.close()
        report["settings"][trim] = rows
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
