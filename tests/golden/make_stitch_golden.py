#!/usr/bin/env python3
"""Generates tests/golden/stitch_golden.json by EXECUTING the reference's own
``VocoderServer.synthesize`` (/root/reference/dual_npu/vocoder_server.py:73-121) and its
PCM16 line (:175) with a deterministic fake ``_inference_chunk`` (tests/helpers.py).
Run in the build container (where /root/reference exists):  python tests/golden/make_stitch_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from helpers import fake_chunk_fn, load_reference_server, reference_server_with  # noqa: E402

CASES_N = [1, 2, 15, 16, 17, 63, 64, 65, 80, 96, 97, 100, 111, 112, 113, 128, 144, 145, 160, 200, 401, 1000]
CHUNK_SAMPLES = [122880, 122325]     # transconv_trim = right / both


def main():
    mod = load_reference_server()
    assert mod is not None, "needs /root/reference"
    out = {"generator": "tests/golden/make_stitch_golden.py", "cases": []}
    for Lc in CHUNK_SAMPLES:
        fn = fake_chunk_fn(Lc)
        srv = reference_server_with(mod, fn)
        for n in CASES_N:
            codes = (np.arange(n * 16, dtype=np.int64).reshape(n, 16) * 7919 + n) % 2048
            audio = srv.synthesize(codes)
            pcm = np.clip(audio * 32767, -32768, 32767).astype(np.int16)   # :175
            out["cases"].append({
                "n": n, "chunk_samples": Lc, "len": int(len(audio)),
                "f32_sha256": hashlib.sha256(np.ascontiguousarray(audio, dtype="<f4").tobytes()).hexdigest(),
                "pcm_sha256": hashlib.sha256(np.ascontiguousarray(pcm, dtype="<i2").tobytes()).hexdigest(),
                "head": [float(x) for x in audio[:4]], "tail": [float(x) for x in audio[-4:]],
            })
    with open(os.path.join(HERE, "stitch_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
