"""CPU oracle of the chunker / overlap-crossfade stitcher / PCM16 conversion / wire format.
TEST INFRASTRUCTURE, NOT PRODUCT CODE (see vocoder_oracle.py header for who may import it).

PARITY STATUS: **pinned.**  Unlike the model graph, this part of the path is reference
code that runs in this image.  ``tests/test_stitch_oracle.py`` executes
``/root/reference/dual_npu/vocoder_server.py``'s own ``synthesize`` (when that tree is
present) and ``tests/golden/stitch_*.npz`` holds outputs generated from it by
``tests/golden/make_stitch_golden.py`` for the GPU box where the tree is absent.

Restated from /root/reference/dual_npu/vocoder_server.py:
  synthesize()        :73-121   (single-chunk branch :77-81, loop :83-119)
  float -> int16      :175
  wire format         :8-12, :143-178
"""
from __future__ import annotations

import struct
from typing import Callable, List, Tuple

import numpy as np

SAMPLES_PER_TOKEN = 1920          # vocoder_server.py:30
OVERLAP = 16                      # vocoder_server.py:84


def window_starts(n_tokens: int, max_tokens: int = 64) -> List[int]:
    """Start frame of every window the reference infers (:88-119: ``while chunk_start < n``,
    step = max_tokens - OVERLAP; the comment "56" at :86 is wrong, the value is 48)."""
    if n_tokens <= max_tokens:
        return [0]
    step = max_tokens - OVERLAP
    return list(range(0, n_tokens, step))


def synthesize(codes: np.ndarray, infer_chunk: Callable[[np.ndarray], np.ndarray],
               max_tokens: int = 64) -> np.ndarray:
    """Plain restatement of ``VocoderServer.synthesize`` with python-slice semantics."""
    n = len(codes)
    if n <= max_tokens:
        padded = np.zeros((1, max_tokens, 16), dtype=np.int64)     # pad with code 0 (:78)
        padded[0, :n, :] = codes[:, :16]
        return infer_chunk(padded)[: n * SAMPLES_PER_TOKEN]
    ov = OVERLAP * SAMPLES_PER_TOKEN
    # np.linspace(1, 0, ov, dtype=f32): computed in float64 then cast (:108)
    fade_out = np.linspace(1.0, 0.0, ov, dtype=np.float32)
    fade_in = (1.0 - fade_out).astype(np.float32)
    result = np.zeros(0, dtype=np.float32)
    for start in window_starts(n, max_tokens):
        ln = min(start + max_tokens, n) - start
        padded = np.zeros((1, max_tokens, 16), dtype=np.int64)
        padded[0, :ln, :] = codes[start:start + ln, :16]
        a = infer_chunk(padded)[: ln * SAMPLES_PER_TOKEN]
        if start == 0:
            result = a
        elif len(result) >= ov and len(a) >= ov:
            blended = result[-ov:] * fade_out + a[:ov] * fade_in
            result = np.concatenate([result[:-ov], blended, a[ov:]])
        else:
            # short last window (1..15 frames): appended un-blended (:116-117)
            result = np.concatenate([result, a])
    return result


def out_samples(n_tokens: int, chunk_samples: int, max_tokens: int = 64) -> int:
    """Length of ``synthesize``'s result given the model emits ``chunk_samples`` per window."""
    spt = SAMPLES_PER_TOKEN
    if n_tokens <= max_tokens:
        return min(chunk_samples, n_tokens * spt)
    ov = OVERLAP * spt
    total = 0
    for start in window_starts(n_tokens, max_tokens):
        ln = min(start + max_tokens, n_tokens) - start
        a = min(chunk_samples, ln * spt)
        if start == 0:
            total = a
        elif total >= ov and a >= ov:
            total = total + a - ov
        else:
            total = total + a
    return total


def to_pcm16(audio: np.ndarray) -> np.ndarray:
    """``np.clip(audio * 32767, -32768, 32767).astype(np.int16)`` (:175): float32 multiply,
    truncation toward zero."""
    return np.clip(np.asarray(audio, dtype=np.float32) * 32767, -32768, 32767).astype(np.int16)


def pack_request(codes: np.ndarray) -> bytes:
    """Client half of the wire format (/root/reference/dual_npu/tts_client.py:81-86)."""
    codes = np.ascontiguousarray(codes, dtype=np.int64)
    return struct.pack("<i", len(codes)) + codes.tobytes()


def unpack_reply(buf: bytes) -> np.ndarray:
    (n,) = struct.unpack("<i", buf[:4])
    return np.frombuffer(buf[4:4 + 2 * n], dtype=np.int16)
