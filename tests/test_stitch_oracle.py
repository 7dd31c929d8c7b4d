"""The chunker / stitcher / PCM16 restatement (oracle/stitch_oracle.py) and the C library's
host-side plan, pinned against the reference: live against
/root/reference/dual_npu/vocoder_server.py when that tree exists, and always against the
golden digests generated from it (tests/golden/stitch_golden.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import fake_chunk_fn, load_reference_server, reference_server_with
from oracle import stitch_oracle as SO

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stitch_golden.json")


def _case_codes(n):
    return (np.arange(n * 16, dtype=np.int64).reshape(n, 16) * 7919 + n) % 2048


def _sha(a, dt):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dt).tobytes()).hexdigest()


def test_oracle_matches_golden_digests():
    cases = json.load(open(GOLDEN))["cases"]
    assert len(cases) >= 40
    for c in cases:
        fn = fake_chunk_fn(c["chunk_samples"])
        audio = SO.synthesize(_case_codes(c["n"]), fn, 64)
        assert len(audio) == c["len"], c
        assert audio.dtype == np.float32
        assert _sha(audio, "<f4") == c["f32_sha256"], c["n"]
        assert _sha(SO.to_pcm16(audio), "<i2") == c["pcm_sha256"], c["n"]
        assert SO.out_samples(c["n"], c["chunk_samples"]) == c["len"]


def test_known_lengths_from_survey_appendix_b():
    # SURVEY.md Appendix B (executed reference): n -> output frames at 1920 samples per frame
    table = {1: 1, 63: 63, 64: 64, 65: 65, 96: 96, 97: 98, 100: 104, 111: 126, 112: 112, 113: 113,
             200: 208, 7500: 7512}
    for n, frames in table.items():
        assert SO.out_samples(n, 64 * 1920) == frames * 1920, n
    assert SO.window_starts(7500) == list(range(0, 7500, 48))
    assert len(SO.window_starts(7500)) == 157


def test_pcm16_truncates_toward_zero():
    x = np.array([0.99999, -1, 1, 0.5, -0.5, 1.2, -1.2, 0.0, 1e-6], dtype=np.float32)
    assert SO.to_pcm16(x).tolist() == [32766, -32767, 32767, 16383, -16383, 32767, -32768, 0, 0]


@pytest.mark.parametrize("Lc", [122880, 122325, 61440, 40000])
def test_oracle_equals_live_reference(Lc, have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present on this box (golden digests cover it)")
    mod = load_reference_server()
    fn = fake_chunk_fn(Lc)
    srv = reference_server_with(mod, fn)
    for n in list(range(1, 70)) + list(range(90, 130)) + [143, 144, 145, 159, 160, 161, 208, 400]:
        codes = _case_codes(n)
        ref = srv.synthesize(codes)
        got = SO.synthesize(codes, fn, 64)
        assert got.shape == ref.shape and np.array_equal(got, ref), (Lc, n)


def test_extra_code_columns_are_ignored(have_reference):
    # the reference slices codes_array[:, :16] (vocoder_server.py:79,94)
    fn = fake_chunk_fn(122880)
    codes = np.concatenate([_case_codes(70), np.full((70, 3), 5, dtype=np.int64)], axis=1)
    a = SO.synthesize(codes, fn, 64)
    b = SO.synthesize(codes[:, :16], fn, 64)
    assert np.array_equal(a, b)


# ---- the C library's planner (host logic, no GPU) -------------------------------------

def _assemble(meta, ov, chunks, fo, fi, total):
    """What the one-launch stitch kernel computes, in numpy, from the plan's meta."""
    out = np.zeros(total, dtype=np.float32)
    for w, (dst, a_len, blended, next_blended, prev_a, _start) in enumerate(meta):
        cur = chunks[w][:a_len]
        lo = 0
        if blended:
            r = chunks[w - 1][prev_a - ov: prev_a]
            out[dst:dst + ov] = r * fo + cur[:ov] * fi
            lo = ov
        hi = a_len - ov if next_blended else a_len
        out[dst + lo: dst + hi] = cur[lo:hi]
    return out


@pytest.mark.parametrize("Lc", [122880, 122325])
def test_c_plan_reproduces_reference_output(backend, Lc):
    fn = fake_chunk_fn(Lc)
    fo, fi = backend.fade_tables(16 * 1920)
    ro = np.linspace(1.0, 0.0, 16 * 1920, dtype=np.float32)
    assert np.array_equal(fo, ro) and np.array_equal(fi, 1.0 - ro)
    for n in [1, 17, 64, 65, 96, 97, 111, 112, 113, 200, 333]:
        codes = _case_codes(n)
        meta, total, pairwise = backend.plan(64, Lc, n)
        assert pairwise
        assert total == SO.out_samples(n, Lc)
        assert [m[5] for m in meta] == SO.window_starts(n)
        chunks = []
        for m in meta:
            s = int(m[5])
            ln = min(s + 64, n) - s
            padded = np.zeros((1, 64, 16), dtype=np.int64)
            padded[0, :ln] = codes[s:s + ln]
            chunks.append(fn(padded))
        got = _assemble(meta, 16 * 1920, chunks, fo, fi, total)
        ref = SO.synthesize(codes, fn, 64)
        assert np.array_equal(got, ref), n


def test_c_plan_flags_non_pairwise_regime(backend):
    # a model emitting fewer than two overlaps per window: blends read blended samples
    meta, total, pairwise = backend.plan(64, 40000, 200)
    assert not pairwise
    assert total == SO.out_samples(200, 40000)
