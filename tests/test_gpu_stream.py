"""Carried-state decode (opt-in): voc_stream_decode_* decodes one sequence piece by piece, every causal
layer carrying its left context from call to call, and must reproduce the decoder run ONCE over the whole sequence
-- the oracle (oracle/vocoder_oracle.py, the CPU restatement of the upstream decoder) on [1, n, 16] codes with no
windows at all -- under transconv_trim = "right".  Gate as for a window: SNR >= 60 dB and max-abs <= 1e-4."""
import numpy as np
import pytest

from oracle import vocoder_oracle as VO

pytestmark = pytest.mark.gpu


def _codes(cfg, n, seed):
    return np.random.default_rng(seed).integers(0, cfg.codebook_size, (n, 16), dtype=np.int64)


def _check(tag, ref, got):
    snr = VO.snr_db(ref, got)
    mx = float(np.abs(ref.astype(np.float64) - got).max())
    print(f"{tag}: SNR {snr:.1f} dB max-abs {mx:.3e} rms {float(np.sqrt((ref.astype(np.float64) ** 2).mean())):.3f}")
    assert snr >= 60.0 and mx <= 1e-4, (tag, snr, mx)


def test_small_architecture_any_split_equals_the_unchunked_oracle(pkg, backend):
    """Sliding window 5 < piece lengths < sequence: the attention history, every convolution halo (up to 54 rows at
    dilation 9) and one-frame pieces are all exercised."""
    cfg = pkg.VocoderConfig.tiny(transconv_trim="right", chunk_frames=8, sliding_window=5, decoder_dim=128)
    w = pkg.init_weights(cfg, 0)
    n = 50
    codes = _codes(cfg, n, 3)
    ref, _ = VO.forward(codes[None], VO.Weights(w), cfg)
    ref = ref.numpy()[0]
    assert ref.shape == (n * 1920,)
    voc = backend.Vocoder(cfg, w, wave=4)
    for split in ([50], [7, 13, 1, 29], [1] * 5 + [45], [25, 25]):
        voc.stream_reset()
        parts, at = [], 0
        for m in split:
            parts.append(voc.stream_decode(codes[at:at + m]))
            at += m
        got = np.concatenate(parts)
        assert got.shape == ref.shape
        _check(f"tiny split {split[:4]}", ref, got)
    # the int16 form uses the reference's truncating conversion (vocoder_server.py:175)
    voc.stream_reset()
    pcm = voc.stream_decode(codes, pcm16=True)
    from oracle import stitch_oracle as SO
    d = np.abs(pcm.astype(np.int32) - SO.to_pcm16(ref).astype(np.int32))
    assert int(d.max()) <= 2
    voc.close()


def test_internal_segmentation_is_invisible(pkg, backend):
    """A call longer than what the activation pools hold for one sequence (wave * chunk_frames - 1 frames) is cut
    into segments inside the library; the result equals the oracle all the same."""
    cfg = pkg.VocoderConfig.tiny(transconv_trim="right", chunk_frames=8, sliding_window=6, decoder_dim=128)
    w = pkg.init_weights(cfg, 1)
    n = 40
    codes = _codes(cfg, n, 5)
    ref, _ = VO.forward(codes[None], VO.Weights(w), cfg)
    voc = backend.Vocoder(cfg, w, wave=2)                        # 15 frames per segment
    got = voc.stream_decode(codes)
    _check("tiny, 3 internal segments", ref.numpy()[0], got)
    voc.close()


def test_production_architecture_two_pieces(pkg, backend):
    """Default architecture (decoder 1536 -> 96 channels, 8 transformer layers, window 72), 150 frames = 12 s decoded
    as 64 + 86 frames on the tensor-core path with the fused residual units, against the oracle on all 150 frames."""
    cfg = pkg.VocoderConfig(transconv_trim="right")
    w = pkg.init_weights(cfg, 0)
    n = 150
    codes = _codes(cfg, n, 11)
    ref, _ = VO.forward(codes[None], VO.Weights(w), cfg)
    ref = ref.numpy()[0]
    voc = backend.Vocoder(cfg, w, wave=4)
    voc.set_option("gemm", "tc")
    got = np.concatenate([voc.stream_decode(codes[:64]), voc.stream_decode(codes[64:])])
    assert got.shape == ref.shape == (n * 1920,)
    _check("production, 64 + 86 frames", ref, got)
    assert voc.simt_launches == 0
    # a new sequence after reset gives the same bits
    voc.stream_reset()
    again = np.concatenate([voc.stream_decode(codes[:64]), voc.stream_decode(codes[64:])])
    assert np.array_equal(got, again)
    # and the windowed reference mode is untouched by the stream state
    a = voc.synthesize(codes[:70])
    voc.stream_decode(codes[:10])
    assert np.array_equal(a, voc.synthesize(codes[:70]))
    voc.close()


def test_production_architecture_live_pieces(pkg, backend):
    """The live-streaming case (the talker emits 12.5 frames/s): the default architecture decoded in pieces of 1, 4, 1, 3,
    1, 16, 1, 4 and 1 frames -- block 0 then sees 32-row inputs, the fused residual units a few tiles, every layer its
    halo from the previous call -- against the oracle run once on all 32 frames.  A piece length that recurs runs eagerly
    the first time, is captured as a CUDA graph the second and replayed from the third on (the position reaches the
    attention kernel through device memory), so the one-frame pieces cover all three."""
    cfg = pkg.VocoderConfig(transconv_trim="right")
    w = pkg.init_weights(cfg, 0)
    pieces = [1, 4, 1, 3, 1, 16, 1, 4, 1]
    n = sum(pieces)
    codes = _codes(cfg, n, 13)
    ref, _ = VO.forward(codes[None], VO.Weights(w), cfg)
    ref = ref.numpy()[0]
    voc = backend.Vocoder(cfg, w, wave=4)
    voc.set_option("gemm", "tc")
    out, at = [], 0
    for k in pieces:
        out.append(voc.stream_decode(codes[at:at + k]))
        assert out[-1].shape == (k * 1920,)
        at += k
    got = np.concatenate(out)
    assert got.shape == ref.shape == (n * 1920,)
    _check("production, pieces 1 + 4 + 1 + 3 + 1 + 16 + 1 + 4 + 1", ref, got)
    assert voc.simt_launches == 0
    # the same stream again: every piece length has been seen, so this pass is graph replays and captures; same bits
    voc.stream_reset()
    again, at = [], 0
    for k in pieces:
        again.append(voc.stream_decode(codes[at:at + k]))
        at += k
    assert np.array_equal(got, np.concatenate(again))
    voc.close()


def test_stream_needs_the_causal_trim(pkg, backend):
    cfg = pkg.VocoderConfig.tiny(chunk_frames=8)                 # transconv_trim = "both": one input step of look-ahead
    voc = backend.Vocoder(cfg, pkg.init_weights(cfg, 0), wave=2)
    with pytest.raises(backend.VocoderError) as e:
        voc.stream_decode(_codes(cfg, 4, 0))
    assert e.value.code == backend.VOC_E_STATE
    voc.close()
