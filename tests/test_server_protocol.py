"""Wire-protocol parity over a real AF_UNIX socket.  Our VocoderServer (with a fake backend --
no GPU here) and the reference's VocoderServer (same fake model) are driven by the
reference's own client half (dual_npu/tts_client.py:78-108) and must answer byte for byte."""
import importlib
import importlib.util
import os
import socket
import struct
import threading
import time

import numpy as np
import pytest

from helpers import REFERENCE, fake_chunk_fn, load_reference_server, reference_server_with
from oracle import stitch_oracle as SO

LC = 122325


class FakeBackend:
    """Stands in for backend.Vocoder: the oracle stitcher over a fake model."""
    max_tokens = 64

    def __init__(self):
        self.fn = fake_chunk_fn(LC)

    def synthesize_pcm16(self, codes):
        if int(codes.max()) >= 2048 or int(codes.min()) < 0:
            raise RuntimeError("audio code outside [0, codebook_size)")
        return SO.to_pcm16(SO.synthesize(codes, self.fn, 64))

    def close(self):
        pass


def _our_server(sock_path):
    mod = importlib.import_module("qwen3-tts-axera-russian_b200.vocoder_server")
    srv = mod.VocoderServer.__new__(mod.VocoderServer)
    srv.socket_path = sock_path
    srv.is_onnx = False
    srv.voc = FakeBackend()
    srv.max_tokens = 64
    srv._running = True
    return srv


def _start(srv):
    t = threading.Thread(target=srv.serve, daemon=True)
    t.start()
    for _ in range(200):
        if os.path.exists(srv.socket_path):
            return t
        time.sleep(0.01)
    raise RuntimeError("server did not come up")


def _raw_request(sock_path, payload, read_reply=True):
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.settimeout(20)
    c.connect(sock_path)
    c.sendall(payload)
    if not read_reply:
        c.shutdown(socket.SHUT_WR)
    data = b""
    while True:
        d = c.recv(65536)
        if not d:
            break
        data += d
    c.close()
    return data


def _ref_client():
    path = os.path.join(REFERENCE, "dual_npu", "tts_client.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_tts_client", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cl = mod.Qwen3TTSClient.__new__(mod.Qwen3TTSClient)
    return cl


def _codes(n):
    return (np.arange(n * 16, dtype=np.int64).reshape(n, 16) * 7919 + n) % 2048


def test_round_trip_and_error_behaviour(tmp_path):
    sp = str(tmp_path / "voc.sock")
    srv = _our_server(sp)
    th = _start(srv)
    try:
        assert oct(os.stat(sp).st_mode & 0o777) == oct(0o666)          # vocoder_server.py:131
        for n in (1, 64, 100):
            reply = _raw_request(sp, SO.pack_request(_codes(n)))
            pcm = SO.unpack_reply(reply)
            (ns,) = struct.unpack("<i", reply[:4])
            assert ns == len(pcm) == SO.out_samples(n, LC)
            assert np.array_equal(pcm, SO.to_pcm16(SO.synthesize(_codes(n), fake_chunk_fn(LC), 64)))
        # malformed: n <= 0, n > 10000, short body, out-of-range code  ->  closed without a reply
        assert _raw_request(sp, struct.pack("<i", 0)) == b""
        assert _raw_request(sp, struct.pack("<i", 10001)) == b""
        assert _raw_request(sp, struct.pack("<i", -5)) == b""
        assert _raw_request(sp, struct.pack("<i", 4) + b"\x00" * 100, read_reply=False) == b""
        bad = _codes(3); bad[1, 2] = 5000
        assert _raw_request(sp, SO.pack_request(bad)) == b""
        # and the server is still alive
        assert len(_raw_request(sp, SO.pack_request(_codes(2)))) == 4 + 2 * 2 * 1920
    finally:
        srv._running = False
        th.join(timeout=5)
    assert not os.path.exists(sp)                                       # unlinked on shutdown


def test_byte_identical_to_reference_server(tmp_path, have_reference):
    if not have_reference:
        pytest.skip("/root/reference not present on this box")
    mod = load_reference_server()
    client = _ref_client()
    ours_path, ref_path = str(tmp_path / "a.sock"), str(tmp_path / "b.sock")
    ours = _our_server(ours_path)
    ref = reference_server_with(mod, fake_chunk_fn(LC), socket_path=ref_path)
    t1, t2 = _start(ours), _start(ref)
    try:
        for n in (1, 17, 64, 65, 97, 111, 200):
            codes = _codes(n)
            res = {}
            client.voc_socket = ours_path
            client._vocoder_chunk(codes.tolist(), "ours", res)          # tts_client.py:78-108
            client.voc_socket = ref_path
            client._vocoder_chunk(codes.tolist(), "ref", res)
            assert res["ours"].dtype == np.int16 and len(res["ours"]) > 0
            assert np.array_equal(res["ours"], res["ref"]), n
            assert _raw_request(ours_path, SO.pack_request(codes)) == _raw_request(ref_path, SO.pack_request(codes))
        for payload in (struct.pack("<i", 0), struct.pack("<i", 10001)):
            assert _raw_request(ours_path, payload) == _raw_request(ref_path, payload) == b""
    finally:
        ours._running = False
        ref._running = False
        t1.join(timeout=5); t2.join(timeout=5)


def test_server_rejects_non_b200_models():
    mod = importlib.import_module("qwen3-tts-axera-russian_b200.vocoder_server")
    with pytest.raises(RuntimeError):
        mod.VocoderServer("vocoder_traced_64.onnx", install_signal_handlers=False)
