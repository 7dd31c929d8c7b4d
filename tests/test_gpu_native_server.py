"""The native server (csrc/voc_server.cpp) on a B200: the reference's wire protocol
(/root/reference/dual_npu/vocoder_server.py:8-12,123-190) spoken by the client half restated from
dual_npu/tts_client.py:84-105, concurrent connections coalesced into one batched synthesis, replies
bit-identical to the in-process backend, and the reference's close-without-reply error behaviour."""
import os
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SERVER = os.path.join(ROOT, "qwen3-tts-axera-russian_b200", "voc_server")


def client_request(sock_path, codes, raw_header=None, timeout=60.0):
    """tts_client._vocoder_chunk (dual_npu/tts_client.py:78-108): returns int16 PCM, or an empty array
    when the server closes without a reply."""
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    s.connect(sock_path)
    try:
        return _exchange(s, codes, raw_header)
    except (ConnectionResetError, BrokenPipeError):
        # the reference client turns any exception into an empty result (dual_npu/tts_client.py:106-108);
        # a server that closes on a bad header while the body is still in flight answers with a reset
        return np.zeros(0, dtype=np.int16)
    finally:
        s.close()


def _exchange(s, codes, raw_header):
    if True:
        n = len(codes)
        s.sendall(raw_header if raw_header is not None else struct.pack("<i", n))
        s.sendall(np.ascontiguousarray(codes, dtype="<i8").tobytes())
        hdr = b""
        while len(hdr) < 4:
            piece = s.recv(4 - len(hdr))
            if not piece:
                return np.zeros(0, dtype=np.int16)
            hdr += piece
        (n_samples,) = struct.unpack("<i", hdr)
        data = bytearray()
        while len(data) < n_samples * 2:
            piece = s.recv(min(1 << 20, n_samples * 2 - len(data)))
            if not piece:
                break
            data += piece
        return np.frombuffer(bytes(data), dtype="<i2")


@pytest.fixture(scope="module")
def server(pkg, backend, tmp_path_factory):
    if not os.path.exists(SERVER):
        pytest.fail(f"{SERVER} is not built (python -c 'import __graft_entry__ as g; g.build()')")
    d = tmp_path_factory.mktemp("native_server")
    cfg = pkg.VocoderConfig.tiny(decoder_dim=256, chunk_frames=64, xf_layers=1)
    w = pkg.init_weights(cfg, 0)
    model = str(d / "tiny.b200voc")
    weights_mod = __import__("importlib").import_module("qwen3-tts-axera-russian_b200.weights")
    weights_mod.save_model(model, cfg, w)
    sock = str(d / "voc.sock")
    proc = subprocess.Popen([SERVER, "--model", model, "--socket", sock, "--wave", "8", "--window-us", "3000"],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    t0 = time.time()
    while not os.path.exists(sock):
        assert proc.poll() is None, proc.stdout.read()
        assert time.time() - t0 < 120, "server did not come up"
        time.sleep(0.05)
    voc = backend.Vocoder(cfg, w, wave=8)
    yield cfg, voc, sock, proc
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=20)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, _ = proc.communicate()
    voc.close()
    print(out[-2000:])
    assert "Vocoder Server stopped." in out


def _codes(cfg, n, seed):
    return np.random.default_rng(seed).integers(0, cfg.codebook_size, (n, 16), dtype=np.int64)


def test_socket_is_world_writable(server):
    _, _, sock, _ = server
    assert (os.stat(sock).st_mode & 0o777) == 0o666          # vocoder_server.py:131


@pytest.mark.parametrize("n", [1, 64, 65, 100, 200])
def test_single_request_matches_in_process_backend(server, n):
    cfg, voc, sock, _ = server
    codes = _codes(cfg, n, n)
    got = client_request(sock, codes)
    assert np.array_equal(got, voc.synthesize_pcm16(codes))


def test_concurrent_requests_are_coalesced_and_exact(server):
    """The streaming client's pattern: one connection per 64-token chunk from separate threads
    (dual_npu/tts_client.py:188-197)."""
    cfg, voc, sock, _ = server
    lens = [64] * 12 + [17, 100, 200, 48]
    reqs = [_codes(cfg, n, 1000 + i) for i, n in enumerate(lens)]
    out = [None] * len(reqs)

    def work(i):
        out[i] = client_request(sock, reqs[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(reqs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r, g in zip(reqs, out):
        assert np.array_equal(g, voc.synthesize_pcm16(r))


def test_bad_requests_are_closed_without_reply(server):
    cfg, voc, sock, _ = server
    good = _codes(cfg, 10, 7)
    assert len(client_request(sock, good, raw_header=struct.pack("<i", 0))) == 0          # n <= 0
    assert len(client_request(sock, good, raw_header=struct.pack("<i", 10001))) == 0      # n > 10000
    bad = good.copy()
    bad[3, 5] = cfg.codebook_size                                                        # ORT's Gather would throw
    res = [None, None]

    def work(i, c):
        res[i] = client_request(sock, c)

    th = [threading.Thread(target=work, args=(0, bad)), threading.Thread(target=work, args=(1, good))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(res[0]) == 0
    assert np.array_equal(res[1], voc.synthesize_pcm16(good))      # the healthy neighbour is still served
    # a client that disconnects mid-body must not wedge the server
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.connect(sock)
    s.sendall(struct.pack("<i", 64) + b"\x00" * 100)
    s.close()
    assert np.array_equal(client_request(sock, good), voc.synthesize_pcm16(good))


@pytest.mark.parametrize("n", [40, 100, 200])
def test_server_reply_matches_the_oracle(server, pkg, n):
    """The native server against the CPU restatement itself (not against the same library in-process): the reply to
    a request equals stitch_oracle.synthesize driven by vocoder_oracle's chunk function, converted with the
    reference's truncating PCM16 rule, to within 1 LSB (the float outputs agree to ~2e-5, so a truncation boundary
    can fall between them)."""
    from oracle import stitch_oracle as SO
    from oracle import vocoder_oracle as VO
    cfg, voc, sock, _ = server
    w = pkg.init_weights(cfg, 0)
    codes = _codes(cfg, n, 300 + n)
    got = client_request(sock, codes)
    orc = VO.OracleVocoder(cfg, w)
    ref = SO.to_pcm16(SO.synthesize(codes, orc._inference_chunk, cfg.chunk_frames))
    assert got.shape == ref.shape
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    print(f"server vs oracle n={n}: max diff {int(d.max())} LSB, differing samples {float((d > 0).mean()):.4f}")
    assert int(d.max()) <= 1
