"""The native server's wire behaviour on a box without a GPU: the real binary (csrc/voc_server.cpp) runs with a
fake backend LD_PRELOADed in front of libvoc_b200.so (tests/stub/voc_stub.c) and is driven by the reference's
own client half (dual_npu/tts_client.py:78-108) when /root/reference is present, else by its restatement.
What is checked is framing, reply lengths (the reference's window arithmetic incl. the short-last-window
quirk), concurrent connections being answered each with its own audio, and close-without-reply on bad input
(/root/reference/dual_npu/vocoder_server.py:8-12,123-190)."""
import importlib.util
import os
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

from helpers import REFERENCE
from oracle import stitch_oracle as SO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qwen3-tts-axera-russian_b200")
SERVER = os.path.join(PKG, "voc_server")
LC = 122325


def fake_pcm(codes, total):
    seed = np.uint32(2166136261)
    with np.errstate(over="ignore"):
        for c in np.asarray(codes, dtype=np.int64).reshape(-1):
            seed = np.uint32((int(seed) ^ int(c)) & 0xffffffff) * np.uint32(16777619)
        x = (np.arange(total, dtype=np.uint64) + np.uint64(int(seed))).astype(np.uint32)
        x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d)
        x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b)
        x ^= x >> np.uint32(16)
    return (x & np.uint32(0xffff)).astype(np.uint16).view(np.int16)


def restated_client(sock_path, codes):
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(30)
    try:
        s.connect(sock_path)
        s.sendall(struct.pack("<i", len(codes)))
        s.sendall(np.ascontiguousarray(codes, dtype="<i8").tobytes())
        hdr = b""
        while len(hdr) < 4:
            p = s.recv(4 - len(hdr))
            if not p:
                return np.zeros(0, dtype=np.int16)
            hdr += p
        (n,) = struct.unpack("<i", hdr)
        data = bytearray()
        while len(data) < 2 * n:
            p = s.recv(min(1 << 20, 2 * n - len(data)))
            if not p:
                break
            data += p
        return np.frombuffer(bytes(data), dtype="<i2")
    except (ConnectionResetError, BrokenPipeError):
        return np.zeros(0, dtype=np.int16)
    finally:
        s.close()


def reference_client(sock_path):
    """tts_client.Qwen3TTSClient._vocoder_chunk bound to `sock_path`, or None when the tree is absent."""
    path = os.path.join(REFERENCE, "dual_npu", "tts_client.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ref_tts_client", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cli = mod.Qwen3TTSClient.__new__(mod.Qwen3TTSClient)
    cli.voc_socket = sock_path

    def call(codes):
        res = {}
        cli._vocoder_chunk(np.asarray(codes, dtype=np.int64).tolist(), 0, res)       # tts_client.py:78-108
        return np.asarray(res[0])
    return call


@pytest.fixture(scope="module")
def server(tmp_path_factory, backend):
    if not os.path.exists(SERVER):
        pytest.fail("voc_server is not built: run build()")
    d = tmp_path_factory.mktemp("native_proto")
    stub = str(d / "libvoc_stub.so")
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "stub", "voc_stub.c"), "-o", stub,
                    "-L" + PKG, "-lvoc_b200", "-Wl,-rpath," + PKG], check=True)
    model = d / "fake.b200voc"
    model.write_bytes(b"stub")
    sock = str(d / "v.sock")
    env = dict(os.environ, LD_PRELOAD=stub)
    proc = subprocess.Popen([SERVER, "--model", str(model), "--socket", sock, "--window-us", "5000",
                             "--recv-timeout-ms", "700", "--send-timeout-ms", "1500", "--max-conns", "24"], env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    t0 = time.time()
    while not os.path.exists(sock):
        assert proc.poll() is None, proc.stdout.read()
        assert time.time() - t0 < 30
        time.sleep(0.02)
    yield sock
    proc.terminate()
    out, _ = proc.communicate(timeout=20)
    assert "Vocoder Server stopped." in out, out[-1000:]


def _codes(n, seed):
    return np.random.default_rng(seed).integers(0, 2048, (n, 16), dtype=np.int64)


@pytest.mark.parametrize("n", [1, 64, 65, 97, 112, 200])
def test_reply_framing_and_lengths(server, n):
    codes = _codes(n, n)
    got = restated_client(server, codes)
    want_len = len(SO.synthesize(codes, lambda p: np.zeros(LC, np.float32), 64))     # the reference's arithmetic
    assert len(got) == want_len
    assert np.array_equal(got, fake_pcm(codes, want_len))


def test_reference_client_is_served(server):
    ref = reference_client(server)
    if ref is None:
        pytest.skip("/root/reference not present on this box")
    codes = _codes(100, 5)
    got = ref(codes)
    want_len = len(SO.synthesize(codes, lambda p: np.zeros(LC, np.float32), 64))
    assert len(got) == want_len == 199125          # 122325 + (99840 - 30720) + 7680: the short last window is appended
    assert np.array_equal(got, fake_pcm(codes, want_len))


def test_concurrent_connections_each_get_their_own_audio(server):
    reqs = [_codes(n, 50 + i) for i, n in enumerate([64] * 10 + [3, 130, 200, 48, 64, 64])]
    out = [None] * len(reqs)
    th = [threading.Thread(target=lambda i=i: out.__setitem__(i, restated_client(server, reqs[i]))) for i in range(len(reqs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r, g in zip(reqs, out):
        assert np.array_equal(g, fake_pcm(r, len(g))) and len(g) > 0


def test_bad_input_is_closed_without_reply_and_neighbours_survive(server):
    good, bad = _codes(20, 1), _codes(20, 2)
    bad[7, 3] = 2048
    res = [None, None]
    th = [threading.Thread(target=lambda: res.__setitem__(0, restated_client(server, bad))),
          threading.Thread(target=lambda: res.__setitem__(1, restated_client(server, good)))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(res[0]) == 0
    assert np.array_equal(res[1], fake_pcm(good, len(res[1]))) and len(res[1]) == 20 * 1920
    for hdr in (0, -5, 10001):
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(10)
        s.connect(server)
        s.sendall(struct.pack("<i", hdr))
        try:
            assert s.recv(4) == b""
        except ConnectionResetError:
            pass
        s.close()
    assert len(restated_client(server, good)) == 20 * 1920


def test_a_stalled_sender_is_dropped_and_a_deaf_receiver_blocks_nobody(server):
    """Hardening beyond the reference (which serialises connections): a peer that announces a body and never sends
    it is dropped at the receive deadline; a peer that never reads its reply does not delay other requests."""
    stall = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    stall.settimeout(10)
    stall.connect(server)
    stall.sendall(struct.pack("<i", 10000) + b"\x00" * 64)           # 1.28 MB announced, 64 bytes sent
    deaf = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    deaf.connect(server)
    big = _codes(2000, 77)                                            # reply of ~7.7 MB, never read
    deaf.sendall(struct.pack("<i", len(big)) + big.astype("<i8").tobytes())
    t0 = time.time()
    good = _codes(20, 3)
    assert len(restated_client(server, good)) == 20 * 1920           # served while both others are pending
    assert time.time() - t0 < 5.0
    try:
        assert stall.recv(4) == b""                                   # closed without a reply at the deadline
    except ConnectionResetError:
        pass
    assert time.time() - t0 < 8.0
    stall.close()
    deaf.close()
    assert len(restated_client(server, good)) == 20 * 1920


def test_connections_beyond_the_cap_are_refused(server):
    held = []
    try:
        for _ in range(24):
            s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            s.settimeout(10)
            s.connect(server)
            held.append(s)
        time.sleep(0.2)
        extra = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        extra.settimeout(10)
        extra.connect(server)
        try:
            extra.sendall(struct.pack("<i", 1))
            assert extra.recv(4) == b""                               # accepted and closed at once
        except (ConnectionResetError, BrokenPipeError):
            pass
        extra.close()
    finally:
        for s in held:
            s.close()
    time.sleep(0.2)
    good = _codes(20, 3)
    assert len(restated_client(server, good)) == 20 * 1920
