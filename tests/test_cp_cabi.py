"""include/cp_b200.h <-> the library <-> the ctypes binding, and the host-side pieces of the code-predictor mirror
(no GPU needed): exported symbols, config / weight inventory agreement with the oracle and with the reference's export
script, loud failure without a GPU."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest

from helpers import REFERENCE
from oracle import code_predictor_oracle as CPO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpm(backend):
    return importlib.import_module("qwen3-tts-axera-russian_b200.code_predictor")


def test_every_declared_symbol_is_exported(backend, cpm):
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "cp_b200.h")).read(), flags=re.S)
    assert "torch" not in src and "at::" not in src
    names = sorted(set(re.findall(r"\b(cp_[a-z0-9_]+)\s*\(", src)))
    lib = ctypes.CDLL(backend.LIB_PATH)
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cp_b200.h but not exported"
    assert sorted(cpm.SIGNATURES) == names


def test_config_and_inventory_agree_with_the_oracle(cpm):
    for a, b in ((cpm.CPConfig(), CPO.CPConfig()), (cpm.CPConfig.tiny(), CPO.CPConfig.tiny())):
        assert a.to_json() == b.to_json()
        assert cpm.weight_shapes(a) == CPO.weight_shapes(b)
    wa, wb = cpm.init_weights(cpm.CPConfig.tiny(), 3), CPO.init_weights(CPO.CPConfig.tiny(), 3)
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    assert cpm.CPConfig.from_weights(cpm.init_weights(cpm.CPConfig(layers=1, groups=2, vocab=64))) == \
        cpm.CPConfig(layers=1, groups=2, vocab=64)


def test_inventory_matches_the_reference_export_script(cpm, have_reference):
    """Every array name scripts/export_code_predictor_weights.py writes is one this backend expects, and vice versa."""
    if not have_reference:
        pytest.skip("reference tree not present")
    src = open(os.path.join(REFERENCE, "scripts", "export_code_predictor_weights.py")).read()
    keys = set(re.findall(r'np_weights\[f?"([^"]+)"\] =', src))
    norm = {re.sub(r"\{[^}]+\}", "N", k) for k in keys}
    mine = {re.sub(r"\d+", "N", k) for k in cpm.weight_shapes(cpm.CPConfig())}
    assert mine <= norm, mine - norm
    assert norm == mine, norm ^ mine


def test_create_fails_loudly_without_gpu(cpm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cpm.CodePredictorError) as e:
        cpm.CodePredictor(cpm.CPConfig.tiny())
    assert e.value.code == cpm.CP_E_CUDA and "no CPU fallback" in str(e.value)


def test_bad_config_is_rejected(cpm):
    lib = cpm.load_library()
    assert not lib.cp_create(b'{"hidden": 1022}', 0)
    assert b"multiples of 4" in lib.cp_last_error(None)
    assert not lib.cp_create(b'{"heads": 16, "kv_heads": 5}', 0)
    assert not lib.cp_create(b'{"max_positions": 8}', 0)
    assert lib.cp_step(None, None, 1, 0, None) == cpm.CP_E_INVALID
    assert lib.cp_predict(None, None, None, 0.1, 50, 0, None) == cpm.CP_E_INVALID


class _OracleBackedCP:
    """Stand-in for code_predictor.CodePredictor on a machine without a GPU: the same methods on the CPU oracle, so that
    the server mirror's host logic (framing, prefill variants, host sampler, error handling) is exercised here."""

    def __init__(self, cfg, w):
        import torch
        self.cfg, self.W, self.kv, self.out, self.torch = cfg, CPO.Weights(w), None, None, torch
        self.cache_len = 0

    def reset(self):
        self.kv, self.cache_len = None, 0

    def step(self, hidden, position):
        x = self.torch.from_numpy(np.asarray(hidden, dtype=np.float32).reshape(-1, self.cfg.hidden))
        assert position == self.cache_len
        with self.torch.no_grad():
            out, self.kv = CPO.step(x, list(range(position, position + x.shape[0])), self.kv, self.W, self.cfg)
        self.cache_len += x.shape[0]
        self.out = out
        return out.numpy()

    def logits(self, g):
        return (self.out[-1] @ self.W[f"lm_head_{g}"].T).numpy()


@pytest.mark.parametrize("batch_prefill", [False, True])
def test_server_mirror_host_logic_equals_the_reference_executing(cpm, have_reference, batch_prefill, tmp_path):
    """code_predictor_server.CodePredictorServer (host sampler, level 1) against the reference's own class executing with
    its ONNX session replaced by the oracle's step: same seeded NumPy generator -> identical 15 codes, for the sequential
    and the 2-token prefill; then the wire protocol end to end over a unix socket."""
    if not have_reference:
        pytest.skip("reference tree not present")
    import importlib.util
    import socket
    import struct
    import threading
    import time
    import torch
    srv_mod = importlib.import_module("qwen3-tts-axera-russian_b200.code_predictor_server")
    cfg = CPO.CPConfig(layers=2, inter=256, vocab=128, head_dim=32)
    w = CPO.init_weights(cfg, 5)
    table = np.random.default_rng(1).standard_normal((32, cfg.hidden)).astype(np.float32)
    hidden = np.random.default_rng(2).standard_normal(cfg.hidden).astype(np.float32)

    # the reference class, constructed without its __init__ (which needs onnxruntime and files)
    spec = importlib.util.spec_from_file_location("ref_cp_server", os.path.join(REFERENCE, "dual_npu", "code_predictor_server.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    ref = ref_mod.CodePredictorServer.__new__(ref_mod.CodePredictorServer)
    ref.temperature, ref.top_k, ref.num_groups, ref.batch_prefill = 0.8, 20, cfg.groups, batch_prefill
    ref.codec_embeddings = [w[f"codec_emb_{i}"] for i in range(cfg.groups)]
    ref.lm_heads = [w[f"lm_head_{i}"] for i in range(cfg.groups)]
    ref.codec_embedding = table
    ref.num_layers, ref.head_dim, ref.num_kv_heads = cfg.layers, cfg.head_dim, cfg.kv_heads
    W = CPO.Weights(w)

    def run(feed):                                               # stands in for sess.run(None, feed)
        x = torch.from_numpy(feed["hidden"][0])
        pos = [int(p) for p in feed["position"]]
        kv = None
        if feed["past_k_0"].shape[2] > 0:
            kv = [(torch.from_numpy(feed[f"past_k_{i}"][0]), torch.from_numpy(feed[f"past_v_{i}"][0])) for i in range(cfg.layers)]
        with torch.no_grad():
            out, nkv = CPO.step(x, pos, kv, W, cfg)
        res = [out.numpy()[None]]
        for k, v in nkv:
            res += [k.numpy()[None], v.numpy()[None]]
        return res
    ref.sess = type("S", (), {"run": staticmethod(lambda _, feed: run(feed))})()
    np.random.seed(77)
    want = ref.predict(hidden, 9)

    srv = srv_mod.CodePredictorServer.__new__(srv_mod.CodePredictorServer)
    srv.socket_path = str(tmp_path / "cp.sock")
    srv.temperature, srv.top_k, srv.batch_prefill, srv.sampler = 0.8, 20, batch_prefill, "host"
    srv.codec_embedding, srv.cp = table, _OracleBackedCP(cfg, w)
    srv.num_groups, srv.num_layers, srv.head_dim, srv.num_kv_heads, srv.hidden_size = cfg.groups, cfg.layers, cfg.head_dim, cfg.kv_heads, cfg.hidden
    srv.codec_embeddings = ref.codec_embeddings
    srv._frame, srv._running = 0, True
    np.random.seed(77)
    assert srv.predict(hidden, 9) == want

    th = threading.Thread(target=srv.serve, daemon=True)
    th.start()
    for _ in range(100):
        if os.path.exists(srv.socket_path):
            break
        time.sleep(0.05)
    np.random.seed(77)
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.connect(srv.socket_path)
    c.sendall(hidden.tobytes() + struct.pack("<i", 9))
    data = b""
    while len(data) < 60:
        piece = c.recv(60 - len(data))
        if not piece:
            break
        data += piece
    c.close()
    assert list(np.frombuffer(data, dtype="<i4")) == want
    # code_0 outside the embedding table: no reply, the server keeps serving
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.connect(srv.socket_path)
    c.sendall(hidden.tobytes() + struct.pack("<i", 10 ** 6))
    assert c.recv(4) == b""
    c.close()
    srv._running = False
    th.join(timeout=5)
    assert not os.path.exists(srv.socket_path)
