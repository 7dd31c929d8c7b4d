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
