"""The C-ABI library loads and exports exactly what include/voc_b200.h declares; host-only
entry points behave; compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "voc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(voc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(backend):
    lib = ctypes.CDLL(backend.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/voc_b200.h but not exported"
    assert sorted(backend.SIGNATURES) == names, "backend.SIGNATURES and the header disagree"
    assert lib.voc_abi_version() == 1


def test_no_torch_types_in_the_boundary():
    src = open(os.path.join(ROOT, "include", "voc_b200.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "at::" not in src and "c10::" not in src


def test_plan_rejects_bad_arguments(backend):
    lib = backend.load_library()
    assert lib.voc_plan(64, 122880, 0, 0, None, None, None) == backend.VOC_E_INVALID
    assert lib.voc_plan(8, 122880, 10, 0, None, None, None) == backend.VOC_E_INVALID
    small = np.zeros(6, dtype=np.int32)
    assert lib.voc_plan(64, 122880, 200, small.size, small.ctypes.data, None, None) == backend.VOC_E_INVALID


def test_create_fails_loudly_without_gpu(backend, pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(backend.VocoderError) as e:
        backend.Vocoder(pkg.VocoderConfig.tiny())
    assert e.value.code == backend.VOC_E_CUDA
    assert "no CPU fallback" in str(e.value)


def test_bad_config_is_rejected(backend):
    lib = backend.load_library()
    assert not lib.voc_create(b'{"transconv_trim": "left"}', 0, 1)
    assert b"transconv_trim" in lib.voc_last_error(None)
    assert not lib.voc_create(b'{"num_quantizers": 8}', 0, 1)
    assert not lib.voc_create(b'not json', 0, 1)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "qwen3-tts-axera-russian_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "stitch_oracle" not in txt and "vocoder_oracle" not in txt, f


def test_create_from_file_fails_loudly(backend, pkg, tmp_path):
    """voc_create_from_file: a missing / malformed container is an error with a message, and a valid one
    still needs a GPU (no CPU fallback)."""
    lib = backend.load_library()
    assert not lib.voc_create_from_file(str(tmp_path / "nope.b200voc").encode(), 0, 1)
    assert b"cannot open" in lib.voc_last_error(None)
    junk = tmp_path / "junk.b200voc"
    junk.write_bytes(b"\x10\x00\x00\x00\x00\x00\x00\x00{not json at all")
    assert not lib.voc_create_from_file(str(junk).encode(), 0, 1)
    import importlib
    W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
    cfg = pkg.VocoderConfig.tiny()
    good = tmp_path / "tiny.b200voc"
    W.save_model(str(good), cfg, pkg.init_weights(cfg, 0))
    import torch
    if not torch.cuda.is_available():
        assert not lib.voc_create_from_file(str(good).encode(), 0, 1)
        assert b"no usable CUDA device" in lib.voc_last_error(None)


def test_native_server_is_built_and_needs_a_gpu(tmp_path, pkg, backend):
    """csrc/voc_server.cpp is compiled by build(); without a B200 it refuses to start (no CPU fallback)."""
    import importlib, os, subprocess, torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "qwen3-tts-axera-russian_b200", "voc_server")
    assert os.path.exists(exe), "run build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=30)
    assert r.returncode == 2 and "--model is required" in r.stderr
    if not torch.cuda.is_available():
        W = importlib.import_module("qwen3-tts-axera-russian_b200.weights")
        cfg = pkg.VocoderConfig.tiny()
        m = tmp_path / "tiny.b200voc"
        W.save_model(str(m), cfg, pkg.init_weights(cfg, 0))
        r = subprocess.run([exe, "--model", str(m), "--socket", str(tmp_path / "s.sock")], capture_output=True, text=True,
                           timeout=60)
        assert r.returncode == 1 and "no usable CUDA device" in (r.stderr + r.stdout)


# ---- the tcgen05 kernel's tile plan (host arithmetic: runs without a GPU) -------------------------------------
# layer table of the production decoder: (name, N, K, taps, M per window)
_LAYERS = [
    ("xf.qkv", 3072, 512, 1, 64), ("xf.o", 512, 1024, 1, 64), ("xf.gate_up", 2048, 512, 1, 64), ("xf.down", 512, 1024, 1, 64),
    ("up.pw2", 1024, 4096, 1, 256), ("conv_in", 1536, 1024, 7, 256),
    ("dec0.convt", 6144, 1536, 2, 256), ("dec0.conv7", 768, 768, 7, 2048), ("dec0.conv1", 768, 768, 1, 2048),
    ("dec1.convt", 1920, 768, 2, 2048), ("dec1.conv7", 384, 384, 7, 10240), ("dec1.conv1", 384, 384, 1, 10240),
    ("dec2.convt", 768, 384, 2, 10240), ("dec3.convt", 288, 192, 2, 40960), ("dec3.conv7", 96, 96, 7, 122880),
]


def test_tile_plan_form_and_pairing_never_depend_on_the_batch(backend):
    """The MMA form (3-pass / concatenated) and cta_group::2 pairing fix the rounding of a layer, so the batch must not
    change them; only the column tile inside the form's family may follow it (tc_gemm.cu: voc_tc_plan_tile)."""
    for name, N, K, taps, M in _LAYERS:
        ref = backend.tc_plan(N, K, taps, M, 1)
        for B in (1, 2, 3, 4, 6, 8, 16, 20, 32, 157, 256):
            for flags in (0, 256, 512):
                t = backend.tc_plan(N, K, taps, M, B, tc_flags=flags)
                assert (t["three_pass"], t["pair"], t["BK"]) == (ref["three_pass"], ref["pair"], ref["BK"]), (name, B, flags, t, ref)
                family = (192, 96) if t["three_pass"] else (128, 64, 32) if N % 128 == 0 else (96,) if N % 96 == 0 else (64, 32)
                assert t["BN"] in family and N % t["BN"] == 0, (name, B, flags, t)
                assert t["p3"] == (t["three_pass"] and t["BN"] == 96), (name, B, flags, t)


def test_tile_plan_widest_tile_for_large_batches_narrower_for_one_window(backend):
    widest = {name: backend.tc_plan(N, K, taps, M, 1, tc_flags=256)["BN"] for name, N, K, taps, M in _LAYERS}
    for name, N, K, taps, M in _LAYERS:
        assert backend.tc_plan(N, K, taps, M, 256)["BN"] == widest[name], name       # the 256-window step is unchanged
        assert backend.tc_plan(N, K, taps, M, 1, tc_flags=512)["BN"] <= widest[name]
    one = {name: backend.tc_plan(N, K, taps, M, 1) for name, N, K, taps, M in _LAYERS}
    # one window: 64 CTAs at C = 768 and 160 tiles on 148 SMs at C = 384 become 96-column 3-pass tiles, the
    # transformer's 512-column projections (4 CTAs with a 64-k-step chain each) 32-column concatenated tiles
    assert one["dec0.conv7"] == {"BN": 96, "BK": 64, "pair": True, "p3": True, "three_pass": True}
    assert one["dec1.conv7"]["BN"] == 96 and one["dec1.conv7"]["pair"]
    assert one["conv_in"]["BN"] == 96 and one["xf.qkv"]["BN"] == 96 and not one["xf.qkv"]["pair"]
    assert one["xf.o"]["BN"] == 32 and one["xf.down"]["BN"] == 32 and not one["xf.o"]["three_pass"]
    # layers whose family has one member keep it
    assert one["dec3.conv7"] == {"BN": 96, "BK": 64, "pair": True, "p3": False, "three_pass": False}
    assert one["dec3.convt"]["BN"] == 96 and not one["dec3.convt"]["p3"]
    with pytest.raises(ValueError):
        backend.tc_plan(100, 64, 1, 128, 1)                                            # N not a multiple of 32


def test_fused_unit_shared_memory_plan(backend):
    """ru_fused.cu's plan at the production shapes: everything fits the 227 KB opt-in, the halo tile is one TMA box, C = 96
    keeps T beside a double-buffered halo ring with >= 4 weight stages (pipelined order of work), C = 192 must overlay T
    on the ring, and a dilation whose halo tile exceeds one box is refused."""
    for Cc in (96, 192):
        for dil in (1, 3, 9):
            p = backend.ru_plan(Cc, 7, dil)
            assert p["smem"] <= 232448 - 5120 and p["halo_stages"] == 2 and p["weight_stages"] >= 2, (Cc, dil, p)
            assert p["box_rows"] >= 128 + 6 * dil and p["box_rows"] % 8 == 0 and p["box_rows"] <= 256
            assert p["alias"] == (Cc == 192), (Cc, dil, p)
            if Cc == 96:
                assert p["weight_stages"] >= 4, (dil, p)
    with pytest.raises(ValueError):
        backend.ru_plan(96, 7, 27)                      # 128 + 162 rows: more than one 256-row box
    with pytest.raises(ValueError):
        backend.ru_plan(384, 7, 1)                      # a tile's T rows would span two column tiles (blocks 0-1)
