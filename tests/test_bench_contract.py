"""The driver's contract for bench.py, checked without a GPU: the reference arm (the CPU oracle timed through the same
script) prints ONE JSON line with the contract's keys and the SAME metric / unit / config as the B200 arm, whose last
recorded line (profiles/, written on a B200) is checked for the keys the contract names."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _last_json_line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert len(lines) == 1, f"exactly one JSON line expected, got {len(lines)}"
    return json.loads(lines[0])


def _recorded_gpu_line():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_v*_final.json")))      # by name: v13 after v10
    if not paths:
        pytest.skip("no recorded B200 bench line under profiles/")
    return _last_json_line(open(paths[-1]).read())


def test_recorded_b200_line_has_the_contract_keys():
    d = _recorded_gpu_line()
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks"} <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"] == "vocoder_xrt" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0          # host buffers in, host buffers out
    assert e["parity"]["pass"] is True and e["parity"]["snr_db"] >= 60.0 and e["parity"]["max_abs"] <= 1e-4
    assert d["gpu_launches"] > 0 and d.get("simt_launches", 0) == 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])


def test_reference_arm_prints_one_line_with_the_same_config():
    """bench.py --impl reference: the oracle port on the host cores, a bounded sample, the B200 arm's metric and config."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _last_json_line(out.stdout)
    assert BASE_KEYS | {"impl"} <= set(d), sorted(BASE_KEYS - set(d))
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == d["unit"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] == d["value"] and c["sample"]
    g = _recorded_gpu_line()
    for k in ("metric", "unit", "higher_is_better", "scaling"):
        assert d[k] == g[k], k
    assert d["config"] == g["config"]                                       # the driver's same_config check
