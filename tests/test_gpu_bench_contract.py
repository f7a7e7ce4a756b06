"""bench.py prints ONE JSON line with the keys the measurement contract names (own arm and --impl reference)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_own_arm_line():
    d = _run("--gpus", "1", "--steps", "3", "--warmup", "3", "--config4-utts", "96")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "audio-s/s" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"]
    own = _run("--gpus", "1", "--steps", "3", "--warmup", "3", "--no-config4", "--no-cpu-baseline")
    assert own["config"] == d["config"]                      # the driver compares the arms' config dicts and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] - 64 * 5.0 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= 1.02 * d["value"]
    assert e["serial"]["value"] > 0 and e["cabi"]["value"] > 0
    assert "synth_stream" in e["call"]                      # the headline e2e goes through the Python API, NumPy to NumPy
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and (r["traffic"] is None or r["traffic"] > 0)
    assert d["gpu_launches"] >= 3 * 20
    c4 = d["config4"]                                        # BASELINE.json configs[3] in miniature (96 of 8192 utterances)
    assert c4["scaling"] == "strong" and c4["utterances"] == 96 and c4["value"] > 0 and c4["lpt_imbalance"] >= 1.0
    assert len(c4["per_rank"]) == 1 and c4["per_rank"][0]["range_reruns"] == 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c


def test_reference_arm_line():
    d = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"].startswith("audio-sec") and d["unit"] == "audio-s/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert "workload" in d["config"]
    own = _run("--gpus", "1", "--steps", "3", "--warmup", "3", "--no-config4", "--no-cpu-baseline")
    assert own["config"] == d["config"]                      # the driver compares the arms' config dicts
