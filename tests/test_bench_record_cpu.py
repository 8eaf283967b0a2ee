"""The JSON line bench.py prints is a contract with the driver: the last validated records committed under profiles/
must carry every key the contract names (own arm and reference arm), with consistent values."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline")


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        pytest.skip("no committed record matches %s" % pattern)
    return json.load(open(files[-1]))


def test_own_arm_record_has_the_contract_keys():
    d = _latest("r02_bench_final[0-9]*_n1.json")
    for k in BASE + ("roofline", "clocks"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["data"] == "synthetic" and d["dtype"] == "bf16" and d["vs_baseline"] is None     # BASELINE.md publishes no B200 number
    assert "workload" in d["config"] and "model" not in d["config"]
    # images per step / step time = value
    assert d["value"] == pytest.approx(d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3), rel=1e-3)
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-3)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"] and e["value"] <= d["value"] * 1.02      # measured through host buffers, not a copy
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] > 0
    k = d["clocks"]
    assert k["sm_mhz"] <= k["sm_max_mhz"] and isinstance(k["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"])


def test_reference_arm_record():
    d = _latest("r02_bench_reference_arm_final[0-9]*.json")
    own = _latest("r02_bench_final[0-9]*_n1.json")
    for k in BASE + ("impl",):
        assert k in d, k
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == own[k]
    assert d["config"]["workload"] == own["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")


def test_metric_is_the_baseline_metric():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    d = _latest("r02_bench_final[0-9]*_n1.json")
    assert d["metric"].split(" (")[0] in base["metric"]
