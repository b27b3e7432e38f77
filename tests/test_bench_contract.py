"""bench.py contract (task statement, section 4): the committed bench line of the measured arm carries every required key with a
sane value, and the reference arm (CPU) prints one JSON line with the same metric / unit and the e2e / cpu_baseline objects."""
import json
import os
import subprocess
import sys

import glob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest_bench_line():
    """the newest committed 1-GPU bench line (profiles/rNN_bench_fast_mode.json)"""
    return json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_fast_mode.json")))[-1]))


def _check_common(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["unit"] == d["unit"]
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("port", "reference")


def test_committed_bench_line_has_the_contract_keys():
    d = _latest_bench_line()
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "bf16" and d["gpu_launches"] > 0
    assert abs(d["value"] - 1024 * 1024 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 1024 * 256 * 3 * 4 and d["e2e"]["d2h_bytes_per_step"] == 1024 * 1024 * 4
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    assert "l2" in d["config"]                      # says how L2 reuse between timed iterations is ruled out


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["dtype"] == "f32"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    ours = _latest_bench_line()
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
