"""bench.py contract on the reference arm (the one arm that runs without a GPU): exactly one JSON line on stdout with the keys the
driver reads; the value is the unmodified reference's makeHeff throughput, MEASURED at the D the line names (never extrapolated)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--D", "250"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    out = json.loads(lines[0])
    for key in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in out, key
    assert out["impl"] == "reference" and out["metric"] == "heff_sigma_builds_per_s" and out["unit"] == "sigma-builds/s"
    assert out["higher_is_better"] is True and out["dtype"] == "f64" and out["vs_baseline"] is None
    assert out["value"] > 0 and abs(out["value"] * out["ms_per_step"] / 1e3 - 1.0) < 1e-6
    assert "workload" in out["config"] and "synth40" in out["config"]["workload"] and "D=250" in out["config"]["workload"]
    assert "D=250" in out["cpu_baseline"]["sample"] and "not extrapolated" in out["cpu_baseline"]["sample"]
    assert out["steps"] >= 1 and out["steps_requested"] == 1
    cb = out["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == out["value"] and "sample" in cb
    e2e = out["e2e"]
    assert e2e["value"] == out["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
