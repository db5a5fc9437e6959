"""bench.py's reference arm (`--impl reference`) runs without a GPU: check that it honours the driver's contract — one JSON
line on stdout with every required key, the tier's `cpu_baseline` / `e2e` objects, and the metric / config that the GPU arm
reports (both arms are built from the same constants, so a drift in one shows here)."""
import ast
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                                   # exactly one JSON line
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "set_ops_per_sec" and d["unit"] == "set-ops/s" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(cb) and cb["kind"] in ("port", "reference") and cb["cores"] >= 1
    assert cb["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # 1024 shards x 63 set-ops per step
    assert abs(d["value"] - 1024 * 63 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_gpu_arm_reports_the_contract_keys():
    """the GPU arm cannot run here; its JSON line is assembled from one dict literal — check the literal's keys"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    keys = set()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Dict):
            ks = {k.value for k in node.keys if isinstance(k, ast.Constant) and isinstance(k.value, str)}
            if {"metric", "roofline", "gpu_launches"} <= ks:
                keys = ks
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in keys, key
    assert 'line["cpu_baseline"] = cpu_rec' in src                 # added on rank 0 at N=1 unless --no-cpu-baseline
    for key in ('"parity_ok"', '"north_star"', '"config4"', '"config3"'):   # driver-visible parity + the sub-records of the other BASELINE configs
        assert key in src, key
