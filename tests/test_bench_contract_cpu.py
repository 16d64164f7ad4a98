"""The driver-facing bench.py contract, as far as it can be checked without a GPU: the
reference arm (`--impl reference`, the CPU restatement of the reference's graph timed on the
host cores) prints one JSON line with the agreed keys, and the product arm refuses to run
without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    p = run("--impl", "reference", "--workload", "cfg1", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "samples/s" and line["higher_is_better"] is True
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["value"] == line["config"]["global_batch"] / (line["ms_per_step"] * 1e-3) or \
        abs(line["value"] - line["config"]["global_batch"] / (line["ms_per_step"] * 1e-3)) < 1e-6 * line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm would run")
    p = run("--steps", "2", "--warmup", "1")
    out = p.stdout.strip().splitlines()
    line = json.loads(out[-1]) if out and out[-1].startswith("{") else None
    # no measurement may come out of a box without a CUDA device
    assert line is None or line.get("gpu_launches", 0) == 0 and not line.get("value")
    assert "no CUDA device" in (p.stderr + p.stdout)
