"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`: the CPU restatement on the host's
cores, a bounded sample of the GPU arm's workload) prints one JSON line with the agreed keys, and the product arm refuses to run --
loudly, with no CPU fallback -- when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    cfg = line["config"]
    assert (cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"]) == (1200, 800, 500, 50), "the GPU arm's workload (BASELINE configs[1])"
    assert "timed_sample" in cfg and "bounded sample" in cfg["workload"], "the label says what is really rendered per step"
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    r = run_bench("--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-c4", timeout=120)
    assert r.returncode != 0, "no CPU fallback: the product arm must not produce a number without a CUDA device"
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]
