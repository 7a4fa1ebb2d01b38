"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port of the path) prints ONE JSON line
with the keys the driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          timeout=900, env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "PQ-quantized pixels/sec" and d["unit"] == "pixels/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the sample is named and flagged as a sample of the headline configuration, not the configuration itself
    assert d["config"]["same_config"] is False and "images per step" in d["config"]["sampling"]
    for k in ("B", "D", "h", "w", "H", "W", "M", "K", "C"):
        assert k in d["config"]
    # pixels of the sample / step time = value
    px = 4 * d["config"]["h"] * d["config"]["w"]
    assert abs(px / (d["ms_per_step"] * 1e-3) - d["value"]) < 1e-6 * d["value"]


def test_reference_arm_other_ranks_do_no_work():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-extras")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
