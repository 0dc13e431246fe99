"""CPU: bench.py's contract on a box without a GPU — the reference arm prints the JSON line the driver parses (same metric /
unit / config keys as the product arm, `impl: reference`, a cpu_baseline describing the run), and the product arm refuses to
run without a CUDA device instead of measuring anything on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT,
                          timeout=600)


def test_reference_arm_prints_the_contract_line():
    out = _bench("--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("rays/sec") and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["gpu_launches"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_refuses_to_run_without_a_gpu():
    out = _bench("--workload", "cfg1", "--steps", "1", "--no-cpu-baseline")
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr + out.stdout
    assert not any(l.startswith("{") for l in out.stdout.splitlines())        # no JSON line = nothing was measured
