"""bench.py's reference arm (the CPU leg the driver runs next to the B200 arm) on the host cores: one JSON line with the
contract's keys, rank 0 only under a multi-rank launch.  CPU only; the B200 arm itself needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env_extra=None):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                          env=env, cwd=ROOT)


@pytest.mark.parametrize("workload", ["lj", "adress", "tetramer"])
def test_reference_arm_line(workload):
    res = run_bench(["--impl", "reference", "--workload", workload, "--steps", "3", "--warmup", "1", "--side", "16"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "atom-steps/s" and d["unit"] == "atom-steps/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and isinstance(d["config"]["workload"], str)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    res = run_bench(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--side", "16"],
                    {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = run_bench(["--steps", "2", "--warmup", "1", "--side", "16", "--no-cpu-baseline", "--no-e2e"])
    assert res.returncode != 0
