"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the agreed keys (rank 0
only under torchrun), and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line():
    r = run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--ref-size", "4", "--cpu-direct-size", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("elements/sec per Newton iteration") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "HEX20 4^3" in cb["sample"]
    assert "Jacobi-PCG" in cb["sample"] and "4x4x4" in d["config"]["workload"]          # same-algorithm CPU leg, sample named
    assert cb["direct"]["value"] > 0 and "SuperLU" in cb["direct"]["sample"]              # the reference's own direct solve beside it
    assert not any("libamaru_b200" in x for x in d["repo_so_mapped"])                     # the CPU arm never maps the product library
    assert d["e2e"] == {"value": d["value"], "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-size", "4"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is visible")
    r = run(["--steps", "1", "--warmup", "3", "--size", "4"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
