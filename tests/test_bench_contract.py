"""bench.py's contract, the parts that need no GPU: the reference arm prints ONE JSON line with the
agreed keys (rank 0 only), and the GPU arm refuses to run -- loudly, no CPU fallback -- without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("contact evals/sec") and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "configs[4]" in d["config"]["workload"] and d["config"]["evals_per_step"] == 1 << 28
    assert d["scaling"] == "strong"
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.shared_config(1)      # the SAME config dict in both arms
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert "819200" in cb["sample"]
    if cb["kind"] == "reference":      # oracle/_ref was built: the C port's number stands beside it
        assert cb["port_value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.parametrize("workload", ["headline", "config2"])
def test_gpu_arm_fails_loudly_without_a_device(workload):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _bench("--workload", workload, "--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert r.stdout.strip() == ""                      # no JSON line, nothing that could be mistaken for a result
    assert "no CPU" in (r.stderr + r.stdout) or "CUDA" in r.stderr
