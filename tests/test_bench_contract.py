"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
(`--impl reference`: the CPU restatement timed on the host cores) prints exactly one JSON line on
stdout with the keys the contract names; ranks other than 0 of a multi-rank launch exit 0 without
work or output; and the product arm fails loudly, not with a CPU fallback, when no device is visible."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run_bench("--impl", "reference", "--config", "c2", "--scale", "0.002", "--steps", "3", "--warmup", "3", "--cpu-budget", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                    # ONE line on stdout, everything else on stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pdhg_iterations_per_sec" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["value"] > 0 and d["ms_per_step"] == pytest.approx(1000.0 / d["value"])
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "c2@scale=0.002" and d["config"]["rows"] == 2000 and d["config"]["cols"] == 4000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "PDHG iterations" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = run_bench("--impl", "reference", "--gpus", "2", "--scale", "0.002", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    from ortools_b200 import pdlp
    if pdlp.backend().device_count() > 0:
        pytest.skip("a CUDA device is visible: the product arm would run")
    r = run_bench("--scale", "0.002", "--steps", "3", "--warmup", "3", "--no-e2e", "--no-cpu")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr or "no CPU fallback" in r.stderr


def test_smoke_fails_loudly_without_a_device():
    """__graft_entry__.smoke() runs the hot path on cuda:0; without a device it must raise, not fall back."""
    from ortools_b200 import pdlp
    if pdlp.backend().device_count() > 0:
        pytest.skip("a CUDA device is visible")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    with pytest.raises(RuntimeError, match="no CUDA device"):
        entry.smoke()
