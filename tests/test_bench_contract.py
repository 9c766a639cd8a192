"""CPU checks of bench.py: the size/byte helpers against SURVEY App. C / 8(d), the JSON line of
the `--impl reference` arm (the CPU restatement on a tiny sample), rank > 0 staying silent, and
the loud failure of the product arm without a GPU."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_size_and_byte_helpers():
    b = _bench_module()
    # (p, s) -> DoFs, cells
    assert b.n_dofs_of(3, 15) == (2738019, 32768)
    assert b.n_dofs_of(4, 18) == (50923779, 262144)
    assert b.n_dofs_of(6, 18) == (171199875, 262144)
    assert b.n_dofs_of(4, 22) == (809244675, 1 << 22)
    # merged Q4 60.2 B/DoF/it, plain 138.67 B/DoF + metadata
    nd, nc = b.n_dofs_of(4, 18)
    assert abs(b.algorithmic_bytes_per_iteration(4, 18) / nd - 60.2) < 0.05
    assert abs(b.algorithmic_bytes_per_iteration(4, 18, merged=False) / nd - (138.667 + 300.0 * nc / nd)) < 1e-2


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=env, timeout=600, cwd=ROOT)


def test_reference_arm_json_line(c_oracle_lib):
    r = _run(None, "--impl", "reference", "--steps", "1", "--warmup", "1", "--degree", "3", "--s", "6", "--cpu-s", "6")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDoF/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "s=6" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
             "--warmup", "1", "--degree", "3", "--s", "6", "--cpu-s", "6")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run({"CUDA_VISIBLE_DEVICES": ""}, "--steps", "1", "--warmup", "1", "--degree", "3", "--s", "6",
             "--no-cpu-baseline")
    assert r.returncode != 0          # no CPU fallback, no JSON line
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
