"""bench.py's reference arm (the CPU leg the driver runs beside the CUDA arm) on a tiny index: the JSON contract."""
import json
import subprocess
import sys

import pytest

import common
from common import co

needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


def _run(*args):
    r = subprocess.run([sys.executable, str(common.ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().split("\n")[-1])


@needs_ref
def test_reference_arm_prints_the_contract_line():
    d = _run("--impl", "reference", "--nvec", "3000", "--nq", "48", "--steps", "1", "--warmup", "0")
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_reference_arm_has_no_exhaustive_mode():
    d = _run("--impl", "reference", "--workload", "c4")
    assert d == {"impl": "reference", "unavailable": "the reference has no exhaustive-scan mode (SURVEY F9)"}


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(common.ROOT / "bench.py"), "--nvec", "3000", "--nq", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
