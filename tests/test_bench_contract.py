"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the compiled reference on
the host and prints the JSON line the driver expects; the GPU arm refuses to run without a CUDA device (no silent
CPU path)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "cymf")


@pytest.mark.skipif(not os.path.isdir(REF), reason="compiled reference (oracle/_ref) not built")
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "wmf_als_epochs_per_sec"
    assert line["unit"] == "epochs/s" and line["higher_is_better"] is True and line["scaling"] == "strong"
    assert 1e-4 < line["value"] < 10 and abs(line["value"] * line["sec_per_epoch"] - 1) < 1e-9
    assert "every 20th user row" in line["config"]["cpu_arm_sample"] == line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "ml-20m" in line["config"]["workload"]


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-extra",
                          "--no-cpu"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and out.stdout.strip() == ""          # no number is ever produced on a CPU
