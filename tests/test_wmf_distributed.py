"""N > 1 coverage of the WMF path: world_size-2 gloo run on CPU (host-side sharding logic + collective
choreography, oracle as the block solver) and, on a multi-GPU box, the NCCL run against a single-GPU fit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "_wmf_dist_worker.py")


def _launch(mode, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_als_choreography_gloo(world):
    r = _launch("gloo", world, 29731 + world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("gloo sharded ALS == oracle") == world


@pytest.mark.gpu
def test_sharded_als_nccl_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _launch("nccl", min(n, 4), 29741)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
