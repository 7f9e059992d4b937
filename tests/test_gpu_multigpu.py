"""-m gpu: strip-partitioned assembly across 2 GPUs (skipped on a single-GPU box) and the C++ adapter example."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["roof", "tension", "roof_pressure"])
def test_strips_two_ranks(case):
    """Two ranks, one matrix: NCCL halo exchange on a 2-GPU box; on a 1-GPU box both ranks assemble their strip on GPU 0
    and exchange over gloo (never skipped).  tension: the collapsed side is one DoF shared by all strips (all-reduce)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", {"roof": "29541", "tension": "29543", "roof_pressure": "29545"}[case], os.path.join(ROOT, "tests", "multigpu_strips.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, KL_CASE=case))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout


def test_patches_two_ranks():
    """Two ranks, one multi-patch matrix split by patches (kl_mp_set_active) with the interface exchange: NCCL on a 2-GPU box,
    gloo with both ranks on GPU 0 otherwise (never skipped)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tests", "multigpu_patches.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ok=True" in r.stdout
