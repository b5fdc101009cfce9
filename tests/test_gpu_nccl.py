"""GPU x2+: the NCCL slab path (one process per GPU under torchrun) against the single-GPU step — skipped on boxes with fewer
than two GPUs (the virtual-rank tests in test_gpu_slabs.py cover the same phases on one device)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("exchange", ["push", "nccl"])
@pytest.mark.parametrize("extra", [["--canonical"], ["--quadratic", "--dims", "128x20x20"]], ids=["linear-canonical", "quadratic-default"])
def test_nccl_slabs_match_single_gpu(extra, exchange):
    """Both forms of exchange 1: the push exchange (direct stores into the neighbour's memory; the default wherever CUDA IPC maps the
    neighbours) and ncclSend / ncclRecv (SPHSM_P2P=0)."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "mg_parity.py"), "--steps", "30"] + extra
    env = dict(os.environ, SPHSM_P2P="1" if exchange == "push" else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["mg_parity"] == "ok" and res["every_particle_owned_once"]
    if exchange == "nccl":
        assert res["push_exchange"] is False
    elif not res["push_exchange"]:
        pytest.skip("CUDA IPC could not map the neighbours on this box: the push exchange is off (NCCL path covered by the other case)")
    if "--canonical" in extra:
        assert res["max_rel_dev_vs_single_gpu"] <= 2e-6


def test_nccl_step_error_is_collective():
    """A halo capacity too small for one cell plane: the rank that overflows records the error, it rides on the next step's
    moment allreduce, and EVERY rank returns an error at the end of that step — nobody is left waiting in a collective."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tools", "mg_parity.py"), "--expect-overflow", "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["mg_error_path"] == "ok", res
