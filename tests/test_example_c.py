"""examples/frame_loop.c: a plain C host against the C-ABI (include/sphsm_b200.h) — compiles and links everywhere; without a
CUDA device it must fail LOUDLY at sphsm_create (no CPU fallback), with one it runs the asynchronous per-frame loop."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "sph_sm_monodomain_b200")


def _build(tmp_path):
    exe = str(tmp_path / "frame_loop")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "frame_loop.c"), "-L", LIBDIR,
           "-lsphsm_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    if _has_gpu():
        pytest.skip("a CUDA device is present: the run itself is the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "sphsm_create" in r.stderr and "failed" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_example_runs_the_async_frame_loop(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "24", "12", "12", "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "frame_loop OK: 3456 particles, 40 frames" in r.stdout and "40 steps taken" in r.stdout, r.stdout + r.stderr
