"""The C ABI seen from C: `include/reach.h` must compile as strict C99 (no C++-isms, no torch types) and a plain C program
must link against libreach_sm100a.so and run a flowpipe (tests/c_abi/flowpipe_smoke.c).  CPU: compile + link, and the
binary fails loudly without a GPU (no CPU fallback).  GPU: the binary's closed-form checks pass."""
import os
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "c_abi" / "flowpipe_smoke.c"


def _build(tmp_path, lib_path):
    exe = tmp_path / "flowpipe_smoke"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(SRC),
           "-L", str(lib_path.parent), "-lreach_sm100a", "-lm", f"-Wl,-rpath,{lib_path.parent}", "-o", str(exe)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_strict_c99_and_a_c_client_links(tmp_path, lib_path):
    exe = _build(tmp_path, lib_path)
    import torch
    if torch.cuda.is_available():
        return                                   # the GPU test below runs it
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, (out.returncode, out.stdout, out.stderr)       # no device: refuses, never computes on the host
    assert "no CUDA device visible" in out.stderr or "CUDA" in out.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_client_runs_a_flowpipe_on_the_gpu(tmp_path, lib_path):
    exe = _build(tmp_path, lib_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=dict(os.environ))
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "flowpipe_smoke: OK" in out.stdout
