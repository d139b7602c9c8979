"""integration/c_caller.c: the C ABI used from plain C (no Python, no Rust) — compile against include/bfgpu.h, link libbfgpu.so, run.
Without a GPU the program executes the guest through `bfgpu_execute` and stops at "no CPU fallback"; on a B200 it goes on to
setup -> commit -> open -> `bfgpu_verify_core_proof` with the reference's FRI parameters (84 queries, 16 proof-of-work bits)."""
import os
import subprocess

import pytest

import zkvm_brainfuck_b200 as bf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_and_run(tmp_path):
    exe = str(tmp_path / "c_caller")
    libdir = os.path.dirname(bf._build.SO)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "integration", "c_caller.c"),
                           "-L" + libdir, "-lbfgpu", "-Wl,-rpath," + libdir, "-o", exe])
    return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_c_caller_executes_and_has_no_cpu_fallback(tmp_path):
    import torch
    r = build_and_run(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "executed: cycles=110 output=Az" in r.stdout
    if not torch.cuda.is_available():
        assert "prover unavailable" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_caller_proves_and_verifies(tmp_path):
    r = build_and_run(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "verifier: accepted" in r.stdout
