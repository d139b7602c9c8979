"""Mutation fuzzing of the native verifier and the bincode serialiser (csrc/verifier.h, host code that parses UNTRUSTED proof words):
no mutant of a valid proof is accepted and no mutant crashes the process — against the shipped library and, when g++ has the
sanitizer runtimes, against an AddressSanitizer + UBSan build of the very same headers (tools/verifier_host_shim.cpp), where any
out-of-bounds read, overflowing shift or signed overflow on a hostile shape word aborts the worker."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FRI = (1, 10, 5)


@pytest.fixture(scope="module")
def proof_file(oracle, tmp_path_factory):
    from oracle import prover as PR, stark as S
    from proofio import serialize
    ex = importlib.import_module("oracle.machine.executor")
    tg = importlib.import_module("oracle.machine.tracegen")
    chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
    prog = ex.Program("++[>+<-]>,.")
    traces, preps = tg.generate_traces(ex.execute(prog, [3])), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), S.FriConfig(*FRI))
    proof.pop("_debug", None)
    words = serialize(proof, pk.names)
    assert bf.verify_core_proof(pk.commit, pk.names, [t.shape[0] for t in pk.traces], words, *FRI) is None
    path = str(tmp_path_factory.mktemp("fuzz") / "proof.npz")
    np.savez(path, words=words, commit=np.asarray(pk.commit, np.uint32), logs=np.array([t.shape[0].bit_length() - 1 for t in pk.traces], np.uint32),
             names=",".join(pk.names), fri=np.array(FRI, np.uint32))
    return path


def run_worker(lib, proof, seed, trials, env=None):
    return subprocess.run([sys.executable, os.path.join(HERE, "fuzz_verifier_worker.py"), lib, proof, str(seed), str(trials)], capture_output=True,
                          text=True, env=env, timeout=900)


def test_no_mutant_is_accepted_by_the_shipped_library(proof_file):
    r = run_worker(bf._build.SO, proof_file, 20261019, 2500)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "accepted 0" in r.stdout


def test_no_mutant_trips_the_sanitizers(proof_file):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no AddressSanitizer runtime in this toolchain")
    out = os.path.join(ROOT, "tools", "bin", "libbfverify_asan.so")
    srcs = [os.path.join(ROOT, "tools", "verifier_host_shim.cpp")] + [os.path.join(ROOT, "zkvm-brainfuck_b200", "csrc", f)
                                                                     for f in ("verifier.h", "gen_air_ext.h", "challenger.h", "kb31.cuh", "gen_air.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-shared", "-fPIC",
                               "-o", out, srcs[0]])
    # the shim carries a copy of the chip table: it must be the library's (gen_air.cuh)
    gen = open(os.path.join(ROOT, "zkvm-brainfuck_b200", "csrc", "gen_air.cuh")).read()
    shim = open(srcs[0]).read()
    import re
    rows = lambda s: re.findall(r'\{"(\w+)", (\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\}', s)  # noqa: E731
    assert rows(gen)[:8] == rows(shim)[:8] and len(rows(gen)) >= 8
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = run_worker(out, proof_file, 7, 1500, env)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
    assert "accepted 0" in r.stdout
