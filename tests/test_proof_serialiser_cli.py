"""The canonical proof serialiser (bfgpu_shard_proof_to_bincode: the bytes `bincode::serialize(&MachineProof)` writes,
reference crates/stark/src/types.rs:32-73,116-119 and crates/core/machine/src/utils/prove.rs:47-56) against an independent
writer that walks the NESTED proof (oracle/prover.py's ShardProof dict) field by field, and the verifier CLI on files."""
import importlib
import json
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf
from proofio import serialize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
FRI = (1, 7, 4)
R32 = (1 << 32) % bf.P


def bincode_from_dict(proof, pk_names, monty):
    """serde/bincode 1.x image of MachineProof { shard_proof: ShardProof }, written from the nested dict"""
    out = bytearray()
    u64 = lambda v: out.extend(struct.pack("<Q", int(v)))
    fe = lambda v: out.extend(struct.pack("<I", (int(v) * R32 % bf.P) if monty else int(v)))
    ext = lambda e: [fe(x) for x in np.asarray(e).ravel()[:4]]

    def vec_ext(v):
        v = np.asarray(v, np.uint64).reshape(-1, 4)
        u64(len(v))
        for e in v:
            ext(e)

    for k in ("main", "permutation", "quotient"):
        for x in proof["commitment"][k]:
            fe(x)
    order = sorted(proof["chip_ordering"], key=lambda n: proof["chip_ordering"][n])
    by_name = {c.name: c for c in chips}
    u64(len(order))
    for name, c in zip(order, proof["opened_values"]):
        for part, both in (("preprocessed", name in pk_names and not by_name[name].local_only), ("main", not by_name[name].local_only), ("permutation", True)):
            vec_ext(c[part]["local"])
            nxt = np.asarray(c[part]["next"], np.uint64).reshape(-1, 4)
            if not both:  # the reference stores zeros of the same width for local-only chips (prover.rs:485-487)
                nxt = np.zeros((len(np.asarray(c[part]["local"]).reshape(-1, 4)), 4), np.uint64)
            vec_ext(nxt)
        u64(len(c["quotient"]))
        for q in c["quotient"]:
            vec_ext(q)
        ext(c["cumulative_sum"])
        u64(c["log_degree"])
    fri = proof["opening_proof"]
    u64(len(fri["commit_phase_commits"]))
    for c in fri["commit_phase_commits"]:
        for x in c:
            fe(x)
    u64(len(fri["query_proofs"]))
    for q in fri["query_proofs"]:
        u64(len(q["input_proof"]))
        for rnd in q["input_proof"]:
            u64(len(rnd["opened_values"]))
            for row in rnd["opened_values"]:
                u64(len(row))
                for x in row:
                    fe(x)
            sib = np.asarray(rnd["opening_proof"]).reshape(-1, 8)
            u64(len(sib))
            for x in sib.ravel():
                fe(x)
        u64(len(q["commit_phase_openings"]))
        for st in q["commit_phase_openings"]:
            ext(st["sibling_value"])
            sib = np.asarray(st["opening_proof"]).reshape(-1, 8)
            u64(len(sib))
            for x in sib.ravel():
                fe(x)
    ext(fri["final_poly"])
    fe(fri["pow_witness"])
    u64(len(order))
    for i, name in enumerate(order):
        u64(len(name))
        out.extend(name.encode())
        u64(i)
    return bytes(out)


@pytest.fixture(scope="module")
def proof(oracle):
    from oracle import prover as PR, stark as S
    prog = ex.Program("++[>+<-]>,.")
    traces, preps = tg.generate_traces(ex.execute(prog, [5])), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    p = PR.prove_shard(chips, pk, traces, ch.clone(), S.FriConfig(*FRI))
    p.pop("_debug", None)
    return pk, p, serialize(p, pk.names)


@pytest.mark.parametrize("monty", [1, 0])
def test_bincode_image_equals_independent_writer(proof, monty):
    pk, p, words = proof
    heights = [t.shape[0] for t in pk.traces]
    blob = bf.proof_to_bincode(pk.names, heights, words, field_repr=monty)
    ref = bincode_from_dict(p, pk.names, bool(monty))
    assert len(blob) == len(ref)
    assert blob == ref
    # Montgomery-representation input words give the same bytes
    mwords_ok = bf.proof_to_bincode(pk.names, heights, words, field_repr=monty)
    assert mwords_ok == blob
    with pytest.raises(bf.BfGpuError):
        bf.proof_to_bincode(pk.names, heights, words[:-3])


def test_cli_verify_and_size_on_files(proof, tmp_path):
    pk, p, words = proof
    pf, vkf = tmp_path / "p.bfproof", tmp_path / "p.vk.json"
    np.ascontiguousarray(words, dtype="<u4").tofile(pf)
    json.dump(dict(commit=[int(x) for x in pk.commit], names=list(pk.names), heights=[int(t.shape[0]) for t in pk.traces],
                   fri=dict(log_blowup=FRI[0], num_queries=FRI[1], pow_bits=FRI[2])), open(vkf, "w"))
    run = lambda *a: subprocess.run([sys.executable, os.path.join(ROOT, "bfprove.py"), *a], capture_output=True, text=True, timeout=120)
    r = run("verify", str(pf), "--vk", str(vkf))
    assert r.returncode == 0 and "accepted" in r.stdout, r.stdout + r.stderr
    r = run("size", str(pf), "--vk", str(vkf))
    assert r.returncode == 0 and f"({len(bf.proof_to_bincode(pk.names, [t.shape[0] for t in pk.traces], words))} bytes" in r.stdout, r.stdout + r.stderr
    bad = words.copy()
    bad[30] = (int(bad[30]) + 1) % bf.P
    np.ascontiguousarray(bad, dtype="<u4").tofile(pf)
    r = run("verify", str(pf), "--vk", str(vkf))
    assert r.returncode == 1 and "rejected" in r.stdout
    hello = os.path.join(ROOT, "tests", "golden", "hello.bf")
    r = run("execute", hello)
    assert r.returncode == 0 and r.stdout.startswith("Hello")
