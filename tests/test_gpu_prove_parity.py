"""GPU parity tests for the shard prover (`MachineProver::{commit, open}`, reference
crates/stark/src/prover.rs:209-553): on the same traces and the same transcript the CUDA prover must produce
the same three commitments, cumulative sums, opened values, FRI commit-phase roots, final polynomial, PoW
witness and query openings as the CPU oracle (oracle/prover.py), and the oracle's restated
`Verifier::verify_shard` (crates/stark/src/verifier.rs:27-216) must accept the GPU proof and reject corruptions.
Programs are the reference's own test programs (crates/core/machine/src/brainfuck/mod.rs:113-189,
crates/test-artifacts/guests/*.bf, committed under tests/golden/)."""
import importlib
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def env(oracle):
    from oracle import prover as PR, stark as S
    ex = importlib.import_module("oracle.machine.executor")
    tg = importlib.import_module("oracle.machine.tracegen")
    chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
    return PR, S, ex, tg, chips


def same(a, b):
    return (np.asarray(a, np.uint64) == np.asarray(b, np.uint64)).all()


def compare_proofs(g, r):
    for k in ("main", "permutation", "quotient"):
        assert same(g["commitment"][k], r["commitment"][k]), k
    assert g["chip_ordering"] == r["chip_ordering"]
    for gc, rc in zip(g["opened_values"], r["opened_values"]):
        assert gc["log_degree"] == rc["log_degree"]
        assert same(gc["cumulative_sum"], rc["cumulative_sum"])
        for part in ("preprocessed", "main", "permutation"):
            assert same(gc[part]["local"], rc[part]["local"]) and same(gc[part]["next"], rc[part]["next"]), part
        for a, b in zip(gc["quotient"], rc["quotient"]):
            assert same(a, b)
    gf, rf = g["opening_proof"], r["opening_proof"]
    assert len(gf["commit_phase_commits"]) == len(rf["commit_phase_commits"])
    for a, b in zip(gf["commit_phase_commits"], rf["commit_phase_commits"]):
        assert same(a, b)
    assert same(gf["final_poly"], rf["final_poly"]) and gf["pow_witness"] == rf["pow_witness"]
    for gq, rq in zip(gf["query_proofs"], rf["query_proofs"]):
        assert gq["index"] == rq["index"]
        for gi, ri in zip(gq["input_proof"], rq["input_proof"]):
            for a, b in zip(gi["opened_values"], ri["opened_values"]):
                assert same(a, b)
            assert same(gi["opening_proof"], ri["opening_proof"])
        for gs, rs in zip(gq["commit_phase_openings"], rq["commit_phase_openings"]):
            assert same(gs["sibling_value"], rs["sibling_value"]) and same(gs["opening_proof"], rs["opening_proof"])


PROGRAMS = [("++-.", []), (">><", []), ("[----]", []), (",.", [7]), ("++[>+<-]>.", []), ("loop.bf", []), ("move.bf", []), ("printa.bf", []),
            ("hello.bf", [])]


@pytest.mark.parametrize("code,stdin", PROGRAMS)
def test_gpu_proof_equals_oracle_proof_and_verifies(ctx, env, code, stdin):
    PR, S, ex, tg, chips = env
    if code.endswith(".bf"):
        code = open(os.path.join(GOLD, code)).read()
    prog = ex.Program(code)
    rec = ex.execute(prog, stdin)
    traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
    queries, pow_bits = 12, 6
    ctx.set_fri_params(1, queries, pow_bits)
    cfg = S.FriConfig(1, queries, pow_bits)
    try:
        prover = bf.CudaProver(ctx)
        pk = prover.setup(preps)
        opk = PR.setup(chips, preps)
        assert same(pk.commit, opk.commit) and pk.names == opk.names
        gch, och = bf.Challenger(ctx), S.Challenger()
        PR.observe_pk(opk, och)
        ref = PR.prove_shard(chips, opk, traces, och.clone(), cfg)
        got = prover.prove(pk, traces, gch)
        compare_proofs(got, ref)
        vk = dict(commit=opk.commit, chip_information=[(n, t.shape[0].bit_length() - 1, lo) for n, t, lo in zip(opk.names, opk.traces, opk.local_only)])
        assert PR.verify_shard(chips, vk, got, och.clone(), cfg) is None
        # a single corrupted opened value must be rejected
        bad = dict(got, opened_values=[dict(c) for c in got["opened_values"]])
        bad["opened_values"][0]["main"] = dict(local=got["opened_values"][0]["main"]["local"].copy(), next=got["opened_values"][0]["main"]["next"])
        bad["opened_values"][0]["main"]["local"][0, 0] = (int(bad["opened_values"][0]["main"]["local"][0, 0]) + 1) % bf.P
        assert PR.verify_shard(chips, vk, bad, och.clone(), cfg) is not None
        pk.free()
    finally:
        ctx.set_fri_params(1, 84, 16)


def test_fibo_e2e_core_gpu_proof_verifies(ctx, env):
    """BASELINE config 1 (`test_e2e_core`, crates/sdk/src/lib.rs:185-196): fibo.bf with stdin [17] -> output [85];
    full-size proof (84 queries, 16 PoW bits) on the GPU, accepted by the restated reference verifier."""
    PR, S, ex, tg, chips = env
    prog = ex.Program(open(os.path.join(GOLD, "fibo.bf")).read())
    rec = ex.execute(prog, [17])
    assert rec.output == [85] and rec.cycles == 33341
    traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
    assert traces["Cpu"].shape == (65536, 31)
    prover = bf.CudaProver(ctx)
    pk = prover.setup(preps)
    proof = prover.prove(pk, traces, bf.Challenger(ctx))
    local_only = dict((c[0], c[4]) for c in prover.chips)
    vk = dict(commit=pk.commit, chip_information=[(n, h.bit_length() - 1, local_only[n]) for n, h in zip(pk.names, pk.heights)])
    och = S.Challenger()
    och.observe_digest(pk.commit)
    for _ in range(7):
        och.observe(0)
    assert PR.verify_shard(chips, vk, proof, och, S.FriConfig()) is None
    total = S.E_ZERO
    for c in proof["opened_values"]:
        total = S.e_add(total, c["cumulative_sum"])
    assert S.e_eq(total, S.E_ZERO)
    pk.free()
