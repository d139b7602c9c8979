"""CPU tests of the machine-level oracle (oracle/prover.py) and of the host-side restatements it consumes
(executor, trace generation, chip constraint programs).  The interpreter outputs are the values the reference's
own tests pin (crates/core/executor/src/executor.rs:335-416: fibo(17) -> 85, hello -> "Hello", ...); proofs are
checked the way every reference machine test does: prove, then verify (crates/core/machine/src/brainfuck/mod.rs:113-189)."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips_mod = importlib.import_module("zkvm-brainfuck_b200.air.chips")


def gold(name):
    return open(os.path.join(GOLD, name)).read()


def test_interpreter_pinned_outputs():
    run = lambda code, stdin=(): ex.execute(ex.Program(code), stdin)
    assert run("++-.").output == [1]
    assert run(",.", [9]).output == [9]
    assert bytes(run(gold("hello.bf")).output) == b"Hello"
    assert bytes(run(gold("printa.bf")).output) == b"A"
    r = run(gold("fibo.bf"), [17])
    assert r.output == [85] and r.cycles == 33341 and len(r.program) == 56
    assert run(gold("loop.bf")).cycles == 17 and run(gold("move.bf")).cycles == 12


def test_chip_shapes_match_reference_layouts():
    got = {c.name: (c.main_width, c.prep_width, len(c.sends), len(c.receives), c.perm_width, len(c.constraints), c.log_quotient_degree, c.local_only)
           for c in chips_mod.machine_chips()}
    # SURVEY.md Appendix A.1 (derived from the reference's cols structs and eval functions)
    assert got == {"Cpu": (31, 0, 14, 2, 9, 20, 1, False), "Program": (1, 6, 0, 1, 2, 0, 1, False), "AddSub": (7, 0, 3, 2, 4, 8, 1, True),
                   "Jump": (45, 0, 0, 1, 2, 44, 1, True), "Memory": (12, 0, 2, 2, 3, 0, 1, False), "Byte": (2, 2, 0, 2, 2, 0, 1, False),
                   "MemoryInstrs": (41, 0, 0, 1, 2, 40, 1, False), "IO": (5, 0, 0, 1, 2, 3, 1, True)}


def test_trace_shapes_fibo():
    prog = ex.Program(gold("fibo.bf"))
    tr = tg.generate_traces(ex.execute(prog, [17]))
    assert {k: v.shape[0] for k, v in tr.items()} == {"Cpu": 65536, "MemoryInstrs": 32768, "AddSub": 16384, "Jump": 8192, "Byte": 65536,
                                                      "Program": 64, "Memory": 16, "IO": 16}


@pytest.fixture(scope="module")
def PR(oracle):
    from oracle import prover
    return prover


@pytest.mark.parametrize("code,stdin", [("++[>+<-]>.", []), (",.>,+.", [200]), ("move.bf", [])])
def test_traces_satisfy_constraints_and_lookups_balance(PR, code, stdin):
    from oracle import stark as S
    if code.endswith(".bf"):
        code = gold(code)
    prog = ex.Program(code)
    traces, preps = tg.generate_traces(ex.execute(prog, stdin)), tg.preprocessed_traces(prog)
    by = {c.name: c for c in chips_mod.machine_chips()}
    rng = np.random.default_rng(0)
    ch = [rng.integers(0, S.P, 4, dtype=np.uint64) for _ in range(2)]
    total = S.E_ZERO
    for name, main in traces.items():
        prep = preps.get(name, np.zeros((main.shape[0], 0), np.uint32))
        perm, cs = PR.generate_permutation_trace(by[name], prep, main, ch)
        assert PR.debug_constraints(by[name], prep, main, perm, ch, cs) == [], name
        total = S.e_add(total, cs)
    assert S.e_eq(total, S.E_ZERO)
    # a corrupted trace cell violates a constraint
    bad = traces["Cpu"].copy()
    bad[1, chips_mod.CPU_LAYOUT["next_pc"]] += 1
    perm, cs = PR.generate_permutation_trace(by["Cpu"], np.zeros((bad.shape[0], 0), np.uint32), bad, ch)
    assert PR.debug_constraints(by["Cpu"], np.zeros((bad.shape[0], 0), np.uint32), bad, perm, ch, cs) != []


def test_oracle_prove_verify_and_rejections(PR):
    from oracle import stark as S
    chips = chips_mod.machine_chips()
    prog = ex.Program("++[>+<-]>.")
    traces, preps = tg.generate_traces(ex.execute(prog)), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    cfg = S.FriConfig(1, 8, 4)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), cfg)
    vk = dict(commit=pk.commit, chip_information=[(n, t.shape[0].bit_length() - 1, lo) for n, t, lo in zip(pk.names, pk.traces, pk.local_only)])
    assert PR.verify_shard(chips, vk, proof, ch.clone(), cfg) is None
    # a proof built from a trace that violates a constraint fails the out-of-domain check
    # constraints(zeta) / Z_H(zeta) == quotient(zeta) (verifier.rs:220-247)
    bad = {k: v.copy() for k, v in traces.items()}
    bad["AddSub"][0, chips_mod.ADDSUB_LAYOUT["carry"]] = 1
    err = PR.verify_shard(chips, vk, PR.prove_shard(chips, pk, bad, ch.clone(), cfg), ch.clone(), cfg)
    assert err is not None and err.startswith("OodEvaluationMismatch")
    # tampering with the proof is rejected
    p2 = dict(proof, opened_values=[dict(c) for c in proof["opened_values"]])
    p2["opened_values"][1]["cumulative_sum"] = S.e_add(proof["opened_values"][1]["cumulative_sum"], S.E_ONE)
    assert PR.verify_shard(chips, vk, p2, ch.clone(), cfg) is not None
    p3 = dict(proof, commitment=dict(proof["commitment"], quotient=proof["commitment"]["main"]))
    assert PR.verify_shard(chips, vk, p3, ch.clone(), cfg) is not None
    vk2 = dict(vk, commit=proof["commitment"]["main"])
    assert PR.verify_shard(chips, vk2, proof, ch.clone(), cfg) is not None
