"""CPU tests of the CPU arm of the AIR-driven prover phases (oracle/fast_air.cpp, used only by bench.py's per-phase CPU prove
baseline): its LogUp permutation traces and quotient values equal the numpy oracle's (oracle/prover.py, which restates
crates/stark/src/permutation.rs:75-148 and quotient.rs:18-165) for every chip of a small program."""
import importlib

import numpy as np
import pytest

ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips_mod = importlib.import_module("zkvm-brainfuck_b200.air.chips")


@pytest.mark.parametrize("code,stdin", [("++[>+<-]>,.", [42]), ("+++[>++<-]>.<,.", [7])])
def test_perm_trace_and_quotient_match_numpy_oracle(oracle, code, stdin):
    from oracle import prover as PR, stark as S
    P = oracle.P
    chips = chips_mod.machine_chips()
    prog = ex.Program(code)
    traces, preps = tg.generate_traces(ex.execute(prog, stdin)), tg.preprocessed_traces(prog)
    rng = np.random.default_rng(11)
    a_l, beta, alpha = (rng.integers(0, P, 4, dtype=np.uint64) for _ in range(3))
    for index, chip in enumerate(chips):
        main = traces[chip.name]
        prep = preps.get(chip.name)
        n = main.shape[0]
        # permutation trace
        ref_perm, ref_cs = PR.generate_permutation_trace(chip, prep, main, (a_l, beta))
        got_perm, got_cs = oracle.air_perm_trace(index, main, prep, a_l, beta)
        assert (got_perm == PR.flatten_to_base(ref_perm)).all(), chip.name
        assert (got_cs == ref_cs).all(), chip.name
        if n < 2:
            continue
        # quotient: LDEs on g * H_2n; the C arm reads bit-reversed rows, the numpy oracle natural order
        log_n = n.bit_length() - 1
        lde_nat = lambda m: oracle.coset_lde_batch(np.ascontiguousarray(m, np.uint32), 1, 3)
        lde_br = lambda m: oracle.coset_lde_batch_bitrev(np.ascontiguousarray(m, np.uint32), 1, 3)
        flat = PR.flatten_to_base(ref_perm)
        ref_q = PR.quotient_values(chip, ref_cs, log_n, lde_nat(prep) if prep is not None else None, lde_nat(main), lde_nat(flat), (a_l, beta), alpha)
        got_q = oracle.air_quotient(index, lde_br(main), lde_br(prep) if prep is not None else None, lde_br(flat), a_l, beta, ref_cs, alpha)
        ref_q = np.asarray(ref_q, np.uint64)
        assert (got_q[0] == ref_q[0::2]).all() and (got_q[1] == ref_q[1::2]).all(), chip.name
