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
        for packed in ([False, True] if oracle.air_packed_available() and n >= 16 else [False]):  # scalar arm and the 16-rows-per-step AVX-512 arm
            got_perm, got_cs = oracle.air_perm_trace(index, main, prep, a_l, beta, packed=packed)
            assert (got_perm == PR.flatten_to_base(ref_perm)).all(), (chip.name, packed)
            assert (got_cs == ref_cs).all(), (chip.name, packed)
        if n < 2:
            continue
        # quotient: LDEs on g * H_2n; the C arm reads bit-reversed rows, the numpy oracle natural order
        log_n = n.bit_length() - 1
        lde_nat = lambda m: oracle.coset_lde_batch(np.ascontiguousarray(m, np.uint32), 1, 3)
        lde_br = lambda m: oracle.coset_lde_batch_bitrev(np.ascontiguousarray(m, np.uint32), 1, 3)
        flat = PR.flatten_to_base(ref_perm)
        ref_q = PR.quotient_values(chip, ref_cs, log_n, lde_nat(prep) if prep is not None else None, lde_nat(main), lde_nat(flat), (a_l, beta), alpha)
        ref_q = np.asarray(ref_q, np.uint64)
        for packed in ([False, True] if oracle.air_packed_available() and n >= 8 else [False]):
            got_q = oracle.air_quotient(index, lde_br(main), lde_br(prep) if prep is not None else None, lde_br(flat), a_l, beta, ref_cs, alpha, packed=packed)
            assert (got_q[0] == ref_q[0::2]).all() and (got_q[1] == ref_q[1::2]).all(), (chip.name, packed)


def test_open_eval_matches_interpolate_coset(oracle):
    from oracle import stark as S
    P = oracle.P
    rng = np.random.default_rng(5)
    for log_n, w in [(3, 2), (6, 5), (9, 31)]:
        m = rng.integers(0, P, (1 << log_n, w), dtype=np.uint32)
        z = rng.integers(0, P, 4, dtype=np.uint64)
        lde = oracle.coset_lde_batch_bitrev(m, 1, 3)
        low_nat = np.asarray(lde[:1 << log_n])[S.bitrev_perm(log_n)]
        ref = S.interpolate_coset(low_nat, S.GEN, z)
        for packed in ([False, True] if oracle.air_packed_available() and log_n >= 4 else [False]):
            got = oracle.open_eval(lde, z, packed=packed)
            assert (got == np.asarray(ref, np.uint64)).all(), (log_n, w, packed)


def test_open_reduce_matches_numpy_formula(oracle):
    """The per-height reduced-opening accumulation of oracle/stark.py::pcs_open, transcribed for one height with two matrices."""
    from oracle import stark as S
    P = oracle.P
    U = np.uint64
    rng = np.random.default_rng(6)
    log_h = 7
    h = 1 << log_h
    alpha = rng.integers(0, P, 4, dtype=U)
    mats = [(rng.integers(0, P, (h, 5), dtype=np.uint32), [rng.integers(0, P, 4, dtype=U), rng.integers(0, P, 4, dtype=U)]),
            (rng.integers(0, P, (h, 3), dtype=np.uint32), [rng.integers(0, P, 4, dtype=U)])]
    xs = S.f_mul(S.powers(S.two_adic_generator(log_h), h), S.GEN)[S.bitrev_perm(log_h)]
    ref = np.zeros((h, 4), U)
    got = np.zeros((h, 4), np.uint32)
    got_p = np.zeros((h, 4), np.uint32)
    num = 0
    for lde, pts in mats:
        w = lde.shape[1]
        ys_all = [rng.integers(0, P, (w, 4), dtype=U) for _ in pts]  # any claimed values: the formula is linear in them
        apow = S.e_powers(alpha, w)
        row_red = np.zeros((h, 4), U)
        for k in range(w):
            row_red = S.e_add(row_red, S.e_scale(np.broadcast_to(apow[k], (h, 4)), lde[:, k].astype(U)))
        before = num
        for z, ys in zip(pts, ys_all):
            off = S.e_pow(alpha, num)
            y_red = S.e_sum(S.e_mul(apow, ys))
            inv_den = S.e_inv(S.e_sub(z, S.e_from_base(xs)))
            ref = S.e_add(ref, S.e_mul(S.e_mul(S.e_sub(y_red, row_red), inv_den), off))
            num += w
        oracle.open_reduce_add(lde, np.array(pts), np.array(ys_all), alpha, before, got, packed=False)
        if oracle.air_packed_available():
            oracle.open_reduce_add(lde, np.array(pts), np.array(ys_all), alpha, before, got_p, packed=True)
    assert (got == ref).all()
    assert not oracle.air_packed_available() or (got_p == ref).all()


@pytest.mark.parametrize("rollin", [False, True])
def test_fri_commit_phase_matches_numpy(oracle, rollin):
    from oracle import stark as S
    if not oracle.fast_available():
        pytest.skip("needs AVX-512")
    P = oracle.P
    U = np.uint64
    rng = np.random.default_rng(7)
    inputs = [rng.integers(0, P, (1 << 8, 4), dtype=U), rng.integers(0, P, (1 << 6, 4), dtype=U), rng.integers(0, P, (1 << 3, 4), dtype=U)]
    betas = rng.integers(0, P, (7, 4), dtype=U)
    # numpy commit phase with the same betas (oracle/stark.py::fri_commit_phase without the low-degree assertion)
    rest = list(inputs)
    folded = rest.pop(0)
    roots = []
    r = 0
    while folded.shape[0] > 2:
        leaves = folded.reshape(-1, 2, 4)
        roots.append(S.ExtTree(leaves).root.copy())
        folded = S.fold_matrix(betas[r], leaves[:, 0], leaves[:, 1])
        if rest and rest[0].shape[0] == folded.shape[0]:
            ro = rest.pop(0)
            if rollin:
                ro = S.e_mul(np.broadcast_to(S.e_mul(betas[r], betas[r]), ro.shape), ro)
            folded = S.e_add(folded, ro)
        r += 1
    g_roots, g_final, _ = oracle.fast_fri_commit_phase(inputs, betas, rollin_beta2=rollin)
    assert (g_roots == np.array(roots)).all()
    assert (g_final == folded[0]).all()


def test_bench_cpu_prove_arm_runs_all_spans(oracle):
    """bench.py's per-phase CPU prove arm on a small program: every span of the reference's prover is exercised and timed."""
    if not oracle.fast_available():
        pytest.skip("needs AVX-512")
    import bench
    chips = chips_mod.machine_chips()
    prog = ex.Program("+++++[>+++[>+>+<<-]<-]>>.")
    traces, preps = tg.generate_traces(ex.execute(prog, [])), tg.preprocessed_traces(prog)
    r = bench.cpu_prove_phases(traces, preps, [c.name for c in chips], local_only=[c.name for c in chips if c.local_only])
    assert set(r["phases_ms"]) == {"commit main", "generate permutation traces", "commit permutation traces", "compute quotient values", "commit quotient", "open"}
    assert all(v > 0 for v in r["phases_ms"].values()) and all(v > 0 for v in r["open_parts_ms"].values())
    assert abs(r["ms"] - sum(r["phases_ms"].values())) < 1e-6 and r["kind"] == "port"
