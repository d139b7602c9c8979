"""LogUp permutation traces and cumulative sums FROM THE DEFINITION, independently of the declarative AIR and of the numpy prover.

The oracle prover (oracle/prover.py) — which the CUDA kernels reproduce word for word (tests/test_gpu_prove_parity.py) — builds every
chip's permutation trace from air/chips.py.  Here the same trace is rebuilt row by row with
  * the lookups of tests/ref_air.py (hand-transcribed from each chip's `eval` in the reference, no import of air/chips.py),
  * `populate_permutation_row` / `generate_permutation_trace` restated literally (crates/stark/src/permutation.rs:27-148: denominator
    alpha + beta^0 * kind + sum_k beta^(k+1) * value_k, sends positive, receives negated, batches of `logup_batch_size` = 2
    (chip.rs:158-160) over sends-then-receives, last column = inclusive running sum of the row sums),
  * big-int F_p^4 arithmetic (tests/pyref.py),
and compared with the oracle's trace column for column, with the cumulative sums of the proof, and the sums must cancel over the
chips (verifier.rs:210-213)."""
import importlib

import numpy as np
import pytest

import pyref
import ref_air as R

P = R.P
REF = {"Cpu": R.cpu_eval, "Jump": R.jump_eval, "AddSub": R.addsub_eval, "MemoryInstrs": R.meminstr_eval, "Memory": R.memory_eval, "IO": R.io_eval,
       "Program": R.program_eval, "Byte": R.byte_eval}


def e_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def e_scale(a, s):
    return [x * s % P for x in a]


def fingerprint(alpha, beta_pow, kind, values):
    d = e_add(alpha, [kind % P, 0, 0, 0])                      # betas.next() = beta^0 multiplies the argument index
    for k, v in enumerate(values):
        d = e_add(d, e_scale(beta_pow[k + 1], v))
    return d


def permutation_trace(name, prep, main, alpha, beta, batch_size=2):
    beta_pow = [[1, 0, 0, 0]]
    for _ in range(8):
        beta_pow.append(pyref.e_mul(beta_pow[-1], beta))
    rows, running = [], [0, 0, 0, 0]
    n = main.shape[0]
    for i in range(n):
        kw = {} if prep is None else {"prep": prep[i]}
        b = REF[name](main[i], main[(i + 1) % n], 0, 0, 0, **kw)
        inter = [(l, True) for l in b.sends] + [(l, False) for l in b.receives]
        row = []
        for s in range(0, len(inter), batch_size):
            acc = [0, 0, 0, 0]
            for (kind, values, mult), is_send in inter[s:s + batch_size]:
                if mult == 0:
                    continue                                  # 0 / denominator
                m = mult if is_send else (-mult) % P
                acc = e_add(acc, e_scale(pyref.e_inv(fingerprint(alpha, beta_pow, kind, values)), m))
            row.append(acc)
        for v in row:
            running = e_add(running, v)
        rows.append(row + [list(running)])
    return rows, running


@pytest.mark.parametrize("code,stdin", [("++[>+<-]>,.", [3]), ("+>+[<->-]<[-],.", [250])])
def test_permutation_traces_and_cumulative_sums_from_the_definition(oracle, code, stdin):
    from oracle import prover as PR, stark as S
    ex = importlib.import_module("oracle.machine.executor")
    tg = importlib.import_module("oracle.machine.tracegen")
    chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
    prog = ex.Program(code)
    traces, preps = tg.generate_traces(ex.execute(prog, stdin)), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), S.FriConfig(1, 4, 2))
    dbg = proof["_debug"]
    alpha, beta = ([int(x) for x in c] for c in dbg["perm_challenges"])
    total = [0, 0, 0, 0]
    by_name = {n: k for n, k in proof["chip_ordering"].items()}
    for chip, prep, main, perm in zip(dbg["ordered"], dbg["preps"], dbg["mains"], dbg["perms"]):
        assert 1 << chip.log_quotient_degree == 2                                  # logup_batch_size (chip.rs:158-160)
        main = np.asarray(main, np.uint64)
        prep = None if prep is None else np.asarray(prep, np.uint64)
        rows, csum = permutation_trace(chip.name, prep, main, alpha, beta)
        perm = np.asarray(perm)
        n_inter = len(rows[0]) - 1
        assert perm.shape == (main.shape[0], n_inter + 1, 4), (chip.name, perm.shape)  # permutation_trace_width (permutation.rs:15-21)
        mine = np.array(rows, dtype=np.uint64)
        bad = np.argwhere(mine != perm.astype(np.uint64))
        assert bad.size == 0, f"{chip.name}: permutation trace differs first at (row, column, coefficient) {bad[0].tolist()}"
        got = [int(x) for x in proof["opened_values"][by_name[chip.name]]["cumulative_sum"]]
        assert got == csum, f"{chip.name}: cumulative sum"
        total = e_add(total, csum)
    assert total == [0, 0, 0, 0]                                                   # the lookup bus balances (verifier.rs:210-213)
