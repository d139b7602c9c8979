"""Quotient values FROM THE DEFINITION at sampled points of the quotient domain, independently of the declarative AIR, of the generated
constraint programs and of the numpy prover.

For a chip with trace domain H (n rows, generator g) the reference evaluates, at every point x = 3 * w_2n^i of the quotient domain
(crates/stark/src/quotient.rs:18-165),  q(x) = fold_alpha(constraints(x)) / Z_H(x)  where
  * the rows "local" / "next" are the trace polynomials at x and g * x (next_step = 2^log_quotient_degree positions further, :41-42),
  * `chip.eval(folder)` runs the chip's own constraints and then `eval_permutation_constraints` (chip.rs:216-229, permutation.rs:157-272),
  * every assert multiplies the accumulator by alpha and adds the constraint (folder.rs:68-89): a Horner fold in call order,
  * selectors (Plonky3 `selectors_on_coset`, P3): Z_H(x) = x^n - 1, is_first_row = Z_H / (x - 1), is_last_row = Z_H / (x - g^-1),
    is_transition = x - g^-1, inv_zeroifier = 1 / Z_H.
Here: the polynomials are evaluated by O(n^2) Lagrange interpolation of the trace columns (tests/pyref.py), the chip constraints and lookups
come from tests/ref_air.py (hand-transcribed `eval`s), the permutation constraints below are restated from permutation.rs, arithmetic is
big-int.  The result must equal `oracle.prover.quotient_values` (which the CUDA quotient kernels reproduce word for word,
tests/test_gpu_prove_parity.py / test_gpu_plug_point2.py) at every sampled index."""
import importlib

import numpy as np

import pyref
import ref_air as R
from test_logup_independent import REF, e_add, e_scale, fingerprint

P = R.P
ONE, ZERO = [1, 0, 0, 0], [0, 0, 0, 0]


def e_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def poly_row(cols, x):
    """The row of trace polynomials at x: one Lagrange evaluation per column."""
    return [pyref.interpolate_eval([int(v) for v in col], x) for col in cols.T]


def ext_row(flat_row, width):
    return [[flat_row[4 * j + k] for k in range(4)] for j in range(width)]


def permutation_constraints(b, perm_local, perm_next, alpha, beta, cumulative_sum, is_first, is_last, is_trans, batch_size=2):
    """eval_permutation_constraints (permutation.rs:157-272) on one point; b = the chip's lookups evaluated on the local row."""
    beta_pow = [ONE]
    for _ in range(8):
        beta_pow.append(pyref.e_mul(beta_pow[-1], beta))
    inter = [(l, True) for l in b.sends] + [(l, False) for l in b.receives]
    out = []
    assert len(perm_local) == -(-len(inter) // batch_size) + 1          # permutation_trace_width
    for entry, s in zip(perm_local[:-1], range(0, len(inter), batch_size)):
        rlcs, mults = [], []
        for (kind, values, mult), is_send in inter[s:s + batch_size]:
            rlcs.append(fingerprint(alpha, beta_pow, kind, values))
            mults.append(mult if is_send else (-mult) % P)
        product, numerator = ONE, ZERO
        for i, (m, rlc) in enumerate(zip(mults, rlcs)):
            product = pyref.e_mul(product, rlc)
            others = ONE
            for j, o in enumerate(rlcs):
                if j != i:
                    others = pyref.e_mul(others, o)
            numerator = e_add(numerator, e_scale(others, m))
        out.append(e_sub(pyref.e_mul(product, entry), numerator))
    sum_local, sum_next = ZERO, ZERO
    for v in perm_local[:-1]:
        sum_local = e_add(sum_local, v)
    for v in perm_next[:-1]:
        sum_next = e_add(sum_next, v)
    phi_local, phi_next = perm_local[-1], perm_next[-1]
    out.append(e_scale(e_sub(phi_local, sum_local), is_first))
    out.append(e_scale(e_sub(e_sub(phi_next, phi_local), sum_next), is_trans))
    out.append(e_scale(e_sub(phi_local, cumulative_sum), is_last))
    return out


def quotient_at(name, prep, main, perm_flat, perm_width, alpha_perm, beta_perm, alpha, cumulative_sum, i):
    n = main.shape[0]
    log_n = n.bit_length() - 1
    g = pyref.two_adic_generator(log_n)
    x = 3 * pow(pyref.two_adic_generator(log_n + 1), i, P) % P           # i-th point of the disjoint quotient domain, shift = generator 3
    xg = x * g % P                                                      # two positions further on a domain twice as long
    z = (pow(x, n, P) - 1) % P
    g_inv = pow(g, -1, P)
    is_first = z * pow((x - 1) % P, -1, P) % P
    is_last = z * pow((x - g_inv) % P, -1, P) % P
    is_trans = (x - g_inv) % P
    local, nxt = poly_row(main, x), poly_row(main, xg)
    kw = {} if prep is None else {"prep": poly_row(prep, x)}
    b = REF[name](local, nxt, is_first, is_last, is_trans, **kw)
    perm_local, perm_next = ext_row(poly_row(perm_flat, x), perm_width), ext_row(poly_row(perm_flat, xg), perm_width)
    acc = ZERO
    for c in b.constraints:                                             # the chip's own constraints, in call order
        acc = e_add(pyref.e_mul(acc, alpha), [c % P, 0, 0, 0])
    for c in permutation_constraints(b, perm_local, perm_next, alpha_perm, beta_perm, cumulative_sum, is_first, is_last, is_trans):
        acc = e_add(pyref.e_mul(acc, alpha), c)
    return e_scale(acc, pow(z, -1, P))


def test_quotient_values_from_the_definition(oracle):
    from oracle import prover as PR, stark as S
    ex = importlib.import_module("oracle.machine.executor")
    tg = importlib.import_module("oracle.machine.tracegen")
    chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
    prog = ex.Program("++[>+<-]>,.")
    traces, preps = tg.generate_traces(ex.execute(prog, [3])), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), S.FriConfig(1, 4, 2))
    dbg = proof["_debug"]
    alpha_perm, beta_perm = ([int(v) for v in c] for c in dbg["perm_challenges"])
    alpha = [int(v) for v in dbg["alpha"]]
    rng = np.random.default_rng(5)
    checked = 0
    for k, (chip, prep, main, perm) in enumerate(zip(dbg["ordered"], dbg["preps"], dbg["mains"], dbg["perms"])):
        if main.shape[0] > 64:
            continue                                                    # Byte (2^16 rows): O(n^2) interpolation; Program covers the preprocessed path
        ld = main.shape[0].bit_length() - 1
        csum = [int(v) for v in proof["opened_values"][proof["chip_ordering"][chip.name]]["cumulative_sum"]]
        nat = lambda lde: np.asarray(lde)[S.bitrev_perm(ld + 1)]        # noqa: E731  (committed LDEs are stored with bit-reversed rows)
        prep_q = nat(pk.data.ldes[pk.chip_ordering[chip.name]]) if chip.name in pk.chip_ordering else np.zeros((2 << ld, 1), np.uint32)
        theirs = np.asarray(PR.quotient_values(chip, np.asarray(csum, np.uint64), ld, prep_q, nat(dbg["main_data"].ldes[k]), nat(dbg["perm_data"].ldes[k]),
                                               dbg["perm_challenges"], dbg["alpha"]))
        assert theirs.shape == (2 << ld, 4)
        main = np.asarray(main, np.uint64)
        prep = None if prep is None else np.asarray(prep, np.uint64)
        perm_flat = np.asarray(perm, np.uint64).reshape(main.shape[0], -1)
        for i in [0, 1, (2 << ld) - 1] + [int(v) for v in rng.integers(0, 2 << ld, 3)]:
            mine = quotient_at(chip.name, prep, main, perm_flat, perm.shape[1], alpha_perm, beta_perm, alpha, csum, i)
            assert mine == [int(v) for v in theirs[i]], f"{chip.name}: quotient value at coset index {i}"
            checked += 1
    assert checked == 7 * 6
