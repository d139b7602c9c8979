"""CPU tests of the generic STARK/PCS oracle layer (oracle/stark.py): extension field, challenger,
domains, TwoAdicFriPcs open -> verify round trips.  PARITY UNPINNED vs Plonky3 (no golden vectors in
the reference): these pin prover/verifier self-consistency and the algebraic definitions."""
import numpy as np
import pytest

from tests import pyref

P = pyref.P


@pytest.fixture(scope="module")
def S(oracle):
    import oracle.stark as S
    return S


def test_ext_field(S):
    rng = np.random.default_rng(1)
    a = rng.integers(0, P, (50, 4), dtype=np.uint64)
    b = rng.integers(0, P, (50, 4), dtype=np.uint64)
    ab = S.e_mul(a, b)
    for i in range(5):
        assert ab[i].tolist() == pyref.e_mul([int(x) for x in a[i]], [int(x) for x in b[i]])
    assert S.e_eq(S.e_mul(a, S.e_inv(a)), np.broadcast_to(S.E_ONE, a.shape))
    assert S.e_inv(a)[0].tolist() == pyref.e_inv([int(x) for x in a[0]])
    assert S.e_eq(S.e_pow(a[0], 5), S.e_mul(a[0], S.e_mul(S.e_mul(a[0], a[0]), S.e_mul(a[0], a[0]))))
    pw = S.e_powers(a[0], 9)
    assert S.e_eq(pw[8], S.e_pow(a[0], 8))
    # X^4 = 3
    x = np.array([0, 1, 0, 0], np.uint64)
    assert S.e_pow(x, 4).tolist() == [3, 0, 0, 0]


def test_challenger_duplex_semantics(S, oracle):
    ch = S.Challenger()
    for v in range(1, 9):
        ch.observe(v)  # 8th observation triggers a duplex
    st = oracle.permute(np.array(list(range(1, 9)) + [0] * 8, np.uint32))
    assert ch.sample() == int(st[7])  # pops from the back of state[0..8]
    assert ch.sample() == int(st[6])
    ch.observe(5)  # clears the output buffer, pending input
    st2 = st.copy()
    st2[0] = 5
    st2 = oracle.permute(st2)
    assert ch.sample() == int(st2[7])
    e = ch.clone().sample_ext()
    assert e.tolist() == [int(st2[6]), int(st2[5]), int(st2[4]), int(st2[3])]
    c2 = ch.clone()
    w = c2.grind(8)
    assert ch.clone().check_witness(8, w)


def test_domains_and_selectors(S):
    d = S.Domain(3)
    q = d.create_disjoint_domain(16)
    assert (q.log_n, q.shift) == (4, 3)
    parts = q.split_domains(2)
    assert [x.shift for x in parts] == [3, 3 * S.two_adic_generator(4) % P] and parts[0].log_n == 3
    sel = d.selectors_on_coset(q)
    xs = S.powers(q.gen(), 16, 3)
    g = d.gen()
    for i in range(16):
        x = int(xs[i])
        zh = (pow(x, 8, P) - 1) % P
        assert int(sel["inv_zeroifier"][i]) == pow(zh, -1, P)
        assert int(sel["is_first_row"][i]) == zh * pow(x - 1, -1, P) % P
        assert int(sel["is_last_row"][i]) == zh * pow(x - pow(g, -1, P), -1, P) % P
        assert int(sel["is_transition"][i]) == (x - pow(g, -1, P)) % P
    # selectors_at_point agrees with the coset vectors when the point is a coset point
    pt = S.e_from_base(int(xs[5]))
    sp = d.selectors_at_point(pt)
    assert sp["is_first_row"].tolist() == [int(sel["is_first_row"][5]), 0, 0, 0]
    assert sp["inv_zeroifier"].tolist() == [int(sel["inv_zeroifier"][5]), 0, 0, 0]


def test_interpolate_coset(S, oracle):
    rng = np.random.default_rng(3)
    n, w = 16, 3
    evals_h = rng.integers(0, P, (n, w), dtype=np.uint32)
    on_coset = oracle.coset_lde_batch(evals_h, 0, 3)  # values on 3*H, natural order
    z = rng.integers(0, P, 4, dtype=np.uint64)
    got = S.interpolate_coset(on_coset, 3, z)
    # compare with Horner evaluation of the coefficients at z
    coef = oracle.idft_batch(evals_h)
    for c in range(w):
        acc = S.E_ZERO.copy()
        for k in range(n - 1, -1, -1):
            acc = S.e_add(S.e_mul(acc, z), S.e_from_base(int(coef[k, c])))
        assert S.e_eq(acc, got[c])


def make_rounds(S, oracle, rng, shapes_per_round, zeta):
    rounds = []
    for shapes in shapes_per_round:
        evals = [rng.integers(0, P, s, dtype=np.uint32) for s in shapes]
        data = oracle.PcsData(evals)
        pts = []
        for (r, _) in shapes:
            dom = S.Domain(r.bit_length() - 1)
            pts.append([zeta, dom.next_point(zeta)] if r > 4 else [zeta])
        rounds.append((data, pts))
    return rounds


def verifier_rounds(S, rounds, opened):
    out = []
    for (data, pts), rv in zip(rounds, opened):
        mats = []
        for lde, p, mv in zip(data.ldes, pts, rv):
            dom = S.Domain((lde.shape[0] >> 1).bit_length() - 1)
            mats.append((dom, list(zip(p, mv))))
        out.append((data.root.copy(), mats))
    return out


def test_pcs_open_verify_roundtrip(S, oracle):
    rng = np.random.default_rng(5)
    cfg = S.FriConfig(1, 10, 6)
    zeta = rng.integers(0, P, 4, dtype=np.uint64)
    rounds = make_rounds(S, oracle, rng, [[(64, 3), (16, 2)], [(64, 5), (32, 1), (4, 2)], [(8, 4)]], zeta)
    ch = S.Challenger()
    ch.observe_slice([1, 2, 3])
    opened, proof = S.pcs_open(cfg, rounds, ch.clone())
    assert len(proof["commit_phase_commits"]) == 6  # 2^7 -> 2
    vr = verifier_rounds(S, rounds, opened)
    assert S.pcs_verify(cfg, vr, proof, ch.clone()) is None
    # opened values are the polynomial evaluations
    coef = oracle.idft_batch(rounds[0][0].evals[0])
    acc = S.E_ZERO.copy()
    for k in range(63, -1, -1):
        acc = S.e_add(S.e_mul(acc, zeta), S.e_from_base(int(coef[k, 1])))
    assert S.e_eq(acc, opened[0][0][0][1])
    # tamper: opened value, final poly, a sibling, the witness, the transcript
    bad = verifier_rounds(S, rounds, opened)
    bad[1][1][0][1][0][1][2][0] = (int(bad[1][1][0][1][0][1][2][0]) + 1) % P
    assert S.pcs_verify(cfg, bad, proof, ch.clone()) is not None
    p2 = dict(proof, final_poly=S.e_add(proof["final_poly"], S.E_ONE))
    assert S.pcs_verify(cfg, vr, p2, ch.clone()) is not None
    p3 = dict(proof, pow_witness=proof["pow_witness"] + 1)
    assert S.pcs_verify(cfg, vr, p3, ch.clone()) in ("InvalidPowWitness", "FinalPolyMismatch", "InputMmcsError", "CommitPhaseMmcsError")
    ch2 = ch.clone()
    ch2.observe(7)
    assert S.pcs_verify(cfg, vr, proof, ch2) is not None


def test_fri_rejects_high_degree(S, oracle):
    """A committed matrix that is NOT a low-degree extension makes the commit phase fail its final check."""
    rng = np.random.default_rng(6)
    cfg = S.FriConfig(1, 4, 2)
    data = oracle.PcsData([rng.integers(0, P, (16, 2), dtype=np.uint32)])
    data.ldes[0][3, 1] = (int(data.ldes[0][3, 1]) + 1) % P  # corrupt one LDE entry in place (tree now inconsistent too)
    zeta = rng.integers(0, P, 4, dtype=np.uint64)
    with pytest.raises(AssertionError, match="low-degree"):
        S.pcs_open(cfg, [(data, [[zeta]])], S.Challenger())
