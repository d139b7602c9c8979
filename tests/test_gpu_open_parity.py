"""GPU parity tests for `Pcs::open` (reference crates/stark/src/prover.rs:460-470): opened values, FRI
commit-phase roots, final polynomial, proof-of-work witness and every query opening must equal the CPU
oracle's (oracle/stark.py) bit for bit on the same committed matrices and the same challenger state, and
the oracle's restated `Pcs::verify` must accept the GPU proof."""
import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
P = bf.P


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


@pytest.fixture(scope="module")
def S(oracle):
    import oracle.stark as S
    return S


def test_challenger_matches_oracle(ctx, S):
    a, b = bf.Challenger(ctx), S.Challenger()
    rng = np.random.default_rng(1)
    for step in range(40):
        if step % 3 == 0:
            v = rng.integers(0, P, int(rng.integers(1, 12)), dtype=np.uint32)
            a.observe_slice(v)
            b.observe_slice(v)
        elif step % 3 == 1:
            assert a.sample_ext().tolist() == b.sample_ext().tolist()
        else:
            assert a.sample_bits(13) == b.sample_bits(13)
    st, ib, ob = a.export()
    assert st.tolist() == b.state.tolist() and ib.tolist() == b.inp and ob.tolist() == b.out
    c = a.clone()
    assert c.sample() == a.sample() == b.sample()


def build_rounds(ctx, oracle, S, rng, shapes_per_round, zeta):
    pcs = bf.TwoAdicFriPcs(ctx)
    g_rounds, o_rounds = [], []
    for shapes in shapes_per_round:
        evals = [rng.integers(0, P, s, dtype=np.uint32) for s in shapes]
        root, data = pcs.commit(evals)
        od = oracle.PcsData(evals)
        assert (root == od.root).all()
        pts = []
        for r, _ in shapes:
            dom = S.Domain(r.bit_length() - 1)
            pts.append([zeta, dom.next_point(zeta)] if r > 4 else [zeta])
        g_rounds.append((data, pts))
        o_rounds.append((od, pts))
    return pcs, g_rounds, o_rounds


def assert_same_opening(got, ref):
    (g_open, g_proof), (r_open, r_proof) = got, ref
    for gr, rr in zip(g_open, r_open):
        for gm, rm in zip(gr, rr):
            for gp, rp in zip(gm, rm):
                assert (np.asarray(gp) == np.asarray(rp)).all()
    assert len(g_proof["commit_phase_commits"]) == len(r_proof["commit_phase_commits"])
    for a, b in zip(g_proof["commit_phase_commits"], r_proof["commit_phase_commits"]):
        assert (a == b).all()
    assert (g_proof["final_poly"] == r_proof["final_poly"]).all()
    assert g_proof["pow_witness"] == r_proof["pow_witness"]
    for gq, rq in zip(g_proof["query_proofs"], r_proof["query_proofs"]):
        assert gq["index"] == rq["index"]
        for gi, ri in zip(gq["input_proof"], rq["input_proof"]):
            for a, b in zip(gi["opened_values"], ri["opened_values"]):
                assert (a == b).all()
            assert (gi["opening_proof"] == ri["opening_proof"]).all()
        for gs, rs in zip(gq["commit_phase_openings"], rq["commit_phase_openings"]):
            assert (gs["sibling_value"] == rs["sibling_value"]).all()
            assert (gs["opening_proof"] == rs["opening_proof"]).all()


@pytest.mark.parametrize("shapes_per_round,queries,pow_bits", [
    ([[(64, 3), (16, 2)], [(64, 5), (32, 1), (4, 2)], [(8, 4)]], 10, 6),
    ([[(16, 1)]], 3, 2),
    ([[(4096, 6), (64, 2)], [(4096, 31), (2048, 41), (512, 7), (16, 5)], [(4096, 36), (2048, 8), (512, 16), (16, 8)], [(4096, 4), (4096, 4), (2048, 4)]], 84, 16),
])
def test_open_matches_oracle_and_verifies(ctx, oracle, S, shapes_per_round, queries, pow_bits):
    rng = np.random.default_rng(len(shapes_per_round) * 100 + queries)
    zeta = rng.integers(0, P, 4, dtype=np.uint64)
    ctx.set_fri_params(1, queries, pow_bits)
    try:
        pcs, g_rounds, o_rounds = build_rounds(ctx, oracle, S, rng, shapes_per_round, zeta)
        gch, och = bf.Challenger(ctx), S.Challenger()
        seed = rng.integers(0, P, 11, dtype=np.uint32)
        gch.observe_slice(seed)
        och.observe_slice(seed)
        cfg = S.FriConfig(1, queries, pow_bits)
        ref = S.pcs_open(cfg, o_rounds, och.clone())
        got = pcs.open(g_rounds, gch)
        assert_same_opening(got, ref)
        # the challenger was advanced identically
        och2 = och.clone()
        S.pcs_open(cfg, o_rounds, och2)
        assert gch.sample() == och2.sample()
        # oracle verifier accepts the GPU proof
        vr = []
        for (data, pts), rv in zip(o_rounds, got[0]):
            mats = []
            for lde, p, mv in zip(data.ldes, pts, rv):
                mats.append((S.Domain((lde.shape[0] >> 1).bit_length() - 1), list(zip(p, mv))))
            vr.append((data.root.copy(), mats))
        assert S.pcs_verify(cfg, vr, got[1], och.clone()) is None
        for d, _ in g_rounds:
            d.free()
    finally:
        ctx.set_fri_params(1, 84, 16)


# ---- transcript options: every Plonky3-internal choice that cannot be confirmed offline is a switch on both sides -------------
@pytest.mark.parametrize("observe,rollin,pow_order", [(1, 0, 0), (0, 0, 0), (1, 1, 0), (1, 0, 1), (0, 1, 1)])
def test_transcript_options_match_oracle(ctx, oracle, S, monkeypatch, observe, rollin, pow_order):
    """bfgpu_set_transcript_option (OBSERVE_OPENED_VALUES / FRI_ROLLIN / POW_ORDER) against the oracle's switches of the same
    meaning: identical openings and proofs for every setting, the oracle verifier with the same setting accepts, and a
    verifier with a DIFFERENT transcript setting rejects (the switch really changes the transcript)."""
    shapes_per_round = [[(64, 3), (16, 2)], [(64, 5), (32, 1), (4, 2)], [(8, 4)]]
    rng = np.random.default_rng(77)
    zeta = rng.integers(0, P, 4, dtype=np.uint64)
    monkeypatch.setattr(S, "OBSERVE_OPENED_VALUES", bool(observe))
    monkeypatch.setattr(S, "FRI_ROLLIN", rollin)
    monkeypatch.setattr(S, "POW_ORDER", pow_order)
    ctx.set_fri_params(1, 9, 5)
    ctx.set_transcript_option("observe_opened_values", observe)
    ctx.set_transcript_option("fri_rollin", rollin)
    ctx.set_transcript_option("pow_order", pow_order)
    try:
        pcs, g_rounds, o_rounds = build_rounds(ctx, oracle, S, rng, shapes_per_round, zeta)
        gch, och = bf.Challenger(ctx), S.Challenger()
        seed = rng.integers(0, P, 5, dtype=np.uint32)
        gch.observe_slice(seed)
        och.observe_slice(seed)
        cfg = S.FriConfig(1, 9, 5)
        ref = S.pcs_open(cfg, o_rounds, och.clone())
        got = pcs.open(g_rounds, gch)
        assert_same_opening(got, ref)
        vr = []
        for (data, pts), rv in zip(o_rounds, got[0]):
            vr.append((data.root.copy(), [(S.Domain((lde.shape[0] >> 1).bit_length() - 1), list(zip(p, mv))) for lde, p, mv in zip(data.ldes, pts, rv)]))
        assert S.pcs_verify(cfg, vr, got[1], och.clone()) is None
        monkeypatch.setattr(S, "OBSERVE_OPENED_VALUES", not observe)
        assert S.pcs_verify(cfg, vr, got[1], och.clone()) is not None
        monkeypatch.setattr(S, "OBSERVE_OPENED_VALUES", bool(observe))
        monkeypatch.setattr(S, "FRI_ROLLIN", 1 - rollin)
        assert S.pcs_verify(cfg, vr, got[1], och.clone()) is not None
        for d, _ in g_rounds:
            d.free()
    finally:
        ctx.set_fri_params(1, 84, 16)
        ctx.set_transcript_option("observe_opened_values", 1)
        ctx.set_transcript_option("fri_rollin", 0)
        ctx.set_transcript_option("pow_order", 0)


# ---- BFGPU_REPR_MONTY: the representation the Rust shim passes (Vec<KoalaBear> memory) through commit and open ----------------------
R32 = (1 << 32) % P
RINV = pow(R32, P - 2, P)


def to_m(a):
    return (np.asarray(a, np.uint64) * np.uint64(R32) % np.uint64(P)).astype(np.uint32)


def from_m(a):
    return (np.asarray(a, np.uint64) * np.uint64(RINV) % np.uint64(P)).astype(np.uint64)


def test_montgomery_representation_commit_and_open(oracle, S):
    shapes_per_round = [[(256, 6), (64, 2)], [(256, 31), (128, 41), (16, 5)], [(256, 4), (256, 4), (128, 4)]]
    rng = np.random.default_rng(4242)
    zeta = rng.integers(0, P, 4, dtype=np.uint64)
    mctx = bf.Context(repr=bf.REPR_MONTY)
    mctx.set_fri_params(1, 12, 6)
    pcs = bf.TwoAdicFriPcs(mctx)
    g_rounds, o_rounds = [], []
    w9 = pow(3, (P - 1) >> 9, P)
    for ri, shapes in enumerate(shapes_per_round):
        evals = [rng.integers(0, P, s, dtype=np.uint32) for s in shapes]
        shifts = [1] * len(shapes) if ri < 2 else [1, pow(w9, P - 2, P), 1]  # one shifted (quotient-chunk style) domain
        root, data = pcs.commit([to_m(e) for e in evals], domain_shifts=to_m(shifts))
        od = oracle.PcsData(evals, domain_shifts=shifts)
        assert (from_m(root) == od.root).all(), "Montgomery-representation commit root"
        assert (from_m(pcs.get_evaluations_on_domain(data, 0, bit_reversed_rows=True)) == od.ldes[0]).all()
        rows, sib = bf.MerkleTreeMmcs(mctx).open_batch(3, data.tree)
        orow, osib = od.tree.open_batch(3)
        assert all((from_m(a) == b).all() for a, b in zip(rows, orow)) and (from_m(sib) == osib).all()
        pts = [[zeta, S.Domain(r.bit_length() - 1).next_point(zeta)] if r > 16 else [zeta] for r, _ in shapes]
        g_rounds.append((data, [[to_m(z) for z in p] for p in pts]))
        o_rounds.append((od, pts))
    gch, och = bf.Challenger(mctx), S.Challenger()
    seed = rng.integers(0, P, 7, dtype=np.uint32)
    gch.observe_slice(to_m(seed))
    och.observe_slice(seed)
    cfg = S.FriConfig(1, 12, 6)
    r_open, r_proof = S.pcs_open(cfg, o_rounds, och)
    g_open, g_proof = pcs.open(g_rounds, gch)
    conv = dict(commit_phase_commits=[from_m(c) for c in g_proof["commit_phase_commits"]], final_poly=from_m(g_proof["final_poly"]),
                pow_witness=g_proof["pow_witness"],  # the witness is serialised canonically in both representations
                query_proofs=[dict(index=q["index"],
                                   input_proof=[dict(opened_values=[from_m(r) for r in ip["opened_values"]], opening_proof=from_m(ip["opening_proof"])) for ip in q["input_proof"]],
                                   commit_phase_openings=[dict(sibling_value=from_m(st["sibling_value"]), opening_proof=from_m(st["opening_proof"]))
                                                          for st in q["commit_phase_openings"]]) for q in g_proof["query_proofs"]])
    assert_same_opening(([[[from_m(v) for v in m] for m in rnd] for rnd in g_open], conv), (r_open, r_proof))
    assert from_m([gch.sample()])[0] == och.sample()
    for d, _ in g_rounds:
        d.free()
    mctx.close()
