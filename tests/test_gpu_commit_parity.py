"""GPU parity tests (run on the B200 box): CUDA path through the C ABI vs the CPU oracle, bit-exact.

Mirrors how the reference exercises this path: `pcs.commit` on lists of trace matrices of mixed
power-of-two heights (reference crates/stark/src/prover.rs:209-236, machine.rs:195-196), with the
Plonky3 trait surface (Dft / Mmcs / Pcs) tested one level at a time.  PARITY UNPINNED at the
Plonky3 boundary (no reference golden vectors): the oracle itself is pinned in the CPU tests.
"""
import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf
from tests import pyref

pytestmark = pytest.mark.gpu
P = bf.P


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


def rand_mat(rng, rows, cols):
    return rng.integers(0, P, (rows, cols), dtype=np.uint32)


def test_permute_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(1)
    st = rng.integers(0, P, (1000, 16), dtype=np.uint32)
    st[0] = 0
    st[1] = P - 1
    got = ctx.permute(st)
    assert (got == oracle.permute_many(st)).all()
    assert got[2].tolist() == pyref.permute(st[2])


def test_permute_montgomery_repr(ctx, oracle):
    """Host words in Montgomery form (what Vec<KoalaBear> memory holds) give the same permutation."""
    rng = np.random.default_rng(2)
    st = rng.integers(0, P, (64, 16), dtype=np.uint32)
    R = (1 << 32) % P
    to_m = lambda a: ((a.astype(np.uint64) * R) % P).astype(np.uint32)
    ctx.set_repr(bf.REPR_MONTY)
    try:
        got_m = ctx.permute(to_m(st))
    finally:
        ctx.set_repr(bf.REPR_CANONICAL)
    assert (got_m == to_m(oracle.permute_many(st))).all()


@pytest.mark.parametrize("cols", [1, 7, 8, 9, 16, 31, 45, 256])
def test_sponge_rows(ctx, oracle, cols):
    rng = np.random.default_rng(cols)
    m = rand_mat(rng, 70, cols)
    got = ctx.hash_rows(m)
    for r in [0, 1, 33, 69]:
        assert (got[r] == oracle.sponge_hash(m[r])).all()


def test_compress(ctx, oracle):
    rng = np.random.default_rng(5)
    l, r = rand_mat(rng, 50, 8), rand_mat(rng, 50, 8)
    got = ctx.compress(l, r)
    for i in range(50):
        assert (got[i] == oracle.compress(l[i], r[i])).all()


@pytest.mark.parametrize("log_n,cols", [(0, 3), (1, 1), (2, 5), (3, 2), (4, 31), (5, 1), (8, 7), (9, 4), (10, 45), (11, 3), (13, 6), (16, 2), (17, 3)])
def test_dft_and_idft(ctx, oracle, log_n, cols):
    rng = np.random.default_rng(100 + log_n)
    m = rand_mat(rng, 1 << log_n, cols)
    dft = bf.Radix2Dit(ctx)
    assert (dft.dft_batch(m) == oracle.dft_batch(m)).all()
    assert (dft.idft_batch(m) == oracle.idft_batch(m)).all()


@pytest.mark.parametrize("log_n,cols,added,shift", [(0, 2, 1, 3), (1, 3, 1, 3), (3, 1, 1, 3), (4, 5, 1, 3), (4, 5, 2, 3), (6, 12, 1, 3),
                                                  (9, 36, 1, 3), (10, 3, 1, 7), (12, 41, 1, 3), (15, 4, 1, 3), (18, 2, 1, 3),
                                                  # round 2: every instantiation of the TMA passes (kernels_ntt3.cuh).  Pass plans: 13 = 7+6,
                                                  # 14 = 7+7, 16 = 8+8, 17 = 6+6+5, 19 = 7+6+6, 20 = 7+7+6; widths with a pitch that is a multiple
                                                  # of four take the fused ingest + first inverse pass (k_ingest_pass, g = 6, 7, 8), with whole
                                                  # and partial 32-column boxes; the others the stand-alone transpose
                                                  (13, 64, 1, 3), (14, 33, 1, 3), (14, 96, 1, 3), (16, 5, 1, 3), (16, 32, 1, 7), (17, 36, 1, 3),
                                                  (19, 8, 1, 3), (20, 4, 1, 3), (13, 12, 2, 3)])
def test_coset_lde(ctx, oracle, log_n, cols, added, shift):
    rng = np.random.default_rng(200 + log_n)
    m = rand_mat(rng, 1 << log_n, cols)
    dft = bf.Radix2Dit(ctx)
    assert (dft.coset_lde_batch(m, added, shift) == oracle.coset_lde_batch(m, added, shift)).all()
    assert (dft.coset_lde_batch(m, added, shift, bit_reversed_rows=True) == oracle.coset_lde_batch_bitrev(m, added, shift)).all()


@pytest.mark.parametrize("env", [{"BFGPU_NTT_TMA": "0"}, {"BFGPU_NTT_TMA": "0", "BFGPU_NTT_TURN": "0"}, {"BFGPU_NTT_TURN": "0"},
                                 {"BFGPU_NTT_INGEST": "0"}, {"BFGPU_NTT_CFWD": "1"}, {"BFGPU_NTT_TMA": "0", "BFGPU_NTT_DUAL": "0"}])
def test_coset_lde_kernel_switches(oracle, env, monkeypatch):
    """The NTT kernel families behind the run-time switches (register-prefetching ntt2 passes, unfused turn, separate transpose,
    round-1 epilogue path) produce the same LDE; a context reads the switches when it is created."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = bf.Context()
    try:
        dft = bf.Radix2Dit(c)
        for log_n, cols in [(13, 40), (16, 8), (17, 4)]:
            m = rand_mat(np.random.default_rng(300 + log_n), 1 << log_n, cols)
            assert (dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True) == oracle.coset_lde_batch_bitrev(m, 1, 3)).all(), (env, log_n, cols)
    finally:
        c.close()


def test_coset_lde_montgomery_repr_fused_ingest(oracle):
    """Montgomery-form caller words (what a Rust `Vec<KoalaBear>` holds) through the fused ingest + first inverse pass (no conversion on load)."""
    R32 = (1 << 32) % P
    rinv = pow(R32, P - 2, P)
    c = bf.Context(repr=bf.REPR_MONTY)
    try:
        for log_n, cols in [(13, 64), (16, 36)]:
            m = rand_mat(np.random.default_rng(400 + log_n), 1 << log_n, cols)
            mm = (m.astype(np.uint64) * np.uint64(R32) % np.uint64(P)).astype(np.uint32)
            out = bf.Radix2Dit(c).coset_lde_batch(mm, 1, R32 * 3 % P, bit_reversed_rows=True)
            assert ((out.astype(np.uint64) * np.uint64(rinv) % np.uint64(P)) == oracle.coset_lde_batch_bitrev(m, 1, 3)).all()
    finally:
        c.close()


def test_coset_lde_against_definition(ctx):
    """Independent of the oracle: LDE values equal Lagrange evaluation (big-int Python)."""
    rng = np.random.default_rng(7)
    m = rand_mat(rng, 8, 2)
    out = bf.Radix2Dit(ctx).coset_lde_batch(m, 1, 3)
    W = pyref.two_adic_generator(4)
    for j in range(16):
        for c in range(2):
            assert int(out[j, c]) == pyref.interpolate_eval([int(v) for v in m[:, c]], 3 * pow(W, j, P) % P)


SHAPES = [
    [(8, 3)],
    [(1, 5)],
    [(2, 17), (2, 1)],
    [(4, 2), (16, 5), (8, 1), (16, 8), (1, 4)],
    [(1024, 31), (512, 41), (256, 7), (128, 45), (1024, 2), (16, 1), (16, 12), (16, 5)],
]


@pytest.mark.parametrize("shapes", SHAPES)
def test_mmcs_commit_layers_and_openings(ctx, oracle, shapes):
    rng = np.random.default_rng(len(shapes) + 40)
    mats = [rand_mat(rng, r, c) for r, c in shapes]
    mmcs = bf.MerkleTreeMmcs(ctx)
    root, tree = mmcs.commit(mats)
    ref = oracle.Tree(mats)
    assert (root == ref.root).all()
    for g, r in zip(tree.layers(), ref.layers()):
        assert (g == r).all()
    max_h = max(r for r, _ in shapes)
    for index in sorted({0, max_h - 1, max_h // 2, (max_h * 5) // 7}):
        rows, sib = mmcs.open_batch(index, tree)
        rrows, rsib = ref.open_batch(index)
        for a, b in zip(rows, rrows):
            assert (a == b).all()
        assert (sib == rsib).all()
        assert oracle.verify_batch(root, shapes, index, rows, sib)


def test_mmcs_rejects_bad_shapes(ctx):
    mmcs = bf.MerkleTreeMmcs(ctx)
    with pytest.raises(bf.BfGpuError, match="power of two"):
        mmcs.commit([np.zeros((6, 2), np.uint32)])
    with pytest.raises(bf.BfGpuError):
        mmcs.commit([np.zeros((0, 2), np.uint32)])


def test_two_adicity_limit(ctx):
    """KoalaBear has two-adicity 24: with blowup 2 the tallest committable matrix has 2^23 rows."""
    with pytest.raises(bf.BfGpuError, match="two-adicity"):
        bf.TwoAdicFriPcs(ctx).commit([np.zeros((1 << 24, 1), np.uint32)])
    with pytest.raises(bf.BfGpuError, match="two-adicity"):
        bf.Radix2Dit(ctx).coset_lde_batch(np.zeros((1 << 23, 1), np.uint32), 2, 3)


def test_maximum_height_commit(ctx, oracle):
    """2^23 x 1: the largest matrix the PCS can commit to (LDE 2^24 rows); root equals the oracle's."""
    rng = np.random.default_rng(99)
    a = rng.integers(0, P, (1 << 23, 1), dtype=np.uint32)
    root, data = bf.TwoAdicFriPcs(ctx).commit([a])
    ref = oracle.PcsData([a])
    assert (root == ref.root).all()
    data.free()


def test_pcs_commit_matches_oracle(ctx, oracle):
    """Shapes of a small `CpuProver::commit` call: 8 chips, heights sorted tallest first."""
    rng = np.random.default_rng(77)
    shapes = [(4096, 31), (4096, 2), (2048, 41), (1024, 7), (512, 45), (64, 1), (16, 12), (16, 5)]
    evals = [rand_mat(rng, r, c) for r, c in shapes]
    pcs = bf.TwoAdicFriPcs(ctx)
    root, data = pcs.commit(evals)
    ref = oracle.PcsData(evals)
    assert (root == ref.root).all()
    for i in range(len(evals)):
        assert (pcs.get_evaluations_on_domain(data, i, bit_reversed_rows=True) == ref.ldes[i]).all()
    # get_evaluations_on_domain in natural order == coset_lde_batch in natural order
    assert (pcs.get_evaluations_on_domain(data, 3) == oracle.coset_lde_batch(evals[3], 1, 3)).all()
    for g, r in zip(data.tree.layers(), ref.tree.layers()):
        assert (g == r).all()
    data.free()


def test_pcs_commit_quotient_chunk_domains(ctx, oracle):
    """Quotient chunks are committed on shifted domains 3*w^i*H (reference prover.rs:391-411)."""
    rng = np.random.default_rng(78)
    w = pyref.two_adic_generator(9)
    evals = [rand_mat(rng, 256, 4), rand_mat(rng, 256, 4), rand_mat(rng, 64, 4)]
    shifts = [3, 3 * w % P, 3]
    pcs = bf.TwoAdicFriPcs(ctx)
    root, data = pcs.commit(evals, domain_shifts=shifts)
    ref = oracle.PcsData(evals, domain_shifts=shifts)
    assert (root == ref.root).all()
    data.free()


def test_pcs_commit_larger_property(ctx, oracle):
    """2^18 x 16: compare the root with the oracle and check linearity of the LDE (size-independent)."""
    rng = np.random.default_rng(79)
    a = rand_mat(rng, 1 << 18, 16)
    pcs = bf.TwoAdicFriPcs(ctx)
    root, data = pcs.commit([a])
    ref = oracle.PcsData([a])
    assert (root == ref.root).all()
    lde = pcs.get_evaluations_on_domain(data, 0, bit_reversed_rows=True)
    assert (lde == ref.ldes[0]).all()
    data.free()


@pytest.mark.parametrize("rows,cols", [(1 << 13, 200), (1 << 12, 256), (1 << 13, 129), (1 << 13, 180)])
def test_pipelined_host_commit_matches_oracle(ctx, oracle, rows, cols):
    """One tall host matrix takes the pipelined path (copy / LDE / incremental leaf sponge per block of columns,
    csrc/bfgpu.cu commit_host_pipelined, 64-column blocks); widths chosen so the last block is 8 columns, a split full
    block, a single column and a split partial block (52 = 32 + 20: a 4-word tail chunk)."""
    rng = np.random.default_rng(rows + cols)
    a = rand_mat(rng, rows, cols)
    pcs = bf.TwoAdicFriPcs(ctx)
    root, data = pcs.commit([a])
    ref = oracle.PcsData([a])
    assert (root == ref.root).all()
    assert (pcs.get_evaluations_on_domain(data, 0, bit_reversed_rows=True) == ref.ldes[0]).all()
    for g, r in zip(data.tree.layers(), ref.tree.layers()):
        assert (g == r).all()
    rows_o, sib = bf.MerkleTreeMmcs(ctx).open_batch(5, data.tree)
    rrows, rsib = ref.tree.open_batch(5)
    assert (rows_o[0] == rrows[0]).all() and (sib == rsib).all()
    data.free()
