"""CPU tests: pin the C oracle's DFT / coset LDE / MerkleTreeMmcs / Pcs::commit against the
O(n^2) definitions and an independent Python tree builder (no reference golden vectors exist:
PARITY UNPINNED)."""
import numpy as np
import pytest

from tests import pyref

P = pyref.P


def rand_mat(rng, rows, cols):
    return rng.integers(0, P, (rows, cols), dtype=np.uint32)


@pytest.mark.parametrize("log_n,cols,shift", [(0, 3, 3), (1, 2, 3), (2, 1, 3), (3, 5, 3), (4, 4, 7), (6, 3, 3)])
def test_lde_fast_vs_naive(oracle, log_n, cols, shift):
    rng = np.random.default_rng(log_n * 31 + cols)
    m = rand_mat(rng, 1 << log_n, cols)
    a = oracle.coset_lde_naive(m, 1, shift)
    b = oracle.coset_lde_batch(m, 1, shift)
    assert (a == b).all()


def test_lde_vs_python_lagrange(oracle):
    rng = np.random.default_rng(7)
    n, cols = 8, 2
    m = rand_mat(rng, n, cols)
    out = oracle.coset_lde_batch(m, 1, 3)
    W = pyref.two_adic_generator(4)
    for j in range(2 * n):
        pt = 3 * pow(W, j, P) % P
        for c in range(cols):
            assert int(out[j, c]) == pyref.interpolate_eval([int(v) for v in m[:, c]], pt)


def test_lde_bitrev_layout_and_coset_halves(oracle):
    """Stored LDE: first half = coset 3*H, second half = coset 3*w_{2n}*H, each bit-reversed."""
    rng = np.random.default_rng(8)
    n, cols = 16, 3
    m = rand_mat(rng, n, cols)
    nat = oracle.coset_lde_batch(m, 1, 3)
    br = oracle.coset_lde_batch_bitrev(m, 1, 3)
    for i in range(2 * n):
        assert (br[pyref.bitrev(i, 5)] == nat[i]).all()
    lo = oracle.coset_lde_batch(m, 0, 3)  # evaluations on 3*H
    for k in range(n):
        assert (br[pyref.bitrev(k, 4)] == lo[k]).all()
    hi = oracle.coset_lde_batch(m, 0, 3 * pyref.two_adic_generator(5) % P)
    for k in range(n):
        assert (br[n + pyref.bitrev(k, 4)] == hi[k]).all()


def test_dft_roundtrip_and_linearity(oracle):
    rng = np.random.default_rng(9)
    a, b = rand_mat(rng, 64, 4), rand_mat(rng, 64, 4)
    assert (oracle.idft_batch(oracle.dft_batch(a)) == a).all()
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    fa, fb, fs = oracle.dft_batch(a), oracle.dft_batch(b), oracle.dft_batch(s)
    assert (((fa.astype(np.uint64) + fb) % P).astype(np.uint32) == fs).all()
    # low-degree-ness: LDE of a degree<n poly, re-interpolated over 2n points, has zero top coeffs
    lde = oracle.coset_lde_batch(a, 1, 1)
    co = oracle.idft_batch(lde)
    assert (co[64:] == 0).all()


def py_tree(mats):
    """Independent MerkleTreeMmcs builder on Python ints."""
    order = sorted(range(len(mats)), key=lambda i: -mats[i].shape[0])  # stable
    hmax = mats[order[0]].shape[0]
    tall = [i for i in order if mats[i].shape[0] == hmax]
    layer = [pyref.sponge(np.concatenate([mats[i][r] for i in tall])) for r in range(hmax)]
    layers = [layer]
    while len(layer) > 1:
        n = len(layer) // 2
        inj = [i for i in order if mats[i].shape[0] == n]
        nxt = []
        for i in range(n):
            d = pyref.compress(layer[2 * i], layer[2 * i + 1])
            if inj:
                d = pyref.compress(d, pyref.sponge(np.concatenate([mats[k][i] for k in inj])))
            nxt.append(d)
        layer = nxt
        layers.append(layer)
    return layers


@pytest.mark.parametrize("shapes", [[(8, 3)], [(8, 3), (8, 9)], [(4, 2), (16, 5), (8, 1), (16, 8), (1, 4)], [(1, 5)], [(2, 17), (2, 1)]])
def test_merkle_matches_python(oracle, shapes):
    rng = np.random.default_rng(len(shapes))
    mats = [rand_mat(rng, r, c) for r, c in shapes]
    t = oracle.Tree(mats)
    ref = py_tree(mats)
    got = t.layers()
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert g.tolist() == r
    assert t.root.tolist() == ref[-1][0]


def test_merkle_open_verify_roundtrip(oracle):
    rng = np.random.default_rng(11)
    shapes = [(32, 7), (8, 2), (32, 4), (16, 12), (2, 3)]
    mats = [rand_mat(rng, r, c) for r, c in shapes]
    t = oracle.Tree(mats)
    for index in [0, 1, 13, 31]:
        rows, sib = t.open_batch(index)
        for m, row in zip(mats, rows):
            assert (row == m[index >> (5 - int(np.log2(m.shape[0])))]).all()
        assert sib.shape == (5, 8)
        assert oracle.verify_batch(t.root, shapes, index, rows, sib)
        bad = [r.copy() for r in rows]
        bad[3][5] ^= 1
        assert not oracle.verify_batch(t.root, shapes, index, bad, sib)
        sib2 = sib.copy()
        sib2[2, 0] = (int(sib2[2, 0]) + 1) % P
        assert not oracle.verify_batch(t.root, shapes, index, rows, sib2)
        assert not oracle.verify_batch(t.root, shapes, index ^ 1, rows, sib)


def test_pcs_commit_is_lde_then_merkle(oracle):
    rng = np.random.default_rng(12)
    evals = [rand_mat(rng, 16, 5), rand_mat(rng, 4, 2), rand_mat(rng, 16, 1)]
    d = oracle.PcsData(evals)
    for e, l in zip(evals, d.ldes):
        assert (l == oracle.coset_lde_batch_bitrev(e, 1, 3)).all()
    t = oracle.Tree([np.array(l) for l in d.ldes])
    assert (t.root == d.root).all()
    # quotient-chunk style domain shift: evaluations given on the coset 3*w*H are extended with
    # shift GENERATOR/domain_shift, so the stored LDE is the same polynomial on 3*K
    w = pyref.two_adic_generator(5)
    shift = 3 * w % P
    d2 = oracle.PcsData([evals[0]], domain_shifts=[shift])
    assert (d2.ldes[0] == oracle.coset_lde_batch_bitrev(evals[0], 1, pow(w, -1, P))).all()
