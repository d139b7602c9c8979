"""The tuned CPU baseline (oracle/fast_commit.c: AVX-512 Montgomery, packed Poseidon2, cache-blocked row NTT — what
bench.py times as the CPU arm) against the slow restatement it must agree with (oracle/dft.c, merkle.c, poseidon2.c,
the checker of every GPU parity test): same permutation outputs, same LDE words, same Merkle root."""
import numpy as np
import pytest

P = 2130706433


@pytest.fixture(scope="module")
def fast(oracle):
    if not oracle.fast_available():
        pytest.skip("host CPU has no AVX-512")
    return oracle


def test_packed_permutation_equals_scalar(fast):
    rng = np.random.default_rng(11)
    st = rng.integers(0, P, (64, 16), dtype=np.uint32)
    st[0] = 0
    st[1] = P - 1
    assert (fast.fast_permute_many(st) == fast.permute_many(st)).all()


@pytest.mark.parametrize("rows,cols", [(16, 16), (16, 1), (32, 8), (64, 24), (256, 5), (1024, 40), (4096, 33), (1 << 13, 256), (1 << 15, 17)])
def test_fast_commit_equals_oracle(fast, rows, cols):
    rng = np.random.default_rng(rows * 31 + cols)
    m = rng.integers(0, P, (rows, cols), dtype=np.uint32)
    ref = fast.PcsData([m])
    root, lde, ph = fast.fast_pcs_commit(m, want_lde=True)
    assert (lde == ref.ldes[0]).all(), "LDE words differ"
    assert (root == ref.root).all(), "Merkle root differs"
    assert all(v >= 0 for v in ph.values())


def test_fast_commit_edge_values_and_refusals(fast):
    m = np.full((64, 16), P - 1, np.uint32)
    m[::3] = 0
    assert (fast.fast_pcs_commit(m)[0] == fast.PcsData([m]).root).all()
    with pytest.raises(RuntimeError):
        fast.fast_pcs_commit(np.zeros((8, 4), np.uint32))   # below one packed leaf group
    with pytest.raises(RuntimeError):
        fast.fast_pcs_commit(np.zeros((48, 4), np.uint32))  # not a power of two
