"""Run-to-run determinism of the asynchronous data paths (GPU): the TMA-fed NTT passes keep a ring of shared-memory slots whose
reuse is ordered by barriers and `cp.async.bulk.wait_group`, the tree top runs on a thread-block cluster — a hazard in either shows
as a difference between repetitions of the same call (compute-sanitizer is not available on the GPU pool, so this is the race test).
The first repetition is also checked against the CPU oracle."""
import hashlib

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
P = bf.P


@pytest.mark.parametrize("log_n,cols", [(13, 36), (16, 32), (17, 8), (18, 64)])
def test_lde_and_commit_repeat_bit_identically(oracle, log_n, cols):
    ctx = bf.Context()
    try:
        m = np.random.default_rng(log_n).integers(0, P, (1 << log_n, cols), dtype=np.uint32)
        dft, pcs = bf.Radix2Dit(ctx), bf.TwoAdicFriPcs(ctx)
        first = dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True)
        assert (first == oracle.coset_lde_batch_bitrev(m, 1, 3)).all()
        root0, d0 = pcs.commit([m])
        d0.free()
        for _ in range(6):
            assert (dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True) == first).all()
            root, d = pcs.commit([m])
            d.free()
            assert (root == root0).all()
    finally:
        ctx.close()


def test_proofs_repeat_bit_identically():
    ctx = bf.Context()
    try:
        prover = bf.CudaProver(ctx)
        seen = {}
        for r in range(6):
            code = "-[>-[>+>+>+<<<-]<-]" if r % 2 else "++[>+<-]>,."
            (words, _), rec = prover.prove_program(code, [] if r % 2 else [42], raw=True)
            h = hashlib.sha256(words.tobytes()).hexdigest()
            assert seen.setdefault(code, h) == h
            rec.free()
    finally:
        ctx.close()
