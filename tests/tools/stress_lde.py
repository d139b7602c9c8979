#!/usr/bin/env python3
"""tests/tools/stress_lde.py [reps] — repeat coset LDEs / commits of several shapes and compare every repetition bit for bit with the
first one (and the first one with the CPU oracle where it is small enough): a race in the shared-memory slot protocol of the TMA
passes or in the cluster tree top would show up as a run-to-run difference.  (compute-sanitizer is not available on the GPU pool.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
import zkvm_brainfuck_b200 as bf

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = bf.Context(0)
dft = bf.Radix2Dit(ctx)
pcs = bf.TwoAdicFriPcs(ctx)
bad = 0
for log_n, cols in [(13, 36), (16, 32), (17, 8), (18, 64), (20, 96), (21, 33)]:
    m = np.random.default_rng(log_n).integers(0, bf.P, (1 << log_n, cols), dtype=np.uint32)
    first = dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True)
    if log_n <= 18:
        assert (first == oracle.coset_lde_batch_bitrev(m, 1, 3)).all(), ("oracle", log_n, cols)
    for r in range(reps):
        out = dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True)
        if not (out == first).all():
            bad += 1
            print("LDE MISMATCH", log_n, cols, "rep", r, int((out != first).sum()), "words", flush=True)
    root0, d0 = pcs.commit([m])
    d0.free()
    for r in range(reps):
        root, d = pcs.commit([m])
        d.free()
        if not (root == root0).all():
            bad += 1
            print("ROOT MISMATCH", log_n, cols, "rep", r, flush=True)
    print("shape", log_n, cols, "done", flush=True)
prover = bf.CudaProver(ctx)
import hashlib
h0 = None
for r in range(reps):
    (words, _), rec = prover.prove_program("++++++++[>-[>-[>+>+<<-]<-]<-]" if r % 2 else "-[>-[>+>+>+<<<-]<-]", [], raw=True)
    h = hashlib.sha256(words.tobytes()).hexdigest()
    if r < 2:
        h0 = (h0 or []) + [h]
    elif h != h0[r % 2]:
        bad += 1
        print("PROOF MISMATCH rep", r, flush=True)
print("stress done, mismatches:", bad)
sys.exit(1 if bad else 0)
