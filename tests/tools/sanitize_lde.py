#!/usr/bin/env python3
"""tests/tools/sanitize_lde.py — small coset LDEs and one small proof for compute-sanitizer runs (memcheck / racecheck / synccheck):
every instantiation family of the TMA passes (fused ingest, INV, TURN, FWD), the cluster tree top and the prover kernels, at
sizes a sanitizer finishes in minutes; results are compared with the CPU oracle so a silent corruption shows as well."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
import zkvm_brainfuck_b200 as bf

ctx = bf.Context(0)
dft = bf.Radix2Dit(ctx)
for log_n, cols in [(12, 8), (13, 36), (16, 4), (17, 4)]:
    m = np.random.default_rng(log_n).integers(0, bf.P, (1 << log_n, cols), dtype=np.uint32)
    ok = (dft.coset_lde_batch(m, 1, 3, bit_reversed_rows=True) == oracle.coset_lde_batch_bitrev(m, 1, 3)).all()
    print("lde", log_n, cols, "ok" if ok else "MISMATCH", flush=True)
    assert ok
pcs = bf.TwoAdicFriPcs(ctx)
evals = [np.random.default_rng(5).integers(0, bf.P, s, dtype=np.uint32) for s in [(4096, 9), (1024, 3), (64, 5)]]
root, data = pcs.commit(evals)
assert (root == oracle.PcsData(evals).root).all()
print("commit ok", flush=True)
data.free()
if "--prove" in sys.argv:
    prover = bf.CudaProver(ctx)
    (words, _), rec = prover.prove_program("++[>+<-]>,.", [42], raw=True)
    print("proof words", len(words), flush=True)
ctx.close()
