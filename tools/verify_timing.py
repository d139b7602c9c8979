#!/usr/bin/env python3
"""tools/verify_timing.py — program -> proof on the GPU, then the native host verifier on the serialised proof."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkvm_brainfuck_b200 as bf
ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
for name, code in (("loop20", "-[>-[>+>+>+<<<-]<-]"), ("loop22", "++++++++[>-[>-[>+>+<<-]<-]<-]")):
    (words, _), rec = prover.prove_program(code, raw=True)
    pk = prover.setup_record(rec)
    t = time.perf_counter()
    r = bf.verify_shard(pk.commit, pk.names, pk.heights, words)
    print(name, "cycles", rec.cycles, "proof bytes", len(words) * 4, "native verify:", r, round((time.perf_counter() - t) * 1e3, 1), "ms")
