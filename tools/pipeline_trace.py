#!/usr/bin/env python3
"""tools/pipeline_trace.py — timestamps of the interpreter thread and the GPU thread inside CudaProver.prove_many."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkvm_brainfuck_b200 as bf
code = "++++++++[>-[>-[>+>+<<-]<-]<-]"
ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
rec = prover.execute(code)
pk = prover.setup_record(rec)
rec.free()
for _ in prover.prove_many([(code, [])] * 4, pk_for=lambda c: pk):
    pass
log = []
orig = prover.execute
def traced(code, stdin=()):
    t0 = time.perf_counter(); r = orig(code, stdin); log.append(("exec", threading.current_thread().name, t0, time.perf_counter())); return r
prover.execute = traced
t00 = time.perf_counter()
last = t00
for _ in prover.prove_many([(code, [])] * 6, pk_for=lambda c: pk):
    now = time.perf_counter(); log.append(("proof", "main", last, now)); last = now
for kind, th, a, b in sorted(log, key=lambda x: x[2]):
    print(f"{kind:6s} {th:12s} start {1e3*(a-t00):8.1f} ms  end {1e3*(b-t00):8.1f} ms  ({1e3*(b-a):.1f} ms)")
