#!/usr/bin/env python3
"""tools/leak_check.py — repeat every public path many times and report device / host memory drift."""
import os, sys, time, resource
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import zkvm_brainfuck_b200 as bf

def free_mb():
    f, t = torch.cuda.mem_get_info()
    return f / 2**20

ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
pcs = bf.TwoAdicFriPcs(ctx)
rng = np.random.default_rng(0)
mats = [rng.integers(0, bf.P, (1 << 12, 40), dtype=np.uint32), rng.integers(0, bf.P, (1 << 8, 9), dtype=np.uint32)]
big = rng.integers(0, bf.P, (1 << 14, 200), dtype=np.uint32)
fibo = open(os.path.join(ROOT, "tests/golden/fibo.bf")).read()
def round_():
    root, data = pcs.commit(mats); data.free()
    root, data = pcs.commit([big]); data.free()
    (w, _), rec = prover.prove_program(fibo, [17], raw=True)
    assert bf.verify_shard(*vk, w) is None
    for _ in prover.prove_many([("-[>-[>+>+>+<<<-]<-]", [])] * 2, pk_for=lambda c: pk20):
        pass
rec = prover.execute(fibo, [17]); pkf = prover.setup_record(rec); vk = (pkf.commit, pkf.names, pkf.heights)
rec20 = prover.execute("-[>-[>+>+>+<<<-]<-]"); pk20 = prover.setup_record(rec20)
for i in range(3):
    round_()
ctx.synchronize()
base_dev, base_host = free_mb(), resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024
for i in range(30):
    round_()
    if i % 10 == 9:
        ctx.synchronize()
        print(f"iter {i+1}: device free {free_mb():.0f} MiB (drift {base_dev - free_mb():+.0f}), host maxrss {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss/1024:.0f} MiB (drift {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss/1024 - base_host:+.0f})")
