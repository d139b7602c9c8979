#!/usr/bin/env python3
"""tests/tools/prove_bench.py — end-to-end shard-prove timings on the GPU (BASELINE.json configs 1-3), with the
per-phase device times recorded by the library.  Usage: python tests/tools/prove_bench.py [fibo hello loop20 loop22]"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import zkvm_brainfuck_b200 as bf

ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
GOLD = os.path.join(ROOT, "tests", "golden")
PROGRAMS = {
    "fibo": (open(os.path.join(GOLD, "fibo.bf")).read(), [17]),      # BASELINE config 1 (test_e2e_core)
    "hello": (open(os.path.join(GOLD, "hello.bf")).read(), []),       # BASELINE config 2 (examples/hello)
    "loop20": ("-[>-[>+>+>+<<<-]<-]", []),                             # ~2^20-row Cpu trace
    "loop22": ("++++++++[>-[>-[>+>+<<-]<-]<-]", []),                   # 2^22-row Cpu trace (north-star size)
}


def main():
    which = sys.argv[1:] or ["hello", "fibo", "loop20"]
    ctx = bf.Context(0)
    prover = bf.CudaProver(ctx)
    for name in which:
        code, stdin = PROGRAMS[name]
        t0 = time.perf_counter()
        prog = ex.Program(code)
        rec = ex.execute(prog, stdin)
        t_exec = time.perf_counter() - t0
        t0 = time.perf_counter()
        traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
        t_trace = time.perf_counter() - t0
        traces = {k: ctx.pinned_copy(v) for k, v in traces.items()}  # trace generators write into page-locked memory
        cells = sum(v.size for v in traces.values())
        pk = prover.setup(preps)
        times = []
        for it in range(4):
            if it == 3:
                ctx.profile_enable(True)
            ctx.synchronize()
            t0 = time.perf_counter()
            buf, decode = prover.prove(pk, traces, bf.Challenger(ctx), raw=True)  # serialised proof; decoding into Python objects is not timed
            ctx.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
        proof = decode()
        phases = {k: round(v[0], 3) for k, v in ctx.profile_read().items() if v[1] or v[0]}
        ctx.profile_enable(False)
        best = min(times[1:])
        out = dict(program=name, cycles=rec.cycles, cpu_rows=int(traces["Cpu"].shape[0]), main_cells=int(cells), exec_s=round(t_exec, 3),
                   tracegen_s=round(t_trace, 3), prove_ms=[round(t, 2) for t in times], best_prove_ms=round(best, 2),
                   cycles_per_s=rec.cycles / (best * 1e-3), rows_per_s=float(traces["Cpu"].shape[0]) / (best * 1e-3), phases_ms_last=phases,
                   pow_witness=proof["opening_proof"]["pow_witness"])
        # program -> proof: native executor + device-side trace generation (no host traces at all)
        ptimes, etimes = [], []
        for it in range(4):
            if it == 3:
                ctx.profile_enable(True)
            ctx.synchronize()
            t0 = time.perf_counter()
            rec2 = prover.execute(code, stdin)
            t1 = time.perf_counter()
            ch = bf.Challenger(ctx)
            bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
            shard = prover.commit_record(rec2)
            buf2 = prover.open_raw(pk, shard, ch.clone())
            ctx.synchronize()
            t2 = time.perf_counter()
            shard.free()
            rec2.free()
            etimes.append((t1 - t0) * 1e3)
            ptimes.append((t2 - t0) * 1e3)
        ph2 = {k: round(v[0], 3) for k, v in ctx.profile_read().items() if v[1] or v[0]}
        ctx.profile_enable(False)
        out.update(program_to_proof_ms=[round(t, 2) for t in ptimes], best_program_to_proof_ms=round(min(ptimes[1:]), 2),
                   native_exec_ms=round(min(etimes[1:]), 2), same_proof=bool((buf2 == buf).all()), phases_ms_program=ph2)
        print(json.dumps(out), flush=True)
        pk.free()
        ctx.free_pinned()
    ctx.close()


if __name__ == "__main__":
    main()
