#!/usr/bin/env python3
"""tests/tools/prove_once.py <program> — setup + 2 proofs of one program (for ncu launch lists)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import zkvm_brainfuck_b200 as bf
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
PROGRAMS = {"fibo": (open(os.path.join(ROOT, "tests/golden/fibo.bf")).read(), [17]), "hello": (open(os.path.join(ROOT, "tests/golden/hello.bf")).read(), []),
            "loop20": ("-[>-[>+>+>+<<<-]<-]", []), "loop22": ("++++++++[>-[>-[>+>+<<-]<-]<-]", [])}
code, stdin = PROGRAMS[sys.argv[1]]
ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
prog = ex.Program(code)
rec = ex.execute(prog, stdin)
traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
traces = {k: ctx.pinned_copy(v) for k, v in traces.items()}
pk = prover.setup(preps)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    proof = prover.prove(pk, traces, bf.Challenger(ctx))
print("ok", proof["opening_proof"]["pow_witness"], ctx.launch_count)
