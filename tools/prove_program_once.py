#!/usr/bin/env python3
"""tools/prove_program_once.py <program> [n] — program -> proof through the native executor and the device-side trace
generators, n times (for ncu launch lists of the whole path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkvm_brainfuck_b200 as bf
PROGRAMS = {"fibo": (open(os.path.join(ROOT, "tests/golden/fibo.bf")).read(), [17]), "hello": (open(os.path.join(ROOT, "tests/golden/hello.bf")).read(), []),
            "loop20": ("-[>-[>+>+>+<<<-]<-]", []), "loop22": ("++++++++[>-[>-[>+>+<<-]<-]<-]", [])}
code, stdin = PROGRAMS[sys.argv[1]]
ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    l0 = ctx.launch_count
    (buf, decode), rec = prover.prove_program(code, stdin, raw=True)
    print("ok", rec.cycles, rec.output[:8], len(buf), "launches", ctx.launch_count - l0)
