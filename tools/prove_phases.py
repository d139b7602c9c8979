#!/usr/bin/env python3
"""tools/prove_phases.py <program> [n] — program -> proof n times, then once more with the library's per-phase CUDA-event
timers; prints the best wall time and the phase table (for variant comparisons: BFGPU_SO=... selects the library)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkvm_brainfuck_b200 as bf
PROGRAMS = {"fibo": (open(os.path.join(ROOT, "tests/golden/fibo.bf")).read(), [17]), "hello": (open(os.path.join(ROOT, "tests/golden/hello.bf")).read(), []),
            "loop20": ("-[>-[>+>+>+<<<-]<-]", []), "loop22": ("++++++++[>-[>-[>+>+<<-]<-]<-]", [])}
code, stdin = PROGRAMS[sys.argv[1]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = bf.Context(0)
prover = bf.CudaProver(ctx)
_rec0 = prover.execute(code, stdin)
pk = prover.setup_record(_rec0)  # the proving key is an input of prove, as in the reference
_rec0.free()
times, etimes, words = [], [], None
for _ in range(n):
    ctx.synchronize()
    t0 = time.perf_counter()
    (buf, decode), rec = prover.prove_program(code, stdin, pk=pk, raw=True)
    rec.free()
    times.append((time.perf_counter() - t0) * 1e3)
    words = buf
ctx.profile_enable(True)
prover.prove_program(code, stdin, pk=pk, raw=True)
ph = {k: round(v[0], 3) for k, v in ctx.profile_read().items() if v[0] or v[1]}
ctx.profile_enable(False)
import hashlib
print(json.dumps({"program": sys.argv[1], "best_ms": round(min(times), 3), "all_ms": [round(t, 2) for t in times], "gpu_phase_sum_ms": round(sum(ph.values()), 3),
                  "phases_ms": ph, "proof_sha256": hashlib.sha256(words.tobytes()).hexdigest()[:16], "lib": os.environ.get("BFGPU_SO", "default")}))
