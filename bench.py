#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200 proving backend (driver contract: one JSON line).

Workload (BASELINE.json configs[3], the largest single-GPU configuration of the commit hot path):
`Pcs::commit` = coset LDE (blowup 2, shift 3, bit-reversed rows) + Poseidon2 MerkleTreeMmcs commit of
a synthetic 2^22 x 256 trace of uniform KoalaBear values.  One "step" = one full commit.

  value : algorithmic LDE bytes (12*R*W) per second, whole job over all ranks, input
          already resident in HBM (row-major, canonical u32), device-timed with CUDA events on the
          stream the kernels run on.
  e2e   : the same call through the C ABI with a pinned HOST matrix (H2D copy inside the timed
          region, root read back to the host).
  roofline      : dominant kernel (Poseidon2 leaf sponge) against the integer-issue peak measured
                  live by a register-only probe; `roofline_hbm`: the LDE kernels against the HBM peak.
  cpu_baseline  : the CPU oracle (`oracle/`, scalar C + OpenMP port of the same algorithm) on a
                  bounded sample; `--impl reference` runs only that arm.

N > 1 (torchrun): every rank commits its own trace (chips / trace matrices shard across GPUs with
no data-path collective), weak scaling, max-over-ranks device time.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P = 2130706433
# SASS thread-instructions per element of the NTT passes of a 2^22-point column (pass plan g = 8, 7, 7): column-loop bodies of
# k_pass<0,3,0,0> x2 + k_pass<0,4,1,0> (584 + 584 + 472) / 16 forward, k_pass<1,4,1,0> + k_pass<1,3,0,0> + k_pass<1,3,0,1> (517 + 641 + 857) / 16
# inverse incl. the fused coset epilogue (cuobjdump -sass, profiles/r1_ntt_instr_counts.txt)
NTT_INSTR_FWD_2P22 = 102.5
NTT_INSTR_INV_2P22 = 125.9
# executed thread-instructions per Poseidon2 permutation in k_leaf_hash: ncu smsp__inst_executed.sum * 32 / permutations
# = 38.01e9 * 32 / 268 435 456 (profiles/r1_leaf_hash_final.md)
P2_INSTR_PER_PERM = 4531
# DRAM bytes of one k_leaf_hash launch at the default workload, from the same ncu --set full capture
LEAF_TRAFFIC_2P22X256 = 8600113000 + 269837312


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-rows", type=int, default=22)
    ap.add_argument("--cols", type=int, default=256)
    ap.add_argument("--cpu-log-rows", type=int, default=16, help="bounded CPU sample height (log2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dist-commit", action="store_true", help="N>1: skip the one-commitment-over-all-ranks measurements")
    ap.add_argument("--no-prove", action="store_true", help="skip the end-to-end shard-prove timings (BASELINE configs 1-3)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(log_rows, cols):
    return f"pcs_commit(coset_lde blowup2 shift3 + poseidon2 merkle) 2^{log_rows}x{cols} KoalaBear"


def algorithmic_bytes(rows, cols):
    return 12 * rows * cols  # read R*W u32 once + write 2R*W u32 once


def num_perms(rows, cols):
    leaves = 2 * rows
    return leaves * ((cols + 7) // 8) + (leaves - 1)


def cpu_commit_sample(log_rows, cols, steps, warmup):
    """Time the CPU oracle's Pcs::commit on a 2^log_rows x cols sample -> (GB/s, seconds/step, threads)."""
    import numpy as np
    import oracle

    # every host thread, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for its workers)
    oracle.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    rng = np.random.default_rng(0xB200)
    m = rng.integers(0, P, (1 << log_rows, cols), dtype=np.uint32)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        d = oracle.PcsData([m])
        dt = time.perf_counter() - t
        del d
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return algorithmic_bytes(1 << log_rows, cols) / sec / 1e9, sec, oracle.get_threads()


def prove_timings(ctx, bf, with_cpu):
    """End-to-end shard proofs (MachineProver::prove minus trace generation) of BASELINE configs 1-3 on this GPU:
    host traces in, proof out, wall-clock of the public API call, best of 3 after one warm-up."""
    import importlib
    import numpy as np
    gold = os.path.join(ROOT, "tests", "golden")
    progs = {"fibo_stdin17 (config 1, test_e2e_core)": (open(os.path.join(gold, "fibo.bf")).read(), [17]),
             "hello (config 2)": (open(os.path.join(gold, "hello.bf")).read(), []),
             "loop 2^20 Cpu rows (config 3)": ("-[>-[>+>+>+<<<-]<-]", []),
             "loop 2^22 Cpu rows (north-star size)": ("++++++++[>-[>-[>+>+<<-]<-]<-]", [])}
    prover = bf.CudaProver(ctx)
    out = {}
    for name, (code, stdin) in progs.items():
        # host traces for the "traces in -> proof out" timing come from the product's own generators (native executor + device
        # trace generation, copied back once); nothing here touches the CPU oracle
        rec = prover.execute(code, stdin)
        pk = prover.setup_record(rec)
        sh0 = prover.commit_record(rec)
        traces = {k: ctx.pinned_copy(v) for k, v in prover.shard_traces(sh0).items()}  # page-locked, as a trace generator would write them
        sh0.free()
        times = []
        for _ in range(4):
            ctx.synchronize()
            t0 = time.perf_counter()
            buf, decode = prover.prove(pk, traces, bf.Challenger(ctx), raw=True)  # serialised proof; decoding into Python objects is not timed
            times.append((time.perf_counter() - t0) * 1e3)
        proof = decode()
        best = min(times[1:])
        entry = {"cycles": rec.cycles, "cpu_rows": int(traces["Cpu"].shape[0]), "committed_main_cells": int(sum(v.size for v in traces.values())),
                 "prove_ms": best, "trace_rows_per_s": float(traces["Cpu"].shape[0]) / (best * 1e-3), "khz": rec.cycles / best,
                 "main_root": [int(x) for x in proof["commitment"]["main"]]}
        # ProverClient::prove end to end on this backend: native executor -> 16 B/cycle records -> device-side trace
        # generation -> commit -> open (setup excluded, as in the reference where the pk is an input of prove)
        ptimes, etimes = [], []
        for _ in range(4):
            ctx.synchronize()
            t0 = time.perf_counter()
            nrec = prover.execute(code, stdin)
            t1 = time.perf_counter()
            ch = bf.Challenger(ctx)
            bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
            shard = prover.commit_record(nrec)
            buf2 = prover.open_raw(pk, shard, ch.clone())
            t2 = time.perf_counter()
            shard.free()
            nrec.free()
            etimes.append((t1 - t0) * 1e3)
            ptimes.append((t2 - t0) * 1e3)
        # one more run with the library's per-phase CUDA-event timers: HBM rooflines of the prover-side kernels from their
        # ALGORITHMIC bytes: quotient reads prep+main+perm LDE rows once and writes 16 B per coset point; the
        # reduced openings read every committed LDE once; the barycentric evaluation reads the low coset (half) of every LDE
        ctx.profile_enable(True)
        nrec = prover.execute(code, stdin)
        ch = bf.Challenger(ctx)
        bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
        shard = prover.commit_record(nrec)
        prover.open_raw(pk, shard, ch.clone())
        ph = {k: v[0] for k, v in ctx.profile_read().items() if v[0] or v[1]}
        ctx.profile_enable(False)
        info = {c[0]: c for c in prover.chips}
        lde_cells = quot_bytes = 0
        for nm, h in zip(shard.names, shard.heights):
            _, mw, pw, ew, _lo = info[nm]
            lde_cells += 2 * h * (pw + mw + 4 * ew + 8)             # prep + main + perm + two 4-column quotient chunks (each 2h x 4)
            quot_bytes += 4 * 2 * h * (pw + mw + 4 * ew) + 16 * 2 * h
        shard.free()
        nrec.free()
        hbm = 6451.5
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", hbm)
        except OSError:
            pass

        def roof(bytes_, ms):
            return {"algorithmic_bytes": int(bytes_), "ms": ms, "GB/s": bytes_ / (ms * 1e-3) / 1e9, "frac_hbm": bytes_ / (ms * 1e-3) / 1e9 / hbm}

        entry["phases_ms"] = ph
        entry["hbm_rooflines"] = {"quotient": roof(quot_bytes, ph["quotient"]), "open_reduce": roof(4 * lde_cells, ph["open_reduce"]),
                                  "open_eval": roof(2 * lde_cells, ph["open_eval"]),
                                  "note": "all three are arithmetic-bound on F_p^4 products (profiles/r1_prover_kernels.md), not HBM-bound"}
        if "2^22" in name:
            # throughput of a stream of proofs: the interpreter of proof k+1 runs on a host thread while the GPU proves proof k
            for _ in prover.prove_many([(code, stdin)] * 4, pk_for=lambda _c: pk):  # warm-up: fills the pool of page-locked record buffers
                pass
            njobs, stamps = 5, []
            for _ in prover.prove_many([(code, stdin)] * njobs, pk_for=lambda _c: pk):
                stamps.append(time.perf_counter())
            entry["pipelined_ms_per_proof"] = (stamps[-1] - stamps[0]) / (njobs - 1) * 1e3
            entry["pipelined_trace_rows_per_s"] = float(traces["Cpu"].shape[0]) / ((stamps[-1] - stamps[0]) / (njobs - 1))
        entry.update({"program_to_proof_ms": min(ptimes[1:]), "native_executor_ms": min(etimes[1:]),
                      "program_proof_equals_trace_proof": bool(buf2.shape == buf.shape and (buf2 == buf).all()),
                      "program_to_proof_khz": rec.cycles / min(ptimes[1:])})
        if with_cpu and name.startswith("hello"):
            # CPU leg (cpu_baseline side of the bench): the oracle proves the same statement from ITS OWN executor and trace
            # generators (numpy + C, single process): parity check + a rough CPU figure
            from oracle import prover as PR, stark as S
            oex = importlib.import_module("oracle.machine.executor")
            otg = importlib.import_module("oracle.machine.tracegen")
            chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
            t0 = time.perf_counter()
            oprog = oex.Program(code)
            otraces, preps = otg.generate_traces(oex.execute(oprog, stdin)), otg.preprocessed_traces(oprog)
            opk = PR.setup(chips, preps)
            och = S.Challenger()
            PR.observe_pk(opk, och)
            ref = PR.prove_shard(chips, opk, otraces, och.clone())
            entry["cpu_oracle_execute_tracegen_prove_ms"] = (time.perf_counter() - t0) * 1e3
            entry["proof_matches_cpu_oracle"] = bool(all((np.asarray(proof["commitment"][k]) == ref["commitment"][k]).all() for k in ("main", "permutation", "quotient"))
                                                     and (np.asarray(proof["opening_proof"]["final_poly"]) == ref["opening_proof"]["final_poly"]).all()
                                                     and proof["opening_proof"]["pow_witness"] == ref["opening_proof"]["pow_witness"])
        out[name] = entry
        pk.free()
        ctx.free_pinned()
    return out


def run_reference(args, rank):
    """Reference arm: the reference's CPU algorithm (oracle port; the Rust prover cannot be built here)
    on all host threads, bounded sample of the same workload."""
    if rank != 0:
        return
    gbs, sec, threads = cpu_commit_sample(args.cpu_log_rows, args.cols, args.steps, args.warmup)
    sample = f"2^{args.cpu_log_rows}x{args.cols} sample of the workload per step (full size would be {1 << (args.log_rows - args.cpu_log_rows)}x longer)"
    line = {
        "impl": "reference",
        "metric": "lde_poseidon2_commit_throughput", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (KoalaBear mod p)",
        "data": "synthetic",
        "config": {"workload": workload_name(args.log_rows, args.cols), "sample": sample},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import ctypes as C
    import numpy as np
    import torch
    import zkvm_brainfuck_b200 as bf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    R, W = 1 << args.log_rows, args.cols
    ctx = bf.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    lib = bf.lib()

    # synthetic trace, resident in HBM: row-major canonical u32 (the layout RowMajorMatrix<KoalaBear> has)
    g = torch.Generator(device="cuda")
    g.manual_seed(0xB200 + rank)
    trace = torch.randint(0, P, (R, W), dtype=torch.int32, device="cuda", generator=g)
    torch.cuda.synchronize()
    mat = bf.Mat(trace.data_ptr(), R, W)
    root = np.zeros(8, np.uint32)
    root_p = root.ctypes.data_as(C.POINTER(C.c_uint32))

    def commit_once():
        h = C.c_void_p()
        ctx.check(lib.bfgpu_pcs_commit(ctx._h, C.byref(mat), None, 1, root_p, C.byref(h)))
        lib.bfgpu_pcs_data_free(h)

    ctx.set_input_space(bf.MEM_DEVICE)
    for _ in range(max(args.warmup, 3)):
        commit_once()
    int32_peak = ctx.int32_peak_probe()  # Ginstr/s (thread-level integer instructions)

    # ---- timed region: K commits, inputs resident in HBM ---------------------------------------
    barrier()
    sampler = ClockSampler(local_rank)
    ctx.profile_enable(True)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            commit_once()
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    phases = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop()
    root_dev = root.copy()
    from importlib import import_module
    shard = import_module("zkvm-brainfuck_b200.shard")
    ms_total = shard.max_over_ranks(ms_total, dist, "cuda")
    ms_step = ms_total / args.steps
    value = world * algorithmic_bytes(R, W) / (ms_step * 1e-3) / 1e9

    # ---- e2e: same call with a pinned HOST matrix ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host = torch.empty((R, W), dtype=torch.int32, pin_memory=True)
        host.copy_(trace)
        torch.cuda.synchronize()
        hmat = bf.Mat(host.data_ptr(), R, W)
        ctx.set_input_space(bf.MEM_HOST)

        def commit_host():
            h = C.c_void_p()
            ctx.check(lib.bfgpu_pcs_commit(ctx._h, C.byref(hmat), None, 1, root_p, C.byref(h)))
            lib.bfgpu_pcs_data_free(h)

        commit_host()
        assert (root == root_dev).all(), "host-path root differs from device-path root"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            commit_host()
        ctx.synchronize()
        dt = time.perf_counter() - t0
        dt = shard.max_over_ranks(dt, dist, "cuda")
        e2e = {"value": world * algorithmic_bytes(R, W) / (dt / args.steps) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": 4 * R * W, "d2h_bytes_per_step": 32, "ms_per_step": dt / args.steps * 1e3}
        del host

    # ---- N > 1: one independent program -> proof per GPU (replica throughput of the whole prover) -----------
    replica_prove = None
    if dist is not None and not args.no_prove:
        code = "++++++++[>-[>-[>+>+<<-]<-]<-]"  # 4 173 897 cycles, Cpu trace 2^22 rows (north-star size)
        prover = bf.CudaProver(ctx)
        rec0 = prover.execute(code)
        pk = prover.setup_record(rec0)
        cycles = rec0.cycles
        rec0.free()

        def prove_once():
            nrec = prover.execute(code)
            ch = bf.Challenger(ctx)
            lib.bfgpu_pk_observe_into(pk._h, ch._h)
            sh = prover.commit_record(nrec)
            buf = prover.open_raw(pk, sh, ch.clone())
            sh.free()
            nrec.free()
            return buf

        first = prove_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prove_once()
        ctx.synchronize()
        dt = shard.max_over_ranks(time.perf_counter() - t0, dist, "cuda") / args.steps
        # stream of proofs per GPU: interpreter of proof k+1 on a host thread while the GPU proves proof k
        for _ in prover.prove_many([(code, [])] * 4, pk_for=lambda _c: pk):
            pass
        barrier()
        njobs, stamps = max(args.steps, 3) + 1, []
        for _ in prover.prove_many([(code, [])] * njobs, pk_for=lambda _c: pk):
            stamps.append(time.perf_counter())
        dtp = shard.max_over_ranks((stamps[-1] - stamps[0]) / (njobs - 1), dist, "cuda")
        replica_prove = {"workload": "program -> proof, 4 173 897 cycles (Cpu trace 2^22 rows), one independent proof per GPU",
                         "ms_per_proof": dt * 1e3, "proofs_per_s": world / dt, "trace_rows_per_s": world * (1 << 22) / dt,
                         "cycles_per_s": world * cycles / dt, "proof_words": int(first.size),
                         "pipelined_ms_per_proof": dtp * 1e3, "pipelined_proofs_per_s": world / dtp, "pipelined_trace_rows_per_s": world * (1 << 22) / dtp}
        pk.free()

    # ---- N > 1: ONE commitment over all ranks (columns -> LDE -> P2P row exchange -> subtrees -> caps) -------
    one_commitment = None
    if dist is not None and not args.no_dist_commit:
        ctx.set_input_space(bf.MEM_DEVICE)
        del trace
        torch.cuda.empty_cache()

        def dist_case(total_cols, exchange, nmats=1):
            """one commitment of `nmats` matrices (nmats = 8: BASELINE config 5, "2^24 rows" as 8 x (2^21 x 64) in ONE commit, the same
            number of cells as 2^22 x 256), every matrix column-sharded over the ranks"""
            rows = R if nmats == 1 else 4 * R // nmats
            c0, nloc = shard.col_range(total_cols, world, rank)
            gg = torch.Generator(device="cuda")
            gg.manual_seed(0xD157 + rank)
            mats = [torch.randint(0, P, (rows, nloc), dtype=torch.int32, device="cuda", generator=gg) for _ in range(nmats)]
            torch.cuda.synchronize()
            roots = []

            def step():
                dc = shard.DistributedCommit(ctx, dist, [rows] * nmats, [total_cols] * nmats, exchange=exchange)
                roots.append(dc.commit([(m.data_ptr(), rows, nloc) for m in mats]).copy())
                dc.free()

            for _ in range(2):
                step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = ctx.launch_count
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
            ms = shard.max_over_ranks(e0.elapsed_time(e1), dist, "cuda") / args.steps
            assert all((r == roots[0]).all() for r in roots)
            del mats
            torch.cuda.empty_cache()
            name = workload_name(args.log_rows, total_cols) if nmats == 1 else f"pcs_commit of {nmats} x (2^{rows.bit_length() - 1} x {total_cols}) in one commitment"
            return {"workload": name, "exchange": exchange, "ms_per_step": ms,
                    "value": algorithmic_bytes(rows * nmats, total_cols) / (ms * 1e-3) / 1e9, "unit": "GB/s",
                    "launches_per_step_per_rank": (ctx.launch_count - l0) // args.steps, "root": [int(x) for x in roots[0]]}

        one_commitment = {
            "note": "one Pcs::commit sharded over the ranks: column shards -> LDE -> row shards stored into peer HBM by "
                    "k_scatter_rows (CUDA IPC over NVLink, overlapped with the next LDE block) -> per-rank subtree -> all-gather of caps; "
                    "'staged' = same with a local pack + NCCL all_to_all_single instead (comparison baseline)",
            "strong_p2p": dist_case(W, "p2p"),
            "weak_p2p": dist_case(W * world, "p2p"),
            "weak_staged_nccl": dist_case(W * world, "staged"),
        }
        if args.log_rows >= 5:  # BASELINE config 5 shape: eight matrices (2^24 rows x 64 columns in total at the default size)
            one_commitment["config5_8_matrices_p2p"] = dist_case(max(W // 4, 8), "p2p", nmats=8)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        leaf_ms, leaf_n = phases["leaf_hash"]
        leaf_ms_per = leaf_ms / max(leaf_n, 1)
        leaf_perms = 2 * R * ((W + 7) // 8)
        leaf_giops = leaf_perms * P2_INSTR_PER_PERM / (leaf_ms_per * 1e-3) / 1e9
        lde_ms = sum(phases[k][0] for k in ("ingest", "intt", "scale", "ntt")) / args.steps
        lde_gbs = algorithmic_bytes(R, W) / (lde_ms * 1e-3) / 1e9
        line = {
            "metric": "lde_poseidon2_commit_throughput", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (KoalaBear mod p, Montgomery on INT32 pipes)", "data": "synthetic",
            "config": {"workload": workload_name(args.log_rows, W), "rows": R, "cols": W, "log_blowup": 1,
                       "parallelism": f"{world} independent trace commits (one per GPU)",
                       "l2": "inputs (4 GiB/step at full size) and LDE (8 GiB) exceed the 126 MB L2; no flush needed"},
            "rows_per_s": world * R / (ms_step * 1e-3),
            "poseidon2_perms_per_s": world * num_perms(R, W) / (ms_step * 1e-3),
            "gpu_launches": launches,
            "phases_ms_per_step": {k: v[0] / args.steps for k, v in phases.items() if v[1]},
            "roofline": {"kernel": "hashk::k_leaf_hash (Poseidon2 sponge, 1 thread/leaf)", "bound": "int32",
                         "achieved": leaf_giops, "peak": int32_peak, "unit": "Ginstr/s", "frac": leaf_giops / int32_peak,
                         "traffic": LEAF_TRAFFIC_2P22X256 if (args.log_rows, W) == (22, 256) else None,
                         "algorithmic_bytes": 8 * R * W + 32 * 2 * R,
                         "note": f"achieved = {leaf_perms} permutations x {P2_INSTR_PER_PERM} SASS integer thread-instructions / {leaf_ms_per:.3f} ms; peak = live register-only IMAD/IADD/LOP3 probe (bfgpu_int32_peak_probe)"},
            "roofline_hbm": {"kernel": "LDE = k_ingest + k_ntt_pass<inv> + k_scale_cosets + k_ntt_pass<fwd>", "bound": "hbm",
                             "achieved": lde_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": lde_gbs / hbm_peak, "traffic": None,
                             "peak_source": hbm_src, "note": f"12*R*W algorithmic bytes / {lde_ms:.3f} ms for the whole LDE"},
            "roofline_ntt_int32": None if args.log_rows != 22 else {
                "kernel": "ntt2::k_pass (3 inverse + 3 forward passes per column at 2^22)", "bound": "int32",
                "achieved": (NTT_INSTR_FWD_2P22 * 2 * R * W + NTT_INSTR_INV_2P22 * R * W) / ((phases["intt"][0] + phases["ntt"][0]) / args.steps * 1e-3) / 1e9,
                "peak": int32_peak, "unit": "Ginstr/s",
                "frac": (NTT_INSTR_FWD_2P22 * 2 * R * W + NTT_INSTR_INV_2P22 * R * W) / ((phases["intt"][0] + phases["ntt"][0]) / args.steps * 1e-3) / 1e9 / int32_peak,
                "note": "the transforms are issue-bound, not HBM-bound: SASS thread-instructions per element (column-loop bodies, cuobjdump) "
                        f"forward {NTT_INSTR_FWD_2P22} per LDE element, inverse {NTT_INSTR_INV_2P22} per trace element; >85 % of them are butterfly arithmetic"},
            "clocks": clocks,
            "root": [int(x) for x in root_dev],
        }
        if e2e:
            line["e2e"] = e2e
        if one_commitment:
            line["one_commitment"] = one_commitment
        if replica_prove:
            line["replica_prove"] = replica_prove
        if world == 1 and not args.no_prove:
            ctx.set_input_space(bf.MEM_HOST)  # traces come from (pinned) host memory
            line["prove"] = prove_timings(ctx, bf, not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            gbs, sec, threads = cpu_commit_sample(args.cpu_log_rows, W, 1, 1)
            line["cpu_baseline"] = {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                                    "sample": f"one commit of a 2^{args.cpu_log_rows}x{W} sample ({sec:.2f} s), oracle C/OpenMP port of the reference algorithm"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
