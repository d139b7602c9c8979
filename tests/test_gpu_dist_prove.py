"""ONE shard proof over several processes/GPUs (csrc/dist_prove.cuh, shard.DistributedProver) against the single-GPU
proof of the same statement: every rank must return, word for word, the proof `bfgpu_machine_open` produces (which the
parity tests pin to the CPU oracle), and the native verifier must accept it.  Ranks share the visible GPUs round-robin
(CUDA IPC also works between processes on one device), so this runs on a 1-GPU box; gloo carries the control plane."""
import os
import socket
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROGRAMS = {
    "tiny": ("++[>+<-]>,.", [42]),                                  # every chip, 16-row traces next to the 2^16-row Byte chip
    "fibo": (open(os.path.join(GOLD, "fibo.bf")).read(), [17]),    # BASELINE config 1: Cpu 2^16 rows
    "loop18": ("+++++[>-[>+>+>+<<<-]<-]", []),                      # ~2^18-row Cpu trace: several sharded FRI rounds above the gather point
}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, fri, from_traces, q, gather_log=None, control_plane="shm"):
    try:
        if gather_log is not None:  # force sharded FRI rounds on small proofs
            os.environ["BFGPU_DIST_FRI_GATHER_LOG"] = str(gather_log)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        import zkvm_brainfuck_b200 as bf
        shard = import_module("zkvm-brainfuck_b200.shard")
        ctx = bf.Context(dev)
        ctx.set_fri_params(*fri)
        code, stdin = PROGRAMS[case]
        prover = bf.CudaProver(ctx)
        dp = shard.DistributedProver(ctx, dist, control_plane=control_plane)
        rec = prover.execute(code, stdin)
        pk = prover.setup_record(rec)
        out = []
        for rep in range(2):  # the second proof recycles the exported buffers and IPC mappings
            ch = bf.Challenger(ctx)
            bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
            if from_traces:
                sh = prover.commit_record(rec)
                traces = prover.shard_traces(sh)
                sh.free()
                words = dp.prove(pk, traces, ch.clone())
            else:
                words = dp.prove_record(pk, rec, ch.clone())
            out.append(words)
        # the single-GPU proof of the same statement on this rank's context
        ch = bf.Challenger(ctx)
        bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
        sh = prover.commit_record(rec)
        ref = prover.open_raw(pk, sh, ch.clone())
        sh.free()
        verdict = bf.verify_shard(pk.commit, pk.names, pk.heights, out[0], *fri)
        same = [bool(w.shape == ref.shape and (w == ref).all()) for w in out]
        first_diff = None
        if not same[0] and out[0].shape == ref.shape:
            first_diff = int(np.flatnonzero(out[0] != ref)[0])
        q.put((rank, None, dict(same=same, verdict=verdict, words=int(ref.size), got=int(out[0].size), first_diff=first_diff, calls=dict(dp.calls),
                                head=out[0][:24].tolist())))
        dist.barrier()
        pk.free()
        ctx.close()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, traceback.format_exc() + repr(e), None))


def _run(world, case, fri=(1, 12, 6), from_traces=False, gather_log=None, control_plane="shm"):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, fri, from_traces, q, gather_log, control_plane)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    try:
        for _ in range(world):
            res.append(q.get(timeout=600))
            if res[-1][1]:
                pytest.fail(res[-1][1])
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    return sorted(res)


@pytest.mark.parametrize("world,case", [(2, "tiny"), (4, "tiny"), (2, "fibo"), (4, "fibo"), (8, "fibo"), (2, "loop18"), (4, "loop18")])
def test_sharded_proof_is_the_single_gpu_proof(world, case):
    res = _run(world, case)
    heads = {tuple(r[2]["head"]) for r in res}
    assert len(heads) == 1, "ranks disagree on the commitments"
    for rank, _, r in res:
        assert r["got"] == r["words"], (rank, r)
        assert all(r["same"]), f"rank {rank}: sharded proof differs from the single-GPU proof (first differing word {r['first_diff']})"
        assert r["verdict"] is None, r["verdict"]


@pytest.mark.parametrize("world,case,gather_log", [(2, "tiny", 6), (4, "fibo", 8), (8, "fibo", 10), (4, "loop18", 13)])
def test_sharded_fri_rounds_above_the_gather_point(world, case, gather_log):
    """the default gather point (2^20 elements) leaves small proofs without a sharded FRI round: lower it so that the per-round subtree +
    cap exchange + sharded fold + owner-answered layer openings run on the test programs too"""
    for rank, _, r in _run(world, case, gather_log=gather_log):
        assert all(r["same"]) and r["verdict"] is None, (rank, r)


def test_sharded_proof_full_parameters_from_host_traces():
    """84 queries / 16 PoW bits (kb31_poseidon2.rs:54-64), traces handed in from the host on every rank"""
    for rank, _, r in _run(2, "fibo", fri=(1, 84, 16), from_traces=True):
        assert all(r["same"]) and r["verdict"] is None, (rank, r)


def test_sharded_proof_through_caller_supplied_callbacks():
    """control plane through the caller's own bfgpu_comm (torch.distributed / gloo callbacks: what a multi-node caller supplies) instead
    of the library's shared-memory communicator"""
    for rank, _, r in _run(4, "tiny", gather_log=6, control_plane="dist"):
        assert all(r["same"]) and r["verdict"] is None, (rank, r)
        assert r["calls"]["all_gather"] >= 10 and r["calls"]["barrier"] >= 6
