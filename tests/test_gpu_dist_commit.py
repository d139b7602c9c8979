"""One `Pcs::commit` spread over several processes/GPUs (csrc/dist_commit.cuh, shard.DistributedCommit) against the
oracle's single commitment of the full matrices: same root, same opened rows and opening proofs.  Ranks share the
visible GPUs round-robin (CUDA IPC also works between two processes on one device), so this runs on a 1-GPU box; the
control plane uses gloo here and NCCL in bench.py."""
import os
import socket
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
P = 2130706433
SHAPES = {
    "chips": [(4096, 31), (4096, 2), (2048, 41), (1024, 7), (512, 45), (64, 1), (16, 12), (16, 5)],  # a CpuProver::commit call
    "wide": [(1 << 14, 200)],        # several 64-column LDE blocks per rank: exercises the LDE/scatter pipeline
    "narrow": [(256, 3), (8, 1), (4, 2)],  # fewer columns than ranks; LDE heights 16 and 8 with 4 ranks -> 4- and 2-row shards
}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _evals(case):
    rng = np.random.default_rng(len(case) * 7 + 1)
    return [rng.integers(0, P, s, dtype=np.uint32) for s in SHAPES[case]]


def _worker(rank, world, port, case, exchange, shifts, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        import zkvm_brainfuck_b200 as bf
        shard = import_module("zkvm-brainfuck_b200.shard")
        ctx = bf.Context(dev)
        evals = _evals(case)
        out = []
        shm = shard.ShmComm(dist)
        for rep in range(2):  # second round reuses cached blocks and IPC mappings
            dc = shard.DistributedCommit(ctx, dist, [m.shape[0] for m in evals], [m.shape[1] for m in evals], exchange=exchange,
                                         comm=shm if rep == 1 else None)  # second round: control plane through shared memory
            local = []
            for i, m in enumerate(evals):
                c0, n = dc.local_cols(i)
                local.append(np.ascontiguousarray(m[:, c0:c0 + n]))
            root = dc.commit(local, domain_shifts=shifts)
            per = dc.rows_per_rank
            opened = []
            for index in sorted({rank * per, rank * per + per - 1, rank * per + (5 % per)}):
                rows, sib = dc.open_batch(index)
                opened.append((index, [r.tolist() for r in rows], sib.tolist()))
            wrong = (rank + 1) % world * per
            try:
                dc.open_batch(wrong)
                refused = False
            except bf.BfGpuError:
                refused = True
            out.append((root.tolist(), opened, refused))
            dc.free()
        q.put((rank, None, out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, traceback.format_exc() + repr(e), None))


def _run(world, case, exchange, shifts=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, exchange, shifts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    for _ in range(world):
        res.append(q.get(timeout=600))
        if res[-1][1]:
            for p in procs:
                p.kill()
            pytest.fail(res[-1][1])
    for p in procs:
        p.join(timeout=120)
    return sorted(res)


@pytest.mark.parametrize("exchange", ["p2p", "staged"])
@pytest.mark.parametrize("world,case", [(2, "chips"), (4, "chips"), (2, "wide"), (4, "narrow")])
def test_distributed_commit_matches_single_commit(world, case, exchange, oracle):
    evals = _evals(case)
    ref = oracle.PcsData(evals)
    res = _run(world, case, exchange)
    for rank, _, reps in res:
        for root, opened, refused in reps:
            assert root == ref.root.tolist()
            assert refused, "a leaf owned by another rank must be refused"
            for index, rows, sib in opened:
                rrows, rsib = ref.tree.open_batch(index)
                assert rows == [r.tolist() for r in rrows]
                assert sib == rsib.tolist()


def test_distributed_commit_shifted_domains(oracle):
    """quotient-chunk style domains (prover.rs:391-411) through the distributed path"""
    evals = _evals("narrow")
    w = pow(3, (P - 1) >> 9, P)
    shifts = [3 * w % P, 3, 3]
    ref = oracle.PcsData(evals, domain_shifts=shifts)
    for rank, _, reps in _run(2, "narrow", "p2p", shifts):
        assert reps[0][0] == ref.root.tolist()


def test_distributed_commit_argument_errors():
    """C-ABI validation of the sharded commit (no second process needed)."""
    import ctypes as C
    import zkvm_brainfuck_b200 as bf
    ctx = bf.Context()
    L = bf.lib()
    rows = (C.c_uint64 * 1)(256)
    cols = (C.c_uint32 * 1)(8)
    h = C.c_void_p()

    def begin(rank, world, r=rows):
        return L.bfgpu_dist_commit_begin(ctx._h, rank, world, r, cols, 1, C.byref(h))

    assert begin(0, 3) == -1 and b"power of two" in L.bfgpu_last_error(ctx._h)
    assert begin(2, 2) == -1
    assert begin(0, 32) == -1
    assert begin(0, 2, (C.c_uint64 * 1)(100)) == -1 and b"power of two" in L.bfgpu_last_error(ctx._h)
    tiny = (C.c_uint64 * 1)(2)  # LDE height 4 < 8 ranks
    assert begin(0, 8, tiny) == -1 and b"below the world size" in L.bfgpu_last_error(ctx._h)
    assert begin(0, 2) == 0
    mat = bf.Mat(None, 256, 4)
    assert L.bfgpu_dist_commit_lde(h, C.byref(mat), None) == -4  # neither peers nor staging set
    cap = (C.c_uint32 * 8)()
    assert L.bfgpu_dist_commit_finish(h, cap) == -4              # no LDE yet
    L.bfgpu_dist_commit_free(h)
    ctx.close()
