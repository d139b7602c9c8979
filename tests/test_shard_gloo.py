"""CPU test of the N>1 host logic with world_size-2 gloo process groups (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import zkvm_brainfuck_b200 as bf
from importlib import import_module

shard = import_module("zkvm-brainfuck_b200.shard")


def test_plan_is_balanced_and_deterministic():
    costs = [65536 * 31, 32768 * 41, 16384 * 7, 8192 * 45, 65536 * 2, 64, 192, 80]
    a, b = shard.plan_units(costs, 4), shard.plan_units(costs, 4)
    assert a == b and set(a) <= set(range(4))
    load = [sum(c for c, o in zip(costs, a) if o == r) for r in range(4)]
    assert max(load) == costs[0]  # the biggest unit alone bounds the makespan
    assert shard.plan_units(costs, 1) == [0] * len(costs)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [10, 7, 5, 3, 1]
    owner = shard.plan_units(costs, world)
    mine = [i for i, o in enumerate(owner) if o == rank]
    roots = np.array([[i * 8 + k for k in range(8)] for i in mine], np.uint32).reshape(-1, 8)
    gathered = shard.gather_roots(roots, dist)
    ms = shard.max_over_ranks(10.0 + rank, dist)
    q.put((rank, owner, [g.tolist() for g in gathered], ms))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_max():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, owner0, g0, ms0), (r1, owner1, g1, ms1) = res
    assert owner0 == owner1 and g0 == g1 and ms0 == ms1 == 11.0
    # every unit's root arrives exactly once, grouped by owning rank
    flat = [row[0] // 8 for grp in g0 for row in grp]
    assert sorted(flat) == [0, 1, 2, 3, 4]
    for rank, grp in enumerate(g0):
        assert all(owner0[row[0] // 8] == rank for row in grp)


# ---- control plane of the sharded prover: the bfgpu_comm callbacks the C library calls (all-gather of host bytes, barrier) ----------
def _comm_worker(rank, world, port, q):
    import ctypes as C
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dp = shard.DistributedProver(None, dist)  # no device context needed for the callbacks themselves
        ag, br = dp._comm.all_gather, dp._comm.barrier
        out = []
        for nbytes in (32, 64, 5 * 64, 4096 + 12):
            send = (C.c_uint8 * nbytes)(*[(rank * 37 + i) % 251 for i in range(nbytes)])
            recv = (C.c_uint8 * (nbytes * world))()
            rc = ag(None, C.cast(send, C.c_void_p), C.cast(recv, C.c_void_p), nbytes)
            out.append((rc, bytes(recv)))
            assert br(None) == 0
        # the library's own shared-memory control plane (csrc/comm_shm.h), called through the same bfgpu_comm layout the prover uses
        AG = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
        BR = C.CFUNCTYPE(C.c_int32, C.c_void_p)

        class Comm(C.Structure):
            _fields_ = [("user", C.c_void_p), ("all_gather", AG), ("barrier", BR)]

        sc = C.cast(dp._shm, C.POINTER(Comm)).contents
        shm_out = []
        for nbytes in (32, 320, (3 << 20) + 17):  # the last one is larger than a slot: goes through in pieces
            send = np.frombuffer(np.random.default_rng(rank * 1000 + nbytes).bytes(nbytes), np.uint8).copy()
            recv = np.zeros(nbytes * world, np.uint8)
            rc = sc.all_gather(sc.user, send.ctypes.data, recv.ctypes.data, nbytes)
            shm_out.append((rc, bool(all((recv[r * nbytes:(r + 1) * nbytes] == np.frombuffer(np.random.default_rng(r * 1000 + nbytes).bytes(nbytes), np.uint8)).all()
                                         for r in range(world)))))
            assert sc.barrier(sc.user) == 0
        dp.close()
        q.put((rank, None, out, dict(dp.calls), shm_out))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:
        import traceback
        q.put((rank, traceback.format_exc() + repr(e), None, None, None))


def test_sharded_prover_control_plane_callbacks_world2():
    """what csrc/dist_prove.cuh relies on: recv = the ranks' byte strings in rank order, identical on every rank; barrier returns 0"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, err, out, calls, shm_out in res:
        assert err is None, err
        assert calls["all_gather"] >= 4 and calls["barrier"] == 4
        assert all(rc == 0 and ok for rc, ok in shm_out), shm_out
        for (rc, blob), nbytes in zip(out, (32, 64, 5 * 64, 4096 + 12)):
            assert rc == 0
            want = b"".join(bytes((r * 37 + i) % 251 for i in range(nbytes)) for r in range(world))
            assert blob == want
