"""Host logic of the multi-GPU commitment (zkvm-brainfuck_b200/shard.py, csrc/dist_commit.cuh), checked on CPU with
world_size-2/4 gloo groups: the row-shard subtrees' caps, all-gathered and hashed up log2(G) levels, give the root of
the single commitment, and local siblings + cap-tree siblings form the global opening proof.  The per-rank device work
is replaced here by the oracle; the GPU version of this test is tests/test_gpu_dist_commit.py."""
import os
import socket
from importlib import import_module

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

shard = import_module("zkvm-brainfuck_b200.shard")
P = 2130706433
SHAPES = [(256, 9), (256, 3), (64, 5), (16, 2), (4, 1)]  # smallest LDE height 8 >= world size 4


def test_col_range_partitions_columns():
    for total in (1, 2, 3, 7, 8, 31, 256):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                c0, n = shard.col_range(total, world, r)
                got += list(range(c0, c0 + n))
            assert got == list(range(total))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _compress_many(oracle):
    return lambda l, r: np.stack([oracle.compress(a, b) for a, b in zip(l, r)])


def _worker(rank, world, port, q):
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    evals = [rng.integers(0, P, s, dtype=np.uint32) for s in SHAPES]
    # what the exchange leaves on this rank: rows [rank*h/G, (rank+1)*h/G) of every bit-reversed LDE
    mine = []
    for m in evals:
        lde = oracle.coset_lde_batch_bitrev(m, 1, 3)
        per = lde.shape[0] // world
        mine.append(np.ascontiguousarray(lde[rank * per:(rank + 1) * per]))
    sub = oracle.Tree(mine)
    caps = np.frombuffer(b"".join(shard._all_gather_bytes(sub.root.tobytes(), dist)), np.uint32).reshape(world, 8)
    layers = shard.cap_tree(caps, _compress_many(oracle))
    per = mine[0].shape[0]
    index = rank * per + (3 % per)
    rows, sib = sub.open_batch(index % per)
    pos, top = rank, []
    for l in layers[:-1]:
        top.append(l[pos ^ 1])
        pos >>= 1
    q.put((rank, layers[-1][0].tolist(), index, [r.tolist() for r in rows], np.concatenate([sib.reshape(-1, 8), np.array(top, np.uint32).reshape(-1, 8)]).tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_caps_of_row_shard_subtrees_give_the_root(world, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    evals = [rng.integers(0, P, s, dtype=np.uint32) for s in SHAPES]
    ref = oracle.PcsData(evals)
    dims = [(2 * r, c) for r, c in SHAPES]
    for rank, root, index, rows, sib in res:
        assert root == ref.root.tolist()
        rrows, rsib = ref.tree.open_batch(index)
        assert rows == [r.tolist() for r in rrows]
        assert sib == rsib.tolist()
        assert oracle.verify_batch(ref.root, dims, index, [np.array(r, np.uint32) for r in rows], np.array(sib, np.uint32))
