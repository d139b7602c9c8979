"""Multi-GPU sharding of the proving hot path: one process per GPU.

Three levels: independent units (`plan_units`: whole traces / proofs per rank, no data-path exchange), ONE commitment over
all ranks (`DistributedCommit`: column-sharded LDE, P2P row exchange into peer HBM, per-rank subtrees, cap tree) and ONE
shard proof over all ranks (`DistributedProver`, csrc/dist_prove.cuh).  `torch.distributed` carries only the control
plane (IPC handles, caps, partial sums, proof pieces, barriers): NCCL or gloo for the tensors of bench.py, a gloo group for
the small host-side exchanges of the sharded prover.
"""
import numpy as np


def plan_units(costs, world_size):
    """Deterministic longest-processing-time assignment of work units to ranks.
    costs: list of per-unit costs (e.g. cells of a trace matrix).  Returns [rank of unit i]."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    owner = [0] * len(costs)
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return owner


def gather_roots(local_roots, dist=None):
    """All-gather the 8-word roots of every rank's units -> list (rank-major) of (n_r, 8) uint32 arrays."""
    local = np.ascontiguousarray(np.asarray(local_roots, np.uint32).reshape(-1, 8))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    import torch
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(counts, torch.tensor([local.shape[0]], dtype=torch.int64))
    m = int(max(c.item() for c in counts))
    buf = torch.zeros((m, 8), dtype=torch.int64)
    buf[:local.shape[0]] = torch.from_numpy(local.astype(np.int64))
    out = [torch.zeros((m, 8), dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(out, buf)
    return [o[:int(c.item())].numpy().astype(np.uint32) for o, c in zip(out, counts)]


def max_over_ranks(value, dist=None, device=None):
    """Device-time reduction used by bench.py: the job time is the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- one commitment over several GPUs -----------------------------------------------------------------------------
def col_range(total_cols, world_size, rank):
    """Columns [c0, c0 + n) of a `total_cols`-wide matrix that `rank` extends when it is matrix 0 of a commitment (matrix i uses
    position (rank + i) mod world: `DistributedCommit.local_cols`; same rule as dist_commit.cuh)."""
    a, b = total_cols * rank // world_size, total_cols * (rank + 1) // world_size
    return a, b - a


def cap_tree(caps, compress):
    """Top of the Merkle tree over the ranks' subtree caps: [caps, ..., root] with root = layers[-1][0].
    `compress(left, right)` maps (n, 8) digest arrays to (n, 8) (TruncatedPermutation, kb31_poseidon2.rs:26)."""
    layers = [np.asarray(caps, np.uint32).reshape(-1, 8)]
    while layers[-1].shape[0] > 1:
        prev = layers[-1]
        layers.append(np.asarray(compress(prev[0::2], prev[1::2]), np.uint32).reshape(-1, 8))
    return layers


def _all_gather_bytes(payload, dist, group=None):
    """All-gather equally sized byte strings through torch.distributed (CUDA tensors under NCCL, CPU under gloo)."""
    import torch
    dev = "cuda" if "nccl" in str(dist.get_backend(group)) else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    return [bytes(o.cpu().numpy().tobytes()) for o in out]


class ShmComm:
    """The library's shared-memory control plane (csrc/comm_shm.h) for one process per GPU on one node: an all-gather of small host
    byte strings and a barrier in a few microseconds.  Collective constructor: every rank of `group` calls it (the segment's name
    is agreed on through `dist` once)."""

    def __init__(self, dist, group=None, slot_bytes=1 << 20):
        import ctypes as C
        import os
        from . import lib
        self._C, self._lib = C, lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        tag = os.urandom(6).hex() if self.rank == 0 else ""
        name = _all_gather_bytes(tag.ljust(12).encode(), dist, group)[0].decode().strip()
        self._h = C.c_void_p()
        rc = self._lib.bfgpu_comm_shm_create(f"/bfgpu-{name}".encode(), self.rank, self.world, slot_bytes, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"bfgpu_comm_shm_create failed ({rc})")
        AG = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
        BR = C.CFUNCTYPE(C.c_int32, C.c_void_p)

        class Comm(C.Structure):
            _fields_ = [("user", C.c_void_p), ("all_gather", AG), ("barrier", BR)]

        self._c = C.cast(self._h, C.POINTER(Comm)).contents

    @property
    def handle(self):
        """bfgpu_comm* for the prover entry points"""
        return self._h

    def all_gather(self, payload):
        n = len(payload)
        recv = self._C.create_string_buffer(n * self.world)
        if self._c.all_gather(self._c.user, payload, recv, n) != 0:
            raise RuntimeError("shared-memory all-gather failed (a peer is gone?)")
        raw = recv.raw
        return [raw[r * n:(r + 1) * n] for r in range(self.world)]

    def barrier(self):
        if self._c.barrier(self._c.user) != 0:
            raise RuntimeError("shared-memory barrier failed (a peer is gone?)")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bfgpu_comm_shm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


class DistributedCommit:
    """`Pcs::commit` (prover.rs:227) of matrices whose COLUMNS are spread over the ranks of a process group: every
    rank extends its columns, the LDE blocks are stored into the peers' row shards over NVLink (exchange="p2p",
    CUDA IPC, no library collective on the data path) or moved by `all_to_all_single` (exchange="staged", the
    comparison baseline), every rank hashes its rows into one subtree, and the caps are all-gathered.
    `root` is identical to the single-GPU commitment of the full matrices."""

    def __init__(self, ctx, dist, rows, total_cols, group=None, exchange="p2p", comm=None):
        """comm: optional `ShmComm` — handles, barrier and caps then go through shared memory instead of three torch.distributed
        collectives (~0.4 ms of a 12 ms commitment on 8 GPUs)."""
        from . import lib, BfGpuError, _u32p, _u64p  # noqa: F401
        import ctypes as C
        self._lib, self._C = lib(), C
        self.ctx, self.dist, self.group = ctx, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if exchange not in ("p2p", "staged"):
            raise ValueError("exchange must be 'p2p' or 'staged'")
        self.exchange = exchange
        self.comm = comm
        self.rows = np.ascontiguousarray(rows, np.uint64)
        self.total_cols = np.ascontiguousarray(total_cols, np.uint32)
        self.n = len(self.rows)
        self._h = C.c_void_p()
        ctx.check(self._lib.bfgpu_dist_commit_begin(ctx._h, self.rank, self.world, self.rows.ctypes.data_as(_u64p),
                                                    self.total_cols.ctypes.data_as(_u32p), self.n, C.byref(self._h)))
        self.root = None
        self._send = self._recv = None

    def local_cols(self, i):
        """(first column, count) of matrix i this rank extends: the even split taken at position (rank + i) mod world (dist_commit.cuh)"""
        c0 = self._C.c_uint32()
        n = self._lib.bfgpu_dist_commit_local_cols(self._h, int(i), self._C.byref(c0))
        return int(c0.value), int(n)

    def commit(self, local_mats, domain_shifts=None):
        """local_mats[i]: this rank's rows[i] x local_cols(i) slice — numpy array (host input space) or a
        (device_pointer, rows, cols) tuple (device input space).  Returns the 8-word root."""
        from . import Mat, _u32, _u32p
        C, L, ctx = self._C, self._lib, self.ctx
        arr, keep = (Mat * self.n)(), []
        for i, m in enumerate(local_mats):
            if isinstance(m, tuple):
                arr[i] = Mat(m[0], m[1], m[2])
            else:
                a = _u32(m)
                keep.append(a)
                arr[i] = Mat(a.ctypes.data if a.size else None, a.shape[0], a.shape[1])
        sh = None
        if domain_shifts is not None:
            sh = _u32(domain_shifts)
        if self.exchange == "p2p":
            handle = (C.c_uint8 * 64)()
            ctx.check(L.bfgpu_dist_commit_recv_handle(self._h, handle))
            handles = b"".join(self.comm.all_gather(bytes(handle)) if self.comm else _all_gather_bytes(bytes(handle), self.dist, self.group))
            ctx.check(L.bfgpu_dist_commit_set_peers(self._h, handles))
        else:
            import torch
            bw = [int(L.bfgpu_dist_commit_block_words(self._h, r)) for r in range(self.world)]
            self._send = torch.empty(max(bw[self.rank] * self.world, 1), dtype=torch.int32, device="cuda")
            self._recv = torch.empty(max(sum(bw), 1), dtype=torch.int32, device="cuda")
            ctx.check(L.bfgpu_dist_commit_set_staging(self._h, C.c_void_p(self._send.data_ptr())))
        ctx.check(L.bfgpu_dist_commit_lde(self._h, arr, sh.ctypes.data_as(_u32p) if sh is not None else None))
        ctx.synchronize()  # this rank's stores into the peers (or the staging buffer) have landed
        if self.exchange == "p2p":
            if self.comm:
                self.comm.barrier()
            else:
                self.dist.barrier(group=self.group)  # ... and so have everybody else's into this rank
        else:
            import torch
            send, recv = self._send[:bw[self.rank] * self.world], self._recv[:sum(bw)]
            on_gpu = "nccl" in str(self.dist.get_backend(self.group))
            if not on_gpu:  # gloo has no CUDA all-to-all: bounce through the host (tests on a single GPU only)
                send, recv_dev, recv = send.cpu(), recv, torch.empty(sum(bw), dtype=torch.int32)
            self.dist.all_to_all_single(recv, send, output_split_sizes=bw, input_split_sizes=[bw[self.rank]] * self.world, group=self.group)
            if not on_gpu:
                recv_dev.copy_(recv)
            torch.cuda.synchronize()
            ctx.check(L.bfgpu_dist_commit_unpack(self._h, C.c_void_p(self._recv.data_ptr())))
        cap = np.zeros(8, np.uint32)
        ctx.check(L.bfgpu_dist_commit_finish(self._h, cap.ctypes.data_as(_u32p)))
        caps = np.frombuffer(b"".join(self.comm.all_gather(cap.tobytes()) if self.comm else _all_gather_bytes(cap.tobytes(), self.dist, self.group)),
                             np.uint32).copy()
        root = np.zeros(8, np.uint32)
        ctx.check(L.bfgpu_dist_commit_root(self._h, caps.ctypes.data_as(_u32p), root.ctypes.data_as(_u32p)))
        self.caps, self.root = caps.reshape(-1, 8), root
        self._send = self._recv = None
        return root

    @property
    def rows_per_rank(self):
        return int(self._lib.bfgpu_dist_commit_rows_per_rank(self._h))

    def owner(self, index):
        return int(index) // self.rows_per_rank

    def open_batch(self, index):
        """Mmcs::open_batch(index) on the rank that owns the leaf: ([row of every matrix], siblings (log2 h, 8))."""
        # what the library writes: log2(rows per rank) local siblings + log2(world) cap-tree siblings (follows the context's log_blowup)
        nl = (self.rows_per_rank.bit_length() - 1) + (self.world.bit_length() - 1)
        total = int(self.total_cols.sum())
        rows = np.zeros(total, np.uint32)
        sib = np.zeros((nl, 8), np.uint32)
        self.ctx.check(self._lib.bfgpu_dist_commit_open_batch(self._h, int(index), rows.ctypes.data_as(self._C.c_void_p),
                                                              sib.ctypes.data_as(self._C.c_void_p)))
        out, o = [], 0
        for c in self.total_cols:
            out.append(rows[o:o + int(c)].copy())
            o += int(c)
        return out, sib

    def free(self):
        if getattr(self, "_h", None):
            self._lib.bfgpu_dist_commit_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass


# ---- one shard proof over several GPUs -----------------------------------------------------------------------------------
class DistributedProver:
    """`MachineProver::prove` (crates/stark/src/prover.rs:560-582) of ONE shard by all ranks of a process group
    (csrc/dist_prove.cuh): every rank calls with the same proving key / record / challenger state and receives the same
    serialised proof, word for word the single-GPU one.  The library's control plane (bfgpu_comm: all-gather of small host
    buffers + barrier) is served by a gloo group created next to the caller's group."""

    def __init__(self, ctx, dist, group=None, control_plane="shm"):
        """control_plane: "shm" = the library's shared-memory communicator (one node; the name is agreed on through `dist` once),
        "dist" = every call through torch.distributed (gloo) callbacks — works across nodes, ~100x the latency per call."""
        import ctypes as C
        from . import lib
        self._C, self._lib, self.ctx, self.dist = C, lib(), ctx, dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._shm = None
        self.control_plane = control_plane
        backend = str(dist.get_backend(group))
        # host-side exchanges: gloo.  (new_group is collective: every rank of the default group constructs its DistributedProver)
        self._cpu_group = group if "gloo" in backend and "nccl" not in backend else dist.new_group(backend="gloo")
        self.calls = {"all_gather": 0, "barrier": 0, "bytes": 0}
        AG = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)
        BR = C.CFUNCTYPE(C.c_int32, C.c_void_p)

        def all_gather(_user, send, recv, nbytes):
            try:
                import torch
                n = int(nbytes)
                src = torch.frombuffer((C.c_uint8 * n).from_address(send), dtype=torch.uint8)
                dst = torch.frombuffer((C.c_uint8 * (n * self.world)).from_address(recv), dtype=torch.uint8)
                self.dist.all_gather(list(dst.view(self.world, n).unbind(0)), src, group=self._cpu_group)
                self.calls["all_gather"] += 1
                self.calls["bytes"] += n
                return 0
            except Exception:  # never unwind through the C frames
                import traceback
                traceback.print_exc()
                return -1

        def barrier(_user):
            try:
                self.dist.barrier(group=self._cpu_group)
                self.calls["barrier"] += 1
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return -1

        class Comm(C.Structure):
            _fields_ = [("user", C.c_void_p), ("all_gather", AG), ("barrier", BR)]

        self._cb = (AG(all_gather), BR(barrier))  # keep the thunks alive
        self._comm = Comm(None, self._cb[0], self._cb[1])
        self._comm_ptr = C.cast(C.pointer(self._comm), C.c_void_p)
        if control_plane == "shm":
            self.shm_comm = ShmComm(dist, self._cpu_group)
            self._shm = self.shm_comm.handle
            self._comm_ptr = self._shm

    def close(self):
        if self._shm is not None:
            self.shm_comm.close()
            self._shm = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def _finish(self, h):
        C, L = self._C, self._lib
        size = L.bfgpu_shard_proof_size(h)
        buf = np.zeros(size, np.uint32)
        self.ctx.check(L.bfgpu_shard_proof_read(h, buf.ctypes.data_as(C.c_void_p)))
        L.bfgpu_shard_proof_free(h)
        return buf

    def prove_record(self, pk, rec, challenger, pow_witness=None):
        """proof words from an execution record (`bf.Record`): device-side trace generation replicated, the rest sharded"""
        C = self._C
        h = C.c_void_p()
        self.ctx.check(self._lib.bfgpu_dist_prove_record(self.ctx._h, self._comm_ptr, self.rank, self.world, pk._h, rec._h, challenger._h,
                                                          -1 if pow_witness is None else int(pow_witness), C.byref(h)))
        return self._finish(h)

    def prove(self, pk, traces, challenger, pow_witness=None):
        """proof words from host traces {chip name: main trace} (every rank passes all of them)"""
        from . import _named_mats
        C = self._C
        named = list(traces.items())
        cn, arr, keep = _named_mats(named)
        h = C.c_void_p()
        self.ctx.check(self._lib.bfgpu_dist_prove(self.ctx._h, self._comm_ptr, self.rank, self.world, pk._h, cn, arr, len(named), challenger._h,
                                                   -1 if pow_witness is None else int(pow_witness), C.byref(h)))
        return self._finish(h)
