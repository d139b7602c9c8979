// comm_shm.h — a native single-node control plane for the sharded prover: bfgpu_comm served out of a POSIX shared-memory segment.
//
// The sharded commitment / proof (dist_commit.cuh, dist_prove.cuh) move their DATA over NVLink; what is left for the caller's
// communicator is ~10 tiny all-gathers (IPC handles, caps, partial sums, proof pieces) and ~8 barriers per proof.  Through
// torch.distributed (gloo, a Python callback per call) each of those costs 150-300 us — 4-5 ms of a 23 ms proof on 8 GPUs.  One
// process per GPU on ONE node can do the same with a shared segment and two atomics: every rank writes its slot, a sense-reversing
// barrier, every rank reads all slots, a second barrier (so nobody overwrites a slot that is still being read): ~2-5 us.
// Multi-node callers keep supplying their own bfgpu_comm (MPI, NCCL + host staging, ...).
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>

struct bfgpu_comm_shm {
    bfgpu_comm comm;  // first member: &comm is what the prover entry points take
    struct Header {
        std::atomic<uint32_t> arrived;
        std::atomic<uint32_t> generation;
        std::atomic<uint32_t> attached;
        uint32_t pad[13];
    };
    Header* hdr = nullptr;
    uint8_t* slots = nullptr;  // world x slot_bytes
    uint64_t slot_bytes = 0, map_bytes = 0;
    uint32_t rank = 0, world = 1;
    double timeout_s = 120.0;
    std::string name;
};

static int32_t shm_barrier(bfgpu_comm_shm* c) {
    const uint32_t gen = c->hdr->generation.load(std::memory_order_acquire);
    if (c->hdr->arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == c->world) {
        c->hdr->arrived.store(0, std::memory_order_relaxed);
        c->hdr->generation.store(gen + 1, std::memory_order_release);
        return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spins = 0; c->hdr->generation.load(std::memory_order_acquire) == gen; spins++) {
        if ((spins & 0x3ff) == 0x3ff) {
            sched_yield();
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > c->timeout_s) return -1;  // a peer died
        }
    }
    return 0;
}
static int32_t shm_cb_barrier(void* user) { return shm_barrier((bfgpu_comm_shm*)user); }
static int32_t shm_cb_all_gather(void* user, const void* send, void* recv, uint64_t bytes) {
    bfgpu_comm_shm* c = (bfgpu_comm_shm*)user;
    for (uint64_t off = 0; off < bytes || (bytes == 0 && off == 0); off += c->slot_bytes) {  // larger payloads go through in slot-sized pieces
        const uint64_t n = bytes - off < c->slot_bytes ? bytes - off : c->slot_bytes;
        memcpy(c->slots + (uint64_t)c->rank * c->slot_bytes, (const uint8_t*)send + off, n);
        if (shm_barrier(c) != 0) return -1;
        for (uint32_t r = 0; r < c->world; r++) memcpy((uint8_t*)recv + (uint64_t)r * bytes + off, c->slots + (uint64_t)r * c->slot_bytes, n);
        if (shm_barrier(c) != 0) return -1;
        if (bytes == 0) break;
    }
    return 0;
}

// Every rank of the job calls this with the SAME name (unique per job, e.g. "/bfgpu-<master port>-<pid of rank 0>"), its rank and the
// world size.  Returns when all ranks have attached; the name is unlinked then, so nothing outlives the processes.
extern "C" int32_t bfgpu_comm_shm_create(const char* name, uint32_t rank, uint32_t world, uint64_t slot_bytes, bfgpu_comm** out) {
    if (!name || !out || world == 0 || rank >= world) return BFGPU_ERR_INVALID;
    *out = nullptr;
    if (slot_bytes < 4096) slot_bytes = 4096;
    slot_bytes = (slot_bytes + 63) & ~(uint64_t)63;
    auto* c = new bfgpu_comm_shm();
    c->name = name;
    c->rank = rank;
    c->world = world;
    c->slot_bytes = slot_bytes;
    c->map_bytes = sizeof(bfgpu_comm_shm::Header) + slot_bytes * world;
    int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)c->map_bytes) != 0) {  // a fresh segment is zero-filled: counters start at 0
        if (fd >= 0) close(fd);
        delete c;
        return BFGPU_ERR_STATE;
    }
    void* p = mmap(nullptr, c->map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) {
        delete c;
        return BFGPU_ERR_STATE;
    }
    c->hdr = (bfgpu_comm_shm::Header*)p;
    c->slots = (uint8_t*)p + sizeof(bfgpu_comm_shm::Header);
    c->comm.user = c;
    c->comm.all_gather = shm_cb_all_gather;
    c->comm.barrier = shm_cb_barrier;
    // wait for everybody, then remove the name
    c->hdr->attached.fetch_add(1, std::memory_order_acq_rel);
    const auto t0 = std::chrono::steady_clock::now();
    while (c->hdr->attached.load(std::memory_order_acquire) < world) {
        sched_yield();
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > c->timeout_s) {
            munmap(p, c->map_bytes);
            shm_unlink(name);
            delete c;
            return BFGPU_ERR_STATE;
        }
    }
    if (shm_barrier(c) != 0) return BFGPU_ERR_STATE;
    if (rank == 0) shm_unlink(name);
    *out = &c->comm;
    return BFGPU_OK;
}
extern "C" void bfgpu_comm_shm_destroy(bfgpu_comm* comm) {
    if (!comm) return;
    auto* c = (bfgpu_comm_shm*)comm->user;
    munmap((void*)c->hdr, c->map_bytes);
    delete c;
}
