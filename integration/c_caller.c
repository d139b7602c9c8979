/* integration/c_caller.c — the boundary used from plain C, no Python and no Rust:
 *
 *     gcc -std=c99 -Iinclude integration/c_caller.c -Lzkvm-brainfuck_b200 -lbfgpu -Wl,-rpath,$PWD/zkvm-brainfuck_b200 -o c_caller
 *     ./c_caller                       # program -> proof -> verify on device 0; without a GPU: execute only, then "no CPU fallback"
 *     ./c_caller 20                    # the same, then 20 more proofs timed from this side of the ABI (ms per proof)
 *
 * The call sequence is the one `ProverClient::{execute, setup, prove, verify}` makes in the reference
 * (crates/sdk/src/lib.rs:19-140 -> crates/prover/src/lib.rs:46-104 -> crates/core/machine/src/utils/prove.rs:24-60):
 *   Executor::run                 bfgpu_execute            (host; works without a device)
 *   StarkMachine::setup           bfgpu_machine_setup_record
 *   challenger + vk.observe_into  bfgpu_challenger_create / bfgpu_pk_observe_into
 *   MachineProver::commit / open  bfgpu_machine_commit_record / bfgpu_machine_open
 *   BfProver::verify              bfgpu_verify_core_proof  (host)
 * tests/test_abi.py builds and runs it (the execute part everywhere, the proving part on a GPU box). */
#define _POSIX_C_SOURCE 199309L /* clock_gettime under -std=c99 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bfgpu.h"

static const char* PROGRAM = "++++++++[>++++++++<-]>+.,.";  /* prints 'A', then echoes the input byte */

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

int main(int argc, char** argv) {
    const uint8_t input[1] = {'z'};
    bfgpu_record* rec = NULL;
    if (bfgpu_execute(NULL, PROGRAM, input, 1, 0, &rec) != BFGPU_OK) {
        fprintf(stderr, "execute: %s\n", bfgpu_record_error(rec));
        bfgpu_record_free(rec);
        return 1;
    }
    uint64_t counts[8];
    bfgpu_record_info(rec, counts);
    uint8_t out[16] = {0};
    if (counts[7] > sizeof out) return 1;
    bfgpu_record_output(rec, out);
    printf("executed: cycles=%llu output=%.*s\n", (unsigned long long)counts[0], (int)counts[7], (const char*)out);
    if (counts[7] != 2 || out[0] != 'A' || out[1] != 'z') return 1;
    const uint64_t n_instr = counts[1];
    bfgpu_record_free(rec);
    rec = NULL;

    bfgpu_ctx* ctx = NULL;
    if (bfgpu_ctx_create(0, &ctx) != BFGPU_OK) {
        /* the library has no CPU path for the prover: this is the expected end of the run on a machine without a B200 */
        printf("prover unavailable: %s\n", bfgpu_last_error(ctx)); /* the handle is returned even on failure, for the message */
        bfgpu_ctx_destroy(ctx);
        return 0;
    }
    int rc = 1;
    bfgpu_pk* pk = NULL;
    bfgpu_shard* shard = NULL;
    bfgpu_challenger* ch = NULL;
    bfgpu_shard_proof* proof = NULL;
    uint32_t vk_commit[8], main_root[8];
    bfgpu_set_fri_params(ctx, 1, 84, 16); /* default_fri_config (kb31_poseidon2.rs:54-64) */
    if (bfgpu_execute(ctx, PROGRAM, input, 1, 0, &rec) != BFGPU_OK) goto done;
    if (bfgpu_machine_setup_record(ctx, rec, vk_commit, &pk) != BFGPU_OK) goto done;
    if (bfgpu_challenger_create(ctx, &ch) != BFGPU_OK) goto done;
    if (bfgpu_pk_observe_into(pk, ch) != BFGPU_OK) goto done;
    if (bfgpu_machine_commit_record(ctx, rec, main_root, &shard) != BFGPU_OK) goto done;
    if (bfgpu_machine_open(ctx, pk, shard, ch, -1, &proof) != BFGPU_OK) goto done;
    {
        const uint64_t n_words = bfgpu_shard_proof_size(proof);
        uint32_t* words = (uint32_t*)malloc(n_words * sizeof(uint32_t));
        if (!words || bfgpu_shard_proof_read(proof, words) != BFGPU_OK) goto done;
        /* verifying key = preprocessed commitment + (name, log2 height) of the preprocessed traces in proving-key order, i.e. sorted by
         * (height desc, name) like machine.rs:182-183: the byte table (2^16 rows) and the program listing (instructions padded to a power
         * of two, at least 16 rows) */
        uint32_t log_prog = 4;
        while ((1ull << log_prog) < n_instr) log_prog++;
        const char* names[2] = {"Byte", "Program"};
        uint32_t logs[2] = {16, log_prog};
        if (log_prog > 16) {
            names[0] = "Program", names[1] = "Byte";
            logs[0] = log_prog, logs[1] = 16;
        }
        char err[256] = {0};
        rc = bfgpu_verify_core_proof(vk_commit, names, logs, 2, words, n_words, BFGPU_REPR_CANONICAL, 1, 84, 16, NULL, 0, err, sizeof err);
        printf("proof: %llu words, verifier: %s\n", (unsigned long long)n_words, rc == BFGPU_OK ? "accepted" : err);
        free(words);
    }
    if (rc == BFGPU_OK && argc > 1) { /* latency of execute -> commit -> open seen by a C caller (the proving key stays resident) */
        const int reps = atoi(argv[1]);
        const double t0 = now_ms();
        for (int k = 0; k < reps && rc == BFGPU_OK; k++) {
            bfgpu_record* r2 = NULL;
            bfgpu_shard* s2 = NULL;
            bfgpu_challenger* c2 = NULL;
            bfgpu_shard_proof* p2 = NULL;
            rc = bfgpu_execute(ctx, PROGRAM, input, 1, 0, &r2);
            if (rc == BFGPU_OK) rc = bfgpu_challenger_create(ctx, &c2);
            if (rc == BFGPU_OK) rc = bfgpu_pk_observe_into(pk, c2);
            if (rc == BFGPU_OK) rc = bfgpu_machine_commit_record(ctx, r2, main_root, &s2);
            if (rc == BFGPU_OK) rc = bfgpu_machine_open(ctx, pk, s2, c2, -1, &p2);
            bfgpu_shard_proof_free(p2);
            bfgpu_shard_free(s2);
            bfgpu_challenger_free(c2);
            bfgpu_record_free(r2);
        }
        if (rc == BFGPU_OK && reps > 0) printf("%d proofs from C: %.3f ms per proof\n", reps, (now_ms() - t0) / reps);
    }
done:
    if (rc != BFGPU_OK) fprintf(stderr, "failed: %s\n", bfgpu_last_error(ctx));
    bfgpu_shard_proof_free(proof);
    bfgpu_shard_free(shard);
    bfgpu_challenger_free(ch);
    bfgpu_pk_free(pk);
    bfgpu_record_free(rec);
    bfgpu_ctx_destroy(ctx);
    return rc == BFGPU_OK ? 0 : 1;
}
