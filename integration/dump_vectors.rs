//! PIN KIT, Rust side — drop this file into the reference as `crates/stark/tests/dump_vectors.rs` (dev-dependencies: serde_json,
//! bf-core-machine, bf-core-executor, bf-test-artifacts as the crate's own tests use) and run
//!
//!     FRI_QUERIES=12 cargo test -r --test dump_vectors -- --nocapture > rust_vectors.json
//!
//! It prints, from the REAL prover (Plonky3 rev 93967fce), the same fields that `tests/golden/vectors.json` of the B200 backend holds
//! (generator: scripts/gen_golden.py, schema below), with every field element as its CANONICAL u32.  Then, in the backend checkout:
//!
//!     python scripts/compare_golden.py rust_vectors.json
//!
//! says which stage first differs and which transcript option (bfgpu_set_transcript_option) would flip it.  This closes the
//! "parity unpinned" gap of the backend: everything it could check offline is self-consistency; this run is the byte pin.
//!
//! The proof-of-work witness: the reference's grind uses rayon `find_any`, so its witness (hence every query) differs per run; the
//! dump therefore records the witness the Rust prover picked, and compare_golden.py re-runs the backend with THAT witness
//! (`fixed_pow_witness`) before comparing query openings.  Nondeterministic `cpu_memory_access` order (hashbrown drain,
//! executor.rs:74-76) permutes the Memory chip's rows: the dump also records the Memory trace so the comparer can feed the same one.
//!
//! NOT compiled in the backend's CI (no Rust toolchain in that image).
use bf_core_executor::{Executor, Program};
use bf_core_machine::{brainfuck::BfAir, utils::setup_logger};
use bf_stark::{koala_bear_poseidon2::KoalaBearPoseidon2, CpuProver, MachineProver, StarkGenericConfig, StarkMachine};
use p3_challenger::{CanObserve, CanSample};
use p3_commit::Pcs;
use p3_dft::TwoAdicSubgroupDft;
use p3_field::{FieldAlgebra, PrimeField32};
use p3_koala_bear::KoalaBear;
use p3_matrix::{dense::RowMajorMatrix, Matrix};
use p3_symmetric::{CryptographicHasher, Permutation, PseudoCompressionFunction};
use serde_json::{json, Value};

type F = KoalaBear;
const P: u64 = 2130706433;

fn c(x: F) -> u32 {
    x.as_canonical_u32()
}
fn cs<I: IntoIterator<Item = F>>(it: I) -> Vec<u32> {
    it.into_iter().map(c).collect()
}
/// numpy.random.default_rng(seed).integers(0, P, (rows, cols)) cannot be reproduced in Rust: the golden generator also stores these
/// inputs under "inputs" so that both sides hash the same matrices.
fn mat(v: &Value) -> RowMajorMatrix<F> {
    let rows = v.as_array().unwrap();
    let w = rows[0].as_array().unwrap().len();
    RowMajorMatrix::new(rows.iter().flat_map(|r| r.as_array().unwrap().iter().map(|x| F::from_canonical_u32(x.as_u64().unwrap() as u32))).collect(), w)
}

#[test]
fn dump_vectors() {
    setup_logger();
    let sc = KoalaBearPoseidon2::new();
    let perm = sc.perm.clone();
    let mut out = serde_json::Map::new();
    // ---- primitives (same inputs as scripts/gen_golden.py) -------------------------------------------------------------------------
    let mut st: [F; 16] = core::array::from_fn(|i| F::from_canonical_u32(i as u32));
    perm.permute_mut(&mut st);
    out.insert("poseidon2_permute_0_to_15".into(), json!(cs(st)));
    let hash = bf_stark::koala_bear_poseidon2::MyHash::new(perm.clone());
    out.insert("sponge_hash_0_to_30".into(), json!(cs(hash.hash_iter((0..31).map(F::from_canonical_u32)))));
    let compress = bf_stark::koala_bear_poseidon2::MyCompress::new(perm.clone());
    let l: [F; 8] = core::array::from_fn(|i| F::from_canonical_u32(i as u32));
    let r: [F; 8] = core::array::from_fn(|i| F::from_canonical_u32(8 + i as u32));
    out.insert("compress_0_to_7_and_8_to_15".into(), json!(cs(compress.compress([l, r]))));
    // ---- LDE and commitments on the committed input matrices (tests/golden/inputs.json, written by gen_golden.py --with-inputs) -------
    if let Ok(text) = std::fs::read_to_string(std::env::var("BFGPU_GOLDEN_INPUTS").unwrap_or("inputs.json".into())) {
        let inputs: Value = serde_json::from_str(&text).unwrap();
        let a = mat(&inputs["seed1_64x3"]);
        let lde = bf_stark::koala_bear_poseidon2::Dft::default().coset_lde_batch(a, 1, F::GENERATOR).to_row_major_matrix();
        out.insert("coset_lde_seed1_64x3_row0_row127".into(), json!([cs(lde.row(0)), cs(lde.row(127))]));
        let mats: Vec<RowMajorMatrix<F>> = ["seed2_1024x31", "seed3_1024x2", "seed4_64x7", "seed5_16x5"].iter().map(|k| mat(&inputs[*k])).collect();
        let pcs = sc.pcs();
        let doms: Vec<_> = mats.iter().map(|m| (<_ as Pcs<_, <KoalaBearPoseidon2 as StarkGenericConfig>::Challenger>>::natural_domain_for_degree(pcs, m.height()), m.clone())).collect();
        let (root, _data) = <_ as Pcs<_, <KoalaBearPoseidon2 as StarkGenericConfig>::Challenger>>::commit(pcs, doms);
        let root: [F; 8] = root.into();
        out.insert("pcs_commit_root_seeds2to5".into(), json!(cs(root)));
    }
    // ---- whole proofs: hello and fibo(17) ----------------------------------------------------------------------------------------------
    let mut proofs = serde_json::Map::new();
    for (name, code, stdin) in [("hello", bf_test_artifacts::HELLO_BF, vec![]), ("fibo", bf_test_artifacts::FIBO_BF, vec![17u8])] {
        let program = Program::from(code).unwrap();
        let machine: StarkMachine<KoalaBearPoseidon2, BfAir<F>> = BfAir::machine(KoalaBearPoseidon2::new());
        let prover = CpuProver::new(machine);
        let (pk, _vk) = prover.setup(&program);
        let mut runtime = Executor::new(program, stdin.clone());
        runtime.run().unwrap();
        let mut challenger = prover.config().challenger();
        let proof = prover.prove(&pk, &mut runtime.record, &mut challenger).unwrap().shard_proof;
        let dig = |h: &bf_stark::Com<KoalaBearPoseidon2>| -> Vec<u32> { let a: [F; 8] = h.clone().into(); cs(a) };
        let mut order: Vec<(&String, &usize)> = proof.chip_ordering.iter().collect();
        order.sort_by_key(|(_, i)| **i);
        let ext = |e: &<KoalaBearPoseidon2 as StarkGenericConfig>::Challenge| -> Vec<u32> { cs(p3_field::FieldExtensionAlgebra::<F>::as_base_slice(e).iter().copied()) };
        proofs.insert(name.into(), json!({
            "stdin": stdin, "cycles": runtime.state.global_clk, "output": runtime.state.output_stream,
            "fri": [1, std::env::var("FRI_QUERIES").ok().and_then(|v| v.parse::<u32>().ok()).unwrap_or(84), 16],
            "preprocessed_commit": dig(&pk.commit),
            "commitments": { "main": dig(&proof.commitment.main_commit), "permutation": dig(&proof.commitment.permutation_commit),
                             "quotient": dig(&proof.commitment.quotient_commit) },
            "chip_ordering": order.iter().map(|(n, i)| ((*n).clone(), **i)).collect::<std::collections::BTreeMap<_, _>>(),
            "cumulative_sums": proof.opened_values.chips.iter().map(|ch| ext(&ch.cumulative_sum)).collect::<Vec<_>>(),
            "fri_commit_phase_commits": proof.opening_proof.fri_proof.commit_phase_commits.iter().map(dig).collect::<Vec<_>>(),
            "final_poly": ext(&proof.opening_proof.fri_proof.final_poly),
            "pow_witness": c(proof.opening_proof.fri_proof.pow_witness),
            "proof_bincode_len": bincode::serialize(&bf_stark::MachineProof { shard_proof: proof.clone() }).unwrap().len(),
            // the Memory chip's row order depends on a randomly seeded hash map (executor.rs:74-76): hand the trace over
            "memory_trace": prover.generate_traces(&runtime.record).into_iter().find(|(n, _)| n == "Memory")
                .map(|(_, t)| (0..t.height()).map(|r| cs(t.row(r))).collect::<Vec<_>>()),
        }));
    }
    out.insert("proofs".into(), Value::Object(proofs));
    let _ = (P, <F as FieldAlgebra>::ONE, <F as PrimeField32>::ORDER_U32);
    println!("{}", serde_json::to_string_pretty(&Value::Object(out)).unwrap());
}
