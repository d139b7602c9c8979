//! `bf-gpu` — the reference-side half of the drop-in boundary: safe wrappers over `bf-gpu-sys` implementing
//!
//!   * plug point #1  `MachineProver<KoalaBearPoseidon2, BfAir<KoalaBear>>` (crates/stark/src/prover.rs:27-150) as [`CudaProver`]:
//!     `setup`, `commit`, `open`, `prove`; traces, LDEs and Merkle trees stay in HBM behind opaque handles;
//!   * plug point #2a `TwoAdicSubgroupDft<KoalaBear>` (alias `Dft`, crates/stark/src/kb31_poseidon2.rs:30) as [`GpuDft`].
//!
//! Selected at compile time through `BfProverComponents::CoreProver` (crates/prover/src/components.rs:11-20):
//! `type CoreProver = bf_gpu::CudaProver<CoreSC, BfAir<<CoreSC as StarkGenericConfig>::Val>>;`
//!
//! NOT COMPILED in the backend's own CI (the build image has no Rust toolchain and Plonky3 is a git dependency that is not
//! vendored): written against the reference sources at the file:line cited on each item and against Plonky3 v0.1.0 as published.
//! The Python/ctypes mirror (`zkvm-brainfuck_b200/__init__.py`) drives the same symbols in the same order and is what the parity
//! tests run; every `unsafe` block here is one call of that table.
//!
//! Representation: `KoalaBear = MontyField31<..>` is `#[repr(transparent)]` over its Montgomery `u32` (the reference itself
//! transmutes `Vec<u32>` <-> `Vec<F>`, crates/core/machine/src/utils/mod.rs:105-114), so `BFGPU_REPR_MONTY` makes every `&[KoalaBear]`
//! an FFI buffer without a copy.
#![allow(clippy::missing_safety_doc)]

use core::ffi::{c_char, CStr};
use std::{ffi::CString, marker::PhantomData, ptr};

use bf_gpu_sys as sys;
use hashbrown::HashMap;
use p3_air::Air;
use p3_challenger::DuplexChallenger;
use p3_dft::TwoAdicSubgroupDft;
use p3_field::{extension::BinomialExtensionField, FieldAlgebra};
use p3_koala_bear::KoalaBear;
use p3_matrix::{dense::RowMajorMatrix, Matrix};
use p3_symmetric::Hash;

use bf_stark::{
    air::MachineAir, koala_bear_poseidon2::KoalaBearPoseidon2, AirOpenedValues, ChipOpenedValues, Com, DebugConstraintBuilder, MachineProof,
    MachineProver, MachineProvingKey, MachineRecord, ShardCommitment, ShardMainData, ShardOpenedValues, ShardProof, StarkGenericConfig,
    StarkMachine, StarkProvingKey, StarkVerifyingKey, Val,
};

type F = KoalaBear;
type EF = BinomialExtensionField<F, 4>;
type SC = KoalaBearPoseidon2;

// ---- context -----------------------------------------------------------------------------------------------------------------
/// `KoalaBearPoseidon2::new()` analogue (kb31_poseidon2.rs:73-85): owns the device context (streams, twiddles, block cache).
pub struct GpuCtx(*mut sys::bfgpu_ctx);
// every entry point serialises on the context's stream; the library is thread-safe w.r.t. distinct handles
unsafe impl Send for GpuCtx {}
unsafe impl Sync for GpuCtx {}

impl GpuCtx {
    pub fn new(device: i32) -> Self {
        let mut h = ptr::null_mut();
        let rc = unsafe { sys::bfgpu_ctx_create(device, &mut h) };
        let ctx = GpuCtx(h);
        ctx.check(rc);
        ctx.check(unsafe { sys::bfgpu_set_repr(ctx.0, sys::BFGPU_REPR_MONTY) });
        // default_fri_config (kb31_poseidon2.rs:54-64): log_blowup 1, $FRI_QUERIES or 84 queries, 16 PoW bits
        let q = std::env::var("FRI_QUERIES").ok().and_then(|v| v.parse().ok()).unwrap_or(84u32);
        ctx.check(unsafe { sys::bfgpu_set_fri_params(ctx.0, 1, q, 16) });
        ctx
    }
    /// The reference's only error type is the unit struct `CpuProverError` (prover.rs:166-168) and its callers `unwrap()`
    /// (crates/core/machine/src/utils/prove.rs:45): a failed call panics with the library's message.
    pub fn check(&self, rc: i32) {
        if rc != sys::BFGPU_OK {
            let msg = unsafe { CStr::from_ptr(sys::bfgpu_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("bfgpu error {rc}: {msg}");
        }
    }
}
impl Drop for GpuCtx {
    fn drop(&mut self) {
        unsafe { sys::bfgpu_ctx_destroy(self.0) }
    }
}

fn as_mat(m: &RowMajorMatrix<F>) -> sys::bfgpu_mat {
    sys::bfgpu_mat { data: m.values.as_ptr() as *const u32, rows: m.height() as u64, cols: m.width() as u64 }
}
fn words_to_field(w: Vec<u32>) -> Vec<F> {
    // MontyField31 is repr(transparent) over u32 and the words are Montgomery residues below p
    unsafe { core::mem::transmute::<Vec<u32>, Vec<F>>(w) }
}

// ---- plug point #2a: TwoAdicSubgroupDft ------------------------------------------------------------------------------------------
/// Replaces `type Dft = Radix2DitParallel<Val>` (kb31_poseidon2.rs:30).  Each call copies in and out (the trait returns a host
/// matrix); the machine prover below keeps everything on the device instead.
pub struct GpuDft(pub std::sync::Arc<GpuCtx>);
impl Clone for GpuDft {
    fn clone(&self) -> Self {
        GpuDft(self.0.clone())
    }
}
impl Default for GpuDft {
    fn default() -> Self {
        GpuDft(std::sync::Arc::new(GpuCtx::new(0)))
    }
}
impl TwoAdicSubgroupDft<F> for GpuDft {
    type Evaluations = RowMajorMatrix<F>;
    fn dft_batch(&self, mat: RowMajorMatrix<F>) -> Self::Evaluations {
        let mut out = vec![0u32; mat.values.len()];
        self.0.check(unsafe { sys::bfgpu_dft_batch(self.0 .0, &as_mat(&mat), out.as_mut_ptr()) });
        RowMajorMatrix::new(words_to_field(out), mat.width())
    }
    fn idft_batch(&self, mat: RowMajorMatrix<F>) -> RowMajorMatrix<F> {
        let mut out = vec![0u32; mat.values.len()];
        self.0.check(unsafe { sys::bfgpu_idft_batch(self.0 .0, &as_mat(&mat), out.as_mut_ptr()) });
        RowMajorMatrix::new(words_to_field(out), mat.width())
    }
    fn coset_lde_batch(&self, mat: RowMajorMatrix<F>, added_bits: usize, shift: F) -> Self::Evaluations {
        let mut out = vec![0u32; mat.values.len() << added_bits];
        let shift: u32 = unsafe { core::mem::transmute(shift) };
        self.0.check(unsafe { sys::bfgpu_coset_lde_batch(self.0 .0, &as_mat(&mat), added_bits as u32, shift, 0, out.as_mut_ptr()) });
        RowMajorMatrix::new(words_to_field(out), mat.width())
    }
}

// ---- challenger: DuplexChallenger's fields are public, the sponge is copied in and out -------------------------------------------
struct GpuChallenger<'a> {
    h: *mut sys::bfgpu_challenger,
    ctx: &'a GpuCtx,
}
impl<'a> GpuChallenger<'a> {
    fn import<P>(ctx: &'a GpuCtx, ch: &DuplexChallenger<F, P, 16, 8>) -> Self {
        let mut h = ptr::null_mut();
        ctx.check(unsafe { sys::bfgpu_challenger_create(ctx.0, &mut h) });
        let st = ch.sponge_state.as_ptr() as *const u32;
        ctx.check(unsafe {
            sys::bfgpu_challenger_import(h, st, ch.input_buffer.as_ptr() as *const u32, ch.input_buffer.len() as u32,
                                         ch.output_buffer.as_ptr() as *const u32, ch.output_buffer.len() as u32)
        });
        GpuChallenger { h, ctx }
    }
    fn export_into<P>(&self, ch: &mut DuplexChallenger<F, P, 16, 8>) {
        let (mut st, mut ib, mut ob) = ([0u32; 16], [0u32; 8], [0u32; 8]);
        let (mut ni, mut no) = (0u32, 0u32);
        self.ctx.check(unsafe { sys::bfgpu_challenger_export(self.h, st.as_mut_ptr(), ib.as_mut_ptr(), &mut ni, ob.as_mut_ptr(), &mut no) });
        ch.sponge_state = unsafe { core::mem::transmute::<[u32; 16], [F; 16]>(st) };
        ch.input_buffer = words_to_field(ib[..ni as usize].to_vec());
        ch.output_buffer = words_to_field(ob[..no as usize].to_vec());
    }
}
impl Drop for GpuChallenger<'_> {
    fn drop(&mut self) {
        unsafe { sys::bfgpu_challenger_free(self.h) }
    }
}

// ---- plug point #1: MachineProver ---------------------------------------------------------------------------------------------------
/// `DeviceProvingKey`: preprocessed traces + their LDE + Merkle tree resident in HBM, plus the host key the
/// reference's verifier-side code still reads.
pub struct GpuProvingKey {
    h: *mut sys::bfgpu_pk,
    pub host: StarkProvingKey<SC>,
}
unsafe impl Send for GpuProvingKey {}
unsafe impl Sync for GpuProvingKey {}
impl Drop for GpuProvingKey {
    fn drop(&mut self) {
        unsafe { sys::bfgpu_pk_free(self.h) }
    }
}
impl MachineProvingKey<SC> for GpuProvingKey {
    fn preprocessed_commit(&self) -> Com<SC> {
        self.host.commit.clone()
    }
    fn observe_into(&self, challenger: &mut <SC as StarkGenericConfig>::Challenger) {
        self.host.observe_into(challenger) // prover.rs:595-601: the commitment, then seven zeros
    }
}

/// `DeviceProverData` of the main commitment (`ShardMainData::main_data`): traces, LDEs and tree on the device.
pub struct GpuShard(*mut sys::bfgpu_shard);
unsafe impl Send for GpuShard {}
unsafe impl Sync for GpuShard {}
impl Drop for GpuShard {
    fn drop(&mut self) {
        unsafe { sys::bfgpu_shard_free(self.0) }
    }
}

pub struct CudaProver<A> {
    machine: StarkMachine<SC, A>,
    ctx: GpuCtx,
    _a: PhantomData<A>,
}

#[derive(Debug, Clone, Copy)]
pub struct CudaProverError;
impl core::fmt::Display for CudaProverError {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "CudaProverError")
    }
}
impl std::error::Error for CudaProverError {}

fn named(traces: &[(String, RowMajorMatrix<F>)]) -> (Vec<CString>, Vec<*const c_char>, Vec<sys::bfgpu_mat>) {
    let names: Vec<CString> = traces.iter().map(|(n, _)| CString::new(n.as_str()).unwrap()).collect();
    let ptrs = names.iter().map(|c| c.as_ptr()).collect();
    let mats = traces.iter().map(|(_, m)| as_mat(m)).collect();
    (names, ptrs, mats)
}

impl<A> MachineProver<SC, A> for CudaProver<A>
where
    A: MachineAir<F> + 'static + Send + Sync,
{
    /// The opened traces are never materialised on the host: the device handle stands in for them.
    type DeviceMatrix = RowMajorMatrix<F>;
    type DeviceProverData = GpuShard;
    type DeviceProvingKey = GpuProvingKey;
    type Error = CudaProverError;

    fn new(machine: StarkMachine<SC, A>) -> Self {
        Self { machine, ctx: GpuCtx::new(0), _a: PhantomData }
    }
    fn machine(&self) -> &StarkMachine<SC, A> {
        &self.machine
    }
    /// `StarkMachine::setup` (machine.rs:154-224) for the verifying key and chip bookkeeping, then the same preprocessed traces are
    /// committed on the device (the two commitments are equal: tests/test_gpu_prove_parity.py pins it against the CPU oracle).
    fn setup(&self, program: &A::Program) -> (Self::DeviceProvingKey, StarkVerifyingKey<SC>) {
        let (pk, vk) = self.machine.setup(program);
        (self.pk_to_device(&pk), vk)
    }
    fn pk_to_device(&self, pk: &StarkProvingKey<SC>) -> Self::DeviceProvingKey {
        let mut by_index: Vec<(&String, &usize)> = pk.chip_ordering.iter().collect();
        by_index.sort_by_key(|(_, i)| **i);
        let traces: Vec<(String, RowMajorMatrix<F>)> = by_index.iter().map(|(n, i)| ((*n).clone(), pk.traces[**i].clone())).collect();
        let (_keep, names, mats) = named(&traces);
        let mut commit = [0u32; 8];
        let mut h = ptr::null_mut();
        self.ctx.check(unsafe { sys::bfgpu_machine_setup(self.ctx.0, names.as_ptr(), mats.as_ptr(), mats.len() as i32, commit.as_mut_ptr(), &mut h) });
        let commit: [F; 8] = unsafe { core::mem::transmute(commit) };
        assert_eq!(Hash::<F, F, 8>::from(commit), pk.commit, "device preprocessed commitment differs from the host one");
        GpuProvingKey { h, host: pk.clone() }
    }
    fn pk_to_host(&self, pk: &Self::DeviceProvingKey) -> StarkProvingKey<SC> {
        pk.host.clone()
    }
    /// prover.rs:209-236.  The library sorts by (height desc, name) itself (`:214`) and keeps traces + LDEs + tree in HBM.
    fn commit(&self, traces: Vec<(String, RowMajorMatrix<F>)>) -> ShardMainData<SC, Self::DeviceMatrix, Self::DeviceProverData> {
        let (_keep, names, mats) = named(&traces);
        let mut root = [0u32; 8];
        let mut h = ptr::null_mut();
        self.ctx.check(unsafe { sys::bfgpu_machine_commit(self.ctx.0, names.as_ptr(), mats.as_ptr(), mats.len() as i32, root.as_mut_ptr(), &mut h) });
        let mut sorted: Vec<&(String, RowMajorMatrix<F>)> = traces.iter().collect();
        sorted.sort_by_key(|(name, t)| (core::cmp::Reverse(t.height()), name.clone()));
        let chip_ordering: HashMap<String, usize> = sorted.iter().enumerate().map(|(i, (n, _))| (n.clone(), i)).collect();
        let root: [F; 8] = unsafe { core::mem::transmute(root) };
        // `traces` of ShardMainData is only read for heights by callers outside the prover (prover/src/lib.rs): keep empty
        // same-height placeholders out of it; the device owns the real ones
        ShardMainData::new(Vec::new(), Hash::from(root), GpuShard(h), chip_ordering)
    }
    /// prover.rs:242-553 in one call: LogUp traces, permutation and quotient commitments, `pcs.open`, on the device, with the caller's
    /// transcript advanced exactly as the reference does.
    fn open(&self, pk: &Self::DeviceProvingKey, data: ShardMainData<SC, Self::DeviceMatrix, Self::DeviceProverData>,
            challenger: &mut <SC as StarkGenericConfig>::Challenger) -> Result<ShardProof<SC>, Self::Error> {
        let ch = GpuChallenger::import(&self.ctx, challenger);
        let mut raw = ptr::null_mut();
        self.ctx.check(unsafe { sys::bfgpu_machine_open(self.ctx.0, pk.h, data.main_data.0, ch.h, -1, &mut raw) });
        ch.export_into(challenger);
        let n = unsafe { sys::bfgpu_shard_proof_size(raw) } as usize;
        let mut words = vec![0u32; n];
        self.ctx.check(unsafe { sys::bfgpu_shard_proof_read(raw, words.as_mut_ptr()) });
        unsafe { sys::bfgpu_shard_proof_free(raw) };
        Ok(decode_shard_proof(&words, &pk.host))
    }
    /// prover.rs:560-582: observe the key, generate traces, commit, open on a clone of the challenger.
    fn prove(&self, pk: &Self::DeviceProvingKey, record: &mut A::Record, challenger: &mut <SC as StarkGenericConfig>::Challenger)
        -> Result<MachineProof<SC>, Self::Error>
    where
        A: for<'a> Air<DebugConstraintBuilder<'a, Val<SC>, <SC as StarkGenericConfig>::Challenge>>,
    {
        pk.observe_into(challenger);
        self.machine.generate_dependencies(record);
        let traces = self.generate_traces(record);
        let data = self.commit(traces);
        let proof = self.open(pk, data, &mut challenger.clone())?;
        Ok(MachineProof { shard_proof: proof })
    }
}

// ---- the flat proof -> ShardProof (layout: include/bfgpu.h at bfgpu_machine_open / bfgpu_pcs_open) ---------------------------------------
/// Decodes by writing the bincode image the library produces for exactly this purpose (`bfgpu_shard_proof_to_bincode`: the bytes
/// `bincode::serialize(&MachineProof)` would write, field elements as Montgomery words = p3-monty-31's serde form) and letting serde
/// build the nested value: no second copy of the layout to keep in sync.
pub fn decode_shard_proof(words: &[u32], pk: &StarkProvingKey<SC>) -> ShardProof<SC> {
    let mut by_index: Vec<(&String, &usize)> = pk.chip_ordering.iter().collect();
    by_index.sort_by_key(|(_, i)| **i);
    let names: Vec<CString> = by_index.iter().map(|(n, _)| CString::new(n.as_str()).unwrap()).collect();
    let name_ptrs: Vec<*const c_char> = names.iter().map(|c| c.as_ptr()).collect();
    let logs: Vec<u32> = by_index.iter().map(|(_, i)| pk.traces[**i].height().trailing_zeros()).collect();
    let mut len = 0u64;
    let mut err = [0 as c_char; 256];
    let call = |out: *mut u8, cap: u64, len: &mut u64, err: &mut [c_char; 256]| unsafe {
        sys::bfgpu_shard_proof_to_bincode(name_ptrs.as_ptr(), logs.as_ptr(), logs.len() as i32, words.as_ptr(), words.len() as u64,
                                          sys::BFGPU_REPR_MONTY, 1, 1, out, cap, len, err.as_mut_ptr(), 256)
    };
    assert_eq!(call(ptr::null_mut(), 0, &mut len, &mut err), sys::BFGPU_OK);
    let mut bytes = vec![0u8; len as usize];
    assert_eq!(call(bytes.as_mut_ptr(), len, &mut len, &mut err), sys::BFGPU_OK);
    let proof: MachineProof<SC> = bincode::deserialize(&bytes).expect("library-written bincode image of MachineProof");
    proof.shard_proof
}

/// `BfProver::verify` (crates/prover/src/verify.rs:10-36) by the library's host verifier, on the flat words of `bfgpu_machine_open`:
/// the Cpu chip must be present with a log degree of at most 22, then `Verifier::verify_shard`.  For deployments that verify without
/// the Rust verifier (and for comparing verdicts with it); `Err` carries the reference's error name
/// (`MissingCpuInFirstShard`, `CpuLogDegreeTooLarge: n`, `InvalidShardProof: <VerificationError>`).
pub fn verify_core_proof(vk: &StarkVerifyingKey<SC>, pk: &StarkProvingKey<SC>, words: &[u32], num_queries: u32) -> Result<(), String> {
    let mut by_index: Vec<(&String, &usize)> = pk.chip_ordering.iter().collect();
    by_index.sort_by_key(|(_, i)| **i);
    let names: Vec<CString> = by_index.iter().map(|(n, _)| CString::new(n.as_str()).unwrap()).collect();
    let name_ptrs: Vec<*const c_char> = names.iter().map(|c| c.as_ptr()).collect();
    let logs: Vec<u32> = by_index.iter().map(|(_, i)| pk.traces[**i].height().trailing_zeros()).collect();
    // Hash<F, F, 8> is [F; 8] and F is repr(transparent) over its Montgomery word
    let commit: [u32; 8] = unsafe { core::mem::transmute_copy(&vk.commit) };
    let mut err = [0 as c_char; 256];
    let rc = unsafe {
        sys::bfgpu_verify_core_proof(commit.as_ptr(), name_ptrs.as_ptr(), logs.as_ptr(), logs.len() as i32, words.as_ptr(), words.len() as u64,
                                     sys::BFGPU_REPR_MONTY, 1, num_queries, 16, ptr::null(), 0, err.as_mut_ptr(), 256)
    };
    if rc == sys::BFGPU_OK {
        Ok(())
    } else {
        Err(unsafe { CStr::from_ptr(err.as_ptr()) }.to_string_lossy().into_owned())
    }
}

/// The pieces of the nested proof, spelled out once for readers who want to see the mapping (not used by `decode_shard_proof`):
/// flat words [0, 24) = the three commitments; word 24 = number of chips; then per chip (index, log_degree, cumulative sum);
/// then the opened values (preprocessed in proving-key order, main, permutation, quotient), the FRI commit-phase commitments,
/// final polynomial, proof-of-work witness, and per query the input openings and commit-phase openings.
#[allow(dead_code)]
fn layout_reference(_c: ShardCommitment<Com<SC>>, _o: ShardOpenedValues<EF>, _a: AirOpenedValues<EF>, _v: ChipOpenedValues<EF>) {}

// silence "unused" for items only mentioned in docs
#[allow(dead_code)]
fn _uses<R: MachineRecord>() -> F {
    F::ZERO
}
