//! `extern "C"` view of `include/miden_b200.h` (the C ABI of the backend).  Layouts are asserted at start-up against
//! `mdn_abi_layout` (see [`check_layout`]), so a header change cannot silently desynchronise the binding.
//!
//! `Felt` is `#[repr(transparent)]` over p3 `Goldilocks`, itself a canonical `u64`
//! (reference crates/field/src/native/mod.rs:58), so `RowMajorMatrix<Felt>::values.as_ptr() as *const u64` is the
//! `MdnMatrix.values` pointer without a copy; `QuadFelt` flattens to two `u64` (c0, c1) exactly as `flatten_to_base`
//! does (crates/lifted-stark/src/prover/mod.rs:403-409).
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_longlong, c_void};

pub const MDN_OK: c_int = 0;
pub const MDN_ERR_INVALID_ARG: c_int = -1;
pub const MDN_ERR_DOMAIN: c_int = -2;
pub const MDN_ERR_CUDA: c_int = -3;
pub const MDN_ERR_UNSUPPORTED: c_int = -4;
pub const MDN_ERR_AUX_BUILDER: c_int = -5;
pub const MDN_ERR_NO_DEVICE: c_int = -6;
pub const MDN_ERR_EXTERNAL_ASSERTION: c_int = -7;
pub const MDN_FLAG_DEVICE_TRACES: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnPcsParams {
    pub log_blowup: u32,
    pub log_folding_arity: u32,
    pub log_final_degree: u32,
    pub folding_pow_bits: u32,
    pub deep_pow_bits: u32,
    pub num_queries: u32,
    pub query_pow_bits: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnChallenger {
    pub sponge_state: [u64; 12],
    pub input_buffer: [u64; 8],
    pub input_len: u32,
    pub output_len: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnLookup {
    pub num_columns: u32,
    pub program_words: u32,
    pub program: *const u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnAir {
    pub width: u32,
    pub aux_width: u32,
    pub num_aux_values: u32,
    pub num_randomness: u32,
    pub log_quotient_degree: u32,
    pub program_words: u32,
    pub program: *const u32,
    pub periodic_values: *const u64,
    pub num_periodic_columns: u32,
    pub log_max_period: u32,
    pub preprocessed_width: u32,
    pub lookup: *const MdnLookup,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnMatrix {
    pub values: *const u64,
    pub log_height: u32,
    pub width: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnStatement {
    pub airs: *const MdnAir,
    pub n_airs: u32,
    pub public_values: *const u64,
    pub n_public_values: u32,
    pub observe_felts: *const u64,
    pub n_observe_felts: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct MdnProof {
    pub log_trace_heights: *const u8,
    pub n_heights: usize,
    pub fields: *const u64,
    pub n_fields: usize,
    pub commitments: *const u64,
    pub n_commitments: usize,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct MdnTimings {
    pub h2d_transpose: f32,
    pub commit_main: f32,
    pub commit_aux: f32,
    pub evaluate_constraints: f32,
    pub commit_quotient: f32,
    pub open: f32,
    pub total: f32,
    pub lde_main: f32,
    pub hash_main: f32,
    pub kernel_ms: [f32; 10],
    pub kernel_regions: [u32; 10],
    pub kernel_launches: u64,
    pub permutations: u64,
    pub leaf_hash_bytes: f64,
    pub ntt_bytes: f64,
}

/// `LiftedAir::build_aux_trace` trampoline (`mdn_aux_builder`).
pub type MdnAuxBuilder = Option<
    unsafe extern "C" fn(
        ctx: *mut c_void,
        instance: u32,
        main: *const MdnMatrix,
        randomness: *const u64,
        aux_out: *mut u64,
        aux_values: *mut u64,
    ) -> c_int,
>;
/// Bootstrap transport of a proof split over several GPUs (`mdn_allgather_fn`).
pub type MdnAllgather =
    Option<unsafe extern "C" fn(ctx: *mut c_void, send: *const u64, recv: *mut u64, n_u64: usize) -> c_int>;
/// `Statement::eval_external` trampoline (`mdn_external_check`).
pub type MdnExternalCheck = Option<
    unsafe extern "C" fn(
        ctx: *mut c_void,
        challenges: *const u64,
        n_challenges: u32,
        aux_values: *const *const u64,
        n_aux_values: *const u32,
        log_trace_heights: *const u8,
        n_airs: u32,
        failed_assertion: *mut u32,
    ) -> c_int,
>;

pub enum MdnSession {}

/// `mdn_hash_kind` (include/miden_b200.h): which `miden_air::config` constructor the session follows.
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum HashKind {
    Poseidon2 = 0,
    Blake3 = 1,
    Keccak = 2,
    /// `rpo_config`: the duplex challenger passed to `prove` must be built over `RpoPermutation256`
    Rpo = 3,
    /// `rpx_config`
    Rpx = 4,
}

/// `mdn_hash_challenger`: p3 `HashChallenger<u8, H, 32>`'s two buffers.
#[repr(C)]
pub struct MdnHashChallenger {
    pub input_buffer: *const u8,
    pub input_len: usize,
    pub output_buffer: *const u8,
    pub output_len: usize,
}

unsafe extern "C" {
    pub fn mdn_session_create(params: *const MdnPcsParams, cuda_device: c_int, out: *mut *mut MdnSession) -> c_int;
    pub fn mdn_session_destroy(s: *mut MdnSession);
    pub fn mdn_last_error(s: *const MdnSession) -> *const c_char;
    pub fn mdn_prove(
        s: *mut MdnSession,
        st: *const MdnStatement,
        traces: *const MdnMatrix,
        challenger: *const MdnChallenger,
        build_aux: MdnAuxBuilder,
        aux_ctx: *mut c_void,
        flags: u32,
        out: *mut MdnProof,
    ) -> c_int;
    pub fn mdn_prove_begin(
        s: *mut MdnSession,
        st: *const MdnStatement,
        traces: *const MdnMatrix,
        challenger: *const MdnChallenger,
        flags: u32,
        main_root: *mut u64,
        randomness_out: *mut u64,
    ) -> c_int;
    pub fn mdn_prove_commit_aux(
        s: *mut MdnSession,
        aux: *const MdnMatrix,
        aux_values: *const *const u64,
        aux_root: *mut u64,
    ) -> c_int;
    pub fn mdn_prove_finish(s: *mut MdnSession, out: *mut MdnProof) -> c_int;
    pub fn mdn_session_set_preprocessed(
        s: *mut MdnSession,
        st: *const MdnStatement,
        preprocessed: *const MdnMatrix,
        commitment_out: *mut u64,
    ) -> c_int;
    pub fn mdn_session_set_shard(s: *mut MdnSession, rank: u32, world: u32, f: MdnAllgather, ctx: *mut c_void) -> c_int;
    pub fn mdn_session_set_external_check(s: *mut MdnSession, f: MdnExternalCheck, ctx: *mut c_void) -> c_int;
    pub fn mdn_session_set_hash(s: *mut MdnSession, kind: c_int) -> c_int;
    pub fn mdn_session_set_hash_challenger(s: *mut MdnSession, c: *const MdnHashChallenger) -> c_int;
    pub fn mdn_session_set_jit(s: *mut MdnSession, min_nodes: u32) -> c_int;
    pub fn mdn_jit_compile_check(program: *const u32, program_words: u32, err: *mut *const c_char) -> c_longlong;
    pub fn mdn_get_timings(s: *mut MdnSession, out: *mut MdnTimings) -> c_int;
    pub fn mdn_abi_layout(out: *mut u32, cap: usize) -> usize;
}

/// Compare this file's `#[repr(C)]` layouts with the library's own `sizeof` / `offsetof` table.
pub fn check_layout() -> Result<(), String> {
    use core::mem::{offset_of, size_of};
    let mine: [u32; 14] = [
        size_of::<MdnPcsParams>() as u32,
        size_of::<MdnChallenger>() as u32,
        size_of::<MdnLookup>() as u32,
        size_of::<MdnAir>() as u32,
        offset_of!(MdnAir, program) as u32,
        offset_of!(MdnAir, periodic_values) as u32,
        offset_of!(MdnAir, preprocessed_width) as u32,
        offset_of!(MdnAir, lookup) as u32,
        size_of::<MdnMatrix>() as u32,
        size_of::<MdnStatement>() as u32,
        size_of::<MdnProof>() as u32,
        size_of::<MdnTimings>() as u32,
        offset_of!(MdnTimings, kernel_ms) as u32,
        offset_of!(MdnTimings, permutations) as u32,
    ];
    let mut theirs = [0u32; 14];
    let n = unsafe { mdn_abi_layout(theirs.as_mut_ptr(), theirs.len()) };
    if n != mine.len() || mine != theirs {
        return Err(format!("libmiden_b200 ABI layout mismatch: binding {mine:?}, library {theirs:?}"));
    }
    Ok(())
}
