/*
 * miden_b200.h -- C ABI of the Blackwell (sm_100a) STARK proving backend for Miden VM.
 *
 * Drop-in boundary: everything beneath `miden_prover::prove_stark()` (reference
 * prover/src/lib.rs:317-355), i.e. `ProverInstance::new(config, statement, None)?.prove(challenger)`
 * (crates/lifted-stark/src/prover/mod.rs:139,157,230-578).  The Rust host keeps trace generation,
 * the AIR (lowered once to an op-list, see mdn_air.program), the LogUp aux-trace builder
 * (a callback) and wincode serialisation of the returned streams.  INTEGRATION.md shows the
 * Rust `extern "C"` binding.
 *
 * Conventions
 *   - every field element is a canonical Goldilocks u64 (< 2^64 - 2^32 + 1); quadratic-extension
 *     elements are two consecutive u64 (c0, c1), u^2 = 7 -- the layout of the reference's
 *     `Felt` / `QuadFelt` (crates/field/src/native/mod.rs:58, flatten_to_base order).
 *   - matrices are row-major exactly like p3 `RowMajorMatrix<Felt>`; the library transposes on
 *     the device.
 *   - functions return 0 on success and a negative mdn_status otherwise; the message is
 *     available through mdn_last_error().  Errors mirror `ProverError` / `ExecutionError::
 *     ProvingError(String)` (prover/mod.rs:582-596, prover/src/lib.rs:336-345).
 *   - a session is bound to one CUDA device and is not thread-safe; sessions are independent.
 *   - there is NO CPU fallback: if no CUDA device is usable, mdn_session_create fails.
 */
#ifndef MIDEN_B200_H
#define MIDEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mdn_session mdn_session;

typedef enum {
    MDN_OK = 0,
    MDN_ERR_INVALID_ARG = -1,      /* malformed statement / trace shape (InstanceError) */
    MDN_ERR_DOMAIN = -2,           /* DomainError: LDE order too large, degree > blowup */
    MDN_ERR_CUDA = -3,             /* device error (message carries the CUDA string) */
    MDN_ERR_UNSUPPORTED = -4,      /* e.g. log_blowup > 4, > 1024 live constraint values */
    MDN_ERR_AUX_BUILDER = -5,      /* aux-trace callback failed */
    MDN_ERR_NO_DEVICE = -6,
    MDN_ERR_EXTERNAL_ASSERTION = -7,   /* ProverError::ExternalAssertionFailed / ::Reduction (prover/mod.rs:383-395) */
} mdn_status;

/* PcsParams::new(log_blowup, log_folding_arity, log_final_degree, folding_pow_bits,
 * deep_pow_bits, num_queries, query_pow_bits) -- crates/lifted-stark/src/pcs/params.rs:53-99.
 * Miden production values: air/src/config.rs:55-67 = {3, 2, 7, 4, 12, 27, 16}. */
typedef struct {
    uint32_t log_blowup;
    uint32_t log_folding_arity;
    uint32_t log_final_degree;
    uint32_t folding_pow_bits;
    uint32_t deep_pow_bits;
    uint32_t num_queries;
    uint32_t query_pow_bits;
} mdn_pcs_params;

/* p3 `DuplexChallenger<Felt, Poseidon2, 12, 8>` state (public fields used at
 * air/src/config.rs:264-271): sponge_state, input_buffer, output_buffer.  output_len counts the
 * unread rate elements; the next sample returns sponge_state[output_len - 1]. */
typedef struct {
    uint64_t sponge_state[12];
    uint64_t input_buffer[8];
    uint32_t input_len;
    uint32_t output_len;
} mdn_challenger;

/* Lowering of `LookupAir::eval` (air/src/lookup/builder.rs) for AIRs whose `build_aux_trace` is
 * `build_logup_aux_trace` (air/src/lookup/aux_builder.rs:49-97).  When `mdn_air.lookup` is set, the aux trace
 * and its committed final are built ON THE DEVICE from the resident main trace right after the randomness is
 * sampled -- no host callback, no aux-trace upload for that AIR.
 *   program words[0..5) = { 0x504B4C4D ("MLKP"), 1, n_nodes, n_interactions, n_consts }
 *   nodes        : as in mdn_air.program, restricted to what a LookupBuilder exposes: MAIN, PUBLIC,
 *                  CHALLENGE (0 = alpha, 1 = beta), CONST, EXT_CONST, ADD, SUB, MUL, NEG, PERIODIC
 *   interactions : 4 words { aux column, flag node | 0xFFFFFFFF, multiplicity node, denominator node } --
 *                  one per `LookupGroup::insert` / `LookupBatch::insert` after `LookupMessage::encode`; it
 *                  contributes multiplicity/denominator to its column on every row where the (0/1) flag is
 *                  non-zero (air/src/lookup/prover.rs:338-362,421-444)
 *   consts       : (lo, hi) word pairs
 * Output (aux_builder.rs:1-20,215-268): f_c(r) = sum of m/d over column c's active interactions at row r;
 * aux[r][c] = f_c(r) for c > 0; aux[r][0] = sum_{r' < r} sum_c f_c(r'); aux value 0 = the total over all rows.
 * Requires aux_width == num_columns (<= 16) and num_aux_values == 1; a zero denominator is an error. */
typedef struct {
    uint32_t num_columns;           /* LookupAir::num_columns() */
    uint32_t program_words;
    const uint32_t* program;
} mdn_lookup;

/* One AIR of the MultiAir (crates/lifted-air/src/air.rs:47-202 `LiftedAir`): shape + constraint
 * program.  `program` is the op-list lowering of `air.eval()`:
 *   words[0..5) = { 0x5249414D ("MAIR"), 1, n_nodes, n_constraints, n_consts }
 *   nodes: 3 words each { op, a, b }; constraints: node ids in emission order;
 *   consts: (lo, hi) word pairs.
 * Ops (leaf vocabulary of crates/ace-codegen/src/dag/lower.rs:109-210):
 *   0 MAIN(a=row offset 0|1, b=col)  1 AUX(a=offset, b=EF col)  2 PUBLIC(a)  3 CHALLENGE(a)
 *   4 AUX_VALUE(a)  5 IS_FIRST_ROW  6 IS_LAST_ROW  7 IS_TRANSITION  8 CONST(a)  9 EXT_CONST(a)
 *   10 ADD(a,b)  11 SUB(a,b)  12 MUL(a,b)  13 NEG(a)  14 PERIODIC(a=periodic column)
 *   15 PREPROCESSED(a=row offset 0|1, b=col)
 * Constraints are folded as acc <- acc*alpha + C_k in emission order
 * (crates/lifted-stark/src/verifier/constraints.rs:83,108). */
typedef struct {
    uint32_t width;                 /* BaseAir::width -- main trace columns */
    uint32_t aux_width;             /* LiftedAir::aux_width -- EF columns */
    uint32_t num_aux_values;        /* LiftedAir::num_aux_values */
    uint32_t num_randomness;        /* LiftedAir::num_randomness */
    uint32_t log_quotient_degree;   /* domain.rs:585-598 (symbolic degree analysis stays host-side) */
    uint32_t program_words;
    const uint32_t* program;
    /* `BaseAir::periodic_columns_matrix()` (crates/lifted-stark/src/prover/periodic.rs:49-98): row-major
     * (1 << log_max_period) x num_periodic_columns, every column repeated to the maximum period.
     * NULL / 0 when the AIR has no periodic columns. */
    const uint64_t* periodic_values;
    uint32_t num_periodic_columns;
    uint32_t log_max_period;
    uint32_t preprocessed_width;    /* BaseAir::preprocessed_width(): 0 = the AIR declares no preprocessed columns */
    const mdn_lookup* lookup;       /* NULL: the aux trace comes from the host (mdn_aux_builder / commit_aux) */
} mdn_air;

/* p3 RowMajorMatrix<Felt>: `values` has (1 << log_height) * width entries.  With
 * MDN_FLAG_DEVICE_TRACES `values` is a device pointer on the session's device. */
typedef struct {
    const uint64_t* values;
    uint32_t log_height;
    uint32_t width;
} mdn_matrix;

/* crates/lifted-air/src/statement.rs `Statement`: AIRs in instance order, shared air_inputs, and
 * the exact felts `Statement::observe` absorbs (AIR-specific, e.g. air/src/lib.rs:817-847). */
typedef struct {
    const mdn_air* airs;
    uint32_t n_airs;
    const uint64_t* public_values;
    uint32_t n_public_values;
    const uint64_t* observe_felts;
    uint32_t n_observe_felts;
} mdn_statement;

/* `LiftedAir::build_aux_trace(main, air_inputs, aux_inputs, challenges)` (prover/mod.rs:357-381),
 * called once per AIR in instance order after the main root is observed.
 *   randomness : 2 * num_randomness u64
 *   aux_out    : (1 << log_height) x (2 * aux_width) row-major, EF flattened to base
 *   aux_values : 2 * num_aux_values u64
 * Return 0 on success.  A NULL builder means all-zero aux traces and values
 * (crates/lifted-stark/src/testing/airs/miden.rs:84-94). */
typedef int (*mdn_aux_builder)(void* ctx, uint32_t instance, const mdn_matrix* main,
                               const uint64_t* randomness, uint64_t* aux_out, uint64_t* aux_values);

/* `StarkProofData { log_trace_heights, transcript: TranscriptData { fields, commitments } }`
 * (crates/lifted-stark/src/proof.rs:57-63, crates/stark-transcript/src/data.rs:11-15).
 * Memory is owned by the session and valid until the next prove on it or its destruction. */
typedef struct {
    const uint8_t* log_trace_heights;
    size_t n_heights;
    const uint64_t* fields;
    size_t n_fields;
    const uint64_t* commitments;    /* 4 u64 per commitment */
    size_t n_commitments;
} mdn_proof;

enum {
    MDN_FLAG_DEVICE_TRACES = 1u,    /* trace matrices already resident in device memory */
};

/* ---- session ------------------------------------------------------------------------------ */
int mdn_session_create(const mdn_pcs_params* params, int cuda_device, mdn_session** out);
void mdn_session_destroy(mdn_session* s);
const char* mdn_last_error(const mdn_session* s);   /* s may be NULL: last create error */

/* ---- one proof on several GPUs ---------------------------------------------------------------------
 * One process per GPU of one NVLink/NVSwitch box; every rank calls mdn_prove (or the staged calls) with the SAME
 * statement / traces / challenger / flags and every rank returns the byte-identical proof.  The work of the ONE proof
 * is partitioned (the reference has no counterpart: it is single-process rayon; the loops that are split are
 * prover/commit.rs:142-180, lmcs/lifted_tree.rs:394-406, prover/constraints/mod.rs:246-259, prover/quotient.rs:163,
 * pcs/deep/prover.rs:214-312, pcs/fri/prover.rs:137-211):
 *   - rank g owns LDE cosets [g*B/G, (g+1)*B/G) of every committed column: forward coset NTTs, leaf sponge,
 *     constraint evaluation, quotient-chunk interpolation, DEEP quotient and FRI folds of those cosets;
 *   - rank g owns the Merkle sub-tree over leaves [g*L/G, (g+1)*L/G) of every commitment (input and FRI trees).
 * Data crosses ranks as peer-memory stores inside the producing kernels (CUDA IPC mappings of each rank's proof
 * arena over NVLink: leaf digests to the sub-tree owner, sub-roots / quotient chunk coefficients / the first small
 * FRI layer / opened values to every rank), ordered by a device-side flag barrier; there is no host round trip and
 * no library collective on the data path.  `fn` is only the bootstrap transport: it must gather `n_u64` words from
 * every rank into `recv` (rank-major) -- e.g. torch.distributed.all_gather -- and carries the 64-byte CUDA IPC
 * handles when an arena slab is created (first proof of a shape) and one word of rendezvous when slabs are
 * released.  world must be a power of two <= min(8, 2^log_blowup); world = 1 turns the partition off.  Collective:
 * every rank must call it at the same point.  The preprocessed bundle (mdn_session_set_preprocessed) and the
 * utility entry points (mdn_coset_lde_batch, mdn_lmcs_commit) are not partitioned. */
typedef int (*mdn_allgather_fn)(void* ctx, const uint64_t* send, uint64_t* recv, size_t n_u64);
int mdn_session_set_shard(mdn_session* s, uint32_t rank, uint32_t world, mdn_allgather_fn fn, void* ctx);

/* ---- Statement::eval_external (crates/lifted-air/src/air.rs:272-288, statement.rs:94-110) -------------------
 * Cross-AIR assertions are host code of the statement.  When a callback is installed the prover calls it once per
 * proof, after the aux traces are built -- on the host or on the device (mdn_air.lookup) -- and BEFORE the aux
 * commitment, exactly where the reference evaluates them (prover/mod.rs:383-395):
 *   challenges  : the shared randomness pool, 2 u64 per EF challenge          aux_values[i] : AIR i's aux values
 *   (instance order, 2 u64 per EF value, n_aux_values[i] of them)            log_trace_heights : instance order
 * Return 0 when every assertion evaluates to zero; > 0 with *failed_assertion = k for
 * `ProverError::ExternalAssertionFailed { assertion: k }`; < 0 for a `ReductionError`.  mdn_prove* then returns
 * MDN_ERR_EXTERNAL_ASSERTION and nothing of the aux phase is committed.  NULL (default) = no assertions. */
typedef int (*mdn_external_check)(void* ctx, const uint64_t* challenges, uint32_t n_challenges,
                                  const uint64_t* const* aux_values, const uint32_t* n_aux_values,
                                  const uint8_t* log_trace_heights, uint32_t n_airs, uint32_t* failed_assertion);
int mdn_session_set_external_check(mdn_session* s, mdn_external_check fn, void* ctx);

/* ---- STARK hash configuration (miden_air::config, air/src/config.rs) ----------------------------------------
 * MDN_HASH_POSEIDON2 (default): `poseidon2_config` (:241-273) -- StatefulSponge<Poseidon2, 12, 8, 4> leaves (alignment 8),
 *   TruncatedPermutation nodes, DuplexChallenger; the pre-bound challenger is the `mdn_challenger` argument of mdn_prove*.
 * MDN_HASH_BLAKE3: `blake3_256_config` (:276-307), the CLI's default hasher (miden-vm/src/cli/prove.rs:51) --
 *   ChainingHasher<Blake3> leaves (state <- blake3(state || little-endian u64 of every felt of the row); alignment 1, so
 *   opened rows and the OOD evaluation lists carry no zero padding), blake3(left || right) nodes,
 *   SerializingChallenger64<Felt, HashChallenger<u8, Blake3, 32>>.  Commitments are 32-byte digests carried as four
 *   little-endian u64 per commitment in `mdn_proof.commitments`.  The pre-bound challenger is the HashChallenger state
 *   installed with mdn_session_set_hash_challenger (after `config.challenger()` + `observe_protocol_params` that is the
 *   input buffer: 32 bytes of relation digest + 8 parameter felts as little-endian u64, and an empty output buffer); the
 *   `challenger` argument of mdn_prove* is ignored and may be NULL.  A preprocessed bundle belongs to the hash it was
 *   committed under.
 * MDN_HASH_KECCAK: `keccak_config` (:309-353) -- SerializingStatefulSponge<StatefulSponge<KeccakF, 25, 17, 4>> leaves (the
 *   overwrite-mode sponge over the canonical u64 of every felt; alignment 17: opened rows and the OOD lists are zero-padded
 *   to multiples of 17), PaddingFreeSponge<KeccakF, 25, 17, 4> nodes, SerializingChallenger64<Felt, HashChallenger<u8,
 *   Keccak256Hash, 32>>.  Digests are four u64 lanes; the challenger is installed like the Blake3 one (its input buffer must
 *   be whole 64-bit words, which `observe_slice(&relation_digest)` + `observe_protocol_params` always gives).
 * MDN_HASH_RPO / MDN_HASH_RPX: `rpo_config` / `rpx_config` (:225-248) -- the Poseidon2 configuration with the permutation
 *   replaced (`alg_config<P>` is generic in it, :255-273): same LMCS, alignment, duplex challenger (pass the `mdn_challenger`
 *   built with that permutation) and proof layout.  Functional coverage: an RPO permutation costs ~9x a Poseidon2 one. */
typedef enum { MDN_HASH_POSEIDON2 = 0, MDN_HASH_BLAKE3 = 1, MDN_HASH_KECCAK = 2, MDN_HASH_RPO = 3, MDN_HASH_RPX = 4 } mdn_hash_kind;
int mdn_session_set_hash(mdn_session* s, mdn_hash_kind kind);
/* p3 `HashChallenger<u8, Blake3 | Keccak256Hash, 32>`: `input_buffer`, `output_buffer` (bytes are sampled from its back). */
typedef struct {
    const uint8_t* input_buffer;
    size_t input_len;
    const uint8_t* output_buffer;
    size_t output_len;
} mdn_hash_challenger;
int mdn_session_set_hash_challenger(mdn_session* s, const mdn_hash_challenger* c);

/* ---- preprocessed columns: Preprocessed::build (crates/lifted-stark/src/preprocessed.rs:63-131) -----
 * `preprocessed[i]` = `BaseAir::preprocessed_trace()` of AIR i (HOST pointer; width 0 where the AIR declares
 * none, otherwise width == airs[i].preprocessed_width).  The declared matrices are sorted by (height, AIR
 * index), LDE'd on the canonical coset of their own height and committed in one aligned LMCS tree that stays
 * on the device and is borrowed by every later proof of this session (prover/mod.rs:118-123), until replaced.
 * `commitment_out` receives `Preprocessed::commitment()` -- the value the verifier is constructed with
 * (verifier/mod.rs:101-121).  A bundle must be installed exactly when some AIR of the proved statement
 * declares preprocessed columns, and each preprocessed height must equal the AIR's main trace height
 * (validate_preprocessed, preprocessed.rs:147-260): otherwise mdn_prove returns MDN_ERR_INVALID_ARG.
 * preprocessed == NULL removes the bundle. */
int mdn_session_set_preprocessed(mdn_session* s, const mdn_statement* st, const mdn_matrix* preprocessed,
                                 uint64_t commitment_out[4]);

/* ---- the drop-in: ProverInstance::prove (prover/mod.rs:230-578) ------------------------------
 * `challenger` is the caller's pre-bound challenger (protocol params observed,
 * prover/src/lib.rs:329-330); the statement felts and instance shape are observed inside, as the
 * reference does (mod.rs:290-291). */
int mdn_prove(mdn_session* s, const mdn_statement* st, const mdn_matrix* traces /* instance order */,
              const mdn_challenger* challenger, mdn_aux_builder build_aux, void* aux_ctx,
              uint32_t flags, mdn_proof* out);

/* ---- the same path, staged (for hosts that prefer to drive the aux build themselves) -------- */
int mdn_prove_begin(mdn_session* s, const mdn_statement* st, const mdn_matrix* traces,
                    const mdn_challenger* challenger, uint32_t flags,
                    uint64_t main_root[4], uint64_t* randomness_out /* 2*max num_randomness */);
int mdn_prove_commit_aux(mdn_session* s, const mdn_matrix* aux /* instance order, base-flattened */,
                         const uint64_t* const* aux_values /* per instance, 2*num_aux_values */,
                         uint64_t aux_root[4]);
int mdn_prove_finish(mdn_session* s, mdn_proof* out);

/* wincode/bincode-default layout of StarkProofData (prover/src/lib.rs:347-354): u64-LE length +
 * height bytes; u64-LE length + u64-LE felts; u64-LE length + 32-byte commitments.  Returns the
 * number of bytes needed; writes only if cap is large enough.  (Layout unpinned in-tree.) */
size_t mdn_proof_serialize(const mdn_proof* p, uint8_t* out, size_t cap);

/* ---- the two trait seams of StarkConfig (crates/lifted-stark/src/config.rs:26-45) ----------- */
/* Dft::coset_lde_batch(mat, added_bits, shift) as used at prover/commit.rs:173.  `out` receives
 * the (1 << (log_height+added_bits)) x width result, row-major, rows in bit-reversed order. */
int mdn_coset_lde_batch(mdn_session* s, const mdn_matrix* mat, uint32_t added_bits, uint64_t shift,
                        uint64_t* out);
/* Lmcs::build_aligned_tree(ldes).root() (lmcs/config.rs:125-139) for matrices given in domain
 * (natural) order, ascending heights.  */
int mdn_lmcs_commit(mdn_session* s, const mdn_matrix* mats_domain_order, uint32_t n_mats,
                    uint64_t root[4]);
/* Poseidon2 permutation on `n` independent 12-element states (crates/crypto/src/hash/
 * algebraic_sponge/poseidon2/mod.rs:31-37), device-computed. */
int mdn_poseidon2_permute(mdn_session* s, uint64_t* states /* n x 12 */, size_t n);

/* ---- host-side transcript helpers -----------------------------------------------------------
 * `CanObserve::observe` / `CanSample::sample` of the duplex challenger for hosts without p3
 * (tests, bench); sequential sponge steps on the CPU, exactly what the reference's host does.
 * Semantics: crates/lib/core/asm/stark/random_coin.masm:103-115,128-135,181-210,272-296. */
void mdn_challenger_observe(mdn_challenger* c, const uint64_t* felts, size_t n);
uint64_t mdn_challenger_sample(mdn_challenger* c);

/* ---- introspection for stage-level parity tests and the benchmark ---------------------------- */
typedef enum {
    MDN_INFO_MAIN_ROOT = 0,        /* 4 u64 */
    MDN_INFO_AUX_ROOT = 1,         /* 4 u64 */
    MDN_INFO_QUOTIENT_ROOT = 2,    /* 4 u64 */
    MDN_INFO_OOD_POINT = 3,        /* 2 u64 */
    MDN_INFO_QUOTIENT_ACC = 4,     /* N_max*D EF values, natural order on gJ */
    MDN_INFO_DEEP_EVALS = 5,       /* L EF values, bit-reversed order */
    MDN_INFO_FRI_ROOTS = 6,        /* 4 u64 per round */
    MDN_INFO_QUERY_INDICES = 7,    /* num_queries u64 */
    MDN_INFO_JIT = 8,              /* per AIR (proof order) of the last proof: 1 = NVRTC kernel, 0 = interpreter */
    MDN_INFO_POOL = 9,             /* cudaMallocAsync pool reserved now / high, used now / high (bytes); proof arena capacity (bytes), slabs, driver allocations so far, live blocks */
    MDN_INFO_BUILD = 10,           /* { generation of the field arithmetic (1 = poseidon2_fast.cuh, 2 = poseidon2_fast2.cuh), generation of the NTT kernels (1 = kernels.cu, 2 = ntt2.cuh) } this library was built with */
} mdn_info;
/* QUOTIENT_ACC / DEEP_EVALS are only recorded (extra device->host copies) after mdn_set_debug(s, 1). */
int mdn_set_debug(mdn_session* s, int enable);
/* Layout self-description for binding authors: { sizeof pcs_params, challenger, lookup, air; offsetof air.program,
 * .periodic_values, .preprocessed_width, .lookup; sizeof matrix, statement, proof, timings; offsetof
 * timings.kernel_ms, .permutations }.  Returns the number of entries. */
size_t mdn_abi_layout(uint32_t* out, size_t cap);

/* ---- run-time specialisation of the constraint evaluator -------------------------------------------
 * AIR programs with at least `min_nodes` nodes (default 256; 0 = never) are lowered to straight-line CUDA,
 * compiled once per program with NVRTC for sm_100a and cached; smaller ones (and all of them when libnvrtc
 * is absent) run on the op-list interpreter.  Both produce identical values. */
int mdn_session_set_jit(mdn_session* s, uint32_t min_nodes);
/* "nvrtc <version>" or why it is unavailable, plus the reason the last JIT attempt was dropped (compile error,
 * or the first-use comparison against the interpreter failed -- the interpreter's result is then kept). */
const char* mdn_jit_status(mdn_session* s);
/* Codegen + NVRTC only, no device needed: cubin size (> 0) or a negative mdn_status with *err set. */
long long mdn_jit_compile_check(const uint32_t* program, uint32_t program_words, const char** err);
/* Copies at most cap u64 into out; returns the number of u64 available (or <0). */
long long mdn_get_info(mdn_session* s, mdn_info what, uint64_t* out, size_t cap);

/* Per-phase device timings of the last prove, in milliseconds (CUDA events on the session's
 * stream).  Names follow the reference's tracing spans (prover/mod.rs:339,412,445,542,561). */
typedef struct {
    float h2d_transpose, commit_main, commit_aux, evaluate_constraints, commit_quotient, open, total;
    float lde_main, hash_main;          /* inside commit_main */
    /* per kernel class, summed over the last prove's launches (CUDA events on the session stream):
     * 0 transpose, 1 NTT/LDE, 2 leaf sponge, 3 Merkle compress, 4 constraints, 5 OOD dot products,
     * 6 DEEP quotient, 7 FRI (leaf+compress+fold), 8 PoW grind, 9 opening gather */
    float kernel_ms[10];
    unsigned kernel_regions[10];        /* timed regions per class */
    unsigned long long kernel_launches; /* kernels launched by the last prove */
    unsigned long long permutations;    /* Poseidon2 permutations executed for commitments */
    double leaf_hash_bytes;             /* algorithmic bytes of the leaf-sponge launches (LDE read + state/digest write) */
    double ntt_bytes;                   /* algorithmic bytes of the LDE: (N + L) * width * 8 per matrix */
} mdn_timings;
int mdn_get_timings(mdn_session* s, mdn_timings* out);

#ifdef __cplusplus
}
#endif
#endif /* MIDEN_B200_H */
