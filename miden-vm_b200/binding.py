"""ctypes view of include/miden_b200.h.  No fallback: importing works without the library (so the
CPU test-suite can check the header/export surface), but every compute entry point raises if
`libmiden_b200.so` cannot be loaded."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDN_LIB_PATH") or os.path.join(HERE, "csrc", "libmiden_b200.so")   # override only for A/B kernel experiments
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


class PcsParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("log_blowup", "log_folding_arity", "log_final_degree",
                                          "folding_pow_bits", "deep_pow_bits", "num_queries", "query_pow_bits")]


class Challenger(C.Structure):
    _fields_ = [("sponge_state", C.c_uint64 * 12), ("input_buffer", C.c_uint64 * 8),
                ("input_len", C.c_uint32), ("output_len", C.c_uint32)]


class Lookup(C.Structure):
    _fields_ = [("num_columns", C.c_uint32), ("program_words", C.c_uint32), ("program", u32p)]


class Air(C.Structure):
    _fields_ = [("width", C.c_uint32), ("aux_width", C.c_uint32), ("num_aux_values", C.c_uint32),
                ("num_randomness", C.c_uint32), ("log_quotient_degree", C.c_uint32),
                ("program_words", C.c_uint32), ("program", u32p),
                ("periodic_values", u64p), ("num_periodic_columns", C.c_uint32), ("log_max_period", C.c_uint32),
                ("preprocessed_width", C.c_uint32), ("lookup", C.POINTER(Lookup))]


class HashChallenger(C.Structure):
    _fields_ = [("input_buffer", C.POINTER(C.c_uint8)), ("input_len", C.c_size_t),
                ("output_buffer", C.POINTER(C.c_uint8)), ("output_len", C.c_size_t)]


HASH_POSEIDON2, HASH_BLAKE3, HASH_KECCAK, HASH_RPO, HASH_RPX = 0, 1, 2, 3, 4


class Matrix(C.Structure):
    _fields_ = [("values", u64p), ("log_height", C.c_uint32), ("width", C.c_uint32)]


class Statement(C.Structure):
    _fields_ = [("airs", C.POINTER(Air)), ("n_airs", C.c_uint32),
                ("public_values", u64p), ("n_public_values", C.c_uint32),
                ("observe_felts", u64p), ("n_observe_felts", C.c_uint32)]


class Proof(C.Structure):
    _fields_ = [("log_trace_heights", C.POINTER(C.c_uint8)), ("n_heights", C.c_size_t),
                ("fields", u64p), ("n_fields", C.c_size_t),
                ("commitments", u64p), ("n_commitments", C.c_size_t)]


class Timings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_transpose", "commit_main", "commit_aux", "evaluate_constraints",
                                         "commit_quotient", "open", "total", "lde_main", "hash_main")] + \
               [("kernel_ms", C.c_float * 10), ("kernel_regions", C.c_uint * 10), ("kernel_launches", C.c_ulonglong),
                ("permutations", C.c_ulonglong), ("leaf_hash_bytes", C.c_double), ("ntt_bytes", C.c_double)]


AUX_BUILDER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.POINTER(Matrix), u64p, u64p, u64p)
ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, u64p, u64p, C.c_size_t)
EXTERNAL_CHECK = C.CFUNCTYPE(C.c_int, C.c_void_p, u64p, C.c_uint32, C.POINTER(u64p), u32p, C.POINTER(C.c_uint8), C.c_uint32, u32p)
FLAG_DEVICE_TRACES = 1

# Every symbol include/miden_b200.h declares (checked by tests/test_abi.py).
EXPORTS = [
    "mdn_session_create", "mdn_session_destroy", "mdn_last_error", "mdn_prove", "mdn_prove_begin",
    "mdn_prove_commit_aux", "mdn_prove_finish", "mdn_proof_serialize", "mdn_coset_lde_batch",
    "mdn_lmcs_commit", "mdn_poseidon2_permute", "mdn_get_info", "mdn_get_timings",
    "mdn_challenger_observe", "mdn_challenger_sample", "mdn_set_debug", "mdn_session_set_shard",
    "mdn_session_set_preprocessed", "mdn_session_set_jit", "mdn_jit_compile_check", "mdn_jit_status", "mdn_abi_layout",
    "mdn_session_set_external_check", "mdn_session_set_hash", "mdn_session_set_hash_challenger",
]

_lib = None


def jit_compile_check(program: np.ndarray) -> int:
    """Lower + NVRTC-compile a constraint program without touching a device; returns the cubin size."""
    err = C.c_char_p()
    prog = np.ascontiguousarray(program, dtype=np.uint32)
    n = lib().mdn_jit_compile_check(prog.ctypes.data_as(u32p), len(prog), C.byref(err))
    if n < 0:
        raise ProverError(n, (err.value or b"").decode())
    return n


class BackendMissing(RuntimeError):
    pass


def lib():
    """Load libmiden_b200.so.  Raises BackendMissing (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendMissing(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        if hasattr(L, "mdn_emulated_build") and os.environ.get("MDN_ALLOW_EMULATOR") != "1":
            # tests/emu builds the kernels for the CPU (one fiber per CUDA thread) so that host logic can be tested
            # without a GPU; it is test infrastructure and must never stand in for the product
            raise BackendMissing(f"{LIB_PATH} is the CPU kernel emulator of tests/emu, not the CUDA backend "
                                 "(there is no CPU fallback; only tests/test_emulated.py may load it)")
        L.mdn_last_error.restype = C.c_char_p
        L.mdn_last_error.argtypes = [C.c_void_p]
        L.mdn_session_create.argtypes = [C.POINTER(PcsParams), C.c_int, C.POINTER(C.c_void_p)]
        L.mdn_session_destroy.argtypes = [C.c_void_p]
        L.mdn_prove.argtypes = [C.c_void_p, C.POINTER(Statement), C.POINTER(Matrix), C.POINTER(Challenger),
                                AUX_BUILDER, C.c_void_p, C.c_uint32, C.POINTER(Proof)]
        L.mdn_prove_begin.argtypes = [C.c_void_p, C.POINTER(Statement), C.POINTER(Matrix), C.POINTER(Challenger),
                                      C.c_uint32, u64p, u64p]
        L.mdn_prove_commit_aux.argtypes = [C.c_void_p, C.POINTER(Matrix), C.POINTER(u64p), u64p]
        L.mdn_prove_finish.argtypes = [C.c_void_p, C.POINTER(Proof)]
        L.mdn_proof_serialize.restype = C.c_size_t
        L.mdn_proof_serialize.argtypes = [C.POINTER(Proof), C.POINTER(C.c_uint8), C.c_size_t]
        L.mdn_coset_lde_batch.argtypes = [C.c_void_p, C.POINTER(Matrix), C.c_uint32, C.c_uint64, u64p]
        L.mdn_lmcs_commit.argtypes = [C.c_void_p, C.POINTER(Matrix), C.c_uint32, u64p]
        L.mdn_poseidon2_permute.argtypes = [C.c_void_p, u64p, C.c_size_t]
        L.mdn_get_info.restype = C.c_longlong
        L.mdn_get_info.argtypes = [C.c_void_p, C.c_int, u64p, C.c_size_t]
        L.mdn_set_debug.argtypes = [C.c_void_p, C.c_int]
        L.mdn_session_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, ALLGATHER, C.c_void_p]
        L.mdn_session_set_hash.argtypes = [C.c_void_p, C.c_int]
        L.mdn_session_set_hash_challenger.argtypes = [C.c_void_p, C.POINTER(HashChallenger)]
        L.mdn_session_set_external_check.argtypes = [C.c_void_p, EXTERNAL_CHECK, C.c_void_p]
        L.mdn_session_set_preprocessed.argtypes = [C.c_void_p, C.POINTER(Statement), C.POINTER(Matrix), u64p]
        L.mdn_abi_layout.restype = C.c_size_t
        L.mdn_abi_layout.argtypes = [u32p, C.c_size_t]
        L.mdn_session_set_jit.argtypes = [C.c_void_p, C.c_uint32]
        L.mdn_jit_status.restype = C.c_char_p
        L.mdn_jit_status.argtypes = [C.c_void_p]
        L.mdn_jit_compile_check.restype = C.c_longlong
        L.mdn_jit_compile_check.argtypes = [u32p, C.c_uint32, C.POINTER(C.c_char_p)]
        L.mdn_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.mdn_challenger_observe.argtypes = [C.POINTER(Challenger), u64p, C.c_size_t]
        L.mdn_challenger_sample.restype = C.c_uint64
        L.mdn_challenger_sample.argtypes = [C.POINTER(Challenger)]
        _lib = L
    return _lib


def ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


class ProverError(RuntimeError):
    """Mirrors `ExecutionError::ProvingError(String)` (reference prover/src/lib.rs:336-345)."""


class Session:
    """One proving session bound to one CUDA device (`mdn_session`)."""

    def __init__(self, params: PcsParams, device: int = 0):
        self._h = C.c_void_p()
        rc = lib().mdn_session_create(C.byref(params), device, C.byref(self._h))
        if rc != 0:
            raise ProverError(f"mdn_session_create failed ({rc}): {lib().mdn_last_error(None).decode()}")

    def close(self):
        if self._h:
            lib().mdn_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise ProverError(f"[{rc}] {lib().mdn_last_error(self._h).decode()}")

    @property
    def handle(self):
        return self._h

    def set_shard(self, rank: int, world: int, allgather_cb):
        """Split every proof of this session over `world` ranks (mdn_session_set_shard); collective."""
        self._allgather_cb = allgather_cb      # keep the ctypes trampoline alive
        self._check(lib().mdn_session_set_shard(self._h, rank, world, allgather_cb, None))

    def set_external_check(self, fn):
        """`Statement::eval_external`: fn(challenges u64[2n], aux_values [per AIR u64[]], log_heights bytes) ->
        None / -1 when every assertion holds, else the index of the first failing assertion."""
        if fn is None:
            self._ext_cb = C.cast(None, EXTERNAL_CHECK)
        else:
            def tramp(ctx, ch, n_ch, vals, n_vals, lh, n_airs, failed):
                try:
                    chal = np.ctypeslib.as_array(ch, shape=(2 * n_ch,)).copy() if n_ch else np.zeros(0, np.uint64)
                    av = [np.ctypeslib.as_array(vals[i], shape=(2 * n_vals[i],)).copy() if n_vals[i] else np.zeros(0, np.uint64) for i in range(n_airs)]
                    k = fn(chal, av, bytes(lh[:n_airs]))
                    if k is None or k < 0:
                        return 0
                    failed[0] = k
                    return 1
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return -1
            self._ext_cb = EXTERNAL_CHECK(tramp)
        self._check(lib().mdn_session_set_external_check(self._h, self._ext_cb, None))

    def set_hash(self, kind: int, challenger_input: bytes = b"", challenger_output: bytes = b""):
        """`blake3_256_config` / `keccak_config` instead of `poseidon2_config` (mdn_session_set_hash) + the pre-bound HashChallenger state."""
        self._check(lib().mdn_session_set_hash(self._h, kind))
        if kind in (HASH_BLAKE3, HASH_KECCAK):
            a = (C.c_uint8 * max(1, len(challenger_input))).from_buffer_copy(challenger_input or b"\0")
            b = (C.c_uint8 * max(1, len(challenger_output))).from_buffer_copy(challenger_output or b"\0")
            hc = HashChallenger(a, len(challenger_input), b, len(challenger_output))
            self._check(lib().mdn_session_set_hash_challenger(self._h, C.byref(hc)))

    def set_jit(self, min_nodes: int):
        """Node threshold above which constraint programs are NVRTC-compiled (0 = interpreter only)."""
        self._check(lib().mdn_session_set_jit(self._h, min_nodes))

    def jit_status(self) -> str:
        return lib().mdn_jit_status(self._h).decode()

    def set_preprocessed(self, statement: Statement, preprocessed):
        """`Preprocessed::build(statement, config)` on the device; returns the commitment (u64[4]).
        `preprocessed` = Matrix array in instance order (width 0 = none), or None to remove the bundle."""
        out = np.zeros(4, dtype=np.uint64)
        self._check(lib().mdn_session_set_preprocessed(self._h, C.byref(statement) if statement is not None else None,
                                                       preprocessed, out.ctypes.data_as(u64p)))
        return out

    def prove(self, statement: Statement, traces, challenger: Challenger, aux_builder=None, flags=0):
        """`ProverInstance::new(config, statement, None)?.prove(challenger)`; returns
        (log_trace_heights bytes, fields u64[], commitments u64[n,4])."""
        proof = Proof()
        cb = aux_builder if aux_builder is not None else C.cast(None, AUX_BUILDER)
        self._check(lib().mdn_prove(self._h, C.byref(statement), traces, C.byref(challenger) if challenger is not None else None, cb, None, flags,
                                    C.byref(proof)))
        return proof_to_numpy(proof)

    def info(self, what: int, cap: int = 0) -> np.ndarray:
        n = lib().mdn_get_info(self._h, what, None, 0)
        if n < 0:
            raise ProverError(f"mdn_get_info({what}) failed")
        out = np.zeros(n, dtype=np.uint64)
        if n:
            lib().mdn_get_info(self._h, what, ptr(out), n)
        return out

    def timings(self) -> Timings:
        t = Timings()
        self._check(lib().mdn_get_timings(self._h, C.byref(t)))
        return t


def proof_to_numpy(proof: Proof):
    heights = bytes(proof.log_trace_heights[: proof.n_heights])
    fields = np.ctypeslib.as_array(proof.fields, shape=(proof.n_fields,)).copy() if proof.n_fields else np.zeros(0, np.uint64)
    comms = (np.ctypeslib.as_array(proof.commitments, shape=(proof.n_commitments * 4,)).copy().reshape(-1, 4)
             if proof.n_commitments else np.zeros((0, 4), np.uint64))
    return heights, fields, comms
