"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs load it."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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


AUX_BUILDER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.POINTER(Matrix), u64p, u64p, u64p)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_threads.restype = C.c_int; L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_threads.restype = C.c_int
        L.orc_fp_mul.restype = C.c_uint64; L.orc_fp_mul.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_fp_inv.restype = C.c_uint64; L.orc_fp_inv.argtypes = [C.c_uint64]
        L.orc_two_adic_generator.restype = C.c_uint64; L.orc_two_adic_generator.argtypes = [C.c_uint32]
        L.orc_lde_shift.restype = C.c_uint64; L.orc_lde_shift.argtypes = [C.c_uint32]
        L.orc_poseidon2_permute.argtypes = [u64p, C.c_size_t]
        L.orc_fri_fold_row.restype = C.c_int
        L.orc_fri_fold_row.argtypes = [C.c_uint32, u64p, C.c_uint64, u64p, u64p]
        L.orc_naive_dft.argtypes = [u64p, C.c_uint32, u64p]
        L.orc_dft.argtypes = [u64p, C.c_uint32, C.c_uint32, C.c_int, u64p]
        L.orc_coset_lde_batch.argtypes = [C.POINTER(Matrix), C.c_uint32, C.c_uint64, u64p]
        L.orc_lmcs_commit.argtypes = [C.POINTER(Matrix), C.c_uint32, u64p, u64p]
        L.orc_challenger_script.argtypes = [C.POINTER(Challenger), u32p, u64p, C.c_size_t, u64p]
        L.orc_prove.restype = C.c_void_p
        L.orc_prove.argtypes = [C.POINTER(PcsParams), C.POINTER(Statement), C.POINTER(Matrix),
                                C.POINTER(Challenger), AUX_BUILDER, C.c_void_p, C.POINTER(Proof)]
        L.orc_prove_pp.restype = C.c_void_p
        L.orc_prove_pp.argtypes = [C.POINTER(PcsParams), C.POINTER(Statement), C.POINTER(Matrix), C.POINTER(Matrix),
                                   C.POINTER(Challenger), AUX_BUILDER, C.c_void_p, C.POINTER(Proof), u64p]
        L.orc_verify_pp.restype = C.c_int
        L.orc_verify_pp.argtypes = [C.POINTER(PcsParams), C.POINTER(Statement), C.POINTER(Proof), C.POINTER(Challenger), u64p]
        L.orc_prove_free.argtypes = [C.c_void_p]
        L.orc_prove_info.restype = C.c_longlong
        L.orc_prove_info.argtypes = [C.c_void_p, C.c_int, u64p, C.c_size_t]
        L.orc_verify.restype = C.c_int
        L.orc_verify.argtypes = [C.POINTER(PcsParams), C.POINTER(Statement), C.POINTER(Proof), C.POINTER(Challenger)]
        _lib = L
    return _lib


def ptr(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)
