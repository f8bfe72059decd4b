"""Shared helpers for the parity tests."""
import ctypes as C

import numpy as np

import oracle_binding as ob
import pkgload

pkg = pkgload.load_pkg()
W = pkg.workload
B = pkg.binding


def oracle_observe(c, felts: np.ndarray):
    """Challenger observe through the ORACLE (used when seeding oracle-side runs)."""
    n = len(felts)
    ops = np.zeros(n, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint64)
    ob.lib().orc_challenger_script(C.cast(C.byref(c), C.POINTER(ob.Challenger)), ops.ctypes.data_as(ob.u32p),
                                   ob.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), n, ob.ptr(out))


def cast(obj, typ):
    """Reinterpret a product-binding struct (or array) as the oracle-binding type of equal layout."""
    return C.cast(C.byref(obj) if not isinstance(obj, C.Array) else obj, C.POINTER(typ))


def oracle_prove(params, wl, challenger, aux_builder=None):
    """Run the oracle prover on a Workload; returns (handle, heights, fields, commitments)."""
    proof = ob.Proof()
    cb = ob.AUX_BUILDER(aux_builder) if aux_builder is not None else C.cast(None, ob.AUX_BUILDER)
    if getattr(wl, "preprocessed", None) is not None:
        wl.oracle_prep_commitment = np.zeros(4, dtype=np.uint64)
        h = ob.lib().orc_prove_pp(cast(params, ob.PcsParams), cast(wl.statement, ob.Statement), cast(wl.matrices, ob.Matrix),
                                  cast(wl.preprocessed_matrices, ob.Matrix), cast(challenger, ob.Challenger), cb, None,
                                  C.byref(proof), ob.ptr(wl.oracle_prep_commitment))
    else:
        h = ob.lib().orc_prove(cast(params, ob.PcsParams), cast(wl.statement, ob.Statement), cast(wl.matrices, ob.Matrix),
                               cast(challenger, ob.Challenger), cb, None, C.byref(proof))
    if not h:
        raise RuntimeError(ob.lib().orc_last_error().decode())
    heights = bytes(proof.log_trace_heights[: proof.n_heights])
    fields = np.ctypeslib.as_array(proof.fields, shape=(proof.n_fields,)).copy()
    comms = np.ctypeslib.as_array(proof.commitments, shape=(proof.n_commitments * 4,)).copy().reshape(-1, 4)
    return h, heights, fields, comms


def oracle_info(h, what):
    n = ob.lib().orc_prove_info(h, what, None, 0)
    out = np.zeros(n, dtype=np.uint64)
    ob.lib().orc_prove_info(h, what, ob.ptr(out), n)
    return out


def oracle_verify(params, wl, challenger, heights, fields, comms, prep_commitment=None):
    hb = (C.c_uint8 * len(heights))(*heights)
    fields = np.ascontiguousarray(fields, dtype=np.uint64)
    comms = np.ascontiguousarray(comms, dtype=np.uint64)
    pf = ob.Proof(hb, len(heights), ob.ptr(fields), len(fields), ob.ptr(comms.reshape(-1)), len(comms))
    if prep_commitment is None and getattr(wl, "preprocessed", None) is not None:
        prep_commitment = wl.oracle_prep_commitment
    if prep_commitment is not None:
        pc = np.ascontiguousarray(prep_commitment, dtype=np.uint64)
        rc = ob.lib().orc_verify_pp(cast(params, ob.PcsParams), cast(wl.statement, ob.Statement), C.byref(pf),
                                    cast(challenger, ob.Challenger), ob.ptr(pc))
    else:
        rc = ob.lib().orc_verify(cast(params, ob.PcsParams), cast(wl.statement, ob.Statement), C.byref(pf),
                                 cast(challenger, ob.Challenger))
    return rc, ob.lib().orc_last_error().decode()
