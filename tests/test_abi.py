"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, and
exports every symbol include/miden_b200.h declares.  No compute is invoked (no GPU here)."""
import ctypes as C
import os
import re

import pytest

import pkgload

pkg = pkgload.load_pkg()
B = pkg.binding
ROOT = pkgload.ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "miden_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mdn_[a-z0-9_]+)\s*\(", hdr)) - {"mdn_aux_builder"})


def test_header_symbols_match_binding_list():
    assert _declared_symbols() == sorted(B.EXPORTS)


def test_library_builds_and_exports_every_symbol():
    import __graft_entry__ as g
    g.build()
    lib = C.CDLL(B.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} not exported by libmiden_b200.so"


def test_ctypes_struct_layouts_match_the_library():
    import numpy as np
    lib = B.lib()
    n = lib.mdn_abi_layout(None, 0)
    v = np.zeros(n, dtype=np.uint32)
    lib.mdn_abi_layout(v.ctypes.data_as(B.u32p), n)
    A, T = B.Air, B.Timings
    expect = [C.sizeof(B.PcsParams), C.sizeof(B.Challenger), C.sizeof(B.Lookup), C.sizeof(A), A.program.offset, A.periodic_values.offset,
              A.preprocessed_width.offset, A.lookup.offset, C.sizeof(B.Matrix), C.sizeof(B.Statement), C.sizeof(B.Proof), C.sizeof(T),
              T.kernel_ms.offset, T.permutations.offset]
    assert list(v) == expect
    import oracle_binding as ob      # the oracle's C API mirrors the same structs
    assert C.sizeof(ob.Air) == C.sizeof(A) and ob.Air.lookup.offset == A.lookup.offset


def test_session_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.ProverError) as e:
        B.Session(pkg.workload.miden_pcs_params(), 0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_host_challenger_matches_oracle(oracle):
    """The product's host-side duplex challenger (used to seed proofs) against the oracle's."""
    import numpy as np
    import helpers as H
    W = pkg.workload
    params = W.miden_pcs_params()
    lib = B.lib()

    def prod_observe(c, felts):
        lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))

    a = W.initial_challenger(params, prod_observe)
    b = W.initial_challenger(params, H.oracle_observe)
    assert bytes(a) == bytes(b)
    extra = np.array([5, 6, 7], dtype=np.uint64)
    prod_observe(a, extra); H.oracle_observe(b, extra)
    assert bytes(a) == bytes(b)
    import oracle_binding as ob
    for _ in range(11):
        va = lib.mdn_challenger_sample(C.byref(a))
        out = np.zeros(1, dtype=np.uint64)
        oracle.orc_challenger_script(C.cast(C.byref(b), C.POINTER(ob.Challenger)), np.array([1], dtype=np.uint32).ctypes.data_as(ob.u32p),
                                     ob.ptr(np.zeros(1, dtype=np.uint64)), 1, ob.ptr(out))
        assert va == int(out[0])
    assert bytes(a) == bytes(b)


def test_proof_serialize_layout():
    """mdn_proof_serialize is pure host code: u64-LE length + height bytes, u64-LE length + u64-LE felts,
    u64-LE length + 32-byte commitments (bincode-default layout expected of wincode)."""
    import numpy as np
    lib = B.lib()
    heights = (C.c_uint8 * 3)(10, 9, 8)
    fields = np.array([1, 2, 0xFFFFFFFF00000000], dtype=np.uint64)
    comms = np.arange(8, dtype=np.uint64)
    pf = B.Proof(heights, 3, B.ptr(fields), 3, B.ptr(comms), 2)
    need = lib.mdn_proof_serialize(C.byref(pf), None, 0)
    assert need == 8 + 3 + 8 + 24 + 8 + 64
    buf = (C.c_uint8 * need)()
    assert lib.mdn_proof_serialize(C.byref(pf), buf, need) == need
    raw = bytes(buf)
    assert raw[:8] == (3).to_bytes(8, "little") and raw[8:11] == bytes([10, 9, 8])
    assert raw[11:19] == (3).to_bytes(8, "little")
    assert raw[19:43] == fields.astype("<u8").tobytes()
    assert raw[43:51] == (2).to_bytes(8, "little")
    assert raw[51:] == comms.astype("<u8").tobytes()
