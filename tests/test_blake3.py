"""Blake3_256 STARK configuration (reference air/src/config.rs:276-307): BLAKE3 itself, the chaining-hasher
LMCS and the hash challenger, pinned on vectors made with the `blake3` package (bindings of the official crate the reference
wraps; tests/golden/make_blake3_vectors.py), for the oracle and -- through the C++ command-line tool and the C ABI -- for the
product; then full proofs, GPU vs oracle, bit for bit (`-m gpu`; the same cases run on the CPU kernel emulator)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import oracle_binding as ob
import pkgload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "blake3_vectors.json")))
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
P = W.P


@pytest.fixture()
def orc_b3():
    ob.build()
    L = ob.lib()
    L.orc_set_hash.restype = C.c_int
    L.orc_set_hash.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.orc_blake3.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    yield L
    L.orc_set_hash(0, None, 0)


def test_oracle_blake3_matches_the_official_crate(orc_b3):
    for v in V["hash"]:
        data = bytes(i % 251 for i in range(v["len"]))
        out = C.create_string_buffer(32)
        orc_b3.orc_blake3(data, len(data), out)
        assert out.raw.hex() == v["digest"], v["len"]


def test_product_blake3_matches_the_official_crate():
    """csrc/blake3.cuh (the word-streaming hasher the kernels and the host transcript use) compiled for the host."""
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp, "test_blake3"])
    for v in V["hash"]:
        if v["len"] % 4:
            continue        # the product hashes digests and little-endian u64 felts: whole words only
        got = subprocess.run([os.path.join(cpp, "test_blake3"), str(v["len"] // 4)], capture_output=True, text=True, check=True).stdout.strip()
        assert got == v["digest"], v["len"]


def _bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def test_oracle_blake3_lmcs_roots(orc_b3):
    """ChainingHasher leaves with state lifting + blake3(left || right) layers against the Python restatement of
    lifted_tree.rs made with the official crate."""
    assert orc_b3.orc_set_hash(1, b"", 0) == 0
    for case in V["lmcs"]:
        mats, keep = (ob.Matrix * len(case["shapes"]))(), []
        for i, ((h, w), rows) in enumerate(zip(case["shapes"], case["rows"])):
            lg = h.bit_length() - 1
            nat = np.zeros((h, max(w, 0)), dtype=np.uint64)       # orc_lmcs_commit takes domain (natural) order
            for r in range(h):
                nat[r] = rows[_bitrev(r, lg)] if w else []
            nat = np.ascontiguousarray(nat)
            keep.append(nat)
            mats[i] = ob.Matrix(nat.ctypes.data_as(ob.u64p) if w else None, lg, w)
        root = np.zeros(4, dtype=np.uint64)
        orc_b3.orc_lmcs_commit(mats, len(case["shapes"]), ob.ptr(root), None)
        assert root.tobytes().hex() == case["root"], case["shapes"]


def test_oracle_hash_challenger_script(orc_b3):
    c = V["challenger"]
    init = bytes.fromhex(c["initial_input_hex"])
    assert orc_b3.orc_set_hash(1, init, len(init)) == 0
    ops = np.array([{"observe": 0, "sample": 1, "bits": 2}[o] for o, _ in c["script"]], dtype=np.uint32)
    args = np.array([a for _, a in c["script"]], dtype=np.uint64)
    out = np.zeros(len(ops), dtype=np.uint64)
    ch = ob.Challenger()
    orc_b3.orc_challenger_script(C.byref(ch), ops.ctypes.data_as(ob.u32p), ob.ptr(args), len(ops), ob.ptr(out))
    assert [int(x) for x in out] == c["results"]


def test_oracle_blake3_prove_verify_tamper(orc_b3):
    import test_airs
    params = W.fast_pcs_params()
    init = W.initial_hash_challenger(params)
    assert orc_b3.orc_set_hash(1, init, len(init)) == 0
    for wl, aux in [(W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2)), None), test_airs.fib_product_workload([7], lqd=1)]:
        ch = W.Challenger()
        h, oh, of, oc = H.oracle_prove(params, wl, ch, aux)
        ob.lib().orc_prove_free(h)
        assert H.oracle_verify(params, wl, ch, oh, of, oc)[0] == 0
        bad = of.copy(); bad[len(bad) // 2] ^= 1
        assert H.oracle_verify(params, wl, ch, oh, bad, oc)[0] != 0
        badc = oc.copy(); badc[0, 0] ^= 1
        assert H.oracle_verify(params, wl, ch, oh, of, badc)[0] != 0
    # the Poseidon2 verifier does not accept a Blake3 proof
    orc_b3.orc_set_hash(0, None, 0)
    chp = W.initial_challenger(params, H.oracle_observe)
    assert H.oracle_verify(params, wl, chp, oh, of, oc)[0] != 0


def _prove_blake3_vs_oracle(orc_b3, params, wl, aux=None, prep=False, debug=True):
    init = W.initial_hash_challenger(params)
    assert orc_b3.orc_set_hash(1, init, len(init)) == 0
    s = B.Session(params, 0)
    try:
        s.set_hash(B.HASH_BLAKE3, init)
        B.lib().mdn_set_debug(s.handle, 1 if debug else 0)
        if prep:
            s.set_preprocessed(wl.statement, wl.preprocessed_matrices)
        got = s.prove(wl.statement, wl.matrices, None, B.AUX_BUILDER(aux) if aux else None)
        ch = W.Challenger()
        h, oh, of, oc = H.oracle_prove(params, wl, ch, aux)
        try:
            names = ["main_root", "aux_root", "quotient_root", "ood_point", "quotient_acc", "deep_evals", "fri_roots", "query_indices"]
            for what, name in enumerate(names):
                if not debug and what in (4, 5):
                    continue
                assert np.array_equal(s.info(what), H.oracle_info(h, what)), f"stage {name} differs"
            assert got[0] == oh and np.array_equal(got[2], oc) and np.array_equal(got[1], of)
        finally:
            ob.lib().orc_prove_free(h)
        assert H.oracle_verify(params, wl, ch, *got)[0] == 0
    finally:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["two_heights", "miden_shape", "host_aux", "logup_device", "preprocessed", "arity8", "blowup2"])
def test_blake3_proofs_bit_exact_vs_oracle(orc_b3, case):
    import test_airs
    if case == "two_heights":
        _prove_blake3_vs_oracle(orc_b3, W.fast_pcs_params(), W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2)))
    elif case == "miden_shape":      # production parameters: 12/4/16-bit grinding through the hash challenger
        _prove_blake3_vs_oracle(orc_b3, W.miden_pcs_params(), W.Workload([10, 9, 8]))
    elif case == "host_aux":
        wl, aux = test_airs.fib_product_workload([7, 5], lqd=1)
        _prove_blake3_vs_oracle(orc_b3, W.fast_pcs_params(), wl, aux)
    elif case == "logup_device":
        _prove_blake3_vs_oracle(orc_b3, W.fast_pcs_params(), test_airs.logup_workload(6, device=True)[0])
    elif case == "preprocessed":     # also: zero-width aux matrices re-hash the chaining state
        _prove_blake3_vs_oracle(orc_b3, W.fast_pcs_params(), test_airs.preprocessed_workload((5, 7), (True, False)), prep=True)
    elif case == "arity8":
        _prove_blake3_vs_oracle(orc_b3, B.PcsParams(3, 3, 2, 2, 3, 7, 4), W.Workload([7, 9], widths=(9, 12), aux_widths=(1, 2)))
    else:
        wl, aux = test_airs.fib_product_workload([6], lqd=1)
        _prove_blake3_vs_oracle(orc_b3, B.PcsParams(1, 2, 1, 1, 2, 6, 3), wl, aux)


@pytest.mark.gpu
def test_blake3_2_16_bit_exact_and_hash_switch(orc_b3):
    """BASELINE config 3's hash on the config-2 size: 2^16 x (51,22,16) under Blake3 bit-exact against the oracle prover, then
    the same session switched back to Poseidon2 still produces the Poseidon2 proof."""
    params = W.miden_pcs_params()
    wl = W.Workload([16, 16, 16])
    _prove_blake3_vs_oracle(orc_b3, params, wl, debug=False)
    orc_b3.orc_set_hash(0, None, 0)
    s = B.Session(params, 0)
    try:
        ch = W.initial_challenger(params, lambda c, f: B.lib().mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(f, dtype=np.uint64)), len(f)))
        small = W.Workload([8, 7, 6])
        a = s.prove(small.statement, small.matrices, ch)
        init = W.initial_hash_challenger(params)
        s.set_hash(B.HASH_BLAKE3, init)
        b = s.prove(small.statement, small.matrices, None)
        s.set_hash(B.HASH_POSEIDON2)
        c = s.prove(small.statement, small.matrices, ch)
        assert np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2]) and not np.array_equal(a[2][:1], b[2][:1])
    finally:
        s.close()
