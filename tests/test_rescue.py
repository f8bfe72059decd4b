"""RPO / RPX STARK configurations (reference air/src/config.rs:225-248: the algebraic configuration with the permutation swapped).

Pins, in order of strength: the reference's OWN 19 `Rpo256::hash_elements` known answers (rescue/rpo/tests.rs:241-430, extracted by
tools/gen_rescue_constants.py) for the oracle and -- through the C++ command-line tool -- for the product's csrc/rescue.cuh; a
pure-Python big-integer restatement (tests/golden/make_rescue_vectors.py, itself pinned on those answers) for the RPX permutation,
the LMCS roots and the duplex challenger under both permutations; then full proofs, GPU vs oracle, bit for bit (`-m gpu`; the
same cases run on the CPU kernel emulator)."""
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
REF = json.load(open(os.path.join(ROOT, "tests", "golden", "rpo_reference_vectors.json")))["hash_elements_prefixes"]
V = json.load(open(os.path.join(ROOT, "tests", "golden", "rescue_vectors.json")))
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
KIND = {"rpo": 3, "rpx": 4}


@pytest.fixture()
def orc():
    ob.build()
    L = ob.lib()
    L.orc_set_hash.restype = C.c_int
    L.orc_set_hash.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.orc_rescue_permute.argtypes = [C.c_int, ob.u64p]
    L.orc_rescue_hash_elements.argtypes = [C.c_int, ob.u64p, C.c_size_t, ob.u64p]
    yield L
    L.orc_set_hash(0, None, 0)


def _tool():
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp, "test_rescue"])
    return os.path.join(cpp, "test_rescue")


def test_oracle_rpo_matches_the_reference_known_answers(orc):
    for n, want in enumerate(REF, 1):
        e, out = np.arange(n, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        orc.orc_rescue_hash_elements(3, ob.ptr(e), n, ob.ptr(out))
        assert [int(x) for x in out] == want, n


def test_product_rpo_matches_the_reference_known_answers():
    lines = subprocess.run([_tool(), "kat"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert [[int(x) for x in l.split()] for l in lines] == REF


@pytest.mark.parametrize("name", ["rpo", "rpx"])
def test_permutations_oracle_and_product_match_the_python_restatement(orc, name):
    cases = V["perm"][name]
    inp = "\n".join(" ".join(str(x) for x in c["in"]) for c in cases)
    got = subprocess.run([_tool(), "perm", str(KIND[name])], input=inp, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for c, line in zip(cases, got):
        st = np.array(c["in"], dtype=np.uint64)
        orc.orc_rescue_permute(KIND[name], ob.ptr(st))
        assert [int(x) for x in st] == c["out"]
        assert [int(x) for x in line.split()] == c["out"]


def _bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


@pytest.mark.parametrize("name", ["rpo", "rpx"])
def test_oracle_lmcs_roots_and_duplex_challenger(orc, name):
    assert orc.orc_set_hash(KIND[name], None, 0) == 0
    for case in V["lmcs"][name]:
        mats, keep = (ob.Matrix * len(case["shapes"]))(), []
        for i, ((h, w), rows) in enumerate(zip(case["shapes"], case["rows"])):
            lg = h.bit_length() - 1
            nat = np.zeros((h, max(w, 0)), dtype=np.uint64)
            for r in range(h):
                nat[r] = rows[_bitrev(r, lg)] if w else []
            nat = np.ascontiguousarray(nat)
            keep.append(nat)
            mats[i] = ob.Matrix(nat.ctypes.data_as(ob.u64p) if w else None, lg, w)
        root = np.zeros(4, dtype=np.uint64)
        orc.orc_lmcs_commit(mats, len(case["shapes"]), ob.ptr(root), None)
        assert [int(x) for x in root] == case["root"], case["shapes"]
    c = V["challenger"][name]
    ch = ob.Challenger()
    for i, v in enumerate(c["capacity"]):
        ch.sponge_state[8 + i] = v
    ops = np.array([{"observe": 0, "sample": 1, "bits": 2}[o] for o, _ in c["script"]], dtype=np.uint32)
    args = np.array([a for _, a in c["script"]], dtype=np.uint64)
    out = np.zeros(len(ops), dtype=np.uint64)
    orc.orc_challenger_script(C.byref(ch), ops.ctypes.data_as(ob.u32p), ob.ptr(args), len(ops), ob.ptr(out))
    assert [int(x) for x in out] == c["results"]


@pytest.mark.parametrize("name", ["rpo", "rpx"])
def test_oracle_prove_verify_tamper(orc, name):
    params = W.fast_pcs_params()
    assert orc.orc_set_hash(KIND[name], None, 0) == 0
    wl = W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2))
    ch = W.initial_challenger(params, H.oracle_observe)
    h, oh, of, oc = H.oracle_prove(params, wl, ch)
    ob.lib().orc_prove_free(h)
    assert H.oracle_verify(params, wl, ch, oh, of, oc)[0] == 0
    bad = of.copy(); bad[len(bad) // 2] ^= 1
    assert H.oracle_verify(params, wl, ch, oh, bad, oc)[0] != 0
    orc.orc_set_hash(0, None, 0)           # the Poseidon2 verifier rejects it
    assert H.oracle_verify(params, wl, W.initial_challenger(params, H.oracle_observe), oh, of, oc)[0] != 0


def _prove_vs_oracle(orc, name, params, wl, aux=None, prep=False):
    assert orc.orc_set_hash(KIND[name], None, 0) == 0
    ch = W.initial_challenger(params, H.oracle_observe)        # the caller's pre-bound duplex state, built with this permutation
    s = B.Session(params, 0)
    try:
        s.set_hash(KIND[name])
        B.lib().mdn_set_debug(s.handle, 1)
        if prep:
            s.set_preprocessed(wl.statement, wl.preprocessed_matrices)
        got = s.prove(wl.statement, wl.matrices, ch, B.AUX_BUILDER(aux) if aux else None)
        h, oh, of, oc = H.oracle_prove(params, wl, ch, aux)
        try:
            names = ["main_root", "aux_root", "quotient_root", "ood_point", "quotient_acc", "deep_evals", "fri_roots", "query_indices"]
            for what, nm in enumerate(names):
                assert np.array_equal(s.info(what), H.oracle_info(h, what)), f"stage {nm} differs"
            assert got[0] == oh and np.array_equal(got[2], oc) and np.array_equal(got[1], of)
        finally:
            ob.lib().orc_prove_free(h)
        assert H.oracle_verify(params, wl, ch, *got)[0] == 0
    finally:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rpo", "rpx"])
@pytest.mark.parametrize("case", ["two_heights", "miden_shape", "host_aux", "preprocessed", "arity8"])
def test_rescue_proofs_bit_exact_vs_oracle(orc, name, case):
    import test_airs
    if case == "two_heights":
        _prove_vs_oracle(orc, name, W.fast_pcs_params(), W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2)))
    elif case == "miden_shape":      # production parameters: grinding through the duplex challenger with this permutation
        _prove_vs_oracle(orc, name, W.miden_pcs_params(), W.Workload([8, 7, 6]))
    elif case == "host_aux":
        wl, aux = test_airs.fib_product_workload([6, 5], lqd=1)
        _prove_vs_oracle(orc, name, W.fast_pcs_params(), wl, aux)
    elif case == "preprocessed":
        _prove_vs_oracle(orc, name, W.fast_pcs_params(), test_airs.preprocessed_workload((5, 6), (True, False)), prep=True)
    else:
        _prove_vs_oracle(orc, name, B.PcsParams(3, 3, 2, 2, 3, 7, 4), W.Workload([6, 7], widths=(9, 12), aux_widths=(1, 2)))


@pytest.mark.gpu
def test_rescue_hash_switch_returns_to_poseidon2(orc):
    params = W.fast_pcs_params()
    small = W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2))
    chp = W.initial_challenger(params, lambda c, f: B.lib().mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(f, dtype=np.uint64)), len(f)))
    s = B.Session(params, 0)
    try:
        a = s.prove(small.statement, small.matrices, chp)
        s.set_hash(B.HASH_RPO)
        b = s.prove(small.statement, small.matrices, chp)
        s.set_hash(B.HASH_RPX)
        c = s.prove(small.statement, small.matrices, chp)
        s.set_hash(B.HASH_POSEIDON2)
        d = s.prove(small.statement, small.matrices, chp)
        assert np.array_equal(a[1], d[1]) and np.array_equal(a[2], d[2])
        assert not np.array_equal(a[2][:1], b[2][:1]) and not np.array_equal(b[2][:1], c[2][:1])
    finally:
        s.close()
