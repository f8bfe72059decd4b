"""Keccak STARK configuration (reference air/src/config.rs:309-353): Keccak-f[1600] and Keccak-256, the
stateful-sponge LMCS (rate 17, alignment 17) and the Keccak-256 hash challenger, pinned on vectors from a pure-Python Keccak
that tests/golden/make_keccak_vectors.py itself pins on hashlib's SHA3-256 and the published Keccak-256 answers -- for the oracle
and, through the C++ command-line tool, for the product's csrc/keccak.cuh; then full proofs, GPU vs oracle, bit for bit
(`-m gpu`; the same cases run on the CPU kernel emulator)."""
import ctypes as C
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H
import oracle_binding as ob
import pkgload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "keccak_vectors.json")))
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
P = W.P
KECCAK = 2


@pytest.fixture()
def orc_kk():
    ob.build()
    L = ob.lib()
    L.orc_set_hash.restype = C.c_int
    L.orc_set_hash.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    L.orc_keccak256.argtypes = [C.c_char_p, C.c_size_t, C.c_uint8, C.c_char_p]
    L.orc_keccak_f.argtypes = [ob.u64p]
    yield L
    L.orc_set_hash(0, None, 0)


def test_oracle_keccak_matches_hashlib_and_vectors(orc_kk):
    out = C.create_string_buffer(32)
    for n in [0, 1, 7, 8, 135, 136, 137, 271, 272, 273, 1000, 4352]:
        data = bytes(i % 251 for i in range(n))
        orc_kk.orc_keccak256(data, n, 6, out)                       # SHA3-256: same permutation and sponge, padding 0x06
        assert out.raw == hashlib.sha3_256(data).digest(), n
    for v in V["hash"]:
        data = bytes(i % 251 for i in range(v["len"]))
        orc_kk.orc_keccak256(data, len(data), 1, out)
        assert out.raw.hex() == v["digest"], v["len"]
    st = np.arange(25, dtype=np.uint64)
    orc_kk.orc_keccak_f(ob.ptr(st))
    assert [f"{int(x):016x}" for x in st] == V["perm"]


def test_product_keccak_matches_hashlib_and_vectors():
    """csrc/keccak.cuh (the lane-streaming hasher the kernels and the host transcript use) compiled for the host."""
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp, "test_keccak"])
    tool = os.path.join(cpp, "test_keccak")
    for nw in [0, 1, 2, 16, 17, 18, 33, 34, 35, 100, 544]:
        data = bytes(i % 251 for i in range(8 * nw))
        got = subprocess.run([tool, "hash", "6", str(nw)], capture_output=True, text=True, check=True).stdout.strip()
        assert got == hashlib.sha3_256(data).hexdigest(), nw
    for v in V["hash"]:
        got = subprocess.run([tool, "hash", "1", str(v["len"] // 8)], capture_output=True, text=True, check=True).stdout.strip()
        assert got == v["digest"], v["len"]
    assert subprocess.run([tool, "perm"], capture_output=True, text=True, check=True).stdout.split() == V["perm"]


def _bitrev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def test_oracle_keccak_lmcs_roots(orc_kk):
    """Stateful-sponge leaves (rate 17, zero-filled partial chunk, state lifting) + PaddingFreeSponge layers against the Python
    restatement of lifted_tree.rs."""
    assert orc_kk.orc_set_hash(KECCAK, b"", 0) == 0
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
        orc_kk.orc_lmcs_commit(mats, len(case["shapes"]), ob.ptr(root), None)
        assert root.tobytes().hex() == case["root"], case["shapes"]


def test_oracle_keccak_hash_challenger_script(orc_kk):
    c = V["challenger"]
    init = bytes.fromhex(c["initial_input_hex"])
    assert orc_kk.orc_set_hash(KECCAK, init, len(init)) == 0
    ops = np.array([{"observe": 0, "sample": 1, "bits": 2}[o] for o, _ in c["script"]], dtype=np.uint32)
    args = np.array([a for _, a in c["script"]], dtype=np.uint64)
    out = np.zeros(len(ops), dtype=np.uint64)
    ch = ob.Challenger()
    orc_kk.orc_challenger_script(C.byref(ch), ops.ctypes.data_as(ob.u32p), ob.ptr(args), len(ops), ob.ptr(out))
    assert [int(x) for x in out] == c["results"]


def test_oracle_keccak_prove_verify_tamper(orc_kk):
    import test_airs
    params = W.fast_pcs_params()
    init = W.initial_hash_challenger(params)
    assert orc_kk.orc_set_hash(KECCAK, init, len(init)) == 0
    for wl, aux in [(W.Workload([6, 5], widths=(9, 20), aux_widths=(1, 2)), None), test_airs.fib_product_workload([7], lqd=1)]:
        ch = W.Challenger()
        h, oh, of, oc = H.oracle_prove(params, wl, ch, aux)
        ob.lib().orc_prove_free(h)
        assert H.oracle_verify(params, wl, ch, oh, of, oc)[0] == 0
        bad = of.copy(); bad[len(bad) // 2] ^= 1
        assert H.oracle_verify(params, wl, ch, oh, bad, oc)[0] != 0
        badc = oc.copy(); badc[0, 0] ^= 1
        assert H.oracle_verify(params, wl, ch, oh, of, badc)[0] != 0
    # neither of the other two verifiers accepts a Keccak proof
    orc_kk.orc_set_hash(1, init, len(init))
    assert H.oracle_verify(params, wl, W.Challenger(), oh, of, oc)[0] != 0
    orc_kk.orc_set_hash(0, None, 0)
    chp = W.initial_challenger(params, H.oracle_observe)
    assert H.oracle_verify(params, wl, chp, oh, of, oc)[0] != 0


def _prove_keccak_vs_oracle(orc_kk, params, wl, aux=None, prep=False, debug=True):
    init = W.initial_hash_challenger(params)
    assert orc_kk.orc_set_hash(KECCAK, init, len(init)) == 0
    s = B.Session(params, 0)
    try:
        s.set_hash(B.HASH_KECCAK, init)
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
@pytest.mark.parametrize("case", ["two_heights", "miden_shape", "host_aux", "logup_device", "preprocessed", "arity2", "arity8", "blowup2", "wide"])
def test_keccak_proofs_bit_exact_vs_oracle(orc_kk, case):
    import test_airs
    if case == "two_heights":
        _prove_keccak_vs_oracle(orc_kk, W.fast_pcs_params(), W.Workload([6, 5], widths=(9, 12), aux_widths=(1, 2)))
    elif case == "miden_shape":      # production parameters and widths 51/22/16: 3, 2 and 1 chunks of 17; grinding through Keccak-256
        _prove_keccak_vs_oracle(orc_kk, W.miden_pcs_params(), W.Workload([10, 9, 8]))
    elif case == "host_aux":
        wl, aux = test_airs.fib_product_workload([7, 5], lqd=1)
        _prove_keccak_vs_oracle(orc_kk, W.fast_pcs_params(), wl, aux)
    elif case == "logup_device":
        _prove_keccak_vs_oracle(orc_kk, W.fast_pcs_params(), test_airs.logup_workload(6, device=True)[0])
    elif case == "preprocessed":
        _prove_keccak_vs_oracle(orc_kk, W.fast_pcs_params(), test_airs.preprocessed_workload((5, 7), (True, False)), prep=True)
    elif case == "arity2":
        _prove_keccak_vs_oracle(orc_kk, B.PcsParams(3, 1, 2, 2, 3, 7, 4), W.Workload([7, 9], widths=(9, 12), aux_widths=(1, 2)))
    elif case == "arity8":           # FRI rows of 16 lanes: one chunk of the rate
        _prove_keccak_vs_oracle(orc_kk, B.PcsParams(3, 3, 2, 2, 3, 7, 4), W.Workload([7, 9], widths=(9, 12), aux_widths=(1, 2)))
    elif case == "wide":             # widths on both sides of the rate: 17 (exact), 18, 34, 35; aux 9 EF columns = 18 lanes
        _prove_keccak_vs_oracle(orc_kk, W.fast_pcs_params(), W.Workload([6, 6, 7, 8], widths=(17, 18, 34, 35), aux_widths=(9, 1, 1, 1)))
    else:
        wl, aux = test_airs.fib_product_workload([6], lqd=1)
        _prove_keccak_vs_oracle(orc_kk, B.PcsParams(1, 2, 1, 1, 2, 6, 3), wl, aux)


@pytest.mark.gpu
def test_keccak_2_16_bit_exact_and_hash_switch(orc_kk):
    """2^16 x (51,22,16) under Keccak bit-exact against the oracle prover, then one session cycling Poseidon2 -> Keccak ->
    Blake3 -> Poseidon2 returns the Poseidon2 proof again."""
    params = W.miden_pcs_params()
    wl = W.Workload([16, 16, 16])
    _prove_keccak_vs_oracle(orc_kk, params, wl, debug=False)
    orc_kk.orc_set_hash(0, None, 0)
    s = B.Session(params, 0)
    try:
        ch = W.initial_challenger(params, lambda c, f: B.lib().mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(f, dtype=np.uint64)), len(f)))
        small = W.Workload([8, 7, 6])
        a = s.prove(small.statement, small.matrices, ch)
        init = W.initial_hash_challenger(params)
        s.set_hash(B.HASH_KECCAK, init)
        b = s.prove(small.statement, small.matrices, None)
        s.set_hash(B.HASH_BLAKE3, init)
        b3 = s.prove(small.statement, small.matrices, None)
        s.set_hash(B.HASH_POSEIDON2)
        c = s.prove(small.statement, small.matrices, ch)
        assert np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2])
        assert not np.array_equal(a[2][:1], b[2][:1]) and not np.array_equal(b3[2][:1], b[2][:1])
    finally:
        s.close()
