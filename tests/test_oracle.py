"""CPU tests that pin the oracle (oracle/) before it is trusted as the parity checker.

Anchors available in the reference for this path: the Poseidon2 permutation KAT,
the field constants restated by the MASM verifier, NaiveDft-style differentials, and the
prove -> verify round trip (the reference's own parity mechanism).  No golden proofs/roots exist
in the reference, so Merkle roots / proof bytes are "parity unpinned" (see oracle/stark.hpp).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers as H
import oracle_binding as ob

W = H.W
P = W.P
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_poseidon2_permutation_kat(oracle):
    # reference crates/crypto/src/hash/algebraic_sponge/poseidon2/test.rs:7-39 (Plonky3's
    # test_default_goldilocks_poseidon2_width_12)
    kat = json.load(open(os.path.join(GOLDEN, "poseidon2_kat.json")))
    st = np.array(kat["input"], dtype=np.uint64)
    oracle.orc_poseidon2_permute(ob.ptr(st), 1)
    assert [int(x) for x in st] == [int(x, 16) for x in kat["output_hex"]]


def test_poseidon2_x8_matches_scalar(oracle):
    # the AVX-512 packing used by the tree builder (oracle/poseidon2_x8.hpp) against the scalar definition,
    # including 0, p-1 and a ragged tail that falls back to the scalar code
    oracle.orc_poseidon2_permute_x8.argtypes = [ob.u64p, C.c_size_t]
    oracle.orc_poseidon2_permute_x8.restype = C.c_int
    rng = np.random.default_rng(7)
    P = 0xFFFFFFFF00000001
    n = 8 * 25 + 3
    a = rng.integers(0, 2**63, size=12 * n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=12 * n, dtype=np.uint64)
    a = np.where(a >= np.uint64(P), a - np.uint64(P), a).astype(np.uint64)
    a[:12] = 0
    a[12:24] = P - 1
    a[24:36] = np.arange(12)
    b, c = a.copy(), a.copy()
    oracle.orc_poseidon2_permute(ob.ptr(b), n)
    oracle.orc_poseidon2_permute_x8(ob.ptr(c), n)
    assert np.array_equal(b, c)


def test_field_constants(oracle):
    # constants.masm:5 ROOT_UNITY; random_coin.masm:426-440 (g = 7, TWO_ADICITY = 32)
    kat = json.load(open(os.path.join(GOLDEN, "field_constants.json")))
    w = oracle.orc_two_adic_generator(32)
    assert w == kat["root_unity_2_32"] == pow(7, (P - 1) >> 32, P)
    assert pow(w, 1 << 31, P) == P - 1
    for bits in (1, 2, 3, 10, 20, 23):
        assert oracle.orc_two_adic_generator(bits) == pow(w, 1 << (32 - bits), P)
    # canonical LDE shift 7^(2^(32 - log_lde)) (crates/lifted-stark/src/domain.rs:358-361)
    for log_lde in (9, 13, 23):
        assert oracle.orc_lde_shift(log_lde) == pow(7, 1 << (32 - log_lde), P)
    # constants.masm:25-27 pin the LDE offset rule for blowup 8: offset^N and lde_g^N do not depend on N
    for log_n in (6, 10, 20):
        assert pow(oracle.orc_lde_shift(log_n + 3), 1 << log_n, P) == kat["quotient_first_shift"]
        assert pow(oracle.orc_two_adic_generator(log_n + 3), 1 << log_n, P) == kat["quotient_shift_ratio"]
    assert oracle.orc_fp_inv(8 * pow(kat["quotient_first_shift"], 7, P) % P) == kat["quotient_first_weight"]
    rng = np.random.default_rng(1)
    for _ in range(200):
        a, b = (int(x) % P for x in rng.integers(0, 2**63, 2, dtype=np.uint64) * 2 + 1)
        assert oracle.orc_fp_mul(a, b) == a * b % P
        assert oracle.orc_fp_mul(a, oracle.orc_fp_inv(a)) == 1
    for a, b in ((P - 1, P - 1), (P - 1, 2), (0, 5), (1 << 63, 1 << 63), (0xFFFFFFFF, 0xFFFFFFFF00000000)):
        assert oracle.orc_fp_mul(a, b) == a * b % P


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 8])
def test_dft_matches_naive(oracle, log_n):
    # mirrors the reference's NaiveDft differentials (quotient.rs:254-265, periodic.rs:194)
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    col = (rng.integers(0, 2**63, n, dtype=np.uint64) % np.uint64(P)).astype(np.uint64)
    naive = np.zeros(n, dtype=np.uint64)
    fast = np.zeros(n, dtype=np.uint64)
    back = np.zeros(n, dtype=np.uint64)
    oracle.orc_naive_dft(ob.ptr(col), log_n, ob.ptr(naive))
    oracle.orc_dft(ob.ptr(col), log_n, 1, 0, ob.ptr(fast))
    oracle.orc_dft(ob.ptr(fast), log_n, 1, 1, ob.ptr(back))
    assert (naive == fast).all() and (back == col).all()


@pytest.mark.parametrize("log_n,added", [(3, 1), (4, 3), (6, 3)])
def test_coset_lde_against_direct_evaluation(oracle, log_n, added):
    # coset_lde_batch output row bitrev(i) must equal f(shift * omega^i) for the interpolant f.
    n, width = 1 << log_n, 3
    rng = np.random.default_rng(7)
    m = (rng.integers(0, 2**63, (n, width), dtype=np.uint64) % np.uint64(P)).astype(np.uint64)
    coeffs = np.zeros_like(m)
    oracle.orc_dft(ob.ptr(m), log_n, width, 1, ob.ptr(coeffs))
    shift = oracle.orc_lde_shift(log_n + added)
    out = np.zeros((n << added, width), dtype=np.uint64)
    mat = ob.Matrix(ob.ptr(m), log_n, width)
    oracle.orc_coset_lde_batch(C.byref(mat), added, shift, ob.ptr(out))
    big = log_n + added
    w = oracle.orc_two_adic_generator(big)
    for i in range(1 << big):
        x = shift * pow(w, i, P) % P
        br = int(format(i, "0%db" % big)[::-1], 2)
        for c in range(width):
            acc = 0
            for k in reversed(range(n)):
                acc = (acc * x + int(coeffs[k, c])) % P
            assert acc == int(out[br, c])
    # trace rows are recovered on the subgroup: f(omega_H^r) = m[r]
    wh = oracle.orc_two_adic_generator(log_n)
    for r in (0, 1, n - 1):
        x = pow(wh, r, P)
        acc = 0
        for k in reversed(range(n)):
            acc = (acc * x + int(coeffs[k, 0])) % P
        assert acc == int(m[r, 0])


def test_challenger_semantics(oracle):
    # random_coin.masm: absorbed-length tag, lazy flush, pop-from-rate[7], squeeze-only permutation.
    def perm(state):
        s = np.array(state, dtype=np.uint64)
        oracle.orc_poseidon2_permute(ob.ptr(s), 1)
        return [int(x) for x in s]

    c = ob.Challenger()
    for i in range(4):
        c.sponge_state[8 + i] = W.RELATION_DIGEST[i]
    ops = np.array([0, 0, 0, 1, 1, 0] + [0] * 8 + [1] * 9, dtype=np.uint32)
    args = np.array([11, 22, 33, 0, 0, 44] + list(range(100, 108)) + [0] * 9, dtype=np.uint64)
    out = np.zeros(len(ops), dtype=np.uint64)
    oracle.orc_challenger_script(C.byref(c), ops.ctypes.data_as(ob.u32p), ob.ptr(args), len(ops), ob.ptr(out))
    st = [0] * 8 + list(W.RELATION_DIGEST)
    st[0:3] = [11, 22, 33]; st[8] = (st[8] + 3) % P          # partial block: zero-fill + tag 3
    st = perm(st)
    assert int(out[3]) == st[7] and int(out[4]) == st[6]      # samples pop rate[7], rate[6]
    # observe(44) then 8 more: first block = [44, 100..106] tag 8, then 107 pending
    st[0:8] = [44] + list(range(100, 107)); st[8] = (st[8] + 8) % P
    st = perm(st)
    st2 = list(st); st2[0] = 107; st2[1:8] = [0] * 7; st2[8] = (st2[8] + 1) % P
    st2 = perm(st2)
    assert [int(x) for x in out[14:22]] == st2[7::-1][:8]
    st3 = perm(st2)                                           # 9th sample: squeeze-only, no tag
    assert int(out[22]) == st3[7]


def test_grind_smallest_witness(oracle):
    c = ob.Challenger()
    ops = np.array([0, 3], dtype=np.uint32)
    args = np.array([5, 6], dtype=np.uint64)
    out = np.zeros(2, dtype=np.uint64)
    c0 = ob.Challenger()
    oracle.orc_challenger_script(C.byref(c0), ops[:1].ctypes.data_as(ob.u32p), ob.ptr(args[:1]), 1, ob.ptr(out[:1]))
    oracle.orc_challenger_script(C.byref(c), ops.ctypes.data_as(ob.u32p), ob.ptr(args), 2, ob.ptr(out))
    wit = int(out[1])
    for w in range(wit + 1):
        cc = ob.Challenger.from_buffer_copy(c0)
        o = np.zeros(2, dtype=np.uint64)
        oracle.orc_challenger_script(C.byref(cc), np.array([0, 2], dtype=np.uint32).ctypes.data_as(ob.u32p),
                                     ob.ptr(np.array([w, 6], dtype=np.uint64)), 2, ob.ptr(o))
        assert (int(o[1]) == 0) == (w == wit)


def _roundtrip(params, wl, aux_builder=None):
    ch = W.initial_challenger(params, H.oracle_observe)
    h, heights, fields, comms = H.oracle_prove(params, wl, ch, aux_builder)
    ob.lib().orc_prove_free(h)
    rc, err = H.oracle_verify(params, wl, ch, heights, fields, comms)
    assert rc == 0, err
    bad = fields.copy(); bad[len(bad) // 2] ^= np.uint64(1)
    assert H.oracle_verify(params, wl, ch, heights, bad, comms)[0] != 0
    badc = comms.copy(); badc[-1, 0] ^= np.uint64(1)
    assert H.oracle_verify(params, wl, ch, heights, fields, badc)[0] != 0
    return heights, fields, comms


def test_prove_verify_roundtrip_mixed_heights():
    # prove -> verify -> tamper, as crates/lifted-stark/src/testing/configs/goldilocks_poseidon2.rs:143-166
    _roundtrip(W.fast_pcs_params(), W.Workload([6, 7, 5], widths=(11, 9, 10), aux_widths=(2, 1, 1)))


def test_prove_verify_roundtrip_miden_params():
    _roundtrip(W.miden_pcs_params(), W.Workload([8, 8, 8]))


def test_prove_verify_single_air_min_height():
    _roundtrip(W.fast_pcs_params(), W.Workload([3], widths=(9,), aux_widths=(0,)))


def test_prove_verify_with_aux_and_transitions():
    import test_airs
    wl, builder = test_airs.fib_product_workload([6, 4])
    _roundtrip(W.fast_pcs_params(), wl, builder)


def test_prove_verify_periodic_columns():
    import test_airs
    _roundtrip(W.fast_pcs_params(), test_airs.periodic_workload(5, lqd=1))
    _roundtrip(W.fast_pcs_params(), test_airs.periodic_workload(4, lqd=3))


@pytest.mark.parametrize("log_hs,with_prep", [((6, 8), (True, True)), ((5, 7), (True, False)), ((5, 7), (False, True)),
                                              ((6, 6, 4), (True, False, True))])
def test_prove_verify_preprocessed(log_hs, with_prep):
    # Preprocessed tree first in the opening batch, shorter than the max LDE when only short AIRs declare one
    import test_airs
    wl = test_airs.preprocessed_workload(log_hs, with_prep)
    params = W.fast_pcs_params()
    heights, fields, comms = _roundtrip(params, wl)
    ch = W.initial_challenger(params, H.oracle_observe)
    wrong = wl.oracle_prep_commitment.copy(); wrong[1] ^= np.uint64(1)
    assert H.oracle_verify(params, wl, ch, heights, fields, comms, prep_commitment=wrong)[0] != 0
    # a trace that violates the preprocessed relation is rejected
    wl2 = test_airs.preprocessed_workload(log_hs, with_prep)
    i = with_prep.index(True)
    wl2.traces[i][2, 0] ^= np.uint64(1)
    h, hh, ff, cc = H.oracle_prove(params, wl2, ch); ob.lib().orc_prove_free(h)
    assert H.oracle_verify(params, wl2, ch, hh, ff, cc)[0] != 0


def test_logup_aux_built_from_lowered_lookup_air():
    # the oracle's build_logup_aux (device-path restatement of build_logup_aux_trace) must equal an independent
    # Python-integer host builder: identical proofs, and the proof verifies (the constraints check the aux trace)
    import test_airs
    params = W.fast_pcs_params()
    wl_dev, _ = test_airs.logup_workload(5, device=True)
    wl_host, builder = test_airs.logup_workload(5, device=False)
    h1, f1, c1 = _roundtrip(params, wl_dev)
    h2, f2, c2 = _roundtrip(params, wl_host, builder)
    assert h1 == h2 and np.array_equal(f1, f2) and np.array_equal(c1, c2)


def test_prove_verify_big_program():
    import test_airs
    _roundtrip(W.fast_pcs_params(), test_airs.big_program_workload(4, n_terms=60))


@pytest.mark.parametrize("log_blowup", [1, 2, 4])
def test_prove_verify_other_blowups(log_blowup):
    import test_airs
    params = H.B.PcsParams(log_blowup, 2, 1, 1, 2, 6, 3)
    wl, builder = test_airs.fib_product_workload([6], lqd=1)
    _roundtrip(params, wl, builder)


@pytest.mark.parametrize("log_arity", [1, 3])
def test_prove_verify_other_fri_arities(log_arity):
    params = H.B.PcsParams(3, log_arity, 2, 2, 3, 7, 4)
    _roundtrip(params, W.Workload([7, 9], widths=(9, 12), aux_widths=(1, 2)))


def test_violated_constraint_is_rejected():
    wl = W.Workload([5], widths=(9,), aux_widths=(1,))
    wl.traces[0][3, 0] = 1   # column 0 must vanish for the product constraint
    params = W.fast_pcs_params()
    ch = W.initial_challenger(params, H.oracle_observe)
    h, heights, fields, comms = H.oracle_prove(params, wl, ch)
    ob.lib().orc_prove_free(h)
    assert H.oracle_verify(params, wl, ch, heights, fields, comms)[0] != 0


def test_oracle_reproduces_committed_golden_proofs():
    """tests/golden/oracle_proofs.json (made by tests/golden/make_golden.py) pins the oracle itself."""
    sys_path = os.path.join(GOLDEN)
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)
    gold = json.load(open(os.path.join(GOLDEN, "oracle_proofs.json")))
    for name, params, wl, builder in mg.cases():
        ch = W.initial_challenger(params, H.oracle_observe)
        h, heights, fields, comms = H.oracle_prove(params, wl, ch, builder)
        try:
            g = gold[name]
            assert [int(x) for x in H.oracle_info(h, 0)] == g["main_root"], name
            assert [int(x) for x in H.oracle_info(h, 2)] == g["quotient_root"], name
            assert mg.digest(heights, fields, comms) == g["proof_sha256"], name
        finally:
            ob.lib().orc_prove_free(h)
