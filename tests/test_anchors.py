"""Cross-anchors between the oracle and the reference's OWN second statement of the protocol: the recursive
verifier (MASM in crates/lib/core/asm/{stark,pcs/fri} plus the VM instructions it is built from).  The Rust
prover cannot be run here, but those instructions are plain in-tree Rust with their constants spelled out,
so the formulas below are restated from them (Python integers) and compared with what the oracle's
prover/verifier use.  TEST INFRASTRUCTURE: the oracle is the checker, nothing here touches the product."""
import ctypes as C
import random

import numpy as np

import oracle_binding as ob

P = 0xFFFFFFFF00000001
W7 = 7   # QuadFelt = F[u]/(u^2 - 7): air/src/constraints/ext_field.rs:11-12

# processor/src/execution/operations/fri_ops/mod.rs:186-194
EIGHT = 8
TWO_INV = 9223372034707292161
TAU_INV = 18446462594437873665
TAU2_INV = 18446744069414584320
TAU3_INV = 281474976710656


def ef_add(x, y): return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)
def ef_sub(x, y): return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)
def ef_mul(x, y): return ((x[0] * y[0] + W7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)
def ef_scale(x, k): return (x[0] * k % P, x[1] * k % P)


def vm_fold2(f_x, f_neg_x, ep):
    """fri_ops/mod.rs:238-240: (f_x + f_neg_x + (f_x - f_neg_x) * ep) * TWO_INV"""
    return ef_scale(ef_add(ef_add(f_x, f_neg_x), ef_mul(ef_sub(f_x, f_neg_x), ep)), TWO_INV)


def vm_fri_ext2fold4(query_values, coset, poe, alpha):
    """op_fri_ext2fold4 (fri_ops/mod.rs:48-143): query_values in stack (bit-reversed) order, natural coset
    index, poe = power of the domain generator at the queried position, alpha = layer challenge."""
    v = [query_values[0], query_values[2], query_values[1], query_values[3]]          # reorder_bitrev4 :179-181
    f_tau = [1, TAU_INV, TAU2_INV, TAU3_INV][coset]                                   # get_tau_factor :197-205
    x = poe * f_tau % P
    x_inv = pow(x, P - 2, P)
    ev = ef_scale(alpha, x_inv)                                                       # compute_evaluation_points :219-223
    es = ef_mul(ev, ev)
    tmp0 = vm_fold2(v[0], v[2], ev)                                                   # fold4 :229-234
    tmp1 = vm_fold2(v[1], v[3], ef_scale(ev, TAU_INV))
    return vm_fold2(tmp0, tmp1, es)


def test_fri_constants_are_the_oracles_roots_of_unity(oracle):
    tau = oracle.orc_two_adic_generator(2)
    assert oracle.orc_fp_inv(tau) == TAU_INV                       # fri_ops/tests.rs:24-31
    assert oracle.orc_fp_inv(oracle.orc_fp_mul(tau, tau)) == TAU2_INV
    assert oracle.orc_fp_inv(oracle.orc_fp_mul(oracle.orc_fp_mul(tau, tau), tau)) == TAU3_INV
    assert oracle.orc_fp_inv(2) == TWO_INV
    assert tau == pow(2, 48, P)                                    # 2^96 = -1: the 4th root of unity is a power of two


def test_fri_fold4_matches_the_vm_instruction(oracle):
    """The oracle folds a physical row [y0, y2, y1, y3] at the row's domain point s (fold/arity4.rs:46-121 as
    restated in oracle/stark.hpp); the recursive verifier folds the same four opened values with
    `fri_ext2fold4`, x = poe * tau^-coset being that point.  Same inputs, same challenge -> same value."""
    rnd = random.Random(2025)
    for case in range(400):
        q = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(4)]
        if case == 0:
            q = [(0, 0)] * 4
        if case == 1:
            q = [(P - 1, P - 1)] * 4
        coset = rnd.randrange(4)
        poe = rnd.randrange(1, P)
        alpha = (rnd.randrange(P), rnd.randrange(P))
        want = vm_fri_ext2fold4(q, coset, poe, alpha)
        s = poe * [1, TAU_INV, TAU2_INV, TAU3_INV][coset] % P
        row = np.array([c for v in q for c in v], dtype=np.uint64)
        beta = np.array(alpha, dtype=np.uint64)
        out = np.zeros(2, dtype=np.uint64)
        assert oracle.orc_fri_fold_row(2, ob.ptr(row), C.c_uint64(pow(s, P - 2, P)), ob.ptr(beta), ob.ptr(out)) == 0
        assert (int(out[0]), int(out[1])) == want, case


def test_fold_arities_agree_on_low_degree_rows(oracle):
    """Size-independent property the reference states for every arity (fold/mod.rs docs): when the 2^a values are
    evaluations of a polynomial of degree < 2^a on s<w>, the fold at beta is f(beta).  Pins the bit-reversed row
    order and the orientation of s_inv for arities 2, 4 and 8 at once."""
    rnd = random.Random(7)
    for la in (1, 2, 3):
        a = 1 << la
        w = oracle.orc_two_adic_generator(la)
        for _ in range(50):
            coef = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(a)]
            s = rnd.randrange(1, P)
            beta = (rnd.randrange(P), rnd.randrange(P))

            def ev_base(x):
                acc = (0, 0)
                for c in reversed(coef):
                    acc = ef_add(ef_scale(acc, x), c)
                return acc
            nat = [ev_base(s * pow(w, k, P) % P) for k in range(a)]
            row = [nat[int(format(j, f"0{la}b")[::-1], 2)] for j in range(a)]
            want = (0, 0)
            for c in reversed(coef):
                want = ef_add(ef_mul(want, beta), c)
            r = np.array([c for v in row for c in v], dtype=np.uint64)
            b = np.array(beta, dtype=np.uint64)
            out = np.zeros(2, dtype=np.uint64)
            assert oracle.orc_fri_fold_row(la, ob.ptr(r), C.c_uint64(pow(s, P - 2, P)), ob.ptr(b), ob.ptr(out)) == 0
            assert (int(out[0]), int(out[1])) == want
