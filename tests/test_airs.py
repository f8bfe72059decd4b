"""Small hand-written AIRs (as op-lists) that exercise everything DummyMidenAir does not:
next-row access, the three row selectors, public values, EF aux columns, randomness, aux values,
mixed base/ext constraints, and a quotient degree below the blowup (upsampling)."""
import ctypes as C

import numpy as np

import pkgload

pkg = pkgload.load_pkg()
W = pkg.workload
AP = pkg.air_program
P = W.P


def ef_mul(x, y):
    return ((x[0] * y[0] + 7 * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def fib_product_program() -> np.ndarray:
    """main: (a, b, c) with a' = b, b' = a + b, c free; aux: p running product of
    (ch0 + a + ch1*b).  Constraint degree 3 -> log_quotient_degree = 1."""
    b = AP.ProgramBuilder()
    a0, b0 = b.main(0, 0), b.main(0, 1)
    a1, b1 = b.main(1, 0), b.main(1, 1)
    tr, first, last = b.is_transition(), b.is_first_row(), b.is_last_row()
    b.assert_zero(first * (a0 - b.public(0)))
    b.assert_zero(first * (b0 - b.public(1)))
    b.assert_zero(tr * (a1 - b0))
    b.assert_zero(tr * (b1 - (a0 + b0)))
    p0, p1 = b.aux(0, 0), b.aux(1, 0)
    factor = b.challenge(0) + a0 + b.challenge(1) * b0
    b.assert_zero_ext(first * (p0 - b.const(1)))
    b.assert_zero_ext(tr * (p1 - p0 * factor))
    b.assert_zero(last * (b0 - b.public(2)))
    b.assert_zero_ext(last * (p0 * factor - b.aux_value(0)))
    b.assert_zero_ext((b.main(0, 2) - b.main(0, 2)) * b.ext_const(3, 5))   # ext constant path, trivially zero
    return b.serialize()


def fib_trace(log_h: int, a=1, b=1) -> np.ndarray:
    n = 1 << log_h
    t = np.zeros((n, 3), dtype=np.uint64)
    for r in range(n):
        t[r] = (a, b, (r * 77 + 5) % P)
        a, b = b, (a + b) % P
    return t


def fib_product_workload(log_heights, with_dummy=False, lqd=1):
    """All instances use the fib/product AIR (each with its own height); public values are those of
    the FIRST instance, so only the first instance's boundary constraints hold unless heights match --
    therefore every instance gets the same height-dependent publics via separate workloads in tests.
    Here: instance i starts from (1, 1) and publics = (1, 1, b_last of instance 0), so only use
    equal heights or a single fib instance plus dummy instances."""
    progs, traces, widths, aux_w, lqds = [], [], [], [], []
    fib_logs = log_heights[:1]
    t0 = fib_trace(fib_logs[0])
    publics = (1, 1, int(t0[-1, 1]))
    progs.append(fib_product_program()); traces.append(t0); widths.append(3); aux_w.append(1); lqds.append(lqd)
    for i, lh in enumerate(log_heights[1:], start=1):
        progs.append(AP.dummy_miden_air()); traces.append(W.synthetic_trace(i, lh, 9)); widths.append(9); aux_w.append(1); lqds.append(3)
    wl = W.Workload(log_heights, widths=widths, aux_widths=aux_w, programs=progs, traces=traces,
                    public_values=publics, log_quotient_degrees=lqds, num_aux_values=[1] * len(log_heights))

    def build_aux(ctx, instance, main, randomness, aux_out, aux_values):
        n = 1 << main.contents.log_height
        if instance != 0:
            for i in range(2 * n):
                aux_out[i] = 0
            aux_values[0] = 0; aux_values[1] = 0
            return 0
        w = main.contents.width
        ch0 = (randomness[0], randomness[1]); ch1 = (randomness[2], randomness[3])
        p = (1, 0)
        for r in range(n):
            aux_out[2 * r], aux_out[2 * r + 1] = p
            a, b_ = main.contents.values[r * w], main.contents.values[r * w + 1]
            f = ((ch0[0] + a + ch1[0] * b_) % P, (ch0[1] + ch1[1] * b_) % P)
            p = ef_mul(p, f)
        aux_values[0], aux_values[1] = p
        return 0

    return wl, build_aux


def periodic_workload(log_h: int, lqd: int = 3):
    """Two periodic columns (periods 4 and 2): col0 = per0 * col1, col2 = col1 + per1."""
    n = 1 << log_h
    per = np.array([[1, 5], [2, 7], [3, 5], [4, 7]], dtype=np.uint64)     # periodic_columns_matrix(): max period 4
    t = W.synthetic_trace(7, log_h, 3)
    for r in range(n):
        c1 = int(t[r, 1])
        t[r, 0] = int(per[r % 4, 0]) * c1 % P
        t[r, 2] = (c1 + int(per[r % 4, 1])) % P
    b = AP.ProgramBuilder()
    b.assert_zero(b.main(0, 0) - b.periodic(0) * b.main(0, 1))
    b.assert_zero(b.main(0, 2) - b.main(0, 1) - b.periodic(1))
    wl = W.Workload([log_h], widths=[3], aux_widths=[0], programs=[b.serialize()], traces=[t],
                    log_quotient_degrees=[lqd], periodic=[per], num_aux_values=[0])
    return wl


def preprocessed_workload(log_hs=(6, 8), with_prep=(True, True), lqd: int = 1):
    """AIRs with preprocessed columns (fixed lookup/selector data, preprocessed.rs): for an AIR with them,
    main[0] = prep[0] * main[1] and main[2] = prep_next[1] + main[1]; otherwise the Fibonacci-free identity
    main[0] = main[1] * main[2].  Heights may differ, so the preprocessed tree is shorter than the max LDE
    domain when only the short AIR declares preprocessed columns (virtual lifting, pcs/prover.rs:22-24)."""
    traces, progs, preps, widths = [], [], [], []
    for i, (lh, wp) in enumerate(zip(log_hs, with_prep)):
        n = 1 << lh
        t = W.synthetic_trace(20 + i, lh, 3)
        b = AP.ProgramBuilder()
        if wp:
            pm = W.synthetic_trace(40 + i, lh, 2)
            for r in range(n):
                c1 = int(t[r, 1])
                t[r, 0] = int(pm[r, 0]) * c1 % P
                t[r, 2] = (int(pm[(r + 1) % n, 1]) + c1) % P
            b.assert_zero(b.main(0, 0) - b.preprocessed(0, 0) * b.main(0, 1))
            b.assert_zero(b.main(0, 2) - b.preprocessed(1, 1) - b.main(0, 1))
            preps.append(pm)
        else:
            for r in range(n):
                t[r, 0] = int(t[r, 1]) * int(t[r, 2]) % P
            b.assert_zero(b.main(0, 0) - b.main(0, 1) * b.main(0, 2))
            preps.append(None)
        traces.append(t); progs.append(b.serialize()); widths.append(3)
    k = len(log_hs)
    return W.Workload(list(log_hs), widths=widths, aux_widths=[0] * k, programs=progs, traces=traces,
                      log_quotient_degrees=[lqd] * k, num_aux_values=[0] * k, preprocessed=preps)


def ef_inv(x):
    n = (x[0] * x[0] - 7 * x[1] * x[1]) % P
    ni = pow(n, P - 2, P)
    return (x[0] * ni % P, (P - x[1]) * ni % P)


def logup_workload(log_h: int, device: bool = True, lqd: int = 2, seed: int = 5):
    """A LogUp AIR in the shape `build_logup_aux_trace` serves (air/src/lookup/aux_builder.rs): main columns
    [x, y, tx, ty, f, m], one periodic selector p (period 4), 3 aux columns, max_message_width 2, 2 bus ids:
      column 0 (accumulator): add(flag f, bus0 (x, y))
      column 1              : insert(always, multiplicity -m, bus0 (tx, ty))
      column 2              : batch(flag p) { add(bus1 (x)); remove(bus1 (tx)) }
    Constraints tie the aux columns to those fractions, so the oracle verifier checks the built aux trace.
    device=True ships the lowered LookupAir (aux built by the backend); device=False returns a host aux
    builder computing the same trace independently with Python integers."""
    import random
    rng = random.Random(seed)
    n = 1 << log_h
    per = np.array([[1], [0], [0], [1]], dtype=np.uint64)
    t = W.synthetic_trace(31, log_h, 6)
    for r in range(n):
        t[r, 4] = rng.randrange(2)
        t[r, 5] = rng.randrange(4)
    MW = 2

    def chal(b):
        return b.challenge(0), b.challenge(1)

    def denoms(b):
        x, y, tx, ty = (b.main(0, c) for c in range(4))
        if isinstance(b, AP.LookupProgramBuilder):
            return b.encode(0, MW, [x, y]), b.encode(0, MW, [tx, ty]), b.encode(1, MW, [x]), b.encode(1, MW, [tx])
        al, be = chal(b)
        g = be * be
        return al + g + x + be * y, al + g + tx + be * ty, al + g * b.const(2) + x, al + g * b.const(2) + tx

    # ---- lowered LookupAir
    lb = AP.LookupProgramBuilder(3)
    d0, d1, d2a, d2b = denoms(lb)
    lb.insert(0, lb.main(0, 4), lb.const(1), d0)
    lb.insert(1, None, -lb.main(0, 5), d1)
    pflag = lb.periodic(0)
    lb.insert(2, pflag, lb.const(1), d2a)
    lb.insert(2, pflag, lb.const(P - 1), d2b)
    # ---- constraints
    b = AP.ProgramBuilder()
    d0, d1, d2a, d2b = denoms(b)
    f, m, p = b.main(0, 4), b.main(0, 5), b.periodic(0)
    a0, a1, a2, a0n = b.aux(0, 0), b.aux(0, 1), b.aux(0, 2), b.aux(1, 0)
    b.assert_zero(f * (f - b.const(1)))
    b.assert_zero_ext(a1 * d1 + m)
    b.assert_zero_ext(a2 * d2a * d2b - (d2b - d2a) * p)
    b.assert_zero_ext(((a0n - a0 - a1 - a2) * d0 - f) * b.is_transition())
    b.assert_zero_ext(a0 * b.is_first_row())
    b.assert_zero_ext(((b.aux_value(0) - a0 - a1 - a2) * d0 - f) * b.is_last_row())
    wl = W.Workload([log_h], widths=[6], aux_widths=[3], programs=[b.serialize()], traces=[t],
                    log_quotient_degrees=[lqd], periodic=[per], num_aux_values=[1],
                    lookups=[(3, lb.serialize())] if device else None)

    def build_aux(ctx, instance, main, randomness, aux_out, aux_values):
        al = (randomness[0], randomness[1]); be = (randomness[2], randomness[3])
        g = ef_mul(be, be); g2 = ((2 * g[0]) % P, (2 * g[1]) % P)
        w = main.contents.width
        v = main.contents.values

        def enc(prefix, e0, e1=None):
            d = ((al[0] + prefix[0] + e0) % P, (al[1] + prefix[1]) % P)
            if e1 is not None:
                d = ((d[0] + be[0] * e1) % P, (d[1] + be[1] * e1) % P)
            return d

        acc = (0, 0)
        for r in range(n):
            x, y, tx, ty, fl, mm = (int(v[r * w + c]) for c in range(6))
            pp = int(per[r % 4, 0])
            f0 = ef_inv(enc(g, x, y)) if fl else (0, 0)
            i1 = ef_inv(enc(g, tx, ty))
            f1 = ((P - mm) * i1[0] % P, (P - mm) * i1[1] % P)
            f2 = (0, 0)
            if pp:
                ia, ib = ef_inv(enc(g2, x)), ef_inv(enc(g2, tx))
                f2 = ((ia[0] - ib[0]) % P, (ia[1] - ib[1]) % P)
            row = [acc, f1, f2]
            for c in range(3):
                aux_out[r * 6 + 2 * c], aux_out[r * 6 + 2 * c + 1] = row[c]
            acc = tuple((acc[q] + f0[q] + f1[q] + f2[q]) % P for q in range(2))
        aux_values[0], aux_values[1] = acc
        return 0

    return wl, (None if device else build_aux)


def big_program_workload(log_h: int, n_terms: int = 300, seed: int = 3, lqd: int = 3):
    """A program with thousands of nodes whose constraints vanish identically (E - E' with E, E' built
    separately), exercising the liveness-based slot allocation of the constraint interpreter."""
    import random
    rng = random.Random(seed)
    width = 12
    b = AP.ProgramBuilder()

    def expr(rs):
        acc = b.const(rs.randrange(P))
        for _ in range(n_terms):
            x = b.main(rs.randrange(2), rs.randrange(width))
            y = b.main(0, rs.randrange(width))
            k = rs.randrange(3)
            term = x * y if k == 0 else (x + y if k == 1 else x - b.const(rs.randrange(P)))
            acc = acc + term if rs.randrange(2) else acc - term
        return acc

    for c in range(4):
        s1 = random.Random(seed * 100 + c)
        s2 = random.Random(seed * 100 + c)
        e1 = expr(s1)
        e2 = expr(s2)
        b.assert_zero(e1 - e2)
        if c == 2:
            b.assert_zero_ext((e1 - e2) * b.challenge(0))
    t = W.synthetic_trace(9, log_h, width)
    wl = W.Workload([log_h], widths=[width], aux_widths=[1], programs=[b.serialize()], traces=[t],
                    log_quotient_degrees=[lqd], num_aux_values=[1])
    return wl
