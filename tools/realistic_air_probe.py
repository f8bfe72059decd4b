#!/usr/bin/env python3
"""Timing-only probe: the 2^20 x (51, 22, 16) proof with constraint programs of the size of the real Miden AIRs
(about 5.2 k field operations per row over the three AIRs) instead of the 19-node
DummyMidenAir of the bench.  The random constraints do not vanish on the random trace, so the proofs are not valid
-- every kernel still does exactly the work it would do on a real trace.  Interpreter vs NVRTC-specialised kernel."""
import ctypes as C, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import pkgload
pkg = pkgload.load_pkg()
W, B, AP = pkg.workload, pkg.binding, pkg.air_program
P = W.P


def random_air(width, aux_width, n_ops, n_constraints, seed):
    rng = random.Random(seed)
    b = AP.ProgramBuilder()
    per = max(2, n_ops // n_constraints)
    for c in range(n_constraints):
        acc = b.main(0, rng.randrange(width))
        for _ in range(per // 2):
            x = b.main(rng.randrange(2), rng.randrange(width))
            k = rng.randrange(4)
            acc = acc * x if k == 0 else (acc + x if k == 1 else (acc - x * b.main(0, rng.randrange(width)) if k == 2 else acc + b.const(rng.randrange(P))))
        if aux_width and c % 7 == 0:
            b.assert_zero_ext(b.aux(0, rng.randrange(aux_width)) * acc + b.challenge(c % 2))
        else:
            b.assert_zero(acc)
    return b.serialize()


lh = int(os.environ.get("LOG_H", "20"))
progs = [random_air(51, 4, 3400, 120, 1), random_air(22, 3, 1300, 60, 2), random_air(16, 1, 500, 30, 3)]
print("nodes per AIR:", [int(p[2]) for p in progs], "constraints:", [int(p[3]) for p in progs])
wl = W.Workload([lh] * 3, programs=progs)
lib = B.lib()
params = W.miden_pcs_params()
def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
ch = W.initial_challenger(params, observe)
sess = B.Session(params, 0)
dev = [torch.from_numpy(t.view(np.int64)).cuda() for t in wl.traces]
mats = (B.Matrix * 3)()
for i in range(3):
    mats[i] = B.Matrix(C.cast(dev[i].data_ptr(), B.u64p), lh, wl.widths[i])
for mode, thr in (("interpreter", 0), ("nvrtc", 1)):
    sess.set_jit(thr)
    t0 = time.time(); sess.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES); first = time.time() - t0
    for _ in range(2):
        sess.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
    t = sess.timings()
    print(f"{mode:11s} total_ms={t.total:.1f} constraints_ms={t.kernel_ms[4]:.2f} first_call_s={first:.1f} jit={[int(v) for v in sess.info(8)]} {sess.jit_status()}", flush=True)
