#!/usr/bin/env python3
"""Constraint-interpreter cost on a large synthetic program (thousands of nodes), 2^18 rows."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pkgload
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
import test_airs
lib = B.lib()
params = W.miden_pcs_params()
def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
ch = W.initial_challenger(params, observe)
sess = B.Session(params, 0)
for n_terms in (50, 150, 400):
    wl = test_airs.big_program_workload(18, n_terms=n_terms)
    nodes = int(wl.programs[0][2])
    for mode, thr in (("interpreter", 0), ("nvrtc", 1)):
        sess.set_jit(thr)
        t0 = time.time()
        sess.prove(wl.statement, wl.matrices, ch)
        first = time.time() - t0
        sess.prove(wl.statement, wl.matrices, ch)
        t = sess.timings()
        print(f"nodes={nodes} {mode:11s} constraints_ms={t.kernel_ms[4]:.2f} total_ms={t.total:.1f} first_call_s={first:.1f} "
              f"jit={list(sess.info(8))} (2^18 rows x 12 cols, 2^21 points)", flush=True)
