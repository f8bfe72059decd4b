#!/usr/bin/env python3
"""Debug helper: JIT vs oracle quotient accumulator under several chunk sizes."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H, oracle_binding as ob, test_airs
W, B = H.W, H.B
lib = B.lib()
params = W.fast_pcs_params()
def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
ch = W.initial_challenger(params, observe)

def run(name, wl, chunk, builder=None):
    os.environ["MDN_JIT_CHUNK"] = str(chunk)
    s = B.Session(params, 0); lib.mdn_set_debug(s.handle, 1); s.set_jit(1)
    cb = B.AUX_BUILDER(builder) if builder else None
    s.prove(wl.statement, wl.matrices, ch, cb)
    g = s.info(4)
    h, *_ = H.oracle_prove(params, wl, ch, builder)
    e = H.oracle_info(h, 4); ob.lib().orc_prove_free(h)
    bad = np.nonzero(g != e)[0]
    print(f"{name:14s} chunk={chunk:6d} jit={list(s.info(8))} mismatches={len(bad)} of {len(g)} first={bad[:6]}", flush=True)

wl_f, bf = test_airs.fib_product_workload([6])
for chunk in (100000, 3, 1):
    run("fib", wl_f, chunk, bf)
wl_l, _ = test_airs.logup_workload(6)
for chunk in (100000, 4):
    run("logup", wl_l, chunk)
for n_terms in (20, 300):
    wl_b = test_airs.big_program_workload(5, n_terms=n_terms)
    for chunk in (100000, 512, 64, 7):
        run(f"big{n_terms}", wl_b, chunk)
