#!/usr/bin/env python3
"""e2e through mdn_prove with PAGEABLE host buffers (what a Rust Vec<Felt> is) vs pinned ones."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pkgload
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
lib = B.lib()
params = W.miden_pcs_params()
wl = W.Workload([20] * 3)
def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
ch = W.initial_challenger(params, observe)
sess = B.Session(params, 0)
pin = [torch.from_numpy(t.view(np.int64)).pin_memory() for t in wl.traces]
pm = (B.Matrix * 3)()
for i in range(3):
    pm[i] = B.Matrix(C.cast(pin[i].data_ptr(), B.u64p), 20, wl.widths[i])
for name, mats in (("pageable", wl.matrices), ("pinned", pm)):
    for _ in range(2):
        sess.prove(wl.statement, mats, ch)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); sess.prove(wl.statement, mats, ch); ts.append((time.perf_counter() - t0) * 1e3)
    print(name, "ms per proof:", [round(x, 1) for x in ts], "h2d_transpose ms", round(sess.timings().h2d_transpose, 1))
