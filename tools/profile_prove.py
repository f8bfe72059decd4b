#!/usr/bin/env python3
"""Run `--proves` proofs of the synthetic workload (device-resident traces) for ncu captures."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pkgload
pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding
ap = argparse.ArgumentParser()
ap.add_argument("--log-height", type=int, default=20)
ap.add_argument("--proves", type=int, default=2)
a = ap.parse_args()
lib = B.lib()
params = W.miden_pcs_params()
wl = W.Workload([a.log_height] * 3)
def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
ch = W.initial_challenger(params, observe)
sess = B.Session(params, 0)
dev = [torch.from_numpy(t.view(np.int64)).cuda() for t in wl.traces]
mats = (B.Matrix * 3)()
for i in range(3):
    mats[i] = B.Matrix(C.cast(dev[i].data_ptr(), B.u64p), a.log_height, wl.widths[i])
for i in range(a.proves):
    sess.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
    t = sess.timings()
    print("prove", i, "total_ms %.2f" % t.total, "launches", t.kernel_launches, [round(x, 2) for x in t.kernel_ms])
