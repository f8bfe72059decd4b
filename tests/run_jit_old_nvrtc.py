"""Subprocess of test_gpu_parity.test_jit_self_check_catches_a_miscompiling_nvrtc: forces the older NVRTC that
torch bundles (it miscompiles the 10 k-node program for sm_100a) and checks that the first-use comparison with
the interpreter retires the kernel and the proof stays bit-exact."""
import ctypes as C
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

cands = glob.glob(os.path.join(sys.prefix, "lib", "python*", "site-packages", "nvidia", "cuda_nvrtc", "lib", "libnvrtc.so.12"))
if not cands:
    print("SKIP no bundled nvrtc"); sys.exit(0)
os.environ["MDN_NVRTC_PATH"] = cands[0]
os.environ["MDN_NVRTC_ALLOW_OLD"] = "1"
import helpers as H, oracle_binding as ob, test_airs
W, B = H.W, H.B
lib = B.lib()
params = W.fast_pcs_params()


def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))


ch = W.initial_challenger(params, observe)
wl = test_airs.big_program_workload(5, n_terms=300)
s = B.Session(params, 0)
s.set_jit(1)
heights, fields, comms = s.prove(wl.statement, wl.matrices, ch)
status = s.jit_status()
h, oh, of, oc = H.oracle_prove(params, wl, ch)
ob.lib().orc_prove_free(h)
assert heights == oh and np.array_equal(fields, of) and np.array_equal(comms, oc), "proof differs from the oracle"
print("STATUS", status, "JIT", [int(v) for v in s.info(8)])
if status.startswith("nvrtc 12.9") or "disagreed" not in status:
    print("NOTE this NVRTC build did not miscompile; self-check path not exercised")
else:
    assert [int(v) for v in s.info(8)] == [0]
    h2, f2, c2 = s.prove(wl.statement, wl.matrices, ch)       # retired kernel: interpreter from now on
    assert np.array_equal(f2, of) and [int(v) for v in s.info(8)] == [0]
    print("SELF-CHECK OK")
