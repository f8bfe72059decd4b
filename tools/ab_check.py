#!/usr/bin/env python3
"""A/B check of a library variant on the GPU box, torch-free so that it starts in seconds.

    python tools/ab_check.py [--lib PATH] [--log-height 20] [--proves 4] [--out gpurun_out/ab.json]

For the library at --lib (default: the in-tree libmiden_b200.so) it
  1. runs the Poseidon2 permutation on the device and compares with the reference KAT
     (tests/golden/poseidon2_kat.json) and with the oracle on random + edge states,
  2. proves small workloads (2^6..2^13, mixed heights) and compares the proof streams with the oracle bit for bit,
  3. proves the 2^--log-height synthetic workload --proves times from host traces, checks that the proofs are
     identical and accepted by the oracle verifier, and reports the per-kernel-class times of the last proof.
Prints one JSON line and writes it to --out.  Exit code 0 only if every check passed.
The oracle is used as the checker only (tests/ infrastructure), never on the product path."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--log-height", type=int, default=20)
ap.add_argument("--proves", type=int, default=4)
ap.add_argument("--out", default=None)
a = ap.parse_args()
if a.lib:
    os.environ["MDN_LIB_PATH"] = os.path.abspath(a.lib)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import pkgload  # noqa: E402

pkg = pkgload.load_pkg()
import helpers as H  # noqa: E402
import oracle_binding as ob  # noqa: E402

W, B = pkg.workload, pkg.binding
P = 0xFFFFFFFF00000001
KERNELS = ["transpose", "ntt_lde", "leaf_sponge", "merkle_compress", "constraints", "ood_dot", "deep", "fri", "pow_grind", "gather"]
res = {"lib": B.LIB_PATH, "checks": {}, "ok": False}
lib = B.lib()
orc = ob.lib()
params = W.miden_pcs_params()
sess = B.Session(params, 0)


def observe(c, felts):
    lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))


def check(name, cond):
    res["checks"][name] = bool(cond)
    if not cond:
        print("CHECK FAILED:", name, flush=True)


# 1. permutation
kat = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon2_kat.json")))
rng = np.random.default_rng(7)
st = (rng.integers(0, 2**63, size=(4096, 12), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(4096, 12), dtype=np.uint64)) % np.uint64(P)
st[0] = np.arange(12)
st[1] = P - 1
st[2] = 0
st[3] = [P - 1, 0, 1, 2**32 - 1, 2**32, 2**63, P - 2**32, 7, 8, 2**64 - 2**33, 3, P - 8]
x, y = st.copy(), st.copy()
check("permute_rc", lib.mdn_poseidon2_permute(sess.handle, B.ptr(x.reshape(-1)), len(x)) == 0)
orc.orc_poseidon2_permute(ob.ptr(y.reshape(-1)), len(y))
check("permute_vs_oracle", np.array_equal(x, y))
check("permute_kat", [f"{int(v):016x}" for v in x[0]] == kat["output_hex"])

# 2. small proofs, bit-exact
for hs in ([6, 6, 6], [10, 9, 8], [13, 11, 12]):
    wl = W.Workload(hs)
    ch = W.initial_challenger(params, observe)
    heights, fields, comms = sess.prove(wl.statement, wl.matrices, ch)
    h, oh, of, oc = H.oracle_prove(params, wl, ch)
    check(f"proof_bit_exact_{hs}", heights == oh and np.array_equal(fields, of) and np.array_equal(comms, oc))
    ob.lib().orc_prove_free(h)

# 3. the bench workload
wl = W.Workload([a.log_height] * 3)
ch = W.initial_challenger(params, observe)
proofs, wall = [], []
for i in range(a.proves):
    t0 = time.perf_counter()
    proofs.append(sess.prove(wl.statement, wl.matrices, ch))
    wall.append((time.perf_counter() - t0) * 1e3)
t = sess.timings()
same = all(p[0] == proofs[0][0] and np.array_equal(p[1], proofs[0][1]) and np.array_equal(p[2], proofs[0][2]) for p in proofs)
check("big_proofs_identical", same)
rc, err = H.oracle_verify(params, wl, ch, *proofs[-1])
check("big_proof_verifies", rc == 0)
res["log_height"] = a.log_height
res["wall_ms_pageable_host_traces"] = [round(w, 2) for w in wall]
res["device_total_ms"] = round(float(t.total), 3)
res["kernels_ms"] = {k: round(float(t.kernel_ms[i]), 3) for i, k in enumerate(KERNELS)}
res["phases_ms"] = {k: round(float(getattr(t, k)), 3) for k in ("h2d_transpose", "commit_main", "commit_aux", "evaluate_constraints", "commit_quotient", "open")}
res["permutations"] = int(t.permutations)
res["kernel_launches"] = int(t.kernel_launches)
res["proof_digest"] = f"{int(np.bitwise_xor.reduce(proofs[-1][1])):016x}"
import hashlib  # noqa: E402
res["proof_sha256"] = hashlib.sha256(bytes(proofs[-1][0]) + np.ascontiguousarray(proofs[-1][1], dtype=np.uint64).tobytes()
                                     + np.ascontiguousarray(proofs[-1][2], dtype=np.uint64).tobytes()).hexdigest()
res["ok"] = all(res["checks"].values())
line = json.dumps(res)
print(line)
if a.out:
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    open(a.out, "w").write(line + "\n")
sys.exit(0 if res["ok"] else 1)
