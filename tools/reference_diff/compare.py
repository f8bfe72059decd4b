#!/usr/bin/env python3
"""Diff a proof dumped by the REFERENCE prover (tools/reference_diff/b200_reference_dump.rs, run where a Rust toolchain
exists) against this repository's proof of the same statement and traces.

    python tools/reference_diff/compare.py --reference ref_10_9_8.json --log-heights 10 9 8          # oracle (CPU)
    python tools/reference_diff/compare.py --reference ref_20.json --log-heights 20 20 20 --gpu      # CUDA backend

This is the check that turns "parity unpinned" (Merkle roots, proof bytes, PoW witness choice) into
pinned: identical `log_trace_heights`, `fields` and `commitments`.  On a mismatch it reports the first differing
stream position; the transcript order tells which protocol item that is -- e.g. commitments[0] is the
main-trace root (LDE + LMCS), the first field after the aux root is the DEEP PoW witness.
TEST INFRASTRUCTURE: uses the oracle as one of the two sides; nothing here is on the product path."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def flatten(x, out):
    """serde_json of p3/Miden types: Felt = u64 number, Hash = array (possibly wrapped in {"value": [...]})."""
    if isinstance(x, bool):
        raise ValueError("unexpected bool")
    if isinstance(x, int):
        out.append(x)
    elif isinstance(x, list):
        for v in x:
            flatten(v, out)
    elif isinstance(x, dict):
        for v in x.values():
            flatten(v, out)
    else:
        raise ValueError(f"unexpected {type(x)}")
    return out


def load_reference(path):
    j = json.load(open(path))
    heights = flatten(j["log_trace_heights"], [])
    tr = j["transcript"]
    return heights, flatten(tr["fields"], []), flatten(tr["commitments"], [])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--log-heights", type=int, nargs="+", required=True)
    ap.add_argument("--gpu", action="store_true", help="prove with libmiden_b200.so instead of the CPU oracle")
    a = ap.parse_args()
    import numpy as np
    import pkgload
    pkg = pkgload.load_pkg()
    import helpers as H
    import oracle_binding as ob
    W, B = pkg.workload, pkg.binding
    params = W.miden_pcs_params()
    wl = W.Workload(a.log_heights)
    if a.gpu:
        lib = B.lib()

        def observe(c, felts):
            lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))
        ch = W.initial_challenger(params, observe)
        heights, fields, comms = B.Session(params, 0).prove(wl.statement, wl.matrices, ch)
    else:
        ch = W.initial_challenger(params, H.oracle_observe)
        h, heights, fields, comms = H.oracle_prove(params, wl, ch)
    ours = (list(heights), [int(v) for v in np.asarray(fields).reshape(-1)], [int(v) for v in np.asarray(comms).reshape(-1)])
    ref = load_reference(a.reference)
    ok = True
    for name, x, y in zip(("log_trace_heights", "fields", "commitments (4 words each)"), ours, ref):
        if x == y:
            print(f"{name}: identical ({len(x)} values)")
            continue
        ok = False
        n = min(len(x), len(y))
        first = next((i for i in range(n) if x[i] != y[i]), n)
        print(f"{name}: DIFFER -- ours {len(x)} values, reference {len(y)}; first difference at index {first}"
              + (f" (ours {x[first]:#018x}, reference {y[first]:#018x})" if first < n else ""))
    print("PARITY WITH THE REFERENCE PROVER: " + ("bit-exact" if ok else "MISMATCH"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
