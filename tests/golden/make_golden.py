#!/usr/bin/env python3
"""Regenerate tests/golden/oracle_proofs.json.

The reference holds no golden proofs for this path and cannot be built here,
so these vectors are produced by the CPU oracle (oracle/) -- they pin the oracle against accidental
change and give the GPU tests a committed target that does not depend on re-running the oracle.
They are NOT reference outputs; DESIGN.md section 3 says which parts of the oracle are pinned
against the reference itself.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np  # noqa: E402

import helpers as H  # noqa: E402
import oracle_binding as ob  # noqa: E402

W = H.W


def cases():
    import test_airs
    yield "miden_shape_8_7_6", W.miden_pcs_params(), W.Workload([8, 7, 6]), None
    yield "single_dummy_5", W.fast_pcs_params(), W.Workload([5], widths=(9,), aux_widths=(1,)), None
    wl, b = test_airs.fib_product_workload([6, 4], lqd=1)
    yield "fib_product_mixed_degrees", W.fast_pcs_params(), wl, b
    yield "periodic_6", W.fast_pcs_params(), test_airs.periodic_workload(6, lqd=3), None


def digest(heights, fields, comms):
    h = hashlib.sha256()
    h.update(bytes(heights))
    h.update(np.ascontiguousarray(fields, dtype="<u8").tobytes())
    h.update(np.ascontiguousarray(comms, dtype="<u8").tobytes())
    return h.hexdigest()


def main():
    out = {}
    for name, params, wl, builder in cases():
        ch = W.initial_challenger(params, H.oracle_observe)
        h, heights, fields, comms = H.oracle_prove(params, wl, ch, builder)
        out[name] = {
            "main_root": [int(x) for x in H.oracle_info(h, 0)],
            "aux_root": [int(x) for x in H.oracle_info(h, 1)],
            "quotient_root": [int(x) for x in H.oracle_info(h, 2)],
            "query_indices": [int(x) for x in H.oracle_info(h, 7)],
            "n_fields": int(len(fields)), "n_commitments": int(len(comms)),
            "proof_sha256": digest(heights, fields, comms),
        }
        ob.lib().orc_prove_free(h)
    json.dump(out, open(os.path.join(HERE, "oracle_proofs.json"), "w"), indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
