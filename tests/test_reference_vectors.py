"""The reference's OWN unit vectors for this path (tests/golden/reference_unit_vectors.json: literal constants from
the Rust test sources, with file:line) against (a) the oracle and (b) the product's host logic.  These are the pieces
of the protocol the reference pins with expected values rather than with prove->verify round trips: the overwrite-mode
sponge and its zero padding, TreeIndices (fold / shrink / missing siblings = the layout of every opening in the proof
streams), PcsParams validation and the FRI shape."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
import pkgload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
V = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_unit_vectors.json")))
pkg = pkgload.load_pkg()
B = pkg.binding


@pytest.fixture(scope="module")
def orc():
    ob.build()
    L = ob.lib()
    L.orc_mock_sponge.restype = C.c_int
    L.orc_mock_sponge.argtypes = [C.c_uint32, C.c_uint32, ob.u64p, C.c_size_t, ob.u64p]
    L.orc_tree_indices.restype = C.c_longlong
    L.orc_tree_indices.argtypes = [C.c_int, ob.u64p, C.c_size_t, C.c_uint32, C.c_uint32, ob.u64p, C.c_size_t, ob.u32p]
    L.orc_pcs_params_check.restype = C.c_int
    L.orc_pcs_params_check.argtypes = [C.POINTER(ob.PcsParams)]
    L.orc_fri_shape.argtypes = [C.POINTER(ob.PcsParams), C.c_uint32, ob.u32p, ob.u64p]
    return L


@pytest.fixture(scope="module")
def host_units():
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-s", "-C", cpp, "test_host_units"])
    exe = os.path.join(cpp, "test_host_units")

    def run(*args):
        return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=30, check=True).stdout.split()
    return run


def _sponge(orc, width, rate, data):
    a = np.array(data, dtype=np.uint64)
    st = np.zeros(width, dtype=np.uint64)
    assert orc.orc_mock_sponge(width, rate, ob.ptr(a) if len(a) else None, len(a), ob.ptr(st)) == 0
    return [int(x) for x in st]


def test_oracle_sponge_on_the_reference_mock_permutation(orc):
    s = V["stateful_sponge_mock"]
    st = _sponge(orc, s["width"], s["rate"], s["input"])
    assert st == s["state_after"] and st[: s["out"]] == s["output"]
    for width, rate, out in s["alignment_semantic"]["shapes"]:
        for n in range(1, 3 * rate + 1):
            data = list(range(1, n + 1))
            padded = data + [0] * ((-n) % rate)
            assert _sponge(orc, width, rate, data)[:out] == _sponge(orc, width, rate, padded)[:out]
    assert _sponge(orc, 4, 2, []) == [0, 0, 0, 0]      # an empty absorb leaves the state untouched (field_sponge.rs:46-53)


def _orc_indices(orc, op, idx, depth, arg=0):
    a = np.array(idx, dtype=np.uint64)
    out = np.zeros(4 * len(idx) * max(1, depth) + 8, dtype=np.uint64)
    d = C.c_uint32()
    n = orc.orc_tree_indices(op, ob.ptr(a) if len(a) else None, len(a), depth, arg, ob.ptr(out), len(out), C.byref(d))
    return n, [int(x) for x in out[: max(n, 0)]], d.value


def test_oracle_tree_indices_match_the_reference_vectors(orc):
    t = V["tree_indices"]
    for c in t["new"]:
        n, got, _ = _orc_indices(orc, 0, c["indices"], c["depth"])
        assert (n >= 0) == c["ok"], c
        if c["ok"]:
            assert got == c["expect"], c
    for c in t["fold_to_depth"]:
        n, got, d = _orc_indices(orc, 1, c["indices"], c["depth"], c["target"])
        assert (n >= 0) == c["ok"], c
        if c["ok"]:
            assert got == c["expect"] and d == c["expect_depth"], c
    for c in t["shrink_depth"]:
        n, got, d = _orc_indices(orc, 2, c["indices"], c["depth"], c["shift"])
        assert got == c["expect"] and d == c["expect_depth"], c
    for c in t["missing_siblings"]:
        n, got, _ = _orc_indices(orc, 3, c["indices"], c["depth"])
        assert [got[i:i + 2] for i in range(0, len(got), 2)] == c["expect"], c
    for depth in range(1, 6):
        n, got, _ = _orc_indices(orc, 3, [0], depth)
        assert n == 2 * depth and [got[2 * i] for i in range(depth)] == [depth - i for i in range(depth)]


def test_product_host_indices_match_the_reference_vectors(host_units):
    t = V["tree_indices"]
    for c in t["fold_to_depth"]:
        if c["ok"]:
            assert [int(x) for x in host_units("fold", c["depth"], c["target"], *c["indices"])] == c["expect"], c
    for c in t["shrink_depth"]:     # shrink_depth(s) == fold_to_depth(depth - s) (tree_indices.rs:128-137)
        assert [int(x) for x in host_units("fold", c["depth"], c["depth"] - c["shift"], *c["indices"])] == c["expect"], c
    for c in t["missing_siblings"]:
        got = [[int(y) for y in x.split(":")] for x in host_units("siblings", c["depth"], *c["indices"])]
        assert got == c["expect"], c
    for depth in range(1, 6):
        got = [int(x.split(":")[0]) for x in host_units("siblings", depth, 0)]
        assert got == [depth - i for i in range(depth)]


ERR = {None: 0, "InvalidFoldingArity": 1, "ZeroBlowup": 2, "ZeroQueries": 3, "FinalDegreeUnreachable": 4}


def test_pcs_params_validation_and_fri_shape(orc):
    for c in V["pcs_params"]["cases"]:
        assert orc.orc_pcs_params_check(C.byref(ob.PcsParams(*c["args"]))) == ERR[c["error"]], c
    for c in V["fri_shape"]["cases"]:
        r, fd = C.c_uint32(), C.c_uint64()
        orc.orc_fri_shape(C.byref(ob.PcsParams(*c["params"])), c["log_lde"], C.byref(r), C.byref(fd))
        assert (r.value, fd.value) == (c["rounds"], c["final_degree"]), c


def test_product_rejects_invalid_pcs_params_before_touching_a_device():
    """mdn_session_create restates PcsParams::new: the invalid parameter sets of the reference's tests are refused
    with MDN_ERR_INVALID_ARG (-1) whether or not a GPU is present; valid ones get past validation (and then fail
    with MDN_ERR_NO_DEVICE (-6) on a box without a GPU)."""
    lib = B.lib()
    for c in V["pcs_params"]["cases"]:
        h = C.c_void_p()
        rc = lib.mdn_session_create(C.byref(B.PcsParams(*c["args"])), 0, C.byref(h))
        if c["error"] is None:
            assert rc in (0, -6), (c, rc, lib.mdn_last_error(None))
            if rc == 0:
                lib.mdn_session_destroy(h)
        else:
            assert rc == -1, (c, rc)
