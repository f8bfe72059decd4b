#!/usr/bin/env python3
"""Occupancy sweep of the Poseidon2 hashing kernels (they are issue-bound; registers per thread decide how many
warps a scheduler can interleave).  Builds one library per HASH_MIN_BLOCKS value (the __launch_bounds__ minimum of
k_leaf_hash / k_compress, 128 threads per block; none of 1..6 spills on sm_100a) and runs tools/ab_check.py on each.

    python tools/tune_hash.py build [--gen "<extra nvcc flags>"]     # here (nvcc cross-compiles), libs go to tools/_tune/
    python tools/tune_hash.py run                                            # on the GPU box: prints leaf/compress ms per variant
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "miden-vm_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_tune")
VALUES = [1, 2, 3, 4, 5, 6]


def build(gen):
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for mb in VALUES:
        so = os.path.join(OUT, f"libmiden_b200_mb{mb}.so")
        cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
               "-fPIC,-Wno-unused-function", "--expt-relaxed-constexpr", f"-DHASH_MIN_BLOCKS={mb}", *gen.split(), "-shared", "-o", so,
               os.path.join(CSRC, "kernels.cu"), os.path.join(CSRC, "session.cu"), "-lcudart", "-ldl"]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        assert p.wait() == 0
    print("built", len(procs), "variants in", OUT)


def run():
    rows = []
    for mb in VALUES:
        so = os.path.join(OUT, f"libmiden_b200_mb{mb}.so")
        if not os.path.exists(so):
            continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_check.py"), "--lib", so, "--proves", "3"],
                           capture_output=True, text=True, timeout=300)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            rows.append((mb, j["ok"], j["kernels_ms"]["leaf_sponge"], j["kernels_ms"]["merkle_compress"], j["device_total_ms"]))
        except Exception:
            rows.append((mb, False, None, None, None))
        print("HASH_MIN_BLOCKS=%d ok=%s leaf_ms=%s compress_ms=%s total_ms=%s" % rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_hash.json"), "w"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        gen = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == "--gen" else ""
        build(gen)
    else:
        run()
