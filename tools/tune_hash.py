#!/usr/bin/env python3
"""Compile-time tuning sweep of the kernels that dominate a proof (they are bound by the two integer pipes, so
registers per thread / resident warps / instruction mix are the levers).  Builds one library per variant and runs
tools/ab_check.py on each (KAT, bit-exact small proofs, 2^20 proof verified by the oracle, kernel-class times).

    python tools/tune_hash.py build [name=flags ...]     # here (nvcc cross-compiles); libraries go to tools/_tune/
    python tools/tune_hash.py run                        # on the GPU box: one line per variant + gpurun_out/tune_*.json

Remaining switches: -DHASH_MIN_BLOCKS=k (default 6: 80 registers), -DP2_LINEAR_FOLD_IMAD, -DP2_FOLD_IMAD (both folds back on the FMA
pipe), -DP2_EAGER_INTERNAL (reduce every lane in every internal round), -DDEEP_PTS=k (points per thread of k_deep, default 2),
-DMDN_GL_SLOW (branchy gl:: arithmetic), -DMDN_GEN1 (first generation).  Every other knob was timed in round 2 and removed
(profiles/r2_tuning.md): hash block sizes 64/256, NTT resident-block bounds, internal rounds unrolled by two, x2*(2^32-1) on the
ALU pipe, three-product squarings, the ALU form of the exact divisions and the reduce-every-product OOD kernel all lost or made
no difference; MDN_GL_FAST, the ALU-pipe fold, the lazy internal rounds, the two-point DEEP kernel and the unreduced OOD
accumulation won and are defaults.
Round-1 results: profiles/r1_summary.md "r1l" (HASH_MIN_BLOCKS 1..8, linear-layer folds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "miden-vm_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_tune")
DEFAULT = ["base=", "eager=-DP2_EAGER_INTERNAL", "foldimad=-DP2_FOLD_IMAD", "glslow=-DMDN_GL_SLOW", "mb5=-DHASH_MIN_BLOCKS=5", "deep1=-DDEEP_PTS=1"]


def build(specs):
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        so = os.path.join(OUT, f"libmiden_b200_{name}.so")
        cmd = ["nvcc", "-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
               "-fPIC,-Wno-unused-function", "--expt-relaxed-constexpr", *flags.split(), "-shared", "-o", so,
               os.path.join(CSRC, "kernels.cu"), os.path.join(CSRC, "session.cu"), "-lcudart", "-ldl"]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
    for name, p in procs:
        err = p.communicate()[1]
        assert p.returncode == 0, f"{name}: {err[-2000:]}"
    print("built", len(procs), "variants in", OUT)


def run():
    libs = sorted(f for f in os.listdir(OUT) if f.startswith("libmiden_b200_") and f.endswith(".so"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for f in libs:
        name = f[len("libmiden_b200_"):-3]
        out = os.path.join(ROOT, "gpurun_out", f"tune_{name}.json")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_check.py"), "--lib", os.path.join(OUT, f), "--proves", "3", "--out", out],
                           capture_output=True, text=True, timeout=300)
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            k = j["kernels_ms"]
            print(f"{name}: ok={j['ok']} leaf={k['leaf_sponge']} compress={k['merkle_compress']} ntt={k['ntt_lde']} fri={k['fri']} sha256={j['proof_sha256'][:12]}", flush=True)
        except Exception:
            print(f"{name}: FAILED rc={r.returncode} {r.stdout[-300:]} {r.stderr[-300:]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build(sys.argv[2:] or DEFAULT)
    else:
        run()
