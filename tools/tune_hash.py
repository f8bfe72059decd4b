#!/usr/bin/env python3
"""Compile-time tuning sweep of the kernels that dominate a proof (they are bound by the two integer pipes, so
registers per thread / resident warps / instruction mix are the levers).  Builds one library per variant and runs
tools/ab_check.py on each (KAT, bit-exact small proofs, 2^20 proof verified by the oracle, kernel-class times).

    python tools/tune_hash.py build [name=flags ...]     # here (nvcc cross-compiles); libraries go to tools/_tune/
    python tools/tune_hash.py run                        # on the GPU box: one line per variant + gpurun_out/tune_*.json

Knobs (all inert by default): -DHASH_MIN_BLOCKS=k (default 6: 80 registers), -DHASH_THREADS_N=n (128), -DP2_INT_UNROLL=u (1),
-DP2_LINEAR_FOLD_IMAD, -DP2_FOLD_ALU, -DP2_EPS_ALU, -DNTT_MIN_BLOCKS=k (unset), -DMDN_GEN1 (first generation),
-DMDN_GL_FAST (carry-flag add/sub/mul for the kernels that use gl:: directly: interpreter, LogUp rows, DEEP, OOD; 35 % fewer
static instructions; logic checked with `make -C tests/emu GEN=-DMDN_GL_FAST` + the emulated parity suite).
Round-1 results: profiles/r1_summary.md "r1l" (HASH_MIN_BLOCKS 1..8, P2_EPS_ALU, linear-layer folds)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "miden-vm_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_tune")
DEFAULT = ["base=", "iu2=-DP2_INT_UNROLL=2", "ht256=-DHASH_THREADS_N=256 -DHASH_MIN_BLOCKS=3", "ht64=-DHASH_THREADS_N=64 -DHASH_MIN_BLOCKS=12",
           "ntt2=-DNTT_MIN_BLOCKS=2", "ntt3=-DNTT_MIN_BLOCKS=3", "ntt4=-DNTT_MIN_BLOCKS=4", "linimad=-DP2_LINEAR_FOLD_IMAD", "glfast=-DMDN_GL_FAST"]


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
