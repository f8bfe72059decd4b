"""`bench.py --impl reference` (the CPU arm the driver runs next to the CUDA arm): one JSON line with the contract's keys, same
`config.workload` string as the CUDA arm would print for that size, and -- under `torch.distributed.run` with two ranks, which
exports OMP_NUM_THREADS=1 -- exactly one line from rank 0 with the OpenMP team restored (the round-1 arm ran single-threaded
there until the driver's timeout).  Small height so that it runs in seconds; no GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check(line, n_gpus):
    assert line["impl"] == "reference" and line["metric"] == "trace cells/sec proved" and line["unit"] == "cells/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == n_gpus and line["gpu_launches"] == 0
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] >= 1
    assert "2^10 x (51,22,16)" in line["config"]["workload"] and "Poseidon2" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1 and "OpenMP threads" in cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1",
                        "--ref-log-height", "10"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = _lines(r.stdout)
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-1500:] + r.stderr[-1500:]
    _check(lines[0], 1)
    assert lines[0]["steps"] == 2 and lines[0]["warmup"] == 1


def test_reference_arm_under_torchrun_two_ranks():
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29741", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--ref-log-height", "10"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    lines = _lines(r.stdout)
    assert r.returncode == 0 and len(lines) == 1, r.stdout[-1500:] + r.stderr[-1500:]      # rank 0 alone prints
    _check(lines[0], 2)
    assert lines[0]["cpu_baseline"]["cores"] == (os.cpu_count() or 1) or lines[0]["cpu_baseline"]["cores"] > 1 or (os.cpu_count() or 1) == 1
