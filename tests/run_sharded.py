#!/usr/bin/env python3
"""torchrun worker: one proof hash-sharded over WORLD_SIZE GPUs must be byte-identical to the same proof
on a single GPU (each rank also proves unsharded as the local reference).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_sharded.py
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import pkgload

pkg = pkgload.load_pkg()
W, B = pkg.workload, pkg.binding


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # On the CPU kernel emulator (tests/test_emulated.py) the same host logic runs over gloo: every rank uses
    # "device" 0 of its own process and the all-gather callback moves host buffers.
    emulated = os.environ.get("MDN_ALLOW_EMULATOR") == "1"
    if emulated:
        dist.init_process_group("gloo")
        local, dev_name = 0, "cpu"
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev_name = f"cuda:{local}"
    lib = B.lib()

    def observe(c, felts):
        lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))

    log_h = int(os.environ.get("SHARD_LOG_H", "12"))
    cases = [(W.miden_pcs_params(), W.Workload([log_h, log_h - 1, log_h - 2]))]
    import test_airs
    wl2, builder = test_airs.fib_product_workload([8, 6], lqd=1)
    cases.append((W.fast_pcs_params(), wl2, builder))
    for case in cases:
        params, wl = case[0], case[1]
        cb = B.AUX_BUILDER(case[2]) if len(case) > 2 else None
        ch = W.initial_challenger(params, observe)
        single = B.Session(params, local)
        ref = single.prove(wl.statement, wl.matrices, ch, cb)
        single.close()
        sh = B.Session(params, local)
        sh.set_shard(rank, world, pkg.parallel.make_allgather_callback(dev_name))
        got = sh.prove(wl.statement, wl.matrices, ch, cb)
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]), \
            f"rank {rank}: sharded proof differs from the single-GPU proof"
        sh.close()
    if rank == 0:
        print(f"SHARDED_OK world={world} backend={'gloo/emulator' if emulated else 'nccl'}")
    # timing at full size (optional)
    if os.environ.get("SHARD_BENCH"):
        params = W.miden_pcs_params()
        wl = W.Workload([20, 20, 20])
        ch = W.initial_challenger(params, observe)
        dev = [torch.from_numpy(t.view(np.int64)).cuda() for t in wl.traces]
        mats = (B.Matrix * 3)()
        for i in range(3):
            mats[i] = B.Matrix(C.cast(dev[i].data_ptr(), B.u64p), 20, wl.widths[i])
        for sharded in (False, True):
            s = B.Session(params, local)
            if sharded:
                s.set_shard(rank, world, pkg.parallel.make_allgather_callback(f"cuda:{local}"))
            for _ in range(3):
                s.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                p = s.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            if sharded:
                import helpers as H
                if rank == 0:
                    rc, err = H.oracle_verify(params, wl, ch, *p)
                    assert rc == 0, err
            tim = s.timings()
            if rank == 0:
                print(f"SHARD_BENCH world={world} sharded={sharded} ms_per_proof={dt * 1e3:.2f} "
                      f"leaf_ms={tim.kernel_ms[2]:.2f} compress_ms={tim.kernel_ms[3]:.2f} cells_per_s={wl.cells / dt:.4g}")
            s.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
