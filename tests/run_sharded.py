#!/usr/bin/env python3
"""torchrun worker: ONE proof split over WORLD_SIZE ranks (mdn_session_set_shard: cosets of the LDE, Merkle sub-trees,
peer-memory stores + device barrier) must be byte-identical to the same proof on a single device.  Every rank also
proves unsplit as its local reference, so "identical on all ranks" and "identical to the single-GPU proof" are both
asserted.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/run_sharded.py

On the CPU kernel emulator (tests/test_emulated.py; MDN_ALLOW_EMULATOR=1 MDN_EMU_SHM=1) the same host logic and the
same kernels run over gloo: every rank is a process with one emulated device, the arena slabs are POSIX shared memory
mapped through the emulated cudaIpc* calls, and the peer stores / flag barrier cross processes exactly like on NVLink.
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


def same(a, b):
    return a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    emulated = os.environ.get("MDN_ALLOW_EMULATOR") == "1"
    if emulated:
        dist.init_process_group("gloo")
        local, dev_name = 0, "cpu"
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dev_name = f"cuda:{local}"
    lib = B.lib()
    allgather = pkg.parallel.make_allgather_callback(dev_name)

    def observe(c, felts):
        lib.mdn_challenger_observe(C.byref(c), B.ptr(np.ascontiguousarray(felts, dtype=np.uint64)), len(felts))

    import test_airs
    log_h = int(os.environ.get("SHARD_LOG_H", "12"))
    lg = world.bit_length() - 1

    # (name, params, workload, aux builder, preprocessed?)
    cases = [("miden-shape mixed heights", W.miden_pcs_params(), W.Workload([log_h, log_h - 1, log_h - 2]), None)]
    wl2, builder = test_airs.fib_product_workload([8, 6], lqd=1)
    cases.append(("fib + dummy, host aux builder, 2 quotient chunks", W.fast_pcs_params(), wl2, builder))
    cases.append(("periodic columns", W.fast_pcs_params(), test_airs.periodic_workload(7, lqd=3), None))
    cases.append(("LogUp aux built on the device", W.fast_pcs_params(), test_airs.logup_workload(7, device=True)[0], None))
    cases.append(("preprocessed columns", W.fast_pcs_params(), test_airs.preprocessed_workload((6, 8), (True, True)), None))
    cases.append(("preprocessed, short tree", W.fast_pcs_params(), test_airs.preprocessed_workload((5, 7), (True, False)), None))
    for lb in (1, 2, 4):
        if lb >= lg:
            wlb, bb = test_airs.fib_product_workload([7], lqd=1)
            cases.append((f"log_blowup {lb}", B.PcsParams(lb, 2, 1, 1, 2, 6, 3), wlb, bb))
    for la in (1, 3):
        cases.append((f"FRI arity 2^{la}", B.PcsParams(3, la, 2, 2, 3, 7, 4), W.Workload([7, 9], widths=(9, 12), aux_widths=(1, 2)), None))
    cases.append(("tiny: zero FRI rounds", W.miden_pcs_params(), W.Workload([4, 6], widths=(9, 9), aux_widths=(1, 1)), None))
    cases.append(("second shape on the same sessions", W.miden_pcs_params(), W.Workload([log_h - 1, log_h - 1], widths=(20, 9), aux_widths=(2, 1)), None))

    # the Blake3_256 configuration through the same partition (digest stores, sub-trees and openings are hash-agnostic)
    cases.append(("blake3: miden-shape mixed heights", W.miden_pcs_params(), W.Workload([log_h - 1, log_h - 2, log_h - 3]), None, "blake3"))
    wl3, builder3 = test_airs.fib_product_workload([8, 6], lqd=1)
    cases.append(("blake3: host aux builder", W.fast_pcs_params(), wl3, builder3, "blake3"))

    # the Keccak configuration: 25-lane states between height groups, alignment 17 in the openings
    cases.append(("keccak: miden-shape mixed heights", W.miden_pcs_params(), W.Workload([log_h - 1, log_h - 2, log_h - 3]), None, "keccak"))
    cases.append(("keccak: preprocessed columns", W.fast_pcs_params(), test_airs.preprocessed_workload((6, 8), (True, True)), None, "keccak"))

    # RPO: the Poseidon2 kernels instantiated with the Rescue permutation, through the same partition
    cases.append(("rpo: two heights", W.fast_pcs_params(), W.Workload([7, 6], widths=(9, 12), aux_widths=(1, 2)), None, "rpo"))

    if os.environ.get("SHARD_SUBSET"):      # one case per feature: mixed heights, host aux, device LogUp, preprocessed, second shape, each hash
        keep = ("miden-shape mixed heights", "fib + dummy", "LogUp aux", "preprocessed columns", "FRI arity 2^3", "second shape", "blake3: miden", "keccak: miden")
        cases = [c for c in cases if c[0].startswith(keep)]

    sessions = {}     # params tuple -> (single, split): sessions are reused so that arena reuse across shapes is exercised

    def sess_for(params, hash_name="poseidon2"):
        key = tuple(getattr(params, f) for f, _ in params._fields_) + (hash_name,)
        if key not in sessions:
            single = B.Session(params, local)
            split = B.Session(params, local)
            if hash_name != "poseidon2":
                for s_ in (single, split):
                    s_.set_hash({"blake3": B.HASH_BLAKE3, "keccak": B.HASH_KECCAK, "rpo": B.HASH_RPO}[hash_name], W.initial_hash_challenger(params))
            split.set_shard(rank, world, allgather)
            for s_ in (single, split):
                lib.mdn_set_debug(s_.handle, 1 if os.environ.get("SHARD_DEBUG_STAGES") else 0)
            sessions[key] = (single, split)
        return sessions[key]

    for case in cases:
        name, params, wl, aux = case[:4]
        hash_name = case[4] if len(case) > 4 else "poseidon2"
        cb = B.AUX_BUILDER(aux) if aux is not None else None
        ch = W.initial_challenger(params, observe) if hash_name in ("poseidon2", "rpo") else None      # any pre-bound duplex state is a valid statement
        single, split = sess_for(params, hash_name)
        if getattr(wl, "preprocessed", None) is not None:
            single.set_preprocessed(wl.statement, wl.preprocessed_matrices)
            split.set_preprocessed(wl.statement, wl.preprocessed_matrices)
        else:
            single.set_preprocessed(None, None)
            split.set_preprocessed(None, None)
        ref = single.prove(wl.statement, wl.matrices, ch, cb)
        got = split.prove(wl.statement, wl.matrices, ch, cb)
        if os.environ.get("SHARD_DEBUG_STAGES"):
            for what in (0, 1, 2, 4, 5, 6, 7):
                a, b_ = single.info(what), split.info(what)
                assert np.array_equal(a, b_), f"rank {rank}: case '{name}': stage {what} differs (first at {int(np.argmax(a != b_)) if len(a) == len(b_) else 'len'})"
        assert same(got, ref), f"rank {rank}: case '{name}': the split proof differs from the single-device proof"
        again = split.prove(wl.statement, wl.matrices, ch, cb)      # arena reuse: the second proof of a shape
        assert same(again, ref), f"rank {rank}: case '{name}': second split proof differs"
        if rank == 0:
            print(f"  ok: {name}", flush=True)
    # a partition finer than one coset per rank is refused before any device work
    if world > 2:
        p1 = B.PcsParams(1, 2, 1, 1, 2, 6, 3)
        s = B.Session(p1, local)
        s.set_shard(rank, world, allgather)
        wlb, bb = test_airs.fib_product_workload([7], lqd=1)
        try:
            s.prove(wlb.statement, wlb.matrices, W.initial_challenger(p1, observe), B.AUX_BUILDER(bb))
            raise AssertionError("world > 2^log_blowup was accepted")
        except B.ProverError as e:
            assert "[-4]" in str(e), e
        s.close()
    for single, split in sessions.values():
        split.close(); single.close()
    if rank == 0:
        print(f"SHARDED_OK world={world} backend={'gloo/emulator' if emulated else 'nccl'}", flush=True)

    # timing at full size (optional)
    if os.environ.get("SHARD_BENCH"):
        params = W.miden_pcs_params()
        wl = W.Workload([20, 20, 20])
        ch = W.initial_challenger(params, observe)
        dev = [torch.from_numpy(t.view(np.int64)).cuda() for t in wl.traces]
        mats = (B.Matrix * 3)()
        for i in range(3):
            mats[i] = B.Matrix(C.cast(dev[i].data_ptr(), B.u64p), 20, wl.widths[i])
        proofs = {}
        for sharded in (False, True):
            s = B.Session(params, local)
            if sharded:
                s.set_shard(rank, world, allgather)
            for _ in range(3):
                s.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                p = s.prove(wl.statement, mats, ch, None, B.FLAG_DEVICE_TRACES)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            proofs[sharded] = p
            tim = s.timings()
            if rank == 0:
                print(f"SHARD_BENCH world={world} sharded={sharded} ms_per_proof={dt * 1e3:.2f} kernels_ms={[round(x, 2) for x in tim.kernel_ms]} "
                      f"cells_per_s={wl.cells / dt:.4g}", flush=True)
            s.close()
        assert same(proofs[True], proofs[False]), f"rank {rank}: 2^20 split proof differs from the single-GPU proof"
        if rank == 0:
            print("SHARD_BENCH proofs identical", flush=True)
    dist.barrier()
    if rank == 0:
        print("SHARDED_OK world", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
