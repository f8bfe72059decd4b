"""world_size-2 gloo test (CPU) of the multi-GPU plumbing used by bench.py --gpus N."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import pkgload
    pkg = pkgload.load_pkg()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = pkg.parallel
    # rank 1 is "slower": the job time is the max over ranks
    secs = 2.0 if rank == 0 else 3.0
    tput = par.aggregate_throughput(cells_per_rank=1000, steps=3, seconds=secs)
    root = np.arange(4, dtype=np.uint64) + np.uint64(0xFFFFFFFF00000000) * np.uint64(rank)
    roots = par.gather_roots(root)
    seeds = par.rank_seed(2025, rank)
    # the all-gather callback handed to mdn_session_set_shard, called the way the C library calls it
    import ctypes as C
    cb = par.make_allgather_callback("cpu")
    send = np.array([rank + 1, 0xFFFFFFFF00000000 + rank, 7, 8], dtype=np.uint64)
    recv = np.zeros(8, dtype=np.uint64)
    u64p = C.POINTER(C.c_uint64)
    assert cb(None, send.ctypes.data_as(u64p), recv.ctypes.data_as(u64p), 4) == 0
    assert recv.tolist() == [1, 0xFFFFFFFF00000000, 7, 8, 2, 0xFFFFFFFF00000001, 7, 8]
    q.put((rank, tput, roots.tolist(), seeds))
    dist.destroy_process_group()


def test_two_rank_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, tput, roots, seed in res:
        assert abs(tput - 2 * 1000 * 3 / 3.0) < 1e-9          # both ranks agree on the slowest time
        assert roots[0] == [0, 1, 2, 3]
        assert roots[1] == [0xFFFFFFFF00000000 + i for i in range(4)]   # uint64 survives the int64 transport
        assert seed == 2025 + rank
