"""Multi-GPU plumbing (one process per GPU, `torch.distributed`).

Round-1 sharding: the unit of work is one proof; rank r proves the r-th trace set of the batch and no
data-path collective is needed (DESIGN.md "Multi-GPU").  The only communication is the timing
reduction (max over ranks) and an optional gather of the 32-byte commitment roots, which is also
the collective the in-proof coset sharding of SURVEY.md section 8(e) will use.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def rank_seed(base_seed: int, rank: int) -> int:
    """Each rank proves a different synthetic trace set (same shape, different values)."""
    return base_seed + rank


def max_over_ranks(seconds: float, device: str = "cpu") -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(cells_per_rank: int, steps: int, seconds: float, device: str = "cpu") -> float:
    """Whole-job cells/s: all ranks' cells divided by the slowest rank's time."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    return world * cells_per_rank * steps / max_over_ranks(seconds, device)


def gather_roots(root: np.ndarray, device: str = "cpu") -> np.ndarray:
    """All-gather of one 4-felt commitment root per rank (int64 view: NCCL/gloo have no uint64)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return root.reshape(1, 4)
    t = torch.from_numpy(root.view(np.int64).copy()).to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


def make_allgather_callback(device: str = "cpu"):
    """`mdn_allgather_fn` for mdn_session_set_shard on top of torch.distributed: gathers `n` u64 words from
    every rank (rank-major).  With the NCCL backend the 32-byte sub-roots travel over NVLink; the
    transport is int64 because NCCL/gloo have no uint64."""
    import ctypes as C
    from .binding import ALLGATHER

    def fn(ctx, send, recv, n):
        try:
            world = dist.get_world_size()
            a = np.ctypeslib.as_array(send, shape=(n,)).view(np.int64).copy()
            t = torch.from_numpy(a).to(device)
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            r = np.ctypeslib.as_array(recv, shape=(world * n,))
            for k, o in enumerate(outs):
                r[k * n:(k + 1) * n] = o.cpu().numpy().view(np.uint64)
            return 0
        except Exception:        # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return 1

    return ALLGATHER(fn)
