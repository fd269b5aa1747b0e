"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPUs, gloo in CPU tests).

The synthesis path shards with NO data-path collective (each rank draws and renders its own views).  The training step
has exactly two exchanges: the gradient all-reduce on one flat fp32 buffer, and -- once per
epoch -- the all-reduce of the per-cell error sums / counts that drive the CCV re-weighting."""
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int = None, world_size: int = None) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n units owned by `rank`."""
    if rank is None:
        rank, world_size = world()
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor, bucket_elems: int = 1 << 24) -> torch.Tensor:
    """In-place sum all-reduce of a flat buffer in launch-latency-sized buckets (64 MB of fp32 by default)."""
    if world()[1] == 1:
        return t
    flat = t.view(-1)
    works = [dist.all_reduce(flat[i:i + bucket_elems], op=dist.ReduceOp.SUM, async_op=True)
             for i in range(0, flat.numel(), bucket_elems)]
    for w in works:
        w.wait()
    return t


def allreduce_cell_errors_(err_sum: torch.Tensor, err_cnt: torch.Tensor, occurrence: torch.Tensor = None):
    """Per-epoch exchange for the CCV re-weighting: sums and counts add, the occurrence map ORs."""
    if world()[1] == 1:
        return err_sum, err_cnt, occurrence
    dist.all_reduce(err_sum, op=dist.ReduceOp.SUM)
    dist.all_reduce(err_cnt, op=dist.ReduceOp.SUM)
    if occurrence is not None:
        occ = occurrence.to(torch.int32)
        dist.all_reduce(occ, op=dist.ReduceOp.MAX)
        occurrence.copy_(occ > 0)
    return err_sum, err_cnt, occurrence
