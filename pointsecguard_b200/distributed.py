"""Multi-GPU host logic: scene blocks shard across the GPUs of one box, one process per GPU.

The reference has no distributed layer; blocks are independent in every
op of the path, so the build adds pure data parallelism (section 8e):

* rank r of W attacks the contiguous slice [r*B/W, (r+1)*B/W) of the global batch;
* every rank seeds the CPU generator identically and draws the FULL ``(B_global,)`` FPS start
  vector of pointnet_util.py:75, then keeps its slice -- so a W-GPU run consumes the generator
  exactly like the 1-GPU run (and like the CPU reference) and produces identical per-block results;
* the only data exchanged is the counter vector of metrics.attack_counters (13x13 confusion matrix
  + 4 scalars, int64): one all-reduce(sum) per evaluated batch -- NCCL over NVLink on the GPUs,
  gloo in the CPU tests;
* NU_attack's accuracy test is a batch-wide sum (nontarget.py:86-87): under sharding the per-step
  hit count is all-reduced (see nu.py); its smoothness term belongs to global block 0 only (:131).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    """This rank's slice of a global batch of ``global_batch`` blocks."""
    global_batch: int
    offset: int
    size: int

    @property
    def owns_block0(self) -> bool:
        return self.offset == 0

    def slice(self, t):
        """Slice dim 0 of a tensor / array laid out over the global batch."""
        return t[self.offset:self.offset + self.size]


def shard_for(global_batch: int, rank: int | None = None, world: int | None = None) -> Shard:
    """Contiguous, balanced partition: the first ``global_batch % world`` ranks get one extra block."""
    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        else:
            rank, world = 0, 1
    if not (0 <= rank < world) or global_batch < 0:
        raise ValueError(f"bad shard request: batch {global_batch}, rank {rank} of {world}")
    q, r = divmod(global_batch, world)
    size = q + (1 if rank < r else 0)
    offset = rank * q + min(rank, r)
    return Shard(global_batch, offset, size)


def draw_starts(level_sizes, T: int, shard: Shard) -> torch.Tensor:
    """FPS start indices for T forwards, drawn on the global CPU generator in the reference's call
    order (per forward: level 1..4, each ``torch.randint(0, N_level, (B_global,))``,
    pointnet_util.py:75), sliced to this rank's blocks.  Returns int32 [len(level_sizes), T, size]."""
    out = torch.empty(len(level_sizes), T, shard.size, dtype=torch.int32)
    for t in range(T):
        for l, n in enumerate(level_sizes):
            full = torch.randint(0, n, (shard.global_batch,), dtype=torch.long)
            out[l, t] = shard.slice(full).to(torch.int32)
    return out


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op in a single-process run)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
