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


def _draw_starts_loop(level_sizes, T: int, global_batch: int) -> torch.Tensor:
    """The reference's own call sequence: per forward, per level, one ``torch.randint(0, N_level, (B,))``
    (pointnet_util.py:75).  int64 [T, levels, B]."""
    out = torch.empty(T, len(level_sizes), global_batch, dtype=torch.long)
    for t in range(T):
        for l, n in enumerate(level_sizes):
            torch.randint(0, n, (global_batch,), dtype=torch.long, out=out[t, l])
    return out


# ---- the same draws in bulk ---------------------------------------------------------------------
# torch.randint on the CPU generator takes one 32-bit mt19937 output per element and returns
# ``output % range`` (ranges below 2**32).  An attack of T forwards makes 4 T such calls; at ~8 us of
# dispatch each that is milliseconds during which the GPU has nothing to do.  The bulk path copies the
# generator's mt19937 state into numpy's MT19937 (same algorithm, same tempering), takes the raw
# outputs in one call, writes the advanced state back, and reduces modulo the level sizes.  It is
# checked once per process against the loop above (values AND the generator's state afterwards) and
# is not used if that check fails.
_STATE_BYTES = 5056          # at::CPUGeneratorImplState: seed u64, left i32, seeded i32, next u64, state u64[624], normal cache
_bulk_ok = None
_bulk_bg = None              # one numpy MT19937 instance, re-pointed at torch's state per call (constructing one seeds it from the OS: 0.2 ms)


def _bulk_raw32(count: int):
    import numpy as np
    st = torch.get_rng_state().numpy().copy()
    if st.shape[0] != _STATE_BYTES:
        raise RuntimeError("unexpected CPU generator state layout")
    left = int(st[8:12].view(np.int32)[0])
    if not (1 <= left <= 624):
        raise RuntimeError("unexpected mt19937 position")
    key = st[24:24 + 624 * 8].view(np.uint64)
    global _bulk_bg
    if _bulk_bg is None:
        _bulk_bg = np.random.MT19937(0)
    bg = _bulk_bg
    bg.state = {"bit_generator": "MT19937", "state": {"key": key.astype(np.uint32), "pos": 625 - left}}
    raw = bg.random_raw(count)
    ns = bg.state["state"]
    key[:] = ns["key"].astype(np.uint64)
    pos = int(ns["pos"])
    st[8:12].view(np.int32)[0] = 625 - pos
    st[16:24].view(np.uint64)[0] = pos
    torch.set_rng_state(torch.from_numpy(st))
    return raw


def _draw_starts_bulk(level_sizes, T: int, global_batch: int) -> torch.Tensor:
    import numpy as np
    L = len(level_sizes)
    raw = _bulk_raw32(T * L * global_batch).reshape(T, L, global_batch)
    mod = np.asarray(level_sizes, dtype=np.uint64).reshape(1, L, 1)
    return torch.from_numpy((raw.astype(np.uint64) % mod).astype(np.int64))


def _bulk_selftest() -> bool:
    saved = torch.get_rng_state()
    try:
        ok = True
        for seed, sizes, T, B in ((12345, [4096, 1000, 256, 64], 37, 5), (7, [65536, 16384, 4096, 1024], 3, 129)):
            torch.manual_seed(seed)
            torch.randint(0, 7, (3,))
            s0 = torch.get_rng_state()
            a = _draw_starts_loop(sizes, T, B)
            tail_a = torch.randint(0, 1 << 30, (700,))
            torch.set_rng_state(s0)
            b = _draw_starts_bulk(sizes, T, B)
            tail_b = torch.randint(0, 1 << 30, (700,))
            ok = ok and torch.equal(a, b) and torch.equal(tail_a, tail_b)
        return ok
    except Exception:
        return False
    finally:
        torch.set_rng_state(saved)


def draw_starts(level_sizes, T: int, shard: Shard) -> torch.Tensor:
    """FPS start indices for T forwards, drawn on the global CPU generator in the reference's call
    order (per forward: level 1..4, each ``torch.randint(0, N_level, (B_global,))``,
    pointnet_util.py:75), sliced to this rank's blocks.  Returns int32 [len(level_sizes), T, size]."""
    global _bulk_ok
    if _bulk_ok is None:
        _bulk_ok = _bulk_selftest()
    if _bulk_ok and max(level_sizes) < (1 << 32):
        full = _draw_starts_bulk(level_sizes, T, shard.global_batch)
    else:
        full = _draw_starts_loop(level_sizes, T, shard.global_batch)
    return full[:, :, shard.offset:shard.offset + shard.size].permute(1, 0, 2).to(torch.int32).contiguous()


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op in a single-process run)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
