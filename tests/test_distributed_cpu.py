"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2: block sharding, the
replicated FPS start draws, and the counter all-reduce (the path's only collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointsecguard_b200 import distributed as D
from pointsecguard_b200 import metrics as MT


def test_shard_partition_is_contiguous_and_balanced():
    for B in (1, 7, 16, 64):
        for W in (1, 2, 3, 8):
            sh = [D.shard_for(B, r, W) for r in range(W)]
            assert sum(s.size for s in sh) == B
            assert sh[0].offset == 0 and all(a.offset + a.size == b.offset for a, b in zip(sh, sh[1:]))
            assert max(s.size for s in sh) - min(s.size for s in sh) <= 1
            assert sum(s.owns_block0 for s in sh if s.size) >= 1
    with pytest.raises(ValueError):
        D.shard_for(4, 2, 2)


def test_sharded_start_draws_equal_the_single_process_draw():
    sizes = [4096, 1024, 256, 64]
    torch.manual_seed(3)
    full = D.draw_starts(sizes, 3, D.Shard(10, 0, 10))
    state_full = torch.get_rng_state()
    parts = []
    for r in range(3):
        torch.manual_seed(3)
        parts.append(D.draw_starts(sizes, 3, D.shard_for(10, r, 3)))
        assert torch.equal(torch.get_rng_state(), state_full)      # every rank consumes the generator identically
    assert torch.equal(torch.cat(parts, dim=2), full)
    # and the single-process draw is the reference's call sequence (pointnet_util.py:75)
    torch.manual_seed(3)
    for t in range(3):
        for l, n in enumerate(sizes):
            assert torch.equal(torch.randint(0, n, (10,), dtype=torch.long).to(torch.int32), full[l, t])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _confusion(pred, lab, mask, target, ncls=13):
    c = np.zeros(ncls * ncls + 4, dtype=np.int64)
    np.add.at(c, lab.reshape(-1) * ncls + pred.reshape(-1), 1)
    c[ncls * ncls] = pred.size
    c[ncls * ncls + 1] = (pred == lab).sum()
    c[ncls * ncls + 2] = mask.sum()
    c[ncls * ncls + 3] = ((pred == target) & mask).sum()
    return c


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                  # same global batch on every rank
        B, N = 6, 512
        lab = rng.integers(0, 13, (B, N))
        pred = np.where(rng.random((B, N)) < 0.7, lab, rng.integers(0, 13, (B, N)))
        mask = lab == 11
        sh = D.shard_for(B)
        assert (sh.offset, sh.size) == (rank * 3, 3) and D.world_size() == world
        torch.manual_seed(5)
        starts = D.draw_starts([512, 128], 2, sh)
        local = torch.from_numpy(_confusion(sh.slice(pred), sh.slice(lab), sh.slice(mask), 7))
        total = D.all_reduce_sum_(local.clone())
        q.put((rank, starts.numpy(), total.numpy(), MT.summarize(total)))
    finally:
        dist.destroy_process_group()


def test_two_rank_counters_equal_the_single_rank_result():
    from oracle import attacks_oracle as AO
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    B, N = 6, 512
    lab = rng.integers(0, 13, (B, N))
    pred = np.where(rng.random((B, N)) < 0.7, lab, rng.integers(0, 13, (B, N)))
    mask = lab == 11
    want = _confusion(pred, lab, mask, 7)
    for rank, starts, total, summary in res:
        assert np.array_equal(total, want)              # N-rank counters == 1-rank counters, exactly
    # the summary reproduces the scripts' arithmetic (NB_nontarget_test_semseg.py:187-212)
    ref = AO.block_metrics(pred, lab.astype(np.float64))
    s = res[0][3]
    assert abs(s["acc"] - ref["acc"]) < 1e-12 and abs(s["miou"] - ref["miou"]) < 1e-9
    assert abs(s["target_acc"] - ((pred == 7) & mask).sum() / mask.sum()) < 1e-12
    # start draws: rank slices concatenate to the single-process draw
    torch.manual_seed(5)
    full = D.draw_starts([512, 128], 2, D.Shard(6, 0, 6)).numpy()
    assert np.array_equal(np.concatenate([res[0][1], res[1][1]], axis=2), full)


def test_bulk_start_draws_equal_the_reference_call_sequence():
    """draw_starts takes the mt19937 outputs in bulk; values and the generator state afterwards must equal the
    reference's per-level torch.randint calls (pointnet_util.py:75), across state-regeneration boundaries."""
    import torch
    from pointsecguard_b200 import distributed as D
    assert D._bulk_selftest()
    for seed, sizes, T, Bg, off, size in ((0, [4096, 1024, 256, 64], 50, 16, 0, 16), (1, [4096, 1024, 256, 64], 7, 128, 48, 16),
                                          (2, [65536, 16384, 4096, 1024], 3, 9, 2, 5), (3, [1000, 333, 77, 5], 11, 4, 0, 4)):
        torch.manual_seed(seed)
        ref = D._draw_starts_loop(sizes, T, Bg)
        tail_ref = torch.rand(5)
        torch.manual_seed(seed)
        got = D.draw_starts(sizes, T, D.Shard(Bg, off, size))
        tail = torch.rand(5)
        assert got.dtype == torch.int32 and tuple(got.shape) == (4, T, size)
        assert torch.equal(got, ref[:, :, off:off + size].permute(1, 0, 2).to(torch.int32))
        assert torch.equal(tail, tail_ref)
    assert D._bulk_ok is True
