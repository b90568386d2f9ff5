"""GPU parity of the geometric primitives: the CUDA kernels (through the drop-in Python API ->
torch.ops.psg.* -> C ABI) against the golden vectors of the executed reference and against the CPU
oracle, bit for bit."""
import os
import zlib

import numpy as np
import pytest
import torch

from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

KINDS = ["uniform", "grid", "clustered", "duplicates", "surface"]
CASES = [(k, 2, 1024, 256, [(0.2, 32), (0.1, 16)], 3) for k in KINDS] + \
        [("sa1", 1, 4096, 1024, [(0.1, 32), (0.05, 16)], 4)]


@pytest.fixture(scope="module")
def PU():
    from pointsecguard_b200.models import pointnet_util
    return pointnet_util


def _xyz(kind, B, N, seed):
    x = syn.make_blocks(B, N, seed, "uniform" if kind == "sa1" else kind)
    return x[:, :3].permute(0, 2, 1).contiguous()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_primitives_match_reference_goldens(golden_dir, PU, case):
    kind, B, N, S, balls, seed = case
    g = dict(np.load(os.path.join(golden_dir, f"geom_{kind}.npz")))
    xyz = _xyz(kind, B, N, seed).cuda()
    torch.manual_seed(seed + 100)               # same CPU-generator draw as the reference made
    fps = PU.farthest_point_sample(xyz, S)
    assert fps.dtype == torch.int64
    assert np.array_equal(fps.cpu().numpy(), g["fps"])
    new_xyz = PU.index_points(xyz, fps)
    assert torch.equal(new_xyz.cpu(), _xyz(kind, B, N, seed)[torch.arange(B)[:, None], fps.cpu()])
    for r, k in balls:
        idx = PU.query_ball_point(r, k, xyz, new_xyz)
        assert np.array_equal(idx.cpu().numpy(), g[f"ball_r{r}_k{k}"]), (kind, r, k)
    # MSG pair sharing one scan gives the same two index sets
    (r0, k0), (r1, k1) = balls
    i0, i1 = torch.ops.psg.ball_query2(r0, k0, r1, k1, xyz, new_xyz)
    assert np.array_equal(i0.cpu().numpy(), g[f"ball_r{r0}_k{k0}"])
    assert np.array_equal(i1.cpu().numpy(), g[f"ball_r{r1}_k{k1}"])
    d = PU.square_distance(xyz, new_xyz)
    assert np.array_equal(d[:, :4].cpu().numpy(), g["sqd_rows"])
    assert np.uint32(zlib.crc32(d.cpu().numpy().tobytes())) == g["sqd_crc"]
    idx, d2, w = torch.ops.psg.three_nn(xyz, new_xyz)
    assert np.array_equal(d2.cpu().numpy(), g["nn_d2"])
    assert np.array_equal(w.cpu().numpy(), g["nn_w"])
    # the reference's sort is unstable: differing indices are allowed only at bit-equal distances
    mine, ref = idx.cpu().numpy(), g["nn_idx"]
    dn = d.cpu().numpy()
    for b, i, k in np.argwhere(mine != ref):
        assert dn[b, i, mine[b, i, k]] == dn[b, i, ref[b, i, k]]


@pytest.mark.parametrize("N,S", [(64, 16), (256, 64), (1024, 256), (4096, 1024), (6000, 300), (12000, 100), (16384, 512), (20000, 64), (40000, 48),
                                 (65536, 96), (70000, 24)])
def test_fps_sizes_against_oracle(PU, N, S):
    from oracle import geom as G
    B = 3
    gcpu = torch.Generator().manual_seed(N)
    xyz = torch.rand(B, N, 3, generator=gcpu)
    xyz[1] = torch.floor(xyz[1] * 16) / 16          # heavy ties: first-max tie-break must hold
    start = torch.randint(0, N, (B,), generator=gcpu)
    ref = G.fps(xyz, S, start)
    out = torch.ops.psg.fps(xyz.cuda(), S, start)
    assert torch.equal(out.cpu(), ref)
    if 8192 < N <= 65536:                        # cluster kernel (default) above; the single-CTA kernels it replaces here
        from pointsecguard_b200 import _lib as L
        L.psg_set_option(b"fps_cluster", 0)
        try:
            out = torch.ops.psg.fps(xyz.cuda(), S, start)
        finally:
            L.psg_set_option(b"fps_cluster", 1)
        assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("N,S,r,k", [(4096, 1024, 0.1, 32), (1024, 256, 0.2, 32), (300, 7, 0.4, 16), (9000, 100, 0.3, 32),
                                     (64, 16, 0.8, 32)])
def test_ball_query_and_three_nn_against_oracle(PU, N, S, r, k):
    from oracle import geom as G
    B = 2
    gcpu = torch.Generator().manual_seed(N + S)
    xyz = torch.rand(B, N, 3, generator=gcpu)
    xyz[1, : N // 2] = xyz[1, N // 2: N // 2 * 2]    # duplicates
    sel = torch.stack([torch.randperm(N, generator=gcpu)[:S] for _ in range(B)])
    new_xyz = xyz[torch.arange(B)[:, None], sel]
    ref = G.ball_query(r, k, xyz, new_xyz)
    out = PU.query_ball_point(r, k, xyz.cuda(), new_xyz.cuda())
    assert torch.equal(out.cpu(), ref)
    ridx, rd2, rw = G.three_nn(xyz, new_xyz)
    idx, d2, w = torch.ops.psg.three_nn(xyz.cuda(), new_xyz.cuda())
    assert torch.equal(idx.cpu(), ridx) and torch.equal(d2.cpu(), rd2) and torch.equal(w.cpu(), rw)


@pytest.mark.parametrize("kind", ["uniform", "surface", "grid", "clustered", "duplicates"])
def test_three_nn_grid_equals_exhaustive_scan(PU, kind):
    """The uniform-grid 3-NN (csrc/neighbors.cu) must return exactly what the exhaustive kernel returns -- indices,
    distances and weights -- on every kind of cloud, including heavy ties (grid / duplicates) and everything in
    one corner (clustered); queries are all points of the cloud, candidates an FPS subset and a random subset."""
    from pointsecguard_b200 import _lib as L
    B, N = 3, 4096
    xyz = _xyz(kind, B, N, 7).cuda()
    gcpu = torch.Generator().manual_seed(11)
    for S, how in ((1024, "fps"), (256, "fps"), (300, "rand"), (2048, "rand"), (5, "rand")):
        if how == "fps":
            sel = PU.farthest_point_sample(xyz, S)
        else:
            sel = torch.stack([torch.randperm(N, generator=gcpu)[:S] for _ in range(B)]).cuda()
        new_xyz = PU.index_points(xyz, sel)
        try:
            L.psg_set_option(b"nn_grid", 1)
            ref = [t.clone() for t in torch.ops.psg.three_nn(xyz, new_xyz)]
            L.psg_set_option(b"nn_grid", 2)
            out = torch.ops.psg.three_nn(xyz, new_xyz)
        finally:
            L.psg_set_option(b"nn_grid", 0)
        for a, b in zip(out, ref):
            assert torch.equal(a, b), (kind, S, how)


def test_ball_query_far_centroid_pads_with_N(PU):
    """A centroid with no point in range leaves N in every slot, as the reference does."""
    xyz = torch.rand(1, 128, 3).cuda()
    far = torch.full((1, 2, 3), 50.0).cuda()
    idx = PU.query_ball_point(0.1, 8, xyz, far)
    assert (idx == 128).all()


def test_index_points_shapes(PU):
    pts = torch.rand(2, 100, 7).cuda()
    idx = torch.randint(0, 100, (2, 5, 3)).cuda()
    out = PU.index_points(pts, idx)
    assert out.shape == (2, 5, 3, 7)
    assert torch.equal(out, pts[torch.arange(2).view(2, 1, 1), idx])
    idx2 = torch.randint(0, 100, (2, 9)).cuda()
    assert torch.equal(PU.index_points(pts, idx2), pts[torch.arange(2).view(2, 1), idx2])


def test_cpu_tensors_are_refused(PU):
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.psg.square_distance(torch.rand(1, 4, 3), torch.rand(1, 4, 3))


@pytest.mark.parametrize("M,R,grp", [(32768, 4096, 32), (12288, 1024, 0), (196608, 1024, 0), (3000, 7, 0), (32768, 16384, 32),
                                     (5000, 9000, 0)])
def test_csr_by_source_is_a_stable_counting_sort(M, R, grp):
    """psg_csr_build_by_source (the deterministic backward of index_points / interpolation): offsets = exclusive counts
    per source point, perm = the slots of every bucket in ASCENDING order; padded ball-query slots (copies of the
    group's first hit, pointnet_util.py:104-106) are left out.  Covers the one-CTA shared-memory kernel (chunk-ordered
    fill, large buckets: 196608 entries on 1024 keys) and the multi-kernel path (R > 8192)."""
    from pointsecguard_b200 import _lib as L
    P = 3
    rng = np.random.default_rng(M + R)
    keys = rng.integers(0, R, (P, M)).astype(np.int32)
    if grp:
        k3 = keys.reshape(P, M // grp, grp)
        nreal = rng.integers(1, grp + 1, (P, M // grp))
        for p in range(P):
            for g in range(M // grp):
                row = np.sort(rng.choice(R, nreal[p, g], replace=False)) if nreal[p, g] <= R else k3[p, g, :nreal[p, g]]
                k3[p, g, :nreal[p, g]] = row
                k3[p, g, nreal[p, g]:] = row[0]                    # padding = copies of the first hit
        keys = k3.reshape(P, M)
    kd = torch.from_numpy(keys).cuda()
    offs = torch.empty(P * (R + 1), dtype=torch.int32, device="cuda")
    perm = torch.full((P * M,), -1, dtype=torch.int32, device="cuda")
    ws = torch.empty(max(L.psg_csr_workspace(P, M, R), 16), dtype=torch.uint8, device="cuda")
    L.psg_csr_build_by_source(kd.data_ptr(), P, M, R, grp, offs.data_ptr(), perm.data_ptr(), ws.data_ptr(),
                              torch.cuda.current_stream().cuda_stream)
    offs, perm = offs.cpu().numpy().reshape(P, R + 1), perm.cpu().numpy().reshape(P, M)
    for p in range(P):
        slots = np.arange(M)
        keep = np.ones(M, bool)
        if grp:
            k = slots % grp
            first = keys[p, slots - k]
            keep = ~((k > 0) & (keys[p] == first))
        ks, sl = keys[p][keep], slots[keep]
        order = np.argsort(ks, kind="stable")
        counts = np.bincount(ks, minlength=R)
        assert np.array_equal(offs[p, 1:] - offs[p, :-1], counts)
        assert offs[p, 0] == 0 and offs[p, R] == keep.sum()
        assert np.array_equal(perm[p, : keep.sum()], sl[order])
