"""Pin the CPU oracle (oracle/) against golden vectors produced by executing the unmodified
reference (oracle/make_golden.py).  CPU only."""
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import attacks_oracle as AO
from oracle import geom as G
from oracle import pointnet2_oracle as PO
from pointsecguard_b200 import synthetic as syn

KINDS = ["uniform", "grid", "clustered", "duplicates", "surface"]
CASES = [(k, 2, 1024, 256, [(0.2, 32), (0.1, 16)], 3) for k in KINDS] + \
        [("sa1", 1, 4096, 1024, [(0.1, 32), (0.05, 16)], 4)]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _xyz(kind, B, N, seed):
    x = syn.make_blocks(B, N, seed, "uniform" if kind == "sa1" else kind)
    return x[:, :3].permute(0, 2, 1).contiguous()


def _assert_nn_idx(mine, ref, dist):
    """The reference takes the first three of an *unstable* torch.sort (pointnet_util.py:302), so
    among exactly equal distances its choice is unspecified.  The oracle fixes
    the tie-break to ascending index; a differing index is accepted only where the two candidates
    are at bit-identical distance."""
    bad = np.argwhere(mine != ref)
    for b, i, k in bad:
        assert dist[b, i, mine[b, i, k]] == dist[b, i, ref[b, i, k]], (b, i, k)


def _is_stable(idx, d2):
    tie = d2[..., 1:] == d2[..., :-1]
    return bool((idx[..., 1:][tie] > idx[..., :-1][tie]).all())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("backend", ["c", "torch"])
def test_geometry_bit_exact(golden_dir, case, backend, monkeypatch):
    kind, B, N, S, balls, seed = case
    g = _load(golden_dir, f"geom_{kind}.npz")
    xyz = _xyz(kind, B, N, seed)
    monkeypatch.setattr(PO, "GEOMETRY", backend)
    start = torch.from_numpy(g["start"]).long()
    fps = PO.farthest_point_sample(xyz, S, start)
    assert np.array_equal(fps.numpy(), g["fps"])
    new_xyz = PO.index_points(xyz, fps)
    for r, k in balls:
        idx = PO.query_ball_point(r, k, xyz, new_xyz)
        assert np.array_equal(idx.numpy(), g[f"ball_r{r}_k{k}"]), (kind, r, k)
    d2, idx = PO.three_nn(xyz, new_xyz)
    assert np.array_equal(d2.numpy(), g["nn_d2"])
    _assert_nn_idx(idx.numpy(), g["nn_idx"], PO.square_distance(xyz, new_xyz).numpy())
    if backend == "c":
        _, _, w = G.three_nn(xyz, new_xyz)
        assert np.array_equal(w.numpy(), g["nn_w"])
        assert _is_stable(idx.numpy(), d2.numpy())
        d = G.square_distance(xyz, new_xyz)
    else:
        d = PO.square_distance(xyz, new_xyz)
    assert np.array_equal(d[:, :4].numpy(), g["sqd_rows"])
    assert np.uint32(zlib.crc32(d.numpy().tobytes())) == g["sqd_crc"]


def test_fps_draws_like_reference(golden_dir):
    """start=None must consume the CPU generator exactly like pointnet_util.py:75."""
    g = _load(golden_dir, "geom_uniform.npz")
    xyz = _xyz("uniform", 2, 1024, 3)
    torch.manual_seed(103)
    fps = PO.farthest_point_sample(xyz, 256)
    assert np.array_equal(fps.numpy(), g["fps"])


@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_model_forward_and_gradient(golden_dir, arch):
    g = _load(golden_dir, f"model_{arch}.npz")
    model = PO.OracleModel(syn.make_state_dict(arch), arch)
    model.trace = []
    x = syn.make_blocks(2, 2048, 0, "uniform").clone().requires_grad_(True)
    torch.manual_seed(0)
    logp, l4 = model(x)
    # index trace identical to the reference's own calls, in call order
    ref_idx = [g[k] for k in sorted(k for k in g if k.startswith("idx"))]
    mine = [v.numpy() for k, v in model.trace if k in ("fps", "ball")]
    assert len(ref_idx) == len(mine)
    for a, b in zip(ref_idx, mine):
        assert np.array_equal(a, b)
    np.testing.assert_allclose(logp.detach().numpy(), g["logp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(l4.detach().numpy(), g["l4"], rtol=1e-5, atol=1e-6)
    y = torch.from_numpy(g["y"]).long()
    cost = torch.nn.functional.cross_entropy(logp.reshape(-1, 13), y.view(-1), reduction="sum") / logp.size(1)
    grad, = torch.autograd.grad(cost, x)
    ref = g["grad"]
    err = np.abs(grad.numpy() - ref).max() / np.abs(ref).max()
    # xyz channels go through the re-derived weight gradient, whose reduction order follows the thread count
    # (measured: 5e-12 absolute at the generating thread count, 3.6e-4 relative single-threaded)
    assert err < 2e-3, err
    errc = np.abs(grad.numpy()[:, 3:] - ref[:, 3:]).max() / np.abs(ref[:, 3:]).max()
    assert errc < 1e-5, errc


def test_attacks_match_reference(golden_dir):
    g = _load(golden_dir, "attack.npz")
    m = PO.OracleModel(syn.make_state_dict("ssg"), "ssg")
    x = syn.make_blocks(2, 4096, 0, "uniform")
    torch.manual_seed(0)
    adv = AO.nb_attack(m, x, g["nb_labels"].astype(np.float64), eps=0.1, alpha=0.05, iters=3)
    same = (adv[:, 3:6].numpy() == g["nb_adv"]).mean()
    assert same > 0.9999, same
    assert torch.equal(adv[:, :3], x[:, :3]) and torch.equal(adv[:, 6:], x[:, 6:])
    dev = (adv[:, 3:6] - x[:, 3:6]).abs().max().item()
    assert 0.1 < dev <= 0.15 + 1e-6           # Q1: returned tensor is the un-projected last step

    x1 = syn.make_blocks(1, 4096, 1, "uniform")
    zl = syn.zband_labels(x1)
    mask = (zl[0] == 11).numpy()
    torch.manual_seed(0)
    adv = AO.tar_nb_attack(m, x1, zl.numpy(), eps=0.5, alpha=0.1, iters=3, target=7, mask=mask)
    same = (adv[:, 3:6].numpy() == g["tnb_adv"]).mean()
    assert same > 0.9999, same

    torch.manual_seed(0)
    adv = AO.nu_attack(m, x1, g["nu_labels"].astype(np.float64), c=0.1, kappa=0, steps=4, lr=0.01)
    np.testing.assert_allclose(adv[:, 3:6].numpy(), g["nu_adv"], rtol=0, atol=2e-5)


def test_tar_nu_and_msg_match_reference(golden_dir):
    g = _load(golden_dir, "attack.npz")
    m = PO.OracleModel(syn.make_state_dict("ssg"), "ssg")
    x1 = syn.make_blocks(1, 4096, 1, "uniform")
    zl = syn.zband_labels(x1)
    mask = (zl[0] == 11).numpy()
    torch.manual_seed(0)
    adv = AO.tar_nu_attack(m, x1, zl.numpy(), c=1, kappa=0, steps=22, lr=0.01, target=7, mask=mask)
    np.testing.assert_allclose(adv.numpy(), g["tnu_adv"], rtol=0, atol=5e-5)

    mm = PO.OracleModel(syn.make_state_dict("msg"), "msg")
    torch.manual_seed(0)
    adv = AO.nb_attack(mm, x1, g["msg_nb_labels"].astype(np.float64), eps=0.1, alpha=0.05, iters=2)
    same = (adv[:, 3:6].numpy() == g["msg_nb_adv"]).mean()
    assert same > 0.9999, same


def test_block_metrics():
    rng = np.random.default_rng(0)
    lab = rng.integers(0, 13, (2, 512))
    pred = lab.copy()
    pred[0, :100] = (pred[0, :100] + 1) % 13
    mt = AO.block_metrics(pred, lab.astype(np.float64))
    assert abs(mt["acc"] - (1 - 100 / 1024)) < 1e-12
    assert mt["seen"].sum() == 1024 and (mt["union"] >= mt["correct"]).all()
