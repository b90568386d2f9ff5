"""GPU parity AT THE SIZES BASELINE.json STATES (configs 1-5), in BOTH MLP modes: the fp32 parity mode and the
tcgen05 TF32 mode that bench.py times -- each compared DIRECTLY with golden vectors made by the reference / the pinned
oracle at the same size (oracle/make_golden_atsize.py), not with each other.

All cases use the trained painted-blocks checkpoints (tests/golden/ckpt_{ssg,msg}_painted.npz, oracle/make_checkpoint.py),
on which the targeted attack reaches a target hit-rate around 0.9, so "matched success rate" is a real gate.

Sign-PGD is chaotic: the unmodified reference and its op-for-op restatement with a different (equally valid) fp32
summation order in ``index_put_(accumulate)`` agree on 100 % of the elements after 3 iterations and on ~87 % after 10
(``sensitivity_identical_fraction`` in atsize_config1.npz); an implementation whose every GEMM rounds differently starts from
more first flips and decorrelates further (measured here: ~0.24 of the elements identical after 10 steps in fp32 mode, with a
per-step sign agreement of 99.99 %).  Whole-trajectory identity therefore cannot be a gate for ANY implementation; it is only
printed.  The sharp at-size gates (asserted below) are:
  * LAST-STEP REPLAY: from the golden trajectory's colours entering its last iteration (``prev``), one attack iteration
    with the same FPS draws must reproduce the golden final step counts rint((adv - ori) / alpha) on
    >= 99.5 % of the elements in fp32 mode (measured 99.99 %) and >= 95 % in TF32 mode (measured 96.1-96.7 %);
  * acc / mIoU / target hit-rate of the adversarial batch: within 0.5 point of the golden run for the 10-iteration MSG case
    (TF32: 1.5 points; its 4 % wrong signs per step move the metrics by about a point: acc 0.474 against 0.484) and the Adam
    attack; for the 10 / 50-iteration sign attacks on the SSG checkpoint within the ORACLE'S OWN scatter: the
    oracle re-run from colours perturbed by 1e-6 ends 2.3 / 2.6 / 1.1 points (acc / mIoU / target hit-rate) away from its
    reference-exact run (atsize_config2_scatter.npz), so +-0.5 point is below what the reference reproduces of itself;
  * NU over coordinates + colours (config 3): the cost of the first steps within rtol 2e-3 (fp32) / 2e-2 (TF32) -- after
    that the FPS picks of moved clouds diverge on ANY rounding difference -- and final acc / mIoU within 0.5 point,
    per-block L2 within 12 % (median within 2 %);
  * indices: bit-exact, read from the engine's own resident buffers (psg_net_read_geometry).
"""
import os

import numpy as np
import pytest
import torch

from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

ORIGIN, TARGET = 11, 7
LEVELS = [4096, 1024, 256, 64]


def _ckpt(golden_dir, arch="ssg"):
    return syn.load_checkpoint(arch)


def _model(arch, sd, mode):
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32, MLP_TF32X3
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.set_mlp_mode({"fp32": MLP_FP32, "tf32": MLP_TF32, "x3": MLP_TF32X3}[mode])
    return m


def _metrics(model, x, labels, mask=None, seed=1):
    from pointsecguard_b200 import metrics as MT
    torch.manual_seed(seed)
    with torch.no_grad():
        logp, _ = model(x)
    c = MT.attack_counters(logp, labels.cuda(), mask.cuda() if mask is not None else None, TARGET if mask is not None else -1)
    return MT.summarize(c.cpu(), 13)


def _steps(adv, x, alpha):
    return np.rint(((adv[:, 3:6] - x[:, 3:6]) / alpha).cpu().numpy()).astype(np.int8)


def _check_metrics(mine, g, prefix, keys=("acc", "miou"), tol=0.005):
    for k in keys:
        want = float(g[f"{prefix}_{k}"])
        t = tol[k] if isinstance(tol, dict) else tol
        assert abs(mine[k] - want) < t, (prefix, k, mine[k], want, t)


def _sign_attack_tolerance(golden_dir):
    """Metric tolerance of the chaotic sign attacks = 1.5 x the largest deviation the ORACLE shows from its own reference-exact
    run when its input colours are perturbed by 1e-6 (atsize_config2_scatter.npz: acc +-2.5 pt, mIoU +-2.9 pt, target
    hit-rate +-1.5 pt) -- the reference does not reproduce its own metrics to the north-star's +-0.5 pt."""
    sc = np.load(os.path.join(golden_dir, "atsize_config2_scatter.npz"))
    ref = np.load(os.path.join(golden_dir, "atsize_config2.npz"))
    return {k: max(0.005, 1.5 * float(np.abs(sc[k] - float(ref["adv_" + k])).max())) for k in ("acc", "miou", "target_acc")}


def _replay_last_step(make_attack, xd, labels_np, g, alpha, iters, sel=None):
    """One iteration from the golden trajectory's state before its last iteration, with the FPS draws of that forward."""
    from pointsecguard_b200 import distributed as D
    B = xd.shape[0]
    code = torch.from_numpy(g["prev"].astype(np.int16)).cuda()              # int8 code: -128 / 127 = clipped to 0 / 1
    col = xd[:, 3:6] + code.float() * alpha
    col = torch.where(code == -128, torch.zeros_like(col), torch.where(code == 127, torch.ones_like(col), col)).contiguous()
    if len(g["prev_fix_idx"]):               # elements off the ori + k * alpha lattice (clipped earlier, stepped back since)
        col.view(-1)[torch.from_numpy(g["prev_fix_idx"].astype(np.int64)).cuda()] = torch.from_numpy(g["prev_fix_val"]).cuda()
    x2 = xd.clone()
    x2[:, 3:6] = col
    torch.manual_seed(0)
    D.draw_starts(LEVELS, iters - 1, D.Shard(B, 0, B))        # the draws of the first iters-1 forwards
    adv = make_attack(1)(x2, labels_np)
    mine = np.rint(((adv[:, 3:6] - xd[:, 3:6]) / alpha).cpu().numpy()).astype(np.int8)
    same = mine == g["steps"]
    return float(same[sel].mean() if sel is not None else same.mean())


REPLAY_FLOOR = {"fp32": 0.995, "tf32": 0.95, "x3": 0.995}     # x3: 3xTF32 per-layer tcgen05 GEMMs, held to the fp32 gates


@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_config1_nb_b4_vs_unmodified_reference(golden_dir, mode):
    """configs[0]: the reference's own CPU case; the golden is torchattacks.NB_attack of the reference itself."""
    from pointsecguard_b200 import torchattacks
    g = np.load(os.path.join(golden_dir, "atsize_config1.npz"))
    assert float(g["oracle_identical_fraction"]) == 1.0              # the op-for-op oracle IS the reference at this size
    sens = float(g["sensitivity_identical_fraction"])
    m = _model("ssg", _ckpt(golden_dir), mode)
    x, labels = syn.make_painted_blocks(4, 4096, 0)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    _check_metrics(_metrics(m, xd, labels), g, "clean")
    mk = lambda it: torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=it)
    torch.manual_seed(0)
    adv = mk(10)(xd, lab)
    same = float((_steps(adv, xd, 0.05) == g["steps"]).mean())
    got = _metrics(m, adv, labels)
    replay = _replay_last_step(mk, xd, lab, g, 0.05, 10)
    print(f"config1 {mode}: last-step replay identical {replay:.5f}; whole trajectory identical {same:.5f} (reference vs its own "
          f"restatement with another summation order: {sens:.5f}); adv acc {got['acc']:.4f} (ref {float(g['adv_acc']):.4f}) "
          f"mIoU {got['miou']:.4f} (ref {float(g['adv_miou']):.4f})")
    assert replay >= REPLAY_FLOOR[mode]
    # (4 blocks instead of the 16 the scatter was measured on: twice the tolerance)
    _check_metrics(got, g, "adv", tol={k: 2 * v for k, v in _sign_attack_tolerance(golden_dir).items()})
    assert got["acc"] < float(g["clean_acc"]) - 0.1                   # the attack did something


@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_config2_tar_nb_b16_50_iterations(golden_dir, mode):
    """configs[1] (the bench workload) at full size against the reference-exact oracle trajectory: last-step replay, acc,
    mIoU and a NON-ZERO target hit-rate."""
    from pointsecguard_b200 import torchattacks
    g = np.load(os.path.join(golden_dir, "atsize_config2.npz"))
    assert float(g["adv_target_acc"]) > 0.5, "the golden itself must show a successful targeted attack"
    m = _model("ssg", _ckpt(golden_dir), mode)
    x, labels = syn.make_painted_blocks(16, 4096, 0)
    mask = labels == ORIGIN
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mk = lambda it: torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=it, target=TARGET, mask=mask)
    torch.manual_seed(0)
    adv = mk(50)(xd, lab)
    steps = _steps(adv, xd, 0.1)
    mk3 = mask.unsqueeze(1).expand(-1, 3, -1).numpy()
    same_masked = float((steps[mk3] == g["steps"][mk3]).mean())
    assert (steps[~mk3] == 0).all() and torch.equal(adv[:, :3], xd[:, :3]) and torch.equal(adv[:, 6:], xd[:, 6:])
    assert torch.equal(adv[:, 3:6][~torch.from_numpy(mk3)], xd[:, 3:6][~torch.from_numpy(mk3)])   # unmasked colours untouched
    got = _metrics(m, adv, labels, mask)
    replay = _replay_last_step(mk, xd, lab, g, 0.1, 50, sel=mk3)
    print(f"config2 {mode}: last-step replay identical on masked points {replay:.5f}; whole trajectory {same_masked:.5f}; "
          f"adv acc {got['acc']:.4f} (oracle {float(g['adv_acc']):.4f}) mIoU {got['miou']:.4f} ({float(g['adv_miou']):.4f}) "
          f"target_acc {got['target_acc']:.4f} ({float(g['adv_target_acc']):.4f})")
    assert replay >= REPLAY_FLOOR[mode]
    assert got["target_acc"] is not None and got["target_acc"] > 0.5
    tol = _sign_attack_tolerance(golden_dir)
    print(f"    metric tolerance from the oracle's own scatter: {tol}")
    _check_metrics(got, g, "adv", ("acc", "miou", "target_acc"), tol=tol)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_config3_nu_coordinates_and_colours_b32_100_steps(golden_dir, mode):
    from pointsecguard_b200 import torchattacks
    g = np.load(os.path.join(golden_dir, "atsize_config3.npz"))
    m = _model("ssg", _ckpt(golden_dir), mode)
    x, labels = syn.make_painted_blocks(32, 4096, 0)
    xd = x.cuda()
    atk = torchattacks.NU_attack(m, c=0.1, kappa=0, steps=100, lr=0.01, field=(0, 6))
    torch.manual_seed(0)
    adv = atk(xd, labels.numpy().astype(np.float64))      # (B = 32: the acc / 4096 early exit of nontarget.py:87 never fires)
    cost = atk.last_cost.cpu().numpy().astype(np.float64)
    rt = 2e-3 if mode != "tf32" else 2e-2
    print(f"config3 {mode}: cost[0..2] {cost[:3]} oracle {g['cost'][:3]}; cost[99] {cost[-1]:.3f} oracle {g['cost'][-1]:.3f}")
    np.testing.assert_allclose(cost[:3], g["cost"][:3], rtol=rt)
    assert abs(cost[-1] - g["cost"][-1]) < 0.05 * abs(g["cost"][-1])
    got = _metrics(m, adv, labels)
    l2 = ((adv - xd) ** 2).flatten(1).sum(1).cpu().numpy()
    print(f"config3 {mode}: adv acc {got['acc']:.4f} (oracle {float(g['adv_acc']):.4f}) mIoU {got['miou']:.4f} "
          f"({float(g['adv_miou']):.4f}); L2 ratio {np.median(l2 / g['l2_per_block']):.4f}")
    _check_metrics(got, g, "adv")
    assert np.all(np.abs(l2 / g["l2_per_block"] - 1.0) < 0.12) and abs(np.median(l2 / g["l2_per_block"]) - 1.0) < 0.02


@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_config4_msg_nb_b64(golden_dir, mode):
    from pointsecguard_b200 import torchattacks
    g = np.load(os.path.join(golden_dir, "atsize_config4.npz"))
    m = _model("msg", _ckpt(golden_dir, "msg"), mode)
    x, labels = syn.make_painted_blocks(64, 4096, 0)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    _check_metrics(_metrics(m, xd, labels), g, "clean")
    mk = lambda it: torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=it)
    torch.manual_seed(0)
    adv = mk(10)(xd, lab)
    same = float((_steps(adv, xd, 0.05) == g["steps"]).mean())
    got = _metrics(m, adv, labels)
    replay = _replay_last_step(mk, xd, lab, g, 0.05, 10)
    print(f"config4 {mode}: last-step replay identical {replay:.5f}; whole trajectory {same:.5f}; adv acc {got['acc']:.4f} "
          f"(oracle {float(g['adv_acc']):.4f}) mIoU {got['miou']:.4f} ({float(g['adv_miou']):.4f})")
    assert replay >= REPLAY_FLOOR[mode]
    # measured: fp32 0.1 pt, 3xTF32 0.4-0.5 pt, TF32 0.9 pt (ten chaotic sign steps after the trajectories part: the
    # last-step replay above is the sharp gate, these are sanity bounds)
    _check_metrics(got, g, "adv", tol={"fp32": 0.005, "x3": 0.01, "tf32": 0.015}[mode])


# ----------------------------------------------------------------------------------------------------------------
# config 5: large blocks.  The engine's OWN index buffers (grid ball query, grid 3-NN, cluster FPS) against the oracle.
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [4096, 16384, 65536])
def test_engine_resident_indices_vs_oracle(golden_dir, N):
    from oracle import geom as G
    m = _model("ssg", _ckpt(golden_dir), "tf32")
    x, _ = syn.make_painted_blocks(1, N, 3)
    xd = x.cuda()
    eng = m.engine(xd.device)
    eng.bind(1, N, 1)
    eng.set_input(xd)
    torch.manual_seed(7)
    starts = eng.draw_starts(1)                       # [4, 1, 1]
    eng.geometry(starts)
    xyz = x[:, :3].permute(0, 2, 1).contiguous()
    cfg = [(1024, 0.1, 32), (256, 0.2, 32), (64, 0.4, 32), (16, 0.8, 32)]
    clouds = [xyz]
    for l, (S, r, K) in enumerate(cfg, start=1):
        fps = G.fps(clouds[-1], S, starts[l - 1, 0].long())
        assert np.array_equal(eng.read_geometry("fps", l).cpu().numpy(), fps.numpy().astype(np.int32)), ("fps", l)
        new_xyz = torch.gather(clouds[-1], 1, fps.unsqueeze(-1).expand(-1, -1, 3))
        assert torch.equal(eng.read_geometry("xyz", l).cpu(), new_xyz)
        ball = G.ball_query(r, K, clouds[-1], new_xyz)
        assert np.array_equal(eng.read_geometry("ball", l).cpu().numpy(), ball.numpy().astype(np.int32)), ("ball", l)
        clouds.append(new_xyz)
    for f in range(4):
        idx, d2, w = G.three_nn(clouds[f], clouds[f + 1])
        assert np.array_equal(eng.read_geometry("nn_idx", f).cpu().numpy(), idx.numpy().astype(np.int32)), ("nn", f)
        assert torch.equal(eng.read_geometry("nn_w", f).cpu(), w), ("nn_w", f)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_forward_and_gradient_n16384_vs_oracle(golden_dir, mode):
    from oracle import pointnet2_oracle as PO
    sd = _ckpt(golden_dir)
    m = _model("ssg", sd, mode)
    x, labels = syn.make_painted_blocks(1, 16384, 4)
    xd = x.cuda().requires_grad_(True)
    torch.manual_seed(3)
    logp, l4 = m(xd)
    xo = x.clone().requires_grad_(True)
    torch.manual_seed(3)
    ref, ref4 = PO.OracleModel(sd, "ssg")(xo)
    rt, at = (1e-3, 2e-4) if mode != "tf32" else (2e-2, 2e-2)
    np.testing.assert_allclose(logp.detach().cpu().numpy(), ref.detach().numpy(), rtol=rt, atol=at)
    assert (logp.detach().cpu().argmax(2) == ref.detach().argmax(2)).float().mean() > (0.9995 if mode != "tf32" else 0.995)
    y = labels.view(-1)
    torch.nn.functional.nll_loss(logp.reshape(-1, 13), y.cuda()).backward()
    torch.nn.functional.nll_loss(ref.reshape(-1, 13), y).backward()
    a, b = xd.grad[:, 3:6].cpu().numpy(), xo.grad[:, 3:6].numpy()
    rel = np.linalg.norm(a - b) / np.linalg.norm(b)
    nz = b != 0
    sign = (np.sign(a[nz]) == np.sign(b[nz])).mean()
    print(f"N=16384 {mode}: max |dlogp| {np.abs(logp.detach().cpu().numpy() - ref.detach().numpy()).max():.2e}, colour-gradient rel {rel:.2e}, sign {sign:.5f}")
    assert rel < (1e-2 if mode != "tf32" else 2e-1) and sign > (0.999 if mode != "tf32" else 0.93)


# ----------------------------------------------------------------------------------------------------------------
# the tcgen05 TF32 path against the REFERENCE goldens directly (the fused SA kernels, tile programs, fp1 + head chain)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_tf32_forward_and_gradient_vs_reference_golden(golden_dir, arch):
    g = dict(np.load(os.path.join(golden_dir, f"model_{arch}.npz")))
    m = _model(arch, syn.make_state_dict(arch), "tf32")
    x = syn.make_blocks(2, 2048, 0, "uniform").cuda().requires_grad_(True)
    torch.manual_seed(0)
    logp, l4 = m(x)
    dl = np.abs(logp.detach().cpu().numpy() - g["logp"]).max()
    d4 = np.abs(l4.cpu().numpy() - g["l4"]).max() / np.abs(g["l4"]).max()
    y = torch.from_numpy(g["y"]).long().cuda()
    cost = torch.nn.functional.cross_entropy(logp.reshape(-1, 13), y.view(-1), reduction="sum") / logp.size(1)
    cost.backward()
    mine, ref = x.grad.cpu().numpy()[:, 3:], g["grad"][:, 3:]
    rel = np.linalg.norm(mine - ref) / np.linalg.norm(ref)
    nz = ref != 0
    sign = (np.sign(mine[nz]) == np.sign(ref[nz])).mean()
    print(f"tf32 {arch} vs reference golden: max |dlogp| {dl:.2e}, l4 rel-to-max {d4:.2e}, grad rel {rel:.2e}, sign {sign:.5f}")
    assert dl < 5e-3 and d4 < 5e-3 and rel < 8e-2 and sign > 0.99


def test_tf32_attacks_vs_reference_golden(golden_dir):
    from pointsecguard_b200 import torchattacks
    g = dict(np.load(os.path.join(golden_dir, "attack.npz")))
    m = _model("ssg", syn.make_state_dict("ssg"), "tf32")
    x = syn.make_blocks(2, 4096, 0, "uniform").cuda()
    torch.manual_seed(0)
    adv = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=3)(x, g["nb_labels"].astype(np.float64))
    same = (adv[:, 3:6].cpu().numpy() == g["nb_adv"]).mean()
    x1 = syn.make_blocks(1, 4096, 1, "uniform").cuda()
    zl = syn.zband_labels(x1.cpu())
    mask = (zl[0] == 11).numpy()
    torch.manual_seed(0)
    adv = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=3, target=7, mask=mask)(x1, zl.numpy().astype(np.float64))
    same_t = (adv[:, 3:6].cpu().numpy() == g["tnb_adv"]).mean()
    print(f"tf32 vs reference golden: NB identical {same:.5f}, tar-NB identical {same_t:.5f}")
    # the default-init random network of these goldens has gradients ~1e-8: TF32 flips more signs there than on the
    # trained checkpoint of the at-size tests
    assert same > 0.9 and same_t > 0.95
