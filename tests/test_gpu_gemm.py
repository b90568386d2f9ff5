"""The shared-MLP GEMM kernels through the C ABI (psg_mlp_forward / psg_mlp_backward) against an
fp64 torch reference: fp32 CUDA-core kernel (mode 0, rtol 1e-5) and tcgen05 TF32 kernel (mode 1,
error bounded by the TF32 input truncation: 2e-3 of the output scale)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # rows, cin (k1, k2), cout
    (256, (16, 0), 32), (1000, (80, 0), 64), (128, (272, 0), 256), (4096, (128, 0), 128),
    (300, (256, 512), 256), (512, (208, 0), 256), (640, (128, 0), 13), (128, (64, 64), 96), (384, (512, 0), 208),
]


def _mk(rows, k1, k2, cout, seed):
    g = torch.Generator().manual_seed(seed)
    a1 = torch.randn(rows, k1, generator=g).cuda()
    a2 = torch.randn(rows, k2, generator=g).cuda() if k2 else None
    w = (torch.randn(cout, k1 + k2, generator=g) / np.sqrt(k1 + k2)).contiguous()
    b = torch.randn(cout, generator=g).contiguous()
    return a1, a2, w, b


@pytest.mark.parametrize("mode", [0, 1], ids=["fp32", "tf32"])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}x{c[1][0]}+{c[1][1]}->{c[2]}" for c in CASES])
def test_mlp_forward_backward(case, mode):
    from pointsecguard_b200 import _lib as L
    from pointsecguard_b200.tlayout import TTensor
    rows, (k1, k2), cout = case
    a1, a2, w, b = _mk(rows, k1, k2, cout, rows + cout)
    st = torch.cuda.current_stream().cuda_stream
    m = L.psg_mlp_create(w.data_ptr(), b.data_ptr(), k1 + k2, cout)
    assert m
    try:
        t1 = TTensor.from_rowmajor(a1)
        t2 = TTensor.from_rowmajor(a2) if k2 else None
        out = TTensor(rows, cout, "cuda")
        L.psg_mlp_forward(m, t1.ptr, t1.wchunks, 0, t1.wchunks, t2.ptr if k2 else None, t2.wchunks if k2 else 0, 0,
                          t2.wchunks if k2 else 0, rows, out.ptr, out.wchunks, 1, mode, st)
        y = out.to_rowmajor()
        a = torch.cat([a1, a2], 1) if k2 else a1
        ref = torch.relu(a.double() @ w.cuda().double().t() + b.cuda().double())
        tol = 1e-5 if mode == 0 else 2e-3
        scale = ref.abs().max().item()
        err = (y.double() - ref).abs().max().item()
        assert err <= tol * scale, (err, scale)
        # dgrad with the ReLU mask of a lower layer
        dy = TTensor.from_rowmajor(torch.randn(rows, cout, device="cuda"))
        dyr = dy.to_rowmajor()
        kp = t1.cpad + (t2.cpad if k2 else 0)
        maskt = TTensor.from_rowmajor(torch.randn(rows, kp, device="cuda"))
        dx = TTensor(rows, kp, "cuda")
        L.psg_mlp_backward(m, dy.ptr, dy.wchunks, rows, dx.ptr, dx.wchunks, maskt.ptr, maskt.wchunks, mode, st)
        got = dx.to_rowmajor()[:, : k1 + k2] if not k2 else dx.to_rowmajor()
        wfull = w.cuda().double()
        refdx = dyr.double() @ wfull
        if k2:   # padded two-source layout: [k1 | pad | k2 | pad]
            full = torch.zeros(rows, kp, dtype=torch.float64, device="cuda")
            full[:, :k1] = refdx[:, :k1]
            full[:, t1.cpad: t1.cpad + k2] = refdx[:, k1:]
            refdx = full
            mk = maskt.to_rowmajor()
        else:
            mk = maskt.to_rowmajor()[:, : k1 + k2]
        refdx = refdx * (mk > 0)
        scale = refdx.abs().max().item()
        err = (got.double() - refdx).abs().max().item()
        assert err <= tol * scale, (err, scale)
    finally:
        L.psg_mlp_destroy(m)


def test_model_tf32_mode_close_to_fp32():
    """Whole network with tcgen05 TF32 MLPs: same geometry (bit-exact indices), log-probabilities
    within 5e-3 absolute of the fp32 path, identical predictions on > 99.5 % of the points."""
    from pointsecguard_b200 import synthetic as syn
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    m = get_model(13)
    m.load_state_dict(syn.make_state_dict("ssg"))
    m = m.cuda().eval()
    x = syn.make_blocks(2, 4096, 0).cuda()
    torch.manual_seed(0)
    ref, _ = m(x)
    m.set_mlp_mode(MLP_TF32)
    torch.manual_seed(0)
    out, _ = m(x)
    m.set_mlp_mode(MLP_FP32)
    err = (out - ref).abs().max().item()
    agree = (out.argmax(2) == ref.argmax(2)).float().mean().item()
    print("tf32 vs fp32: max |dlogp|", err, "argmax agreement", agree)
    assert err < 5e-3 and agree > 0.995


@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_model_tf32_gradient_close_to_fp32(arch):
    """tcgen05 mode (fused set-abstraction kernels, TF32 MLPs) against the fp32 parity mode: same
    indices, log-probabilities within 5e-3, input gradient within 8e-2 relative (Frobenius) with
    > 99 % sign agreement -- the stated, looser TF32 tolerance (2^-11 weight noise
    alone gives 4e-2 / 99.7 %)."""
    from pointsecguard_b200 import synthetic as syn
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(syn.make_state_dict(arch))
    m = m.cuda().eval()
    x0 = syn.make_blocks(2, 4096, 0).cuda()
    res = {}
    for mode in (MLP_FP32, MLP_TF32):
        m.set_mlp_mode(mode)
        x = x0.clone().requires_grad_(True)
        torch.manual_seed(0)
        logp, l4 = m(x)
        y = (logp.detach().argmax(2) + 1) % 13 if mode == MLP_FP32 else res["y"]
        res["y"] = y
        cost = torch.nn.functional.cross_entropy(logp.reshape(-1, 13), y.view(-1), reduction="sum") / logp.size(1)
        cost.backward()
        res[mode] = (logp.detach(), l4, x.grad[:, 3:].clone())
    m.set_mlp_mode(MLP_FP32)
    (lp0, l40, g0), (lp1, l41, g1) = res[MLP_FP32], res[MLP_TF32]
    err = (lp1 - lp0).abs().max().item()
    l4err = ((l41 - l40).abs().max() / l40.abs().max()).item()
    rel = ((g1 - g0).norm() / g0.norm()).item()
    nz = g0 != 0
    sign = (torch.sign(g1[nz]) == torch.sign(g0[nz])).float().mean().item()
    zero_same = ((g1 == 0) == (g0 == 0)).float().mean().item()
    print(f"{arch}: tf32 vs fp32 |dlogp| {err:.2e} l4 rel {l4err:.2e} grad rel {rel:.2e} sign {sign:.5f} zero-pattern {zero_same:.5f}")
    assert err < 5e-3 and l4err < 5e-3 and rel < 8e-2 and sign > 0.99 and zero_same > 0.999


def test_attack_metrics_tf32_vs_fp32_within_half_a_point():
    """north_star gate: accuracy / mIoU / attack success after the attack agree within +-0.5 pt between
    the tcgen05 TF32 path and the fp32 parity path (same seeds, bit-identical geometry)."""
    from pointsecguard_b200 import metrics as MT, synthetic as syn, torchattacks
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    m = get_model(13)
    m.load_state_dict(syn.make_state_dict("ssg", init="he"))  # input-sensitive random network (synthetic.py)
    m = m.cuda().eval()
    x = syn.make_blocks(16, 4096, 3).cuda()                   # 65536 points: the gate is on aggregate metrics
    torch.manual_seed(5)
    lab = m(x)[0].argmax(2)                                  # clean predictions: initial accuracy 100 %
    zl = syn.zband_labels(x.cpu())
    mask = zl == 11
    res = {}
    for mode in (MLP_FP32, MLP_TF32):
        m.set_mlp_mode(mode)
        torch.manual_seed(0)
        adv = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=10)(x, lab.cpu().numpy().astype(np.float64))
        torch.manual_seed(1)
        s1 = MT.summarize(MT.attack_counters(m(adv)[0], lab))
        torch.manual_seed(0)
        advt = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=10, target=7, mask=mask)(x, zl.numpy().astype(np.float64))
        torch.manual_seed(1)
        s2 = MT.summarize(MT.attack_counters(m(advt)[0], zl.cuda(), mask.cuda(), 7))
        res[mode] = (s1, s2, adv)
    m.set_mlp_mode(MLP_FP32)
    (a1, a2, adv0), (b1, b2, adv1) = res[MLP_FP32], res[MLP_TF32]
    print("NB  fp32", a1, "\n    tf32", b1, "\ntar fp32", a2, "\n    tf32", b2)
    same = (adv0 == adv1).float().mean().item()
    print("identical perturbed elements fp32 vs tf32:", same)
    assert abs(a1["acc"] - b1["acc"]) < 0.005 and abs(a1["miou"] - b1["miou"]) < 0.005
    assert abs(a2["acc"] - b2["acc"]) < 0.005 and abs(a2["miou"] - b2["miou"]) < 0.005
    assert abs((a2["target_acc"] or 0) - (b2["target_acc"] or 0)) < 0.005
    assert a1["acc"] < 0.9                                   # the attack did something
