"""The deep levels as one persistent kernel per direction (csrc/deep.cu: FP4 / FP3 GEMMs, interpolations and the segmented
sums around them as phases separated by a grid barrier).  Every phase repeats the stand-alone kernel's arithmetic in the
same order, so the result must be BIT-IDENTICAL to the launch-per-layer path -- forward log-probabilities, input gradient
and whole attacks -- at batch sizes on both sides of the tile-program threshold (B = 3: every FP level below fp1 runs per
layer; B = 16: fp4 / fp3; B = 40: fp4 only)."""
import numpy as np
import pytest
import torch

from pointsecguard_b200 import _lib as L
from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _model(arch):
    from pointsecguard_b200.engine import MLP_TF32
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(syn.load_checkpoint(arch))
    m = m.cuda().eval()
    m.set_mlp_mode(MLP_TF32)
    return m


@pytest.fixture(autouse=True)
def _restore_option():
    yield
    L.psg_set_option(b"deep", 0)


@pytest.mark.parametrize("arch,B", [("ssg", 3), ("ssg", 16), ("ssg", 40), ("msg", 3), ("msg", 16)])
def test_forward_and_gradient_bit_identical(arch, B):
    m = _model(arch)
    x = syn.make_blocks(B, 4096, 2, "uniform").cuda()
    outs = []
    for mode in (3, 1, 0):
        L.psg_set_option(b"deep", mode)
        xg = x.clone().requires_grad_(True)
        torch.manual_seed(0)
        logp, l4 = m(xg)
        logp[:, :, 3].sum().backward()
        outs.append((logp.detach().clone(), l4.clone(), xg.grad.clone()))
    for k in (0, 1):
        for a, b in zip(outs[k], outs[2]):
            assert torch.equal(a, b)
    assert outs[0][2].abs().sum() > 0


@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_attacks_bit_identical(arch):
    from pointsecguard_b200 import torchattacks
    m = _model(arch)
    x, labels = syn.make_painted_blocks(16, 4096, 1)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mask = labels == 11
    res = []
    for mode in (3, 0):
        L.psg_set_option(b"deep", mode)
        torch.manual_seed(0)
        a = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=6)(xd, lab)
        torch.manual_seed(0)
        b = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=12, target=7, mask=mask)(xd, lab)
        res.append((a, b))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert not torch.equal(res[0][0], xd)


def test_repeated_attacks_deterministic():
    """The grid-barrier counter runs on from launch to launch: many back-to-back attacks stay identical."""
    from pointsecguard_b200 import torchattacks
    m = _model("ssg")
    x, labels = syn.make_painted_blocks(16, 4096, 3)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    first = None
    for _ in range(6):
        torch.manual_seed(0)
        a = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=10)(xd, lab)
        if first is None:
            first = a
        assert torch.equal(a, first)
