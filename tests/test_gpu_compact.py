"""Compacted neighbourhood rows (csrc/compact.cu): the fused set-abstraction kernels run on the real ball-query hits only.
The result must be BIT-IDENTICAL to the padded [S][nsample] layout -- forward log-probabilities, input gradient and whole
attacks -- on uniform blocks (few hits), clustered blocks (every ball overflows: nothing to compact) and duplicates."""
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
    L.psg_set_option(b"sa_compact", 1)


@pytest.mark.parametrize("arch", ["ssg", "msg"])
@pytest.mark.parametrize("kind", ["uniform", "clustered", "duplicates", "surface"])
def test_forward_and_gradient_bit_identical(arch, kind):
    m = _model(arch)
    x = syn.make_blocks(3, 4096, 2, kind).cuda()
    outs = []
    for on in (1, 0):
        L.psg_set_option(b"sa_compact", on)
        xg = x.clone().requires_grad_(True)
        torch.manual_seed(0)
        logp, l4 = m(xg)
        logp[:, :, 3].sum().backward()
        outs.append((logp.detach().clone(), l4.clone(), xg.grad.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    assert outs[0][2].abs().sum() > 0


@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_attacks_bit_identical(arch):
    from pointsecguard_b200 import torchattacks
    m = _model(arch)
    x, labels = syn.make_painted_blocks(5, 4096, 1)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mask = labels == 11
    res = []
    for on in (1, 0):
        L.psg_set_option(b"sa_compact", on)
        torch.manual_seed(0)
        a = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=6)(xd, lab)
        torch.manual_seed(0)
        b = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=6, target=7, mask=mask)(xd, lab)
        c = None
        if arch == "ssg":
            torch.manual_seed(0)
            c = torchattacks.NU_attack(m, c=0.1, kappa=0, steps=4, lr=0.01)(xd, lab)
        res.append((a, b, c))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    if arch == "ssg":
        assert torch.equal(res[0][2], res[1][2])
    assert not torch.equal(res[0][0], xd)
