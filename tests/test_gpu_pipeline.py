"""Geometry head start of the norm-bounded attacks (torchattacks/attacks/nontarget.py::_head_start): the first forwards run on
the primary engine while a second engine computes the geometry of the rest on a side stream.  Same FPS draws, same kernels:
the perturbed blocks must be BIT-IDENTICAL to the single-pass loop, for every head length, with several geometry chunks
(iters x B > 1024 problems), in both MLP modes, and the CPU generator must end in the same state."""
import numpy as np
import pytest
import torch

from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _model(arch, mode):
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(syn.load_checkpoint(arch))
    m = m.cuda().eval()
    m.set_mlp_mode(mode)
    return m


@pytest.mark.parametrize("arch,mode", [("ssg", 1), ("ssg", 0), ("msg", 1)])
def test_head_start_bit_identical(arch, mode):
    from pointsecguard_b200 import torchattacks
    m = _model(arch, mode)
    x, labels = syn.make_painted_blocks(6, 4096, 4)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mask = labels == 11
    res = []
    for head in (0, 1, 2, 3):
        m.geometry_head = head
        torch.manual_seed(0)
        a = torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=12)(xd, lab)
        b = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=13, target=7, mask=mask)(xd, lab)
        res.append((a, b, torch.get_rng_state().clone()))
    for r in res[1:]:
        assert torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1]) and torch.equal(r[2], res[0][2])
    assert not torch.equal(res[0][0], xd)


def test_head_start_with_several_geometry_chunks():
    """40 blocks x 60 iterations = 2400 FPS problems: three geometry chunks of 25 forwards, each with its own head start."""
    from pointsecguard_b200 import torchattacks
    m = _model("ssg", 1)
    x, labels = syn.make_painted_blocks(40, 1024, 5)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    res = []
    for head in (0, 2):
        m.geometry_head = head
        torch.manual_seed(0)
        res.append(torchattacks.NB_attack(m, eps=0.1, alpha=0.02, iters=60)(xd, lab))
    assert torch.equal(res[0], res[1])


def test_repeated_attacks_with_head_start_deterministic():
    from pointsecguard_b200 import torchattacks
    m = _model("ssg", 1)
    x, labels = syn.make_painted_blocks(16, 4096, 6)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mask = labels == 11
    first = None
    for _ in range(8):
        torch.manual_seed(0)
        a = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=20, target=7, mask=mask)(xd, lab)
        first = a if first is None else first
        assert torch.equal(a, first)
