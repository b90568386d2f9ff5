"""3xTF32 mode (MLP_TF32X3): the narrow set-abstraction branches run as FUSED kernels with the error-compensated split inside
(csrc/sa_fused.cu, X3 instantiations: hi + lo weights resident, every operand written as A and A_lo, three MMAs per K step),
fp1 + head as one forward + backward tile program with the operand and its residual in tensor memory (csrc/chain_fused.cu,
tile_kernel<1,1,TS,X3>: three passes per layer), the other layers through the per-layer 3xTF32 GEMM (csrc/gemm_tc.cu).
The fused SA kernels contract every dot product in the order of the per-layer kernel with the same three products per K step,
so they must reproduce it BIT FOR BIT (x3_fused = 1 against 0); the head program orders its three passes differently (all of
A_lo W_hi, then A W_lo, then A W_hi), so it is compared at rounding level (x3_fused = 3).  The precision gates themselves
(fp32 tolerances against the reference goldens) are the `x3` cases of tests/test_gpu_model.py and tests/test_gpu_configs.py."""
import numpy as np
import pytest
import torch

from pointsecguard_b200 import _lib as L
from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _model(arch):
    from pointsecguard_b200.engine import MLP_TF32X3
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(syn.load_checkpoint(arch))
    m = m.cuda().eval()
    m.set_mlp_mode(MLP_TF32X3)
    return m


@pytest.fixture(autouse=True)
def _restore_option():
    yield
    L.psg_set_option(b"x3_fused", 3)
    L.psg_set_option(b"sa_compact", 1)


@pytest.mark.parametrize("arch", ["ssg", "msg"])
@pytest.mark.parametrize("kind", ["uniform", "clustered", "duplicates"])
def test_fused_x3_equals_per_layer_x3(arch, kind):
    m = _model(arch)
    x = syn.make_blocks(3, 4096, 2, kind).cuda()
    outs = []
    for fused, compact in ((1, 1), (0, 1), (1, 0), (3, 1)):  # (1, 0): no compacted rows -> the fused x3 kernels step aside
        L.psg_set_option(b"x3_fused", fused)
        L.psg_set_option(b"sa_compact", compact)
        xg = x.clone().requires_grad_(True)
        torch.manual_seed(0)
        logp, l4 = m(xg)
        logp[:, :, 3].sum().backward()
        outs.append((logp.detach().clone(), l4.clone(), xg.grad.clone()))
    for k in (1, 2):
        for a, b in zip(outs[0], outs[k]):
            assert torch.equal(a, b)
    assert outs[0][2].abs().sum() > 0
    # fp1 + head as one 3xTF32 program: same values up to the order of the fp32 additions
    assert torch.equal(outs[3][1], outs[0][1])                                       # l4 features do not pass through the head
    assert (outs[3][0] - outs[0][0]).abs().max().item() < 2e-4 * max(1.0, outs[0][0].abs().max().item())
    g0, g3 = outs[0][2][:, 3:], outs[3][2][:, 3:]
    assert (g3 - g0).norm().item() < 2e-3 * g0.norm().item()


def test_fused_x3_attack_equals_per_layer_x3_and_is_deterministic():
    from pointsecguard_b200 import torchattacks
    m = _model("ssg")
    x, labels = syn.make_painted_blocks(16, 4096, 1)
    xd, lab = x.cuda(), labels.numpy().astype(np.float64)
    mask = labels == 11
    res = []
    for fused in (1, 0, 1, 3, 3):
        L.psg_set_option(b"x3_fused", fused)
        torch.manual_seed(0)
        res.append(torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=20, target=7, mask=mask)(xd, lab))
    assert torch.equal(res[0], res[1]) and torch.equal(res[0], res[2])
    assert torch.equal(res[3], res[4])                    # the head program is deterministic too
    assert not torch.equal(res[0], xd)
