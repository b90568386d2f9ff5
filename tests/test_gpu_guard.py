"""Out-of-bounds hunt without compute-sanitizer (closed on this pool; tools/sanitize.sh is the recipe for a pool where it
runs): the engine workspace -- every activation, index, mask and scratch buffer of the attack path is carved from it --
is allocated between two 8 MiB guard bands of a known pattern (PSG_GUARD), whole attacks run in every mode and layout, and the
bands must come back untouched.  Shared-memory races are hunted by repetition: tests/test_gpu_model.py and
tests/test_gpu_compact.py demand bit-identical results over repeated and re-laid-out runs."""
import numpy as np
import pytest
import torch

from pointsecguard_b200 import _lib as L
from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch", ["ssg", "msg"])
@pytest.mark.parametrize("mode", ["fp32", "tf32", "x3"])
def test_workspace_guard_bands_survive_whole_attacks(monkeypatch, arch, mode):
    from pointsecguard_b200 import torchattacks
    from pointsecguard_b200.engine import MLP_FP32, MLP_TF32, MLP_TF32X3
    monkeypatch.setenv("PSG_GUARD", "8")
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13)
    m.load_state_dict(syn.load_checkpoint(arch))
    m = m.cuda().eval()
    m.set_mlp_mode({"fp32": MLP_FP32, "tf32": MLP_TF32, "x3": MLP_TF32X3}[mode])
    try:
        for B, N, kind in ((3, 4096, "uniform"), (2, 2000, "clustered"), (1, 5000, "duplicates")):
            x = syn.make_blocks(B, N, 4, kind).cuda()
            labels = syn.zband_labels(x.cpu())
            lab = labels.numpy().astype(np.float64)
            for compact in (1, 0):
                L.psg_set_option(b"sa_compact", compact)
                torch.manual_seed(0)
                torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=3)(x, lab)
                assert m.engine(x.device).guard_intact(), (B, N, kind, compact, "NB")
                torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=2, target=7, mask=labels == 11)(x, lab)
                assert m.engine(x.device).guard_intact(), (B, N, kind, compact, "tar-NB")
                # long enough for the geometry head start: the second engine's workspace is guarded too
                torchattacks.NB_attack(m, eps=0.1, alpha=0.02, iters=16)(x, lab)
                assert m.engine(x.device).guard_intact() and m._tail is not None and m._tail.guard_intact(), (B, N, kind, compact, "head start")
                if arch == "ssg":
                    torchattacks.NU_attack(m, c=0.1, kappa=0, steps=2, lr=0.01)(x, lab)
                    assert m.engine(x.device).guard_intact(), (B, N, kind, compact, "NU")
                    torchattacks.NU_attack(m, c=0.1, kappa=0, steps=2, lr=0.01, field=(0, 6))(x, lab)
                    assert m.engine(x.device).guard_intact(), (B, N, kind, compact, "NU xyz")
                xg = x.clone().requires_grad_(True)
                m(xg)[0].sum().backward()
                assert m.engine(x.device).guard_intact(), (B, N, kind, compact, "autograd")
    finally:
        L.psg_set_option(b"sa_compact", 1)
