"""Pin the TRAINING oracle (oracle/train_oracle.py) against golden vectors produced by executing the unmodified
reference training step (oracle/make_golden_train.py: reference get_model.train() + get_loss + torch.optim.Adam,
two steps, B=2 x 1024 painted blocks).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import train_oracle as TO
from pointsecguard_b200 import synthetic as syn

CLASS_WEIGHTS = [1.0, 1.2, 0.8, 1.5, 1.0, 0.7, 1.3, 1.0, 0.9, 1.1, 1.4, 0.6, 1.0]     # oracle/make_golden_train.py


@pytest.mark.parametrize("arch", ["ssg", "msg"])
def test_train_step_matches_reference(golden_dir, arch):
    g = dict(np.load(os.path.join(golden_dir, f"train_{arch}.npz")))
    torch.set_num_threads(1)
    tr = TO.Trainer(syn.make_state_dict(arch, init="he"), arch, lr=1e-3, weight_decay=1e-4)
    w = torch.tensor(CLASS_WEIGHTS)
    torch.manual_seed(11)
    for s in range(2):
        x, y = syn.make_painted_blocks(2, 1024, 50 + s)
        loss, logp = tr.loss_and_grads(x, y, w)
        assert abs(float(loss) - float(g[f"loss{s}"])) < 2e-6 * abs(float(g[f"loss{s}"]))
        np.testing.assert_allclose(logp.numpy(), g[f"logp{s}"], rtol=1e-4, atol=1e-5)
        pn = [str(k) for k in g["pnames"]]
        assert sorted(pn) == sorted(tr.keys)
        gs = np.array([tr.sd[k].grad.double().sum().item() for k in pn])
        ga = np.array([tr.sd[k].grad.double().abs().sum().item() for k in pn])
        np.testing.assert_allclose(ga, g[f"gradabs{s}"], rtol=2e-4, atol=1e-7)
        np.testing.assert_allclose(gs, g[f"gradsum{s}"], rtol=1e-3, atol=1e-5 * max(1.0, float(ga.max())))
        if s == 0:
            for k in g:
                if k.startswith("grad0/"):
                    ref = g[k]
                    np.testing.assert_allclose(tr.sd[k[6:]].grad.numpy(), ref, rtol=1e-3, atol=1e-5 * max(1e-3, float(np.abs(ref).max())))
        tr.opt.step()
    sd = tr.state_dict()
    # the functional oracle consumed the CPU generator exactly like the reference modules (FPS starts, dropout)
    assert np.array_equal(torch.get_rng_state().numpy()[:64], g["rng_after"])
    keys = [str(k) for k in g["keys"]]
    assert sorted(keys) == sorted(sd.keys())
    mine_abs = np.array([sd[k].double().abs().sum().item() for k in keys])
    np.testing.assert_allclose(mine_abs, g["abs"], rtol=1e-5, atol=1e-6)
    for k in g:
        if k.startswith("final/"):
            np.testing.assert_allclose(sd[k[6:]].numpy(), g[k], rtol=1e-4, atol=2e-6)
