#!/usr/bin/env python
"""Cross-check of the three MLP modes on odd problem sizes (rows not a multiple of 128, few blocks, clustered / duplicate clouds):
forward log-probabilities and input gradient of the 3xTF32 and TF32 modes against the CUDA-core fp32 mode, SSG and MSG, and a
short NB attack through the geometry head start.  Not a parity test (tests/ has those) -- a net for edge cases of the fused paths."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointsecguard_b200 import synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_FP32, MLP_TF32, MLP_TF32X3

worst = {"x3_logp": 0.0, "x3_grad": 0.0, "tf32_logp": 0.0, "tf32_grad": 0.0}
rng = np.random.default_rng(0)
for arch in ("ssg", "msg"):
    if arch == "ssg":
        from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    else:
        from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model
    m = get_model(13); m.load_state_dict(syn.load_checkpoint(arch)); m = m.cuda().eval()
    for case in range(int(os.environ.get("FUZZ_CASES", "10"))):
        B = int(rng.choice([1, 2, 3, 5, 7])); N = int(rng.choice([1024, 1100, 1500, 2000, 3001, 4096, 5000, 7777]))
        kind = str(rng.choice(["uniform", "clustered", "duplicates", "surface"]))
        x = syn.make_blocks(B, N, int(rng.integers(0, 100)), kind).cuda()
        res = {}
        for name, mode in (("fp32", MLP_FP32), ("x3", MLP_TF32X3), ("tf32", MLP_TF32)):
            m.set_mlp_mode(mode)
            xg = x.clone().requires_grad_(True)
            torch.manual_seed(case)
            logp, _ = m(xg)
            logp[:, :, 3].sum().backward()
            torch.manual_seed(case)
            lab = logp.detach().argmax(2).cpu().numpy().astype(np.float64)
            adv = torchattacks.NB_attack(m, eps=0.1, alpha=0.02, iters=17)(x, lab)
            assert torch.isfinite(adv).all() and torch.isfinite(logp).all() and torch.isfinite(xg.grad).all(), (arch, B, N, kind, name)
            res[name] = (logp.detach(), xg.grad[:, 3:].clone())
        for name in ("x3", "tf32"):
            dl = (res[name][0] - res["fp32"][0]).abs().max().item()
            dg = ((res[name][1] - res["fp32"][1]).norm() / res["fp32"][1].norm()).item()
            worst[name + "_logp"] = max(worst[name + "_logp"], dl); worst[name + "_grad"] = max(worst[name + "_grad"], dg)
            print(f"{arch} B={B} N={N} {kind:10s} {name}: max |dlogp| {dl:.2e}, grad rel {dg:.2e}")
print("worst:", worst)
assert worst["x3_logp"] < 2e-3 and worst["x3_grad"] < 5e-2, "3xTF32 too far from fp32 somewhere"
