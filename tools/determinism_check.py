import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from pointsecguard_b200 import synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_TF32
from pointsecguard_b200.models.pointnet2_sem_seg import get_model
m = get_model(13); m.load_state_dict(syn.make_state_dict("ssg", init="he")); m = m.cuda().eval(); m.set_mlp_mode(MLP_TF32)
x, labels, mask = bench.make_inputs(16, 0)
lab = labels.numpy().astype(np.float64)
xd = x.cuda()
outs = []
for it in (1, 2, 5) + (50,) * 12:
    torch.manual_seed(0)
    a = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=it, target=7, mask=mask)(xd, lab)
    outs.append((it, a.cpu()))
for i in range(3, len(outs)):
    d = (outs[i][1] != outs[3][1])
    print("iters", outs[i][0], "differs from first 50-run:", int(d.sum()), "of", d.numel())
# one step twice
torch.manual_seed(0); a1 = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=1, target=7, mask=mask)(xd, lab).cpu()
print("1-step equal:", torch.equal(a1, outs[0][1]))
torch.manual_seed(0); a2 = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=2, target=7, mask=mask)(xd, lab).cpu()
print("2-step equal:", torch.equal(a2, outs[1][1]))
# forward logits determinism
torch.manual_seed(1); l1 = m(xd)[0].cpu(); torch.manual_seed(1); l2 = m(xd)[0].cpu(); print("forward equal:", torch.equal(l1, l2))
x2 = xd.clone().requires_grad_(True)
torch.manual_seed(1); g1 = torch.autograd.grad(m(x2)[0][..., 3].sum(), x2)[0].cpu()
torch.manual_seed(1); g2 = torch.autograd.grad(m(x2)[0][..., 3].sum(), x2)[0].cpu()
print("grad equal:", torch.equal(g1, g2), float((g1 - g2).abs().max()))
