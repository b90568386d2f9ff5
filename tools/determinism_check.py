"""Run-to-run determinism of the tcgen05 path: repeats the bench attack and the attacks of every BASELINE configuration
and prints how many output elements differ between repeats (must be 0 everywhere; there are no float atomics on the path)."""
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

# ---- the other BASELINE configurations: repeated runs must be bit-identical ----
def repeat(name, make, x, lab, runs=4):
    first, diff = None, []
    for _ in range(runs):
        torch.manual_seed(0)
        a = make()(x, lab).detach().cpu()
        if first is None:
            first = a
        else:
            diff.append(int((a != first).sum()))
    print(f"{name}: differing elements per repeat {diff}")

x = syn.make_blocks(32, 4096, 0).cuda()
torch.manual_seed(5); labc = m(x)[0].argmax(2).cpu().numpy().astype(np.float64)
repeat("NU colour B=32 x 100", lambda: torchattacks.NU_attack(m, c=0.1, kappa=0, steps=100, lr=0.01), x, labc)
repeat("NU coords+colour B=32 x 40", lambda: torchattacks.NU_attack(m, c=0.1, kappa=0, steps=40, lr=0.01, field=(0, 6)), x, labc, runs=3)
for N in (16384, 65536):
    x = syn.make_blocks(8, N, 0).cuda()
    torch.manual_seed(5); labn = m(x)[0].argmax(2).cpu().numpy().astype(np.float64)
    repeat(f"NB N={N} B=8 x 10", lambda: torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=10), x, labn, runs=3)
from pointsecguard_b200.models.pointnet2_sem_seg_msg import get_model as get_msg
mm = get_msg(13); mm.load_state_dict(syn.make_state_dict("msg", init="he")); mm = mm.cuda().eval(); mm.set_mlp_mode(MLP_TF32)
x = syn.make_blocks(64, 4096, 0).cuda()
torch.manual_seed(5); labm = mm(x)[0].argmax(2).cpu().numpy().astype(np.float64)
repeat("MSG NB B=64 x 10", lambda: torchattacks.NB_attack(mm, eps=0.1, alpha=0.05, iters=10), x, labm, runs=4)
x1 = syn.make_blocks(1, 4096, 3).cuda(); l1 = syn.zband_labels(x1.cpu()); mk1 = (l1 == 11)[0].numpy()
repeat("tar-NU B=1 x 60", lambda: torchattacks.tar_NU_attack(m, c=0.1, kappa=0, steps=60, lr=0.01, target=7, mask=mk1), x1, l1.numpy().astype(np.float64), runs=3)
