"""cProfile of the HOST side of one 50-step attack at the bench configuration (what runs before / between the launches)."""
import os, sys, time, cProfile, pstats, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench
from pointsecguard_b200 import synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_TF32
from pointsecguard_b200.models.pointnet2_sem_seg import get_model
m = get_model(13); m.load_state_dict(syn.make_state_dict("ssg", init="he")); m = m.cuda().eval(); m.set_mlp_mode(MLP_TF32)
x, labels, mask = bench.make_inputs(16, 0)
lab = labels.numpy().astype(np.float64); xd = x.cuda()
atk = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=50, target=7, mask=mask)
for _ in range(3): atk(xd, lab)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); atk(xd, lab); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
