#!/usr/bin/env python
"""Per-kernel-family device time (the library's CUDA-event profiler) of the non-headline BASELINE configurations:
MSG NB B=64, NU colour / coordinates B=32, SSG NB at N = 16384 / 65536."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointsecguard_b200 import _lib as L, synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_TF32


def families(run, steps):
    run(); torch.cuda.synchronize()
    L.psg_prof_enable(1)
    torch.manual_seed(0); run(); torch.cuda.synchronize()
    n = L.psg_prof_ncat(); ms = (C.c_double * n)(); cnt = (C.c_int64 * n)()
    L.psg_prof_collect(ms, cnt); L.psg_prof_enable(0)
    return {L.psg_prof_name(i).decode(): round(ms[i] / steps, 4) for i in range(n) if cnt[i]}


def model(arch):
    from pointsecguard_b200.models import pointnet2_sem_seg, pointnet2_sem_seg_msg
    m = (pointnet2_sem_seg if arch == "ssg" else pointnet2_sem_seg_msg).get_model(13)
    m.load_state_dict(syn.make_state_dict(arch, init="he"))
    m = m.cuda().eval(); m.set_mlp_mode(MLP_TF32)
    return m


out = {}
msg = model("msg")
x = syn.make_blocks(64, 4096, 0).cuda(); torch.manual_seed(5); lab = msg(x)[0].argmax(2).cpu().numpy().astype(np.float64)
out["MSG NB B=64 x10"] = families(lambda: torchattacks.NB_attack(msg, eps=0.1, alpha=0.05, iters=10)(x, lab), 10)
ssg = model("ssg")
x = syn.make_blocks(32, 4096, 0).cuda(); torch.manual_seed(5); lab = ssg(x)[0].argmax(2).cpu().numpy().astype(np.float64)
out["NU colour B=32 x100"] = families(lambda: torchattacks.NU_attack(ssg, c=0.1, kappa=0, steps=100, lr=0.01)(x, lab), 100)
out["NU coords B=32 x40"] = families(lambda: torchattacks.NU_attack(ssg, c=0.1, kappa=0, steps=40, lr=0.01, field=(0, 6))(x, lab), 40)
for N in (16384, 65536):
    x = syn.make_blocks(8, N, 0).cuda(); torch.manual_seed(5); lab = ssg(x)[0].argmax(2).cpu().numpy().astype(np.float64)
    out[f"NB N={N} B=8 x10"] = families(lambda: torchattacks.NB_attack(ssg, eps=0.1, alpha=0.05, iters=10)(x, lab), 10)
for k, v in out.items():
    print(k, "total", round(sum(v.values()), 3), json.dumps(v))
