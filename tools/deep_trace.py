#!/usr/bin/env python
"""Phase trace of the deep kernel (csrc/deep.cu): globaltimer stamps of CTA 0 for the deep launches of ONE attack step at
the bench configuration -- per phase the time CTA 0 works and the time it then waits in the grid barrier (= the slowest CTA)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointsecguard_b200 import _lib as L, synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_TF32
from pointsecguard_b200.models.pointnet2_sem_seg import get_model

m = get_model(13); m.load_state_dict(syn.make_state_dict("ssg", init="he")); m = m.cuda().eval(); m.set_mlp_mode(MLP_TF32)
x, labels, mask = bench.make_inputs(int(os.environ.get("BLOCKS", "16")), 0)
lab = labels.numpy().astype(np.float64); xd = x.cuda()
mk = lambda it: torchattacks.tar_NB_attack(m, eps=bench.EPS, alpha=bench.ALPHA, iters=it, target=bench.TARGET, mask=mask)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    assert L.psg_set_option(k.encode(), int(v)) == 0, k
mk(5)(xd, lab); torch.cuda.synchronize()
NL = 40
buf = torch.zeros(NL, 4, 512, dtype=torch.int64, device="cuda")
L.psg_debug_trace(buf.data_ptr(), NL)
mk(1)(xd, lab); torch.cuda.synchronize()
L.psg_debug_trace(None, 0)
t = buf.cpu().numpy().reshape(NL, -1)
names = {0: "gemm", 1: "interp", 2: "segsum"}
for li in range(NL):
    w = t[li]
    if w[100] != 0xDEE9:
        continue
    nph = int(w[101])
    print(f"launch {li}: {nph} phases, prologue {w[1] - w[0]} ns (incl. waiting for the predecessor), total {w[3 + 2 * (nph - 1)] - w[0]} ns")
    prev = w[1]
    for i in range(nph):
        print(f"   phase {i} {names[int(w[102 + i])]:7s}: work {w[2 + 2 * i] - prev:6d} ns, barrier {w[3 + 2 * i] - w[2 + 2 * i]:6d} ns")
        prev = w[3 + 2 * i]
