#!/usr/bin/env python
"""Fixed cost of one attack call vs its per-step cost at the bench configuration: times attacks of several lengths
(CUDA events, L2 flushed, best of 3) and fits time = a + b * steps; also reports the host time until the call returns
(everything enqueued).  a > 0 is time the GPU spends waiting for the host or in non-amortised geometry."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pointsecguard_b200 import synthetic as syn, torchattacks
from pointsecguard_b200.engine import MLP_TF32
from pointsecguard_b200.models.pointnet2_sem_seg import get_model

m = get_model(13); m.load_state_dict(syn.make_state_dict("ssg", init="he")); m = m.cuda().eval(); m.set_mlp_mode(MLP_TF32)
x, labels, mask = bench.make_inputs(16, 0)
lab = labels.numpy().astype(np.float64)
xd = x.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for K in (10, 25, 50, 64, 100):
    atk = torchattacks.tar_NB_attack(m, eps=0.5, alpha=0.1, iters=K, target=7, mask=mask)
    atk(xd, lab); torch.cuda.synchronize()
    best, host = 1e9, 1e9
    for _ in range(3):
        torch.manual_seed(0); flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); atk(xd, lab); e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); host = min(host, (t1 - t0) * 1e3)
    res[K] = (best, host)
    print(f"K={K}: gpu {best:.3f} ms ({best / K:.4f} ms/step), host enqueue {host:.3f} ms")
ks = np.array(sorted(res)); ts = np.array([res[k][0] for k in ks])
b, a = np.polyfit(ks, ts, 1)
print(f"fit: fixed {a:.3f} ms per attack + {b:.4f} ms per step  ->  {1e3 / b:.0f} steps/s asymptotic")
