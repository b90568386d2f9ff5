#!/usr/bin/env python
"""Run the BASELINE.json configurations natively on one B200 and print one JSON line per config
(steps/s, block-steps/s, attack metrics).  configs[0] is the reference's CPU case (bench.py
--impl reference); configs[1] is bench.py's headline.  This tool covers the rest at their stated sizes:

  [1] SSG tar-NB 50 iters B=16          (same as bench.py, for cross-checking)
  [2] SSG NU over coordinates + RGB, B=32, 100 iters
  [2c] SSG NU colour only (the reference's own field), B=32, 100 iters
  [3] MSG NB 10 iters, B=64 on one GPU (the multi-GPU run shards this batch: bench.py --gpus N)
  [4] SSG NB, B=8, N = 4096 / 16384 / 65536 (tools/primitives_bench.py adds the primitive roofline)

    python tools/run_configs.py [--out profiles/r1_configs.json] [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def timed_attack(atk, x, lab, reps=2):
    atk(x, lab)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.manual_seed(0)
        e0.record()
        adv = atk(x, lab)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return adv, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from pointsecguard_b200 import metrics as MT, synthetic as syn, torchattacks
    from pointsecguard_b200.engine import MLP_TF32
    from pointsecguard_b200.models import pointnet2_sem_seg as S, pointnet2_sem_seg_msg as M

    def model(arch):
        m = (S if arch == "ssg" else M).get_model(13)
        m.load_state_dict(syn.make_state_dict(arch, init="he"))      # input-sensitive random network (synthetic.py)
        m = m.cuda().eval()
        m.set_mlp_mode(MLP_TF32)
        return m

    out = []

    def report(name, B, N, iters, ms, adv, m, labels, mask=None, target=-1, executed=None):
        steps = executed if executed is not None else iters
        torch.manual_seed(1)
        c = MT.attack_counters(m(adv)[0], labels.cuda(), mask.cuda() if mask is not None else None, target)
        rec = {"config": name, "blocks": B, "points": N, "iters": iters, "steps_executed": steps, "ms": ms,
               "steps_per_s": steps / (ms / 1e3), "block_steps_per_s": B * steps / (ms / 1e3), "mlp": "tf32",
               "metrics": MT.summarize(c.cpu())}
        out.append(rec)
        print(json.dumps(rec), flush=True)

    q = 5 if args.quick else 1
    ssg = model("ssg")
    # [1]
    x = syn.make_blocks(16, 4096, 0).cuda(); lab = syn.zband_labels(x.cpu()); mask = lab == 11
    atk = torchattacks.tar_NB_attack(ssg, eps=0.5, alpha=0.1, iters=50 // q, target=7, mask=mask)
    adv, ms = timed_attack(atk, x, lab.numpy().astype(np.float64))
    report("[1] SSG tar-NB", 16, 4096, 50 // q, ms, adv, ssg, lab, mask, 7)
    # [2c] / [2]: untargeted labels = clean prediction (no exit at step 0)
    x = syn.make_blocks(32, 4096, 0).cuda()
    torch.manual_seed(5)
    lab = ssg(x)[0].argmax(2).cpu()
    for name, field in (("[2c] SSG NU colour", None), ("[2] SSG NU coordinates + colour", (0, 6))):
        atk = torchattacks.NU_attack(ssg, c=0.1, kappa=0, steps=100 // q, lr=0.01, field=field)
        adv, ms = timed_attack(atk, x, lab.numpy().astype(np.float64), reps=1)
        report(name, 32, 4096, 100 // q, ms, adv, ssg, lab)
    # [3]
    msg = model("msg")
    x = syn.make_blocks(64, 4096, 0).cuda()
    torch.manual_seed(5)
    lab = msg(x)[0].argmax(2).cpu()
    atk = torchattacks.NB_attack(msg, eps=0.1, alpha=0.05, iters=10)
    adv, ms = timed_attack(atk, x, lab.numpy().astype(np.float64))
    report("[3] MSG NB B=64", 64, 4096, 10, ms, adv, msg, lab)
    # [4]
    for N in (4096, 16384, 65536):
        x = syn.make_blocks(8, N, 0).cuda()
        torch.manual_seed(5)
        lab = ssg(x)[0].argmax(2).cpu()
        atk = torchattacks.NB_attack(ssg, eps=0.1, alpha=0.05, iters=10)
        adv, ms = timed_attack(atk, x, lab.numpy().astype(np.float64))
        report(f"[4] SSG NB N={N}", 8, N, 10, ms, adv, ssg, lab)
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
