"""Train the synthetic *painted-blocks* checkpoint on the CPU with the training oracle (TEST INFRASTRUCTURE).

    python -m oracle.make_checkpoint [steps] [arch]   # from the repo root; writes tests/golden/ckpt_<arch>_painted.npz

From ``synthetic.make_state_dict(arch, init="he")``, ``steps`` (default 300) Adam steps (lr 1e-3, weight decay 1e-4,
train_semseg.py:125-132) on fresh ``synthetic.make_painted_blocks(4, 4096, 1000 + step)`` batches, uniform class
weights, ``torch.manual_seed(4321)`` before the loop (FPS starts, dropout).  The result is a network whose
predictions depend on the colours, so the targeted attacks reach a target hit-rate well above zero and the
"matched success rate" gates of the tests and of bench.py mean something.  The tensors are stored rounded to float16
(half the bytes in the repository); the checkpoint IS the rounded network -- ``synthetic.load_checkpoint`` widens it back
to float32 exactly, for the oracle and for the GPU alike.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import train_oracle as TO                      # noqa: E402
from pointsecguard_b200 import synthetic as syn            # noqa: E402

def main(steps=300, arch="ssg", B=4, N=4096):
    OUT = os.path.join(REPO, "tests", "golden", f"ckpt_{arch}_painted.npz")
    torch.set_num_threads(os.cpu_count() or 1)
    tr = TO.Trainer(syn.make_state_dict(arch, init="he"), arch, lr=1e-3, weight_decay=1e-4)
    torch.manual_seed(4321)
    t0 = time.time()
    for step in range(steps):
        x, y = syn.make_painted_blocks(B, N, 1000 + step)
        loss, logp = tr.step(x, y)
        if step % 20 == 0 or step == steps - 1:
            acc = (logp.max(2)[1] == y).float().mean().item()
            print(f"step {step:4d}  loss {float(loss):.4f}  train acc {acc:.4f}  ({time.time() - t0:.0f} s)", flush=True)
    sd = tr.state_dict()
    np.savez_compressed(OUT, **{k: (v.numpy().astype(np.float16) if v.dtype == torch.float32 else v.numpy()) for k, v in sd.items()})
    print("wrote", OUT, os.path.getsize(OUT) >> 10, "KiB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, sys.argv[2] if len(sys.argv) > 2 else "ssg")
