"""The N-rank == 1-rank check (tools/dist_check.py) as a test the driver runs: two ranks under torchrun attack their shards
of a global batch (NB and NU) and evaluate whole scenes; rank 0 re-runs everything alone and demands bit-identical
perturbed blocks, equal all-reduced counters and equal vote pools.  NCCL when two GPUs are visible; on a one-GPU box the
two ranks share the GPU and exchange over gloo (same sharding code, other transport)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_equal_one_rank():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    env = dict(os.environ, PSG_DIST_BACKEND=backend, OMP_NUM_THREADS="2")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(REPO, "tools", "dist_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=REPO, env=env)
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stderr[-3000:]
    assert "NB: shards == full batch: True; counters equal: True" in p.stdout
    assert "NU: shards == full batch: True; counters equal: True" in p.stdout
    assert "global counters equal: True; this rank's vote pools equal: True" in p.stdout
