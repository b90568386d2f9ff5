"""tests/golden/knn.npz: outputs of the reference's dense kNN graph (ResGCN/gcn_lib/dense/torch_edge.py) on seeded inputs,
made by executing the unmodified reference file.  Its module-level ``from torch_cluster import knn_graph`` (used only by the
sparse DilatedKnnGraph, not by the dense path) is satisfied with an empty stub because torch_cluster is not installed.
Build container only:      python -m oracle.make_golden_knn"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pointsecguard_b200 import synthetic as syn          # noqa: E402

REF_FILE = "/root/reference/ResGCN/gcn_lib/dense/torch_edge.py"


def load_reference():
    stub = types.ModuleType("torch_cluster")
    stub.knn_graph = None
    sys.modules.setdefault("torch_cluster", stub)
    spec = importlib.util.spec_from_file_location("ref_torch_edge", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def inputs():
    """(name, x [B, C, N, 1], k, dilation)"""
    x3 = syn.make_blocks(2, 2048, 5, "uniform")[:, :3].contiguous().unsqueeze(-1)           # xyz, as ResGCN's first layer
    xg = syn.make_blocks(2, 1024, 6, "grid")[:, :3].contiguous().unsqueeze(-1)              # many exact ties
    g = torch.Generator().manual_seed(9)
    x9 = syn.make_blocks(2, 1024, 7, "uniform").contiguous().unsqueeze(-1)                  # all 9 input channels
    x64 = torch.randn(1, 64, 1024, 1, generator=g)                                          # feature space of the deeper blocks
    return [("xyz", x3, 16, 1), ("grid", xg, 16, 1), ("c9", x9, 32, 2), ("c64", x64, 16, 1)]


def main():
    R = load_reference()
    out = {}
    torch.set_num_threads(1)
    for name, x, k, dil in inputs():
        e = R.DenseDilatedKnnGraph(k // dil, dil)(x)
        full = R.dense_knn_matrix(x, k)
        xt = x.transpose(2, 1).squeeze(-1)
        d = R.pairwise_distance(xt)
        out[name + "_edge"] = e.numpy().astype(np.int16)
        out[name + "_knn"] = full[0].numpy().astype(np.int16)
        out[name + "_d2"] = torch.gather(d, 2, full[0]).numpy()
        out[name + "_drow"] = d[:, :2].numpy()
    path = os.path.join(REPO, "tests", "golden", "knn.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
