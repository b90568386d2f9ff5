"""CPU oracle of the dense kNN graph -- TEST INFRASTRUCTURE ONLY.

Restates ResGCN/gcn_lib/dense/torch_edge.py: pairwise_distance (:32-43), dense_knn_matrix (:45-59), DenseDilated.forward
(:19-29, deterministic branch).  Parity status: PINNED against tests/golden/knn.npz, which oracle/make_golden_knn.py produced
by executing the unmodified reference file (with a stub for its unused ``torch_cluster`` import, which is not installed).
``torch.topk`` leaves the order of exactly equal distances unspecified; ``knn_stable`` fixes it to ascending index, which is
what the GPU kernel produces."""
from __future__ import annotations

import torch


def pairwise_distance(x):
    """:32-43"""
    x_inner = -2 * torch.matmul(x, x.transpose(2, 1))
    x_square = torch.sum(torch.mul(x, x), dim=-1, keepdim=True)
    return x_square + x_inner + x_square.transpose(2, 1)


def dense_knn_matrix(x, k=16):
    """:45-59"""
    with torch.no_grad():
        x = x.transpose(2, 1).squeeze(-1)
        batch_size, n_points, n_dims = x.shape
        _, nn_idx = torch.topk(-pairwise_distance(x.detach()), k=k)
        center_idx = torch.arange(0, n_points).repeat(batch_size, k, 1).transpose(2, 1)
    return torch.stack((nn_idx, center_idx), dim=0)


def knn_stable(x, k):
    """The k nearest by (distance, index): x [B, N, C] -> (idx [B,N,k], d2 [B,N,k])."""
    d = pairwise_distance(x)
    ds, idx = torch.sort(d, dim=-1, stable=True)
    return idx[:, :, :k], ds[:, :, :k]


def dilated(edge_index, dilation):
    """:28 (non-stochastic)"""
    return edge_index[:, :, :, ::dilation]
