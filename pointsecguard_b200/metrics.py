"""Script-level metrics of the attack scripts as one GPU histogram + (multi-GPU) one all-reduce.

Reference: PointNet/NB_nontarget_test_semseg.py:187-212 (acc, adv_acc, per-class seen / correct /
union -> block mIoU, 13 x 5 numpy passes over D2H copies per batch) and NB_target_test_semseg.py:187-190
(target_acc over the masked points).  Here one kernel builds the 13x13 confusion matrix plus four
scalars into an int64 buffer; that buffer is the only thing ranks exchange (NCCL all-reduce, sum).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def attack_counters(logp: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor | None = None, target: int = -1,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """logp [B,N,C] float32 (CUDA), labels [B,N] -> int64 [C*C + 4] counters (accumulated into ``out``)."""
    if not logp.is_cuda:
        raise RuntimeError("attack_counters needs CUDA tensors; there is no CPU fallback")
    Bn = logp.shape[0] * logp.shape[1]
    ncls = logp.shape[2]
    logp = logp.detach().contiguous()
    lab = labels.to(device=logp.device, dtype=torch.int32).contiguous()
    m = None
    if mask is not None:
        m = mask.to(device=logp.device)
        if m.dim() == 1:
            m = m.view(1, -1).expand(logp.shape[0], -1)
        m = m.to(torch.uint8).contiguous()
    if out is None:
        out = torch.zeros(ncls * ncls + 4, dtype=torch.int64, device=logp.device)
    L.psg_confusion_matrix(logp.data_ptr(), lab.data_ptr(), m.data_ptr() if m is not None else None, int(target), Bn, ncls,
                           out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return out


def summarize(counters, ncls: int = 13) -> dict:
    """acc / per-class IoU / mIoU / target hit-rate exactly as the scripts derive them from their
    seen / correct / union counters (NB_nontarget_test_semseg.py:205-212)."""
    c = np.asarray(counters.cpu() if torch.is_tensor(counters) else counters, dtype=np.int64)
    conf = c[: ncls * ncls].reshape(ncls, ncls)
    rows, correct, nmask, hit = (int(v) for v in c[ncls * ncls:])
    seen = conf.sum(1)
    tp = np.diag(conf)
    union = seen + conf.sum(0) - tp
    iou = tp / (union.astype(np.float64) + 1e-6)
    miou = float(np.mean(iou[seen != 0])) if (seen != 0).any() else 0.0
    return {"points": rows, "acc": correct / max(rows, 1), "miou": miou,
            "target_acc": (hit / nmask) if nmask else None, "masked_points": nmask}
