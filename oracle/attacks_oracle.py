"""CPU oracle: functional restatement of the reference's four colour attacks.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
PointNet/attacks/torchattacks/attacks/nontarget.py (NB_attack :10-42, NU_attack :44-135) and
target.py (tar_NB_attack :7-45, tar_NU_attack :52-175) including the behaviours SURVEY.md
Appendix A lists (un-projected return value Q1, `_targeted` = +1 Q3, all-channel clamp Q4,
acc / 4096 Q9 ...).  ``model`` is any callable x[B,9,N] -> (logp[B,N,13], _), e.g.
oracle.pointnet2_oracle.OracleModel or the reference's own get_model on CPU.

Parity status: PINNED against tests/golden/attack_*.npz (made by executing the unmodified
reference, oracle/make_golden.py).

Generalisations the reference cannot express are opt-in arguments and keep
the reference arithmetic otherwise:
  * ``field``: channel slice that is perturbed (reference: 3:6 only);
  * tar_nb_attack with a [B,N] mask and B > 1 (reference: B == 1 and a [N] mask).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

TARGETED_SIGN = 1   # attack.py:13; the scripts never call set_attack_mode (Q3)


def _labels(labels):
    return torch.as_tensor(np.asarray(labels), dtype=torch.int64) if not torch.is_tensor(labels) \
        else labels.to(torch.int64)


def nb_attack(model, images, labels, eps=0.3, alpha=2 / 255, iters=40, field=slice(3, 6), snapshots=None):
    """nontarget.py:18-42: L-inf sign ascent on sum-CE / N; returns the un-projected last step.
    ``snapshots``: optional list that receives the projected colours entering every iteration (test hook)."""
    col = images[:, field].clone().detach()
    ori = col.clone()
    adv = images.clone().detach()
    y = _labels(labels).to(images.device)
    for _ in range(iters):
        if snapshots is not None:
            snapshots.append(col.detach().clone())
        col.requires_grad_(True)
        adv[:, field] = col
        logp, _ = model(adv)
        cost = F.cross_entropy(logp.reshape(-1, logp.size(2)), y.view(-1), reduction="sum") / logp.size(1)
        g, = torch.autograd.grad(cost, col)
        adv = adv.detach()
        adv[:, field] = adv[:, field] + alpha * g.sign()
        eta = torch.clamp(adv[:, field] - ori, -eps, eps)
        col = torch.clamp(ori + eta, 0, 1).detach()
    return adv


def tar_nb_attack(model, images, labels, eps=0.3, alpha=2 / 255, iters=40, target=None, mask=None,
                  field=slice(3, 6), snapshots=None):
    """target.py:18-45: masked sign *descent* on mean CE to ``target`` over all points.
    mask: bool [N] (reference, B == 1) or bool [B,N] (generalised, per-block masks)."""
    m = torch.as_tensor(np.asarray(mask)) if not torch.is_tensor(mask) else mask
    B, _, N = images.shape
    m = m.view(1, N).expand(B, N) if m.dim() == 1 else m
    m3 = m.view(B, 1, N)
    col = images[:, field].clone().detach()
    ori = col.clone()
    adv = images.clone().detach()
    tgt = torch.full((B * N,), int(target), dtype=torch.int64)
    for _ in range(iters):
        if snapshots is not None:
            snapshots.append(col.detach().clone())
        col.requires_grad_(True)
        adv[:, field] = torch.where(m3, col, adv[:, field])
        logp, _ = model(adv)
        cost = F.cross_entropy(logp.reshape(-1, logp.size(2)), tgt)
        g, = torch.autograd.grad(cost, col)
        adv = adv.detach()
        stepped = adv[:, field] - alpha * g.sign()
        adv[:, field] = torch.where(m3, stepped, adv[:, field])
        eta = torch.clamp(adv[:, field] - ori, -eps, eps)
        col = torch.clamp(ori + eta, 0, 1).detach()
    return adv


def _atanh_space(c):
    x = c * 2 - 1
    return 0.5 * torch.log((1 + x) / (1 - x))          # nontarget.py:111-117 (Q6: +-inf at 0/1)


def _cw_f(logp, y, kappa):
    """nontarget.py:120-128: clamp(p_label - max_other p, min=-kappa) on softmax(log-probs)."""
    p = F.softmax(logp, dim=2)
    oh = F.one_hot(y, p.size(2)).to(p.dtype)
    other = ((1 - oh) * p).max(2)[0]
    own = (oh * p).max(2)[0]
    return torch.clamp(TARGETED_SIGN * (own - other), min=-kappa)


def _smooth(adv, images, k):
    """nontarget.py:130-135: k smallest colour-space distances, batch element 0 only (Q8)."""
    a = adv[0, 3:6].transpose(1, 0)
    o = images[0, 3:6].transpose(1, 0)
    return torch.cdist(a, o).sort(1)[0][:, :k]


def nu_attack(model, images, labels, c=1e-4, kappa=0, steps=1000, lr=0.01, early_exit=True,
              return_trace=False, field=slice(3, 6), box=None):
    """nontarget.py:52-106.  ``field`` / ``box`` generalise the colour slice of the reference (3:6 in
    [0,1]) to other channels inside a per-channel tanh-space box [lo, hi]; nothing else changes."""
    images = images.clone().detach()
    y = _labels(labels)
    if box is None:
        w0 = _atanh_space(images[:, field].clone())
        lo = hi = None
    else:
        lo = torch.tensor(box[0], dtype=images.dtype).view(1, -1, 1)
        hi = torch.tensor(box[1], dtype=images.dtype).view(1, -1, 1)
        w0 = _atanh_space((images[:, field].clone() - lo) / (hi - lo))
    w = w0.detach().requires_grad_(True)
    best = images.clone()
    opt = torch.optim.Adam([w], lr=lr)
    trace = []
    for step in range(steps):
        col = 0.5 * (torch.tanh(w) + 1)
        if box is not None:
            col = lo + (hi - lo) * col
        adv = best.clone().detach()
        adv[:, field] = col
        l2 = ((adv - images) ** 2).flatten(1).sum(1).sum()
        logp, _ = model(adv)
        f = _cw_f(logp, y, kappa).sum()
        sm = _smooth(adv, images, 10).sum()
        cost = f + c * sm + c * l2
        acc = (logp.max(2)[1] == y).sum().item() / 4096          # Q9
        opt.zero_grad()
        cost.backward()
        opt.step()
        best = adv.clone().detach()
        trace.append((float(cost.detach()), acc))
        if early_exit and acc < 1 / 13:
            break
    return (best, trace) if return_trace else best


def tar_nu_attack(model, images, labels, c=1e-4, kappa=0, steps=1000, lr=0.01, target=None, mask=None,
                  return_trace=False):
    """target.py:62-133, mask bool [N]; batch semantics as in the reference."""
    m = torch.as_tensor(np.asarray(mask)) if not torch.is_tensor(mask) else mask
    images = images.clone().detach()
    y = _labels(labels)
    w = _atanh_space(images[:, 3:6][:, :, m].clone()).detach().requires_grad_(True)
    best = images.clone()
    prev = torch.full([steps], 1e10)
    opt = torch.optim.Adam([w], lr=lr)
    nmask = int(m.sum())
    trace = []
    for step in range(steps):
        col = 0.5 * (torch.tanh(w) + 1)
        adv = best.clone().detach()
        adv[:, 3:6][:, :, m] = col
        l2 = ((adv - images) ** 2).flatten(1).sum(1).sum()
        logp, _ = model(adv)
        if target is None:
            f = _cw_f(logp, y, kappa).sum()                                   # non_f :149-156
        else:
            f = _cw_f(logp, torch.full_like(y, int(target)), kappa).sum()     # tar_f :159-168 (Q3)
        sm = _smooth(adv, images, 5).sum()
        cost = f + c * sm + c * l2
        prev[step] = cost.detach()
        pred = logp.max(2)[1]
        if target is None:
            tacc = (pred[:, m] == y[:, m]).sum().item() / nmask
        else:
            tacc = (pred[:, m] == int(target)).sum().item() / nmask
        opt.zero_grad()
        cost.backward()
        opt.step()
        best = adv.clone().detach()
        trace.append((float(cost.detach()), tacc))
        if (target is None and tacc < 1 / 13) or (target is not None and tacc > 0.9):
            break
        if step > 0 and step % 50 == 0:                                       # :123-125
            lr = lr / 2
            opt = torch.optim.Adam([w], lr=lr)
        if step > 10 and step % 10 == 0 and cost.item() >= prev[step - 10]:   # :127-132 (Q4)
            sel = best[:, 3:6][:, :, m]
            best[:, 3:6][:, :, m] = sel + torch.empty_like(sel).uniform_(0, 1)
            best = torch.clamp(best, 0, 1)
    return (best, trace) if return_trace else best


# --------------------------------------------------------------------------------------------
# script-level metrics  (NB_nontarget_test_semseg.py:187-212)
# --------------------------------------------------------------------------------------------
def block_metrics(pred, labels, num_classes=13):
    """acc, per-class seen / correct / union and the block mIoU as the scripts compute them."""
    pred = np.asarray(pred).reshape(-1)
    lab = np.asarray(labels).reshape(-1).astype(np.int64)
    seen = np.array([(lab == l).sum() for l in range(num_classes)])
    correct = np.array([((pred == l) & (lab == l)).sum() for l in range(num_classes)])
    union = np.array([((pred == l) | (lab == l)).sum() for l in range(num_classes)])
    iou = correct / (union.astype(np.float64) + 1e-6)
    miou = float(np.mean(iou[seen != 0])) if (seen != 0).any() else 0.0
    acc = float((pred == lab).sum()) / pred.size
    return {"acc": acc, "seen": seen, "correct": correct, "union": union, "miou": miou}


def add_vote(vote_label_pool, point_idx, pred_label, weight):
    """NB_nontarget_test_semseg.py:55-62 (vectorised restatement of the double loop)."""
    pi = np.asarray(point_idx).astype(np.int64).reshape(-1)
    pl = np.asarray(pred_label).astype(np.int64).reshape(-1)
    w = np.asarray(weight).reshape(-1) != 0
    np.add.at(vote_label_pool, (pi[w], pl[w]), 1)
    return vote_label_pool


def scene_metrics(vote_label_pool, whole_scene_label, num_classes=13):
    """:216-238: argmax of the pool, per-class seen / correct / union, scene mIoU over the seen classes."""
    return block_metrics(np.argmax(vote_label_pool, 1), whole_scene_label, num_classes)
