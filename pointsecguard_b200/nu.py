"""Host side of the norm-unbounded colour attacks -- NU_attack (reference
PointNet/attacks/torchattacks/attacks/nontarget.py:52-106) and tar_NU_attack (target.py:62-133).

The per-step work (tanh-space colours -> forward -> C&W f -> input gradient -> smoothness term ->
cost / accuracy test -> Adam) is one C call, ``psg_nu_step`` (csrc/net.cu, csrc/nu.cu), with no host
round trip.  The host keeps only what the reference decides on the host:

* the FPS start draws on the CPU generator, in the reference's order (pointnet_util.py:75);
* the early exit: the device latches a ``done`` flag at the step whose accuracy test fires and
  freezes the attack state; the host reads it once per *chunk* of steps and rewinds the CPU
  generator to where the reference left it (the draws of the steps it never ran are un-done);
* tar-NU's schedule (target.py:123-132): every 50 steps halve ``lr`` and re-create Adam; every 10
  steps after step 10 compare the cost with the one 10 steps earlier and, if it did not drop, add
  uniform noise to the masked colours (overwritten by tanh(w) at the next step) and clamp **all
  nine channels** to [0,1] -- which silently changes xyz for the rest of the attack (Q4), so the
  geometry is rebuilt from the clamped image.  The noise draw is made on the CPU generator and
  discarded, which keeps later FPS start draws aligned with the CPU reference.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L
from . import distributed as D

_MAX_PROBLEMS = 1024     # FPS problems (= forwards x blocks) whose geometry is resident at once


def _draw_chunk(eng, T):
    """FPS starts for T forwards + the generator state after each forward's draws."""
    starts, states = [], []
    for t in range(T):
        starts.append(eng.draw_starts(1))
        states.append(torch.get_rng_state())
    return torch.cat(starts, dim=1), states


# Default tanh-space box of the coordinate + colour field (an extension of the reference, whose field is
# colour only): block-centred x, y in [-0.5, 0.5], z in [0, 3] metres (S3DISDataLoader.py:154-161) with
# a 10 % margin so that points on the block boundary do not map to +-inf (Q6), colours in [0, 1].
COORD_COLOR_BOX = ([-0.55, -0.55, -0.15, 0.0, 0.0, 0.0], [0.55, 0.55, 3.15, 1.0, 1.0, 1.0])


def nu_attack(atk, images, labels, mask, target, neighbour, masked_variant=None):
    """Shared driver.  ``masked_variant`` False -> NU_attack semantics (nontarget.py), True ->
    tar_NU_attack semantics (target.py); default: masked iff a mask is given.

    ``atk.field = (c0, c1)`` (default: the reference's colours, 3:6) selects the perturbed channels;
    a field that includes coordinates makes every step rebuild the geometry from the step's image and
    adds the geometric gradient (csrc/geomgrad.cu) to the coordinate channels."""
    eng = atk._engine(images)
    dev = images.device
    B, Cc, N = images.shape
    field = getattr(atk, "field", None) or (3, 6)
    c0, nc = int(field[0]), int(field[1]) - int(field[0])
    moves_xyz = c0 < 3
    box = getattr(atk, "box", None)
    if box is None:
        box = ([COORD_COLOR_BOX[0][c] for c in range(c0, c0 + nc)], [COORD_COLOR_BOX[1][c] for c in range(c0, c0 + nc)]) \
            if c0 + nc <= 6 else ([0.0] * nc, [1.0] * nc)
    tar_variant = (mask is not None) if masked_variant is None else masked_variant
    if tar_variant and mask is None:
        raise ValueError("tar_NU_attack needs a mask (target.py:53)")
    img = images.detach().to(torch.float32).contiguous()            # `images` of the reference (never modified)
    base = img                                                       # best_adv_images minus the colours
    lab = atk._labels_i32(labels, dev)
    msk = atk._mask_u8(mask, B, N, dev) if mask is not None else None
    steps = int(atk.steps)
    st = torch.cuda.current_stream().cuda_stream

    w = torch.empty(B, nc, N, dtype=torch.float32, device=dev)
    am, av = torch.empty_like(w), torch.empty_like(w)
    adv = img.clone()
    cost = torch.zeros(max(steps, 1), dtype=torch.float32, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.psg_nu_scratch_floats(B, N), dtype=torch.float32, device=dev)
    buf = L.NuBuffers(w.data_ptr(), am.data_ptr(), av.data_ptr(), adv.data_ptr(), base.data_ptr(), img.data_ptr(),
                      msk.data_ptr() if msk is not None else None, lab.data_ptr(), cost.data_ptr(), status.data_ptr(),
                      scratch.data_ptr())
    if (c0, nc) != (3, 3) or getattr(atk, "box", None) is not None:
        buf.field_c0, buf.field_nc = c0, nc
        for j in range(nc):
            buf.box_lo[j], buf.box_hi[j] = float(box[0][j]), float(box[1][j])

    if tar_variant:
        m_cpu = torch.as_tensor(np.asarray(mask)) if not torch.is_tensor(mask) else mask.cpu()
        m_cpu = m_cpu.to(torch.bool)
        # target.py:101-105: hits over the masked points / mask.sum() (the reference's mask is [N])
        denom = max(float(m_cpu.sum().item()), 1.0)
        thr, above = (1.0 / 13.0, 0) if target is None else (0.9, 1)
        masked_only = 1
    else:
        denom, thr, above, masked_only = 4096.0, 1.0 / 13.0, 0, 0      # nontarget.py:87,95 (Q9)
    tgt = -1 if target is None else int(target)
    sign = float(atk._targeted)                                        # +1 unless set_attack_mode was called (Q3)

    chunk_cap = max(1, _MAX_PROBLEMS // B)
    # Sharded over ranks (distributed.py): the accuracy test and the cost are sums over the GLOBAL
    # batch, so every step ends with one tiny all-reduce and the host decides, as the reference's
    # per-step .item() does; the smoothness term belongs to the rank that owns global block 0.
    sharded = eng.shard is not None and D.world_size() > 1
    host_cost = {}
    thr_dev = thr
    if sharded:
        chunk_cap = 1
        thr_dev = 1e30 if above else -1.0                              # the device latch never fires
        if not eng.shard.owns_block0:
            neighbour = 0
        if tar_variant and m_cpu.dim() == 2:
            dn = torch.tensor([denom], dtype=torch.float64, device=dev)
            denom = max(float(D.all_reduce_sum_(dn).item()), 1.0)
    if moves_xyz:
        chunk_cap = 1                                                  # geometry changes with every step
    eng.bind(B, N, min(chunk_cap, max(steps, 1)))
    eng.set_input(img)
    L.psg_nu_init(eng._net, C.byref(buf), st)
    L.psg_net_set_xyz_grad(eng._net, 1 if moves_xyz else 0)

    lr = float(atk.lr)
    adam_k, reset = 0, 0
    step = 0
    while step < steps:
        # a chunk ends where the host has to look at device results (tar-NU's every-10-steps test)
        end = min(steps, step + chunk_cap)
        if tar_variant:
            end = min(end, max(20, (step + 9) // 10 * 10) + 1)
        T = end - step
        starts, states = _draw_chunk(eng, T)
        starts_dev = None
        if moves_xyz:
            starts_dev = starts.to(device=dev, dtype=torch.int32).contiguous()     # [4, 1, B]: rebuilt inside the step
        else:
            eng.geometry(starts)
        lr_at = []                       # learning rate in force when each step of the chunk started
        for i in range(T):
            s = step + i
            lr_at.append(lr)
            adam_k += 1
            step_size = lr / (1.0 - 0.9 ** adam_k)
            bc2 = math.sqrt(1.0 - 0.999 ** adam_k)
            L.psg_nu_step(eng._net, C.byref(buf), i, s, tgt, int(neighbour), float(atk.c), float(atk.kappa), sign,
                          step_size, bc2, reset, denom, thr_dev, above, masked_only,
                          starts_dev.data_ptr() if starts_dev is not None else None, st)
            reset = 0
            if tar_variant and s > 0 and s % 50 == 0:                 # target.py:123-125
                lr = lr / 2
                atk.lr = lr
                adam_k, reset = 0, 1
        last = end - 1
        if sharded:
            red = torch.stack([status[2].double(), cost[last].double()])
            hits, gcost = D.all_reduce_sum_(red).cpu().tolist()       # the step's only exchange (2 numbers)
            host_cost[last] = gcost
            acc = hits / denom
            if (acc > thr) if above else (acc < thr):
                atk.lr = lr_at[-1]       # the reference returns before the halving of the exit step (target.py:118-125)
                break
        else:
            stat = status.cpu()                                       # the chunk's only host sync
            if int(stat[0]):
                torch.set_rng_state(states[int(stat[1]) - step])      # un-draw the steps the reference never ran
                if tar_variant:
                    atk.lr = lr_at[int(stat[1]) - step]               # ... and un-halve the rate (target.py:118-125)
                break
        if tar_variant and last > 10 and last % 10 == 0:              # target.py:127-132
            c2 = [host_cost[last], host_cost[last - 10]] if sharded else cost[[last, last - 10]].cpu().tolist()
            if float(c2[0]) >= float(c2[1]):
                noise = _bingo_noise(m_cpu, B, eng.shard if sharded else None)   # CPU-generator draw, like the CPU reference's
                if last == steps - 1:
                    # the loop ends here: the reference returns the noised, all-channel-clamped image
                    _apply_noise(adv, noise, m_cpu)
                    L.psg_clamp(adv.data_ptr(), adv.numel(), 0.0, 1.0, st)
                elif base is img:
                    # the noise is overwritten by tanh(w) at the next step; what persists is the clamp
                    # of all nine channels (Q4): xyz changes, so the geometry is rebuilt from it
                    base = img.clone()
                    L.psg_clamp(base.data_ptr(), base.numel(), 0.0, 1.0, st)
                    buf.base = base.data_ptr()
                    eng.set_input(base)
        step = end
    L.psg_net_set_xyz_grad(eng._net, 0)
    atk.model._generation += 1
    atk.last_cost = cost          # per-step cost of this call (device tensor; entries past an early exit stay 0)
    return adv


def _bingo_noise(m_cpu, B, shard=None):
    """torch.empty_like(best[:, 3:6][:, :, mask]).uniform_(0, 1) of target.py:131 (under sharding:
    drawn for the global batch and sliced, so every rank consumes the generator identically)."""
    if m_cpu.dim() == 1:
        full = torch.empty(shard.global_batch if shard else B, 3, int(m_cpu.sum())).uniform_(0, 1)
        return shard.slice(full) if shard else full
    return [torch.empty(3, int(m_cpu[b].sum())).uniform_(0, 1) for b in range(B)]


def _apply_noise(adv, noise, m_cpu):
    m = m_cpu.to(adv.device)
    if m_cpu.dim() == 1:
        col = adv[:, 3:6]
        col[:, :, m] = col[:, :, m] + noise.to(adv.device)
    else:
        for b, nz in enumerate(noise):
            col = adv[b, 3:6]
            col[:, m[b]] = col[:, m[b]] + nz.to(adv.device)
