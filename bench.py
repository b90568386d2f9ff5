#!/usr/bin/env python
"""bench.py -- PGD attack steps/sec on 4096-pt PointNet++ sem-seg blocks (BASELINE.json metric).

A *step* is one PGD iteration (one forward + one input-gradient backward + the fused perturbation
update) for a batch of 16 S3DIS-shaped 4096-point blocks per GPU: BASELINE.json configs[1]
(PointNet++ SSG sem-seg, norm-bounded targeted PGD, B=16x4096, random-init weights, synthetic data).
The timed region is one ``tar_NB_attack(iters=K)`` call: the batched geometry pass (FPS, ball query,
3-NN, CSRs for all K forwards), the FPS start draws on the CPU generator, and the K-step loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); every rank attacks its own 16 blocks (weak
scaling, no data-path collective) and the per-class counters are all-reduced over NCCL afterwards.
``--impl reference`` times the CPU restatement of the reference's own PyTorch path (oracle/, with
the reference's op-for-op geometry) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

B_PER_GPU = 16
N_POINTS = 4096
EPS, ALPHA, TARGET, ORIGIN = 0.5, 0.1, 7, 11          # NB_target_test_semseg.py:177, :48-49
FLOPS_PER_BLOCK_STEP = 3.880e9                        # forward + dgrad, SSG, N=4096
METRIC = "pgd_attack_steps_per_sec"
UNIT = "steps/s (1 step = fwd + input-grad bwd + update of 16 blocks x 4096 pts per GPU)"


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.maxclk = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.maxclk = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.02)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxclk, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxclk, "reasons": sorted(self.reasons)}


# stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there at any NCCL_DEBUG level from VERSION up, whatever NCCL_DEBUG_FILE says), so main() points fd 1 at
# stderr for the whole run and the line goes out through a private duplicate of the original stdout.
_JSON_FD = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def make_inputs(B, seed):
    from pointsecguard_b200 import synthetic as syn
    x = syn.make_blocks(B, N_POINTS, seed, "uniform")
    labels = syn.zband_labels(x)
    mask = labels == ORIGIN                                  # per-block masks, [B,N]
    return x, labels, mask


# --------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's PyTorch path)
# --------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, sample_blocks=2):
    """Times `steps` PGD iterations of the reference algorithm on `sample_blocks` blocks on all host
    threads and scales to the 16-block step.  Returns (steps_per_s, cores, sample string)."""
    from oracle import attacks_oracle as AO
    from oracle import pointnet2_oracle as PO
    from pointsecguard_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    PO.GEOMETRY = "torch"          # op-for-op the reference's geometry (Python FPS loop, matmul + sort)
    model = PO.OracleModel(syn.make_state_dict("ssg", init="he"), "ssg")
    x, labels, mask = make_inputs(sample_blocks, 0)
    torch.manual_seed(0)
    if warmup > 0:
        AO.tar_nb_attack(model, x, labels.numpy(), eps=EPS, alpha=ALPHA, iters=min(warmup, 1), target=TARGET, mask=mask)
    t0 = time.perf_counter()
    AO.tar_nb_attack(model, x, labels.numpy(), eps=EPS, alpha=ALPHA, iters=steps, target=TARGET, mask=mask)
    dt = time.perf_counter() - t0
    sps = steps / dt * sample_blocks / B_PER_GPU
    sample = (f"tar-NB SSG, {sample_blocks} of the 16 blocks x {steps} iters on {cores} threads "
              f"({dt:.1f} s), reference op-for-op geometry; scaled to 16-block steps")
    return sps, cores, sample, dt


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    sps, cores, sample, dt = cpu_reference(min(steps, 48), args.warmup, sample_blocks=4)
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / sps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: PointNet++ SSG sem_seg, norm-bounded targeted PGD, B=16x4096 per GPU, "
                               "random-init weights, synthetic S3DIS-shaped blocks"},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def _flops_rows_chain(widths):
    """forward FLOPs per row of a conv chain [c0 -> c1 -> ...]"""
    return sum(2 * a * b for a, b in zip(widths[:-1], widths[1:]))


def run_native(args, rank, world, local_rank):
    import ctypes as C
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pointsecguard_b200 import _lib as L
    from pointsecguard_b200 import distributed as D
    from pointsecguard_b200 import synthetic as syn, torchattacks
    from pointsecguard_b200 import metrics as MT
    from pointsecguard_b200.engine import MLP_TF32
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model

    for opt in ("sa_ng", "clusters", "fp_min_tiles", "fp_slabs", "ts"):                 # A/B switches of the library (experiments only)
        if os.environ.get("PSG_OPT_" + opt.upper()):
            L.psg_set_option(opt.encode(), int(os.environ["PSG_OPT_" + opt.upper()]))
    model = get_model(13)
    # init="he": a random network whose predictions depend on the input, so attack_metrics are informative
    model.load_state_dict(syn.make_state_dict("ssg", init="he"))
    model = model.to(dev).eval()
    if args.mlp == "tf32":
        model.set_mlp_mode(MLP_TF32)
    # weak scaling: a global batch of 16 blocks per GPU; every rank builds the same global batch from
    # the seed, keeps its contiguous slice, and draws FPS starts for the GLOBAL batch (sliced), so an
    # N-GPU run attacks exactly the blocks a 1-GPU run of the same global batch would, with the same draws
    shard = D.shard_for(B_PER_GPU * world, rank, world)
    model.set_shard(shard if world > 1 else None)
    x_all, labels_all, mask_all = make_inputs(B_PER_GPU * world, seed=0)
    x_host, labels, mask = shard.slice(x_all), shard.slice(labels_all), shard.slice(mask_all)
    x_pin = x_host.contiguous().pin_memory()                      # [B,9,N] contiguous pinned host copy
    x_dev = x_host.to(dev)
    adv_host = torch.empty(x_pin.shape, dtype=torch.float32).pin_memory()      # e2e: the perturbed blocks land here
    lab_np = labels.numpy().astype(np.float64)
    K, W = args.steps, args.warmup
    mk = lambda iters: torchattacks.tar_NB_attack(model, eps=EPS, alpha=ALPHA, iters=iters, target=TARGET, mask=mask)
    atk = mk(K)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: W untimed steps, then one untimed attack at the timed geometry-chunk size
    if W > 0:
        mk(W)(x_dev, lab_np)
    mk(min(K, 64))(x_dev, lab_np)

    # ---- device-resident timing: exactly K steps ----
    # clocks are sampled on rank 0 only (its line carries them; eight ranks polling NVML at once stalled each other's
    # first launches: +2.6 ms per attack at 8 GPUs, none at 200 steps), and the sampler thread is already running --
    # NVML handle, first query -- when the timed region starts
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.manual_seed(0)
    flush.fill_(1)
    if clocks is not None:
        clocks.start()
        time.sleep(0.03)
    barrier()
    l0 = L.psg_launch_count()
    e0.record()
    adv = atk(x_dev, lab_np)
    e1.record()
    barrier()
    launches = L.psg_launch_count() - l0
    clk = clocks.finish() if clocks is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * K / (ms_total / 1e3)

    # ---- end to end through the public API from pinned host memory, result read back ----
    torch.manual_seed(0)
    flush.fill_(2)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    xd = x_pin.to(dev, non_blocking=True)
    adv2 = atk(xd, lab_np)
    adv_host.copy_(adv2, non_blocking=True)          # result read back into pinned host memory
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * K / (float(ms2.item()) / 1e3)
    h2d = x_pin.numel() * 4 / K
    d2h = adv_host.numel() * 4 / K
    same_as_resident = bool(torch.equal(adv_host, adv.cpu()))

    # ---- metrics + the only collective: per-class counters all-reduced over NCCL ----
    torch.manual_seed(1)
    logp_adv, _ = model(adv)
    counters = MT.attack_counters(logp_adv, labels.to(dev), mask.to(dev), TARGET)
    D.all_reduce_sum_(counters)
    summary = MT.summarize(counters.cpu(), 13)

    line = None
    if rank == 0:
        # ---- live per-kernel-family timing (CUDA event pairs on the launching stream) ----
        L.psg_prof_enable(1)
        kp = K        # same attack length as the timed run: the batched geometry pass amortises over the same number of forwards
        torch.manual_seed(0)
        mk(kp)(x_dev, lab_np)
        ncat = L.psg_prof_ncat()
        msb = (C.c_double * ncat)()
        cnt = (C.c_int64 * ncat)()
        L.psg_prof_collect(msb, cnt)
        L.psg_prof_enable(0)
        fam = {L.psg_prof_name(i).decode(): {"ms_per_step": msb[i] / kp, "launches_per_step": cnt[i] / kp}
               for i in range(ncat) if cnt[i]}
        pk = peaks()
        # fp32 CUDA-core MLPs have no tensor peak; the TF32 tcgen05 peak is half the measured bf16 one
        peak = pk["bf16_tflops_sustained"] / 2.0
        rows0 = B_PER_GPU * N_POINTS
        # dominant kernel: fp1 + head forward + backward (chain_fused.cu), one launch per step:
        # 2 x (128->128->128->128 fp1, 128->128 conv1, 128->13 conv2) FLOPs per point
        chain_flops = 2 * _flops_rows_chain([128, 128, 128, 128, 128, 13]) * rows0
        if "head_chain" in fam:
            dom = fam["head_chain"]
            dur_ms = dom["ms_per_step"] / max(dom["launches_per_step"], 1)
            ach = chain_flops / (dur_ms / 1e3) / 1e12
            roofline = {
                "bound": "tensor", "kernel": "chain_kernel: fp1 + conv1 + conv2 forward, loss gradient, dgrad chain (1 launch/step)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "peak_source": f"{pk['source']} bf16 sustained / 2 (TF32 dense rate)",
                "flops_per_launch": chain_flops, "ms_per_launch": dur_ms,
                "traffic": _ncu_traffic("chain_kernel"),
            }
        else:
            g = fam.get("gemm_fwd", {}).get("ms_per_step", 0) + fam.get("gemm_bwd", {}).get("ms_per_step", 0)
            ach = FLOPS_PER_BLOCK_STEP * B_PER_GPU / (g / 1e3) / 1e12 if g else 0.0
            roofline = {"bound": "tensor", "kernel": "fp32 CUDA-core GEMMs (all layers, parity mode)", "achieved": ach,
                        "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                        "peak_source": f"{pk['source']} bf16 sustained / 2 (TF32 dense rate)"}
        mlp_ms = sum(fam.get(k, {}).get("ms_per_step", 0) for k in
                     ("gemm_fwd", "gemm_bwd", "sa_fused_fwd", "sa_fused_bwd", "fp_fused_fwd", "fp_fused_bwd", "head_chain"))
        roofline["all_mlp_kernels"] = {
            "ms_per_step": mlp_ms, "tflops": FLOPS_PER_BLOCK_STEP * B_PER_GPU / (mlp_ms / 1e3) / 1e12 if mlp_ms else None,
            "frac": FLOPS_PER_BLOCK_STEP * B_PER_GPU / (mlp_ms / 1e3) / 1e12 / peak if mlp_ms else None}
        roofline["mlp_mode"] = args.mlp
        # HBM-side primitives (SURVEY.md 8d byte formulas, all four levels), achieved GB/s vs measured copy bandwidth
        npts = [N_POINTS, 1024, 256, 64, 16]
        byt = {"fps": sum(12 * npts[l] + 8 * npts[l + 1] for l in range(4)),
               "ball_query": sum(12 * npts[l] + 12 * npts[l + 1] + 8 * npts[l + 1] * 32 for l in range(4)),
               "three_nn": sum(12 * npts[l] + 12 * npts[l + 1] + 36 * npts[l] for l in range(4))}
        roofline["primitives_hbm"] = {
            k: {"GBps": byt[k] * B_PER_GPU / (fam[k]["ms_per_step"] / 1e3) / 1e9,
                "frac_of_hbm": byt[k] * B_PER_GPU / (fam[k]["ms_per_step"] / 1e3) / 1e9 / pk["hbm_gbs"]}
            for k in byt if k in fam}
        # ---- CPU baseline (bounded sample of the same workload on the host cores) ----
        if args.no_cpu or world > 1:          # the CPU leg runs on rank 0 at N = 1 only
            cpu = None
        else:
            sps, cores, sample, _ = cpu_reference(40, 1, sample_blocks=4)
            cpu = {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.mlp == "fp32" else "tf32", "data": "synthetic",
            "config": {"workload": "configs[1]: PointNet++ SSG sem_seg, norm-bounded targeted PGD (tar_NB_attack eps=0.5 "
                                   "alpha=0.1 target=7, mask=z-band class 11), B=16x4096 per GPU, random-init weights, "
                                   "synthetic S3DIS-shaped blocks",
                       "blocks_per_gpu": B_PER_GPU, "points": N_POINTS, "mlp": args.mlp,
                       "tolerance": "fp32: logits rtol 1e-3; tf32: |dlogp| < 5e-3, grad rel < 8e-2, sign agreement > 99 % "
                                    "(tests/test_gpu_gemm.py), FPS / ball-query / 3-NN indices bit-exact in both modes",
                       "l2": "256 MB buffer written before each timed attack (L2 flush); steps inside an attack run "
                             "back to back as in the reference loop"},
            "block_steps_per_s": value * B_PER_GPU,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "wall_s": wall, "identical_to_resident_run": same_as_resident},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "kernel_families_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in fam.items()},
            "cpu_baseline": cpu,
            "attack_metrics": summary,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def _ncu_traffic(kernel):
    """dram bytes (read + write) per launch of `kernel` from the committed ncu --set full summary, or None."""
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mlp", default=os.environ.get("PSG_MLP", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1")
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
