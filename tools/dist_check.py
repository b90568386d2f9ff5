#!/usr/bin/env python
"""Multi-GPU consistency check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Every rank attacks its shard of a global batch (NB and NU); rank 0 also runs the whole batch alone and
checks that the gathered shards reproduce it bit for bit, that the all-reduced counters equal the
single-GPU counters exactly, and that the sharded NU (whose accuracy test is a batch-wide sum exchanged
every step) takes the same early exit as the unsharded one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # PSG_DIST_BACKEND=gloo lets several ranks share one GPU (NCCL refuses that): the host-side sharding logic, the
    # all-reduced counters and the per-step NU exchange are the same code, only the transport differs
    backend = os.environ.get("PSG_DIST_BACKEND", "nccl")
    local = local % torch.cuda.device_count()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)
    from pointsecguard_b200 import distributed as D, metrics as MT, synthetic as syn, torchattacks
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    m = get_model(13)
    m.load_state_dict(syn.make_state_dict("ssg", init="he"))
    m = m.to(dev).eval()
    G = 2 * world
    x = syn.make_blocks(G, 2048, 11)
    torch.manual_seed(5)
    lab = m(x.to(dev))[0].argmax(2).cpu()
    sh = D.shard_for(G)
    ok = True

    def gather(t):
        if backend != "nccl":                       # gloo gathers host tensors
            t = t.cpu()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return torch.cat(parts).to(dev)

    for name, make in (("NB", lambda mm: torchattacks.NB_attack(mm, eps=0.1, alpha=0.05, iters=4)),
                       ("NU", lambda mm: torchattacks.NU_attack(mm, c=0.1, kappa=0, steps=6, lr=0.01))):
        m.set_shard(sh)
        torch.manual_seed(3)
        adv = make(m)(sh.slice(x).to(dev), sh.slice(lab).numpy().astype(np.float64))
        torch.manual_seed(4)
        cnt = MT.attack_counters(m(adv)[0], sh.slice(lab).to(dev))
        D.all_reduce_sum_(cnt)
        full = gather(adv)
        m.set_shard(None)
        if rank == 0:
            dist_backup = dist.get_world_size
            # the single-GPU run must not see the process group as a sharded world
            torch.manual_seed(3)
            import pointsecguard_b200.distributed as DD
            ws = DD.world_size
            DD.world_size = lambda: 1
            ref = make(m)(x.to(dev), lab.numpy().astype(np.float64))
            torch.manual_seed(4)
            rc = MT.attack_counters(m(ref)[0], lab.to(dev))
            DD.world_size = ws
            same = torch.equal(full, ref)
            csame = torch.equal(cnt, rc)
            print(f"{name}: shards == full batch: {same}; counters equal: {csame}; acc {MT.summarize(rc.cpu())['acc']:.4f}", flush=True)
            ok = ok and same and csame
        dist.barrier()
    # ---- whole-scene evaluation (slicer -> attack -> votes -> IoU): scenes strided over the ranks, counters all-reduced ----
    import tempfile
    from pointsecguard_b200 import scene_eval
    from pointsecguard_b200.data_utils.S3DISDataLoader import ScannetDatasetWholeScene
    nscenes = 2 * world + 1
    box = [None]
    if rank == 0:                                             # one directory for all ranks: the same os.listdir order everywhere
        box[0] = tempfile.mkdtemp()
        for i in range(nscenes):
            np.save(os.path.join(box[0], f"Area_5_room_{i:02d}.npy"), syn.make_room(2400 + 100 * i, 40 + i, "tiny"))
    dist.broadcast_object_list(box, 0)
    d = box[0]
    ds = ScannetDatasetWholeScene(d + "/", block_points=1024)
    mk = lambda: torchattacks.NB_attack(m, eps=0.1, alpha=0.05, iters=2)
    res = scene_eval.evaluate_dataset(m, ds, mk, batch_size=4, seed=100)
    if rank == 0:
        import pointsecguard_b200.distributed as DD
        ws, rk = DD.world_size, DD.rank
        DD.world_size, DD.rank = (lambda: 1), (lambda: 0)
        saved = DD.all_reduce_sum_
        DD.all_reduce_sum_ = lambda t: t
        ref = scene_eval.evaluate_dataset(m, ds, mk, batch_size=4, seed=100)
        DD.world_size, DD.rank, DD.all_reduce_sum_ = ws, rk, saved
        same = torch.equal(res["counters"], ref["counters"])
        mine_ok = all(torch.equal(res["scenes"][i]["pool"].pool, ref["scenes"][i]["pool"].pool) for i in res["scenes"])
        print(f"scene_eval: {nscenes} scenes over {world} ranks: global counters equal: {same}; this rank's vote pools equal: {mine_ok}; "
              f"adv scene mIoU {ref['adv_scene']['miou_seen']:.4f} (clean {ref['scene']['miou_seen']:.4f})", flush=True)
        ok = ok and same and mine_ok
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev if backend == "nccl" else "cpu")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
