"""Whole-scene attack evaluation on the GPU (slicer -> model / attack -> votes -> IoU) against the script loop
restated on the CPU from the oracle pieces (NB_nontarget_test_semseg.py:138-241)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from pointsecguard_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _oracle_scene_loop(om, room, bp, batch_size, eps, alpha, iters):
    from oracle import attacks_oracle as AO
    from oracle import scene_slicer_oracle as SO
    lw = SO.label_weights([room[:, 6]])
    P = room.shape[0]
    pool, adv_pool = np.zeros((P, 13)), np.zeros((P, 13))
    adv_scene = np.zeros((P, 6))
    data, label, smpw, pidx = SO.slice_room(room, lw, bp)
    preds, adv_preds, labs = [], [], []
    for s in range(0, data.shape[0], batch_size):
        e = min(s + batch_size, data.shape[0])
        x = torch.Tensor(data[s:e]).float().transpose(2, 1)
        seg = om(x)[0]
        adv = AO.nb_attack(om, x, label[s:e].astype(np.float64), eps=eps, alpha=alpha, iters=iters)
        adv_seg = om(adv)[0]
        adv_scene[pidx[s:e].reshape(-1).astype(int)] = adv.transpose(1, 2)[:, :, :6].detach().numpy().reshape(-1, 6)
        p, ap = seg.argmax(2).numpy(), adv_seg.argmax(2).numpy()
        AO.add_vote(pool, pidx[s:e], p, smpw[s:e])
        AO.add_vote(adv_pool, pidx[s:e], ap, smpw[s:e])
        preds.append(p); adv_preds.append(ap); labs.append(label[s:e])
    blk = AO.block_metrics(np.concatenate(preds), np.concatenate(labs))
    adv_blk = AO.block_metrics(np.concatenate(adv_preds), np.concatenate(labs))
    return {"pool": pool, "adv_pool": adv_pool, "adv_scene": adv_scene, "block": blk, "adv_block": adv_blk,
            "scene": AO.scene_metrics(pool, room[:, 6]), "adv_scene_m": AO.scene_metrics(adv_pool, room[:, 6]), "index": pidx}


def test_scene_evaluation_matches_the_restated_script_loop():
    from oracle import pointnet2_oracle as PO
    from pointsecguard_b200 import scene_eval, torchattacks
    from pointsecguard_b200.data_utils.S3DISDataLoader import ScannetDatasetWholeScene
    from pointsecguard_b200.models.pointnet2_sem_seg import get_model
    sd = syn.make_state_dict("ssg", init="he")
    m = get_model(13)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    om = PO.OracleModel(sd, "ssg")
    room = syn.make_room(2600, 3, "tiny")
    # labels the network can partly predict: its own clean prediction is not available per scene point, keep z-bands
    bp, bs, eps, alpha, iters = 1024, 4, 0.1, 0.05, 2
    d = tempfile.mkdtemp()
    np.save(os.path.join(d, "Area_5_synthetic_1.npy"), room)
    ds = ScannetDatasetWholeScene(d + "/", block_points=bp)
    np.random.seed(21); torch.manual_seed(22)
    ref = _oracle_scene_loop(om, room, bp, bs, eps, alpha, iters)
    np.random.seed(21); torch.manual_seed(22)
    got = scene_eval.evaluate_scene(m, ds, 0, lambda: torchattacks.NB_attack(m, eps=eps, alpha=alpha, iters=iters), batch_size=bs)
    # votes: identical up to the few points whose fp32 arg-max differs between the CPU and the GPU execution
    for a, b in ((got["pool"].pool, ref["pool"]), (got["adv_pool"].pool, ref["adv_pool"])):
        a = a.cpu().numpy()
        assert a.sum() == b.sum()                                    # same number of votes cast
        assert (a != b).any(1).mean() < 0.01
    same = (got["adv_whole_scene"].cpu().numpy() == ref["adv_scene"].astype(np.float32)).all(1).mean()
    assert same > 0.99, same
    for key_g, key_r in (("block", "block"), ("adv_block", "adv_block")):
        assert abs(got[key_g]["acc"] - ref[key_r]["acc"]) < 0.005 and abs(got[key_g]["miou"] - ref[key_r]["miou"]) < 0.005
    assert abs(got["scene"]["miou_seen"] - ref["scene"]["miou"]) < 0.005 and abs(got["scene"]["acc"] - ref["scene"]["acc"]) < 0.005
    assert abs(got["adv_scene"]["miou_seen"] - ref["adv_scene_m"]["miou"]) < 0.005
    # dataset-level driver: one scene, one rank == the scene result
    np.random.seed(21); torch.manual_seed(22)
    allr = scene_eval.evaluate_dataset(m, ds, lambda: torchattacks.NB_attack(m, eps=eps, alpha=alpha, iters=iters), batch_size=bs)
    assert torch.equal(allr["scenes"][0]["pool"].pool, got["pool"].pool)
    assert abs(allr["adv_scene"]["miou_seen"] - got["adv_scene"]["miou_seen"]) < 1e-12


def test_scatter_last_wins_is_numpy_assignment():
    from pointsecguard_b200.scene_eval import _scatter_last_wins
    rng = np.random.default_rng(0)
    idx = rng.integers(0, 500, 4000)
    rows = rng.random((4000, 6)).astype(np.float32)
    ref = np.zeros((700, 6), np.float32)
    ref[idx] = rows
    dst = torch.zeros(700, 6, device="cuda")
    _scatter_last_wins(dst, torch.from_numpy(idx).cuda(), torch.from_numpy(rows).cuda())
    assert np.array_equal(dst.cpu().numpy(), ref)
