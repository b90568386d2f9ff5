"""pointsecguard_b200 -- B200-native (sm_100a) implementation of the PointNet++ semantic-segmentation
attack hot path of C0ldstudy/PointSecGuard.

Layout (mirrors the reference's PointNet/ tree for the hot path only):
    models/pointnet_util.py            geometric primitives + SA / SA-MSG / FP modules
    models/pointnet2_sem_seg.py        SSG sem-seg network (get_model, get_loss)
    models/pointnet2_sem_seg_msg.py    MSG sem-seg network
    torchattacks/                      NB_attack, NU_attack, tar_NB_attack, tar_NU_attack
    engine.py, ops.py, _lib.py         host side of the C ABI (include/psg_b200.h)
    csrc/                              hand-written sm_100a CUDA kernels + the C ABI
"""
__version__ = "0.1.0"
