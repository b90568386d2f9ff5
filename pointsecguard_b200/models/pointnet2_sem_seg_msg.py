"""Drop-in for PointNet/models/pointnet2_sem_seg_msg.py (MSG sem-seg network, :6-50)."""
import torch.nn as nn
import torch.nn.functional as F

from pointsecguard_b200.models._semseg_base import SemSegBase
from pointsecguard_b200.models.pointnet_util import PointNetFeaturePropagation, PointNetSetAbstractionMsg


class get_model(SemSegBase):
    arch = "msg"

    def __init__(self, num_classes):
        super().__init__()
        self.sa1 = PointNetSetAbstractionMsg(1024, [0.05, 0.1], [16, 32], 9, [[16, 16, 32], [32, 32, 64]])
        self.sa2 = PointNetSetAbstractionMsg(256, [0.1, 0.2], [16, 32], 32 + 64, [[64, 64, 128], [64, 96, 128]])
        self.sa3 = PointNetSetAbstractionMsg(64, [0.2, 0.4], [16, 32], 128 + 128, [[128, 196, 256], [128, 196, 256]])
        self.sa4 = PointNetSetAbstractionMsg(16, [0.4, 0.8], [16, 32], 256 + 256, [[256, 256, 512], [256, 384, 512]])
        self.fp4 = PointNetFeaturePropagation(512 + 512 + 256 + 256, [256, 256])
        self.fp3 = PointNetFeaturePropagation(128 + 128 + 256, [256, 256])
        self.fp2 = PointNetFeaturePropagation(32 + 64 + 256, [256, 128])
        self.fp1 = PointNetFeaturePropagation(128, [128, 128, 128])
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)
        self._init_runtime()


class get_loss(nn.Module):
    """pointnet2_sem_seg_msg.py:44-50: weighted NLL = F.nll_loss(pred, target, weight=weight), as its own deterministic kernel pair
    (csrc/train.cu: ordered double-precision reduction forward, one elementwise kernel backward)."""

    def forward(self, pred, target, trans_feat, weight):
        from pointsecguard_b200.train import nll_loss
        return nll_loss(pred, target, weight)
