"""Drop-in for PointNet/models/pointnet2_sem_seg.py (SSG sem-seg network, :6-49)."""
import torch.nn as nn
import torch.nn.functional as F

from pointsecguard_b200.models._semseg_base import SemSegBase
from pointsecguard_b200.models.pointnet_util import PointNetFeaturePropagation, PointNetSetAbstraction


class get_model(SemSegBase):
    arch = "ssg"

    def __init__(self, num_classes):
        super().__init__()
        self.sa1 = PointNetSetAbstraction(1024, 0.1, 32, 9 + 3, [32, 32, 64], False)
        self.sa2 = PointNetSetAbstraction(256, 0.2, 32, 64 + 3, [64, 64, 128], False)
        self.sa3 = PointNetSetAbstraction(64, 0.4, 32, 128 + 3, [128, 128, 256], False)
        self.sa4 = PointNetSetAbstraction(16, 0.8, 32, 256 + 3, [256, 256, 512], False)
        self.fp4 = PointNetFeaturePropagation(768, [256, 256])
        self.fp3 = PointNetFeaturePropagation(384, [256, 256])
        self.fp2 = PointNetFeaturePropagation(320, [256, 128])
        self.fp1 = PointNetFeaturePropagation(128, [128, 128, 128])
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)
        self._init_runtime()


class get_loss(nn.Module):
    """pointnet2_sem_seg.py:43-49: weighted NLL = F.nll_loss(pred, target, weight=weight), as its own deterministic kernel pair
    (csrc/train.cu: ordered double-precision reduction forward, one elementwise kernel backward)."""

    def forward(self, pred, target, trans_feat, weight):
        from pointsecguard_b200.train import nll_loss
        return nll_loss(pred, target, weight)
