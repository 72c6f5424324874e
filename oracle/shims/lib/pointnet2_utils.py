"""oracle/shims/lib/pointnet2_utils.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU stand-in for the reference's absent ``lib.pointnet2_utils`` extension module, so that the
UNMODIFIED reference model (``/root/reference/scripts/ActiveSceneFlow/TFlowV3_Occlussion.py``)
can be imported and run on CPU to generate golden vectors (oracle/gen_golden.py), and so the
oracle port can be timed as the CPU baseline.  Call surface = SURVEY.md section 2.3 / 8(b);
call sites: ASF/utils/utils.py:226-233,291-302, ASF/utils/soflow.py:30,387-406,1241-1249,1459-1470.
"""
import torch

from oracle import point_ops as _ops


def furthest_point_sample(xyz, npoint):
    return _ops.furthest_point_sample(xyz, npoint)


def gather_operation(features, idx):
    return _ops.gather_operation(features, idx)


def knn(k, unknown, known):
    return _ops.knn(k, unknown, known)


def three_nn(unknown, known):
    return _ops.three_nn(unknown, known)


def grouping_operation(features, idx):
    return _ops.grouping_operation(features, idx)


def three_interpolate(features, idx, weight):
    return _ops.three_interpolate(features, idx, weight)


def ball_query(radius, nsample, xyz, new_xyz):
    return _ops.ball_query(radius, nsample, xyz, new_xyz)[0]


class QueryAndGroup(torch.nn.Module):
    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        g = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            return g
        gf = grouping_operation(features, idx)
        return torch.cat([g, gf], 1) if self.use_xyz else gf


class GroupAll(torch.nn.Module):
    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        g = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return g
        gf = features.unsqueeze(2)
        return torch.cat([g, gf], 1) if self.use_xyz else gf
