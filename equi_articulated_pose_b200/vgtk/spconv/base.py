"""Spherical point-cloud containers (reference: vgtk/vgtk/spconv/base.py:4-44).

xyz [B,3,N]; feats is LOGICALLY [B,C,N,A] like the reference.  The B200 kernels keep it
channels-last in memory ([B,N,A,C]); `feats` is then a permuted view, so reference-style
consumers (1x1 Conv2d, means over dims, ...) keep working without a copy."""
from vgtk.point3d import PointSet


class SphericalPointCloud():
    def __init__(self, xyz, feats, anchors):
        self._xyz = PointSet(xyz)
        self._feats = feats
        self._anchors = anchors

    @property
    def xyz(self):
        return self._xyz.data

    @property
    def feats(self):
        return self._feats

    @property
    def anchors(self):
        return self._anchors


class SphericalPointCloudPose(SphericalPointCloud):
    def __init__(self, xyz, feats, anchors, pose):
        super().__init__(xyz, feats, anchors)
        self._pose = pose      # [B,N,4,4] per-point rigid pose

    @property
    def pose(self):
        return self._pose
