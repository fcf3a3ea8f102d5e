"""Point container used by SphericalPointCloud (reference: vgtk/vgtk/point3d/base.py:15-42)."""


class PointSet():
    def __init__(self, p):
        # p: [(b,) 3 or 4, n]
        self._p = p

    @property
    def is_hom(self):
        return self._p.shape[-2] == 4

    @property
    def n_batch(self):
        return self._p.shape[0]

    @property
    def n_point(self):
        return self._p.shape[-1]

    @property
    def device(self):
        return self._p.device

    @property
    def data(self):
        return self._p
