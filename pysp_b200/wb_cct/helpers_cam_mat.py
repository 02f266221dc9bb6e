"""Camera-matrix containers (reference: wb_cct/helpers_cam_mat.py:22-38)."""
import numpy as np

from ..colour import bradford_adapt_matrix  # noqa: F401  (re-exported under the reference's name)


class ChromacityMat:
    def __init__(self, mat, xyz):
        self.mat = np.array(mat, copy=True)
        self.mat.setflags(write=False)
        self.xyz = np.array(xyz, copy=True)
        self.xyz.setflags(write=False)


class MatXyzToCamera(ChromacityMat):
    """XYZ -> camera matrix with the XYZ white it was optimised for."""

    def __init__(self, mat, xyz, series=None):
        super().__init__(mat, xyz)
        self.series = series

    def interpolate(self, next, blend):
        t = np.clip(blend, 0.0, 1.0)
        return self.mat * (1 - t) + (next.mat * t)
