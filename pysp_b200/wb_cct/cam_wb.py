"""White-balance controller: the three accessors the develop path uses.

Reference: CameraWhiteBalanceController.get_reciprocal_multipliers / get_matrix / copy
(wb_cct/cam_wb.py:236-260).  The CCT/Duv solver of the reference (cam_wb.py:42-234, needs
colour-science) is O(1) host work outside the hot path and is out of scope: this
controller is constructed from an already-chosen XYZ->camera matrix, its white and the neutral.
Any object with the same three methods (e.g. the reference's own controller) is accepted wherever a
`cam_wb` is expected.
"""
import numpy as np

from .helpers_cam_mat import MatXyzToCamera


class CameraWhiteBalance:
    def __init__(self, mat_xyz_to_cam, white_xyz, neutral=None):
        mat = np.asarray(mat_xyz_to_cam)
        xyz = np.asarray(white_xyz, dtype=np.float64)
        self._mat = MatXyzToCamera(mat, xyz)
        if neutral is None:
            neutral = mat.astype(np.float64) @ xyz
        # float32, as produced on the reference's EXIF path (wb_cct/helpers_exif.py:79)
        self._neutral = np.asarray(neutral, dtype=np.float32)

    def get_reciprocal_multipliers(self):
        return np.copy(1.0 / self._neutral)

    def get_matrix(self):
        return self._mat

    def copy(self):
        return CameraWhiteBalance(self._mat.mat, self._mat.xyz, self._neutral)
