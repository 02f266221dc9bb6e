"""Host-side colour set-up (3x3 matrices, float64 NumPy) for the develop path.

O(1) work per image, evaluated with the reference's own formulas so the device receives bit-identical
constants: colorize/rgb_space.py:19-56 (primaries -> XYZ, Bradford-adapted to the camera white),
wb_cct/helpers_cam_mat.py:7-20 (Bradford), wb_cct/standard_ill.py:33 (D65),
colorize/transform.py:40-49 (detinted camera -> RGB matrix).  `colour.xy_to_XYZ` (colour-science) is
restated as xy -> (x/y, 1, (1-x-y)/y).
"""
import numpy as np

_XY_D65 = (0.31272, 0.32903)
_BRADFORD = np.array([[0.8951000, 0.2664000, -0.1614000],
                      [-0.7502000, 1.7135000, 0.0367000],
                      [0.0389000, -0.0685000, 1.0296000]])


def xy_to_XYZ(xy):
    x, y = float(xy[0]), float(xy[1])
    return np.array([x / y, 1.0, (1.0 - x - y) / y], dtype=np.float64)


def bradford_adapt_matrix(current_xyz, target_xyz):
    """Von-Kries scaling in Bradford LMS taking `current_xyz` white to `target_xyz`."""
    lms_now = np.matmul(_BRADFORD, current_xyz)
    lms_want = np.matmul(_BRADFORD, target_xyz)
    gain = lms_want / lms_now
    scale = np.array([[gain[0], 0, 0], [0, gain[1], 0], [0, 0, gain[2]]])
    return np.matmul(np.linalg.inv(_BRADFORD), np.matmul(scale, _BRADFORD))


class ArbitraryRgbColorspace:
    """Linear RGB space given by xy primaries and a white point."""

    def __init__(self, primary_xy_r, primary_xy_g, primary_xy_b, white_xy=_XY_D65):
        self._prim = (primary_xy_r, primary_xy_g, primary_xy_b)
        self._white = xy_to_XYZ(white_xy)

    def mat_to_xyz(self, destination_whitepoint=None):
        cols = [(p[0] / p[1], 1, (1 - p[0] - p[1]) / p[1]) for p in self._prim]
        m = np.array([[cols[0][0], cols[1][0], cols[2][0]],
                      [cols[0][1], cols[1][1], cols[2][1]],
                      [cols[0][2], cols[1][2], cols[2][2]]])
        s = np.linalg.inv(m) @ self._white
        m[:, 0] *= s[0]
        m[:, 1] *= s[1]
        m[:, 2] *= s[2]
        if destination_whitepoint is None:
            return m
        dest = np.array(destination_whitepoint)
        assert dest.shape[0] == 3 and dest.ndim == 1
        return bradford_adapt_matrix(self._white, dest) @ m

    def mat_to_rgb(self, source_whitepoint=None):
        return np.linalg.inv(self.mat_to_xyz(source_whitepoint))


class LinRgbColorspace:
    REC709 = ArbitraryRgbColorspace((0.64, 0.33), (0.3, 0.6), (0.15, 0.06))
    REC2020 = ArbitraryRgbColorspace((0.708, 0.292), (0.170, 0.797), (0.131, 0.046))


def cam_to_rgb_matrix(cam_xyz_matrix, destination_colorspace=LinRgbColorspace.REC709):
    """Row-major float64 M such that rgb_out = M @ cam_rgb (the reference applies it as
    np.dot(rgb, M.T)): detinted inverse of (XYZ->camera) @ (RGB->XYZ adapted to the camera white)."""
    to_xyz = destination_colorspace.mat_to_xyz(np.asarray(cam_xyz_matrix.xyz).tolist())
    fwd = np.matmul(np.asarray(cam_xyz_matrix.mat), to_xyz)
    fwd = fwd / fwd.sum(axis=1)[:, np.newaxis]
    return np.linalg.inv(fwd)
