from .helpers_cam_mat import MatXyzToCamera, ChromacityMat, bradford_adapt_matrix  # noqa: F401
from .cam_wb import CameraWhiteBalance  # noqa: F401
