from .chan_distortion_corr import apply_opcode_3_warp, get_opcode_3_block, stack_warp_prior  # noqa: F401
from .dng_warp_rectilinear_coords import compute_offset_remapping_table, compute_remapping_table  # noqa: F401
