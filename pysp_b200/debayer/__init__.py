from .ahd import debayer as debayer_ahd  # noqa: F401
from .edge_assisted_gaussian import debayer as debayer_eag  # noqa: F401
