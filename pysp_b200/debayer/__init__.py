from .ahd import debayer as debayer_ahd  # noqa: F401
