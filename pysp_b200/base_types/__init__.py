from .image_base import (BayerPattern, RawDemosaicData, RawCameraData_BaseType, RawBayerData_BaseType,  # noqa: F401
                         RawRggbBayerData_BaseType)
