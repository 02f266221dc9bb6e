"""Quality / pattern selectors (reference: const.py:3-9)."""
from enum import Enum, auto


class QualityDemosaic(Enum):
    Draft = auto()
    Fast = auto()
    Best = auto()


class PatternDemosaic(Enum):
    Rgbg = auto()
