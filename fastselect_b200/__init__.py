"""fastselect_b200 -- B200-native (sm_100a) Relief-family feature scoring.

Drop-in for the GPU backend of ``fast_select.{ReliefF, SURF, MultiSURF, TuRF}``
(reference: src/fast_select/__init__.py:1-10).  The estimators call hand-written CUDA
through the C ABI in ``include/fastselect_b200.h``; nothing here falls back to a CPU.
"""
from ._relief import MultiSURF, ReliefF, SURF
from ._turf import TuRF

__all__ = ["ReliefF", "SURF", "MultiSURF", "TuRF"]
__version__ = "0.1.0"
