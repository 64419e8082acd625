"""parllel_b200 -- B200-native (sm_100a) batch post-processing for parllel.

Drop-in replacements for the one data-parallel hot path of balazsgyenes/parllel: the batch
``Transform`` pipeline (``EstimateAdvantage``, ``NormalizeAdvantage``, ``NormalizeRewards``,
``NormalizeObservations``, ``ClipRewards``) and ``ReplayBuffer.sample_batch()``.  Host code is
Python + PyTorch (device memory, streams, ``torch.distributed``); all arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``include/parllel_b200.h``.
"""
from .arrays import DeviceArray
from .tree import ArrayDict, dict_map
from .spec import BatchSpec

__all__ = ["ArrayDict", "BatchSpec", "DeviceArray", "dict_map"]
__version__ = "0.1.0"
