"""Minibatch sources of the algorithms, B200-native: the SAC replay ring (``ReplayBuffer``) and PPO's
shuffled epoch loader (``BatchedDataLoader``), both gathering with one CUDA launch over all leaves."""
from . import batched_dataloader as _loader, replay as _replay

ReplayBuffer = _replay.ReplayBuffer
BatchedDataLoader = _loader.BatchedDataLoader

__all__ = ["BatchedDataLoader", "ReplayBuffer"]
