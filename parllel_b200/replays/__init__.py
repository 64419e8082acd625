"""B200-native drop-ins for parllel's replay containers (parllel/replays/__init__.py)."""
from .batched_dataloader import BatchedDataLoader
from .replay import ReplayBuffer

__all__ = ["BatchedDataLoader", "ReplayBuffer"]
