"""B200-native drop-ins for parllel's batch transforms (parllel/transforms/__init__.py)."""
from .advantage import EstimateAdvantage
from .clip_rewards import ClipRewards
from .multi_advantage import EstimateMultiAgentAdvantage
from .norm_advantage import NormalizeAdvantage
from .norm_obs import NormalizeObservations
from .norm_rewards import NormalizeRewards
from .pipeline import FusedBatchTransforms
from .running_mean_std import RunningMeanStd
from .transform import Transform

__all__ = [
    "EstimateAdvantage",
    "EstimateMultiAgentAdvantage",
    "FusedBatchTransforms",
    "ClipRewards",
    "NormalizeAdvantage",
    "NormalizeObservations",
    "NormalizeRewards",
    "RunningMeanStd",
    "Transform",
]
