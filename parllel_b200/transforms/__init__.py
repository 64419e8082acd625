"""Batch post-processing operators of parllel, B200-native.

The public names match ``parllel.transforms`` so scripts only change the import path; two names are
additions: ``FusedBatchTransforms`` (runs a sampler's transform list as one fused chain) and the
device-resident ``RunningMeanStd``.
"""
from . import (advantage as _advantage, base as _base, clip_rewards as _clip, multi_advantage as _multi,
               norm_advantage as _norm_adv, norm_obs as _norm_obs, norm_rewards as _norm_rew,
               pipeline as _pipeline, running_mean_std as _rms)

Transform = _base.Transform
EstimateAdvantage = _advantage.EstimateAdvantage
EstimateMultiAgentAdvantage = _multi.EstimateMultiAgentAdvantage
NormalizeAdvantage = _norm_adv.NormalizeAdvantage
NormalizeObservations = _norm_obs.NormalizeObservations
NormalizeRewards = _norm_rew.NormalizeRewards
ClipRewards = _clip.ClipRewards
RunningMeanStd = _rms.RunningMeanStd
FusedBatchTransforms = _pipeline.FusedBatchTransforms

__all__ = sorted(name for name, obj in globals().items() if isinstance(obj, type) and not name.startswith("_"))
