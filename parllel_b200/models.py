"""Device-side replacement for ``RunningMeanStdModel`` (parllel/torch/models.py:193-230).

Same interface (``nn.Module`` with ``mean`` / ``var`` / ``count`` buffers, ``update(x)``), but the batch
moments come from the column-moments kernel and, under ``torch.distributed``, the ranks' moments are
all-gathered and merged exactly in rank order.  The reference averages the per-rank means AND variances
(``all_reduce`` of ``[mean, var]`` divided by the world size, :212-218), which drops the between-rank
variance term; with one process the two agree.
"""
from __future__ import annotations

import torch
from torch import nn

from .transforms.running_mean_std import RunningMeanStd


class RunningMeanStdModel(nn.Module):
    def __init__(self, shape, device=None) -> None:
        super().__init__()
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        self._stats = RunningMeanStd(self.shape, initial_count=0.0, device=device)
        dev = self._stats.device
        self.register_buffer("mean", torch.zeros(self.shape, device=dev))
        self.register_buffer("var", torch.ones(self.shape, device=dev))
        self.register_buffer("count", torch.zeros((), device=dev))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs) -> None:
        """Checkpoint restore: the registered buffers are what a state dict carries (as in the reference,
        where they ARE the state), so after loading them the float64 working statistics are rebuilt from
        them -- otherwise the next ``update()`` would merge into fresh statistics and overwrite the restore."""
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._stats.load_state(self.mean.detach().double().cpu().numpy(), self.var.detach().double().cpu().numpy(),
                               float(self.count.item()))

    def update(self, x: torch.Tensor) -> None:
        """``x``: ``[T, B, *shape]``, ``[B, *shape]`` or ``[*shape]`` float32 CUDA tensor (models.py:207-209)."""
        import torch.distributed as dist
        lead = x.dim() - len(self.shape)
        assert 0 <= lead <= 2 and tuple(x.shape[lead:]) == self.shape, f"bad input shape {tuple(x.shape)}"
        rows = x.reshape((-1,) + self.shape) if self.shape else x.reshape(-1, 1)
        stats = self._stats
        stats.sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.shape:
            stats.update(rows.contiguous(), f32_flow=False)
        else:
            stats.update(rows.contiguous().reshape(1, -1), f32_flow=False)
        n = stats._n
        self.mean.copy_(stats.mean_tensor[:n].reshape(self.shape))
        self.var.copy_(stats.var_tensor[:n].reshape(self.shape))
        self.count.copy_(stats.count_tensor[0])
