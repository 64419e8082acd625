"""Batch diagnostics the learners log next to the transformed batch, as device reductions.

* ``explained_variance``  -- parllel/torch/utils.py:135-156
* ``valid_mask``          -- parllel/samplers/recurrent.py:102-103,141-142

Both run through the C ABI (``plb_batch_moments_f32`` / ``plb_residual_moments_f32`` /
``plb_valid_prefix_u8``); there is no CPU path.

* ``valid_mean``          -- parllel/torch/utils.py:42-62.  The one exception: the learners call it on loss terms, so
  it has to stay inside autograd; it is a handful of torch reductions on whatever device the loss lives on, kept here
  so that code which imports it from parllel's utilities finds it next to ``explained_variance``.
"""
from __future__ import annotations

import torch

from . import _lib
from ._leaves import rows_2d, stream_ptr

_scratch: dict = {}


def _buffers(device: torch.device):
    buf = _scratch.get(device)
    if buf is None:
        lib = _lib.load()
        buf = (torch.zeros(lib.plb_batch_moments_workspace_bytes(), dtype=torch.uint8, device=device),
               torch.zeros(6, dtype=torch.float64, device=device))
        _scratch[device] = buf
    return buf


def explained_variance(y_pred: torch.Tensor, y_true: torch.Tensor) -> float:
    """``1 - Var[y_true - y_pred] / Var[y_true]`` (``nan`` when ``Var[y_true]`` is zero), as a Python
    float like the reference (utils.py:151-156).  Both variances use the same denominator, so the ratio
    is the ratio of the centred sums of squares, reduced in float64 on the device in two launches."""
    assert y_true.ndim == y_pred.ndim
    if not (y_true.is_cuda and y_pred.is_cuda):
        raise TypeError("explained_variance needs CUDA tensors")
    if y_true.shape != y_pred.shape:
        y_pred, y_true = torch.broadcast_tensors(y_pred, y_true)
    yt = y_true.detach().to(torch.float32).contiguous().reshape(1, -1)
    yp = y_pred.detach().to(torch.float32).contiguous().reshape(1, -1)
    lib = _lib.load()
    ws, out = _buffers(yt.device)
    s = stream_ptr(yt.device)
    n = yt.shape[1]
    _lib.check(lib.plb_batch_moments_f32(yt.data_ptr(), n, 1, n, 0, 0, out.data_ptr(), ws.data_ptr(), ws.numel(), s),
               "plb_batch_moments_f32")
    _lib.check(lib.plb_residual_moments_f32(yt.data_ptr(), n, yp.data_ptr(), n, 1, n, 0, 0, out.data_ptr() + 24,
                                            ws.data_ptr(), ws.numel(), s), "plb_residual_moments_f32")
    count, _, m2_true, _, _, m2_res = out.cpu().tolist()
    var_true = m2_true / (count - 1.0) if count > 1.0 else float("nan")
    if var_true != var_true or abs(var_true) <= 1e-8:  # torch.allclose(var_y, 0): atol 1e-8
        return float("nan")
    return 1.0 - m2_res / m2_true


def valid_mask(done: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """``valid[0] = True; valid[t+1] = valid[t] & ~done[t]`` over a ``[T, B]`` batch of done flags."""
    if not done.is_cuda or done.dtype not in (torch.bool, torch.uint8) or done.dim() != 2:
        raise TypeError("valid_mask needs a [T, B] CUDA bool tensor")
    T, B = done.shape
    if out is None:
        out = torch.empty((T, B), dtype=torch.bool, device=done.device)
    dptr, dld, _ = rows_2d(done)
    optr, old, _ = rows_2d(out)
    _lib.check(_lib.load().plb_valid_prefix_u8(dptr, dld, T, B, optr, old, stream_ptr(done.device)),
               "plb_valid_prefix_u8")
    return out


def valid_mean(tensor: torch.Tensor, valid: torch.Tensor | None = None, dim=None) -> torch.Tensor:
    """Masked mean of a loss term, differentiable (parllel/torch/utils.py:42-62, same three cases):

    * no mask: the plain mean (over ``dim`` if given, else over everything);
    * mask, no ``dim``: the mean of the selected elements; a ``[T, B]`` mask selects whole trailing slices of a
      ``[T, B, ...]`` tensor;
    * mask and ``dim``: the mask is extended over the trailing axes, masked-out entries count as zero, and the sum
      along ``dim`` is divided by the number of NON-ZERO entries left -- the reference counts with
      ``count_nonzero`` on the masked values, so a valid element that happens to be exactly 0 is not counted; kept.
    """
    if valid is None:
        return tensor.mean(dim=() if dim is None else dim)
    if dim is None:
        return tensor[valid].mean()
    mask = valid.reshape(valid.shape + (1,) * (tensor.ndim - valid.ndim))
    kept = torch.where(mask, tensor, torch.zeros((), dtype=tensor.dtype, device=tensor.device))
    return kept.sum(dim=dim) / kept.count_nonzero(dim=dim)
