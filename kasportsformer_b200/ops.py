"""torch custom ops over the C-ABI: `torch.ops.kasf.forward` and `torch.ops.kasf.metrics`.

The drop-in nn.Module calls the CUDA library through these registered operators (not through bare ctypes calls), so
the dispatcher knows them: they show up in profiler traces and `torch.library.opcheck`, take part in
`torch.compile` / export graphs as opaque calls with a shape-only fake implementation, and refuse CPU tensors with the
library's own "no CPU path" error.  Operands are tensors and plain ints only; the packed weight blob (and, for
precision "exact", the fp32 weight image) are ordinary uint8 / float32 device tensors owned by the module.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi


@torch.library.custom_op("kasf::forward", mutates_args=(), device_types="cuda")
def kasf_forward(x: torch.Tensor, blob: torch.Tensor, image: Optional[torch.Tensor], n_layers: int, n_frames: int,
                 return_rep: bool, precision: int, flags: int, num_heads: int = 8) -> torch.Tensor:
    """KASportsFormer.forward (reference model/KASportsFormer.py:320-347) -> kasf_forward_ex.
    x float32 [B, n_frames, 17, 3] contiguous; returns [B, n_frames, 17, 3] (or [.., 512] with return_rep)."""
    cfg = dict(n_layers=n_layers, n_frames=n_frames, dim_feat=128, dim_rep=512, num_heads=num_heads, mlp_ratio=4,
               num_joints=17, neighbour_num=4)
    return _capi.forward(cfg, blob, x, return_rep, precision="exact" if precision == 1 else "fast", image=image,
                         two_tiles=bool(flags & _capi.FLAG_TWO_TILES))


@kasf_forward.register_fake
def _(x, blob, image, n_layers, n_frames, return_rep, precision, flags, num_heads=8):
    return x.new_empty(x.shape[0], x.shape[1], 17, 512 if return_rep else 3)


@torch.library.custom_op("kasf::metrics", mutates_args=(), device_types="cuda")
def kasf_metrics(pred: torch.Tensor, pred_flip: Optional[torch.Tensor], gt: torch.Tensor, res: torch.Tensor,
                 factor: torch.Tensor, actions: torch.Tensor, n_actions: int) -> torch.Tensor:
    """The evaluation epilogue (reference train_and_evaluate_sp.py:46-103, utils/error_calc.py) -> kasf_metrics:
    per-action partial sums, float64 [n_actions, 22]."""
    return _capi.metrics(pred, gt, res, factor, actions, n_actions, pred_flip=pred_flip)


@kasf_metrics.register_fake
def _(pred, pred_flip, gt, res, factor, actions, n_actions):
    return pred.new_empty(n_actions, _capi.METRIC_COLS, dtype=torch.float64)
