"""GPU-resident evaluation epilogue + the one collective of the path.

Replaces the per-clip numpy loop of the reference (train_and_evaluate_sp.py:40-127): predictions stay
on the device, `kasf_metrics` accumulates per-action partial sums in fp64, and -- when several ranks
each evaluate a shard of the clips -- one `all_gather` of the [n_actions, 22] table (a few hundred
bytes; NCCL over NVLink on GPUs, gloo in the CPU tests) is the only communication of the whole path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _capi

COLS = _capi.METRIC_COLS


def finalize_metrics(sums: np.ndarray, expect_frames: Optional[int] = None) -> Dict[str, object]:
    """Per-action sums [A,22] -> the reference's result dict (train_and_evaluate_sp.py:105-127):
    mean over frames per action, then mean over the actions that occurred.  expect_frames: the number of frames that
    were fed; kasf_metrics skips clips whose action index is out of range, which shows up here as missing frames."""
    sums = np.asarray(sums, dtype=np.float64)
    if expect_frames is not None and int(round(sums[:, 3].sum())) != int(expect_frames):
        raise ValueError(f"{int(round(sums[:, 3].sum()))} frames accumulated, {expect_frames} expected: "
                         "an action index was outside [0, n_actions)")
    seen = sums[:, 3] > 0
    s = sums[seen]
    mp = s[:, 0] / s[:, 3]
    pm = s[:, 1] / s[:, 3]
    ac = s[:, 2] / np.maximum(s[:, 4], 1.0)
    jt = s[:, 5:22] / s[:, 3:4]
    return {"mpjpe": float(mp.mean()), "p_mpjpe": float(pm.mean()), "acceleration_error": float(ac.mean()),
            "mpjpe_activity": mp.tolist(), "mpjpe_joint": jt.mean(axis=0), "activity_index": np.nonzero(seen)[0].tolist()}


def gather_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather the per-rank [A,22] tables and add them (every rank gets the global table)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    parts = [torch.empty_like(sums) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, sums.contiguous(), group=group)
    return torch.stack(parts).sum(dim=0)


@torch.no_grad()
def evaluate_batch(model, x: torch.Tensor, gt: torch.Tensor, res: torch.Tensor, factor: torch.Tensor,
                   actions: torch.Tensor, n_actions: int, flip: bool = True,
                   sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One batch of the reference eval loop on the device: (flip-TTA) forward + metric partial sums.
    With flip=True the flipped clips ride in the same forward as a 2B batch (…_sp.py:46-51)."""
    if flip:
        xx = torch.cat([x, _capi.joint_flip(x)], dim=0)
        yy = model(xx)
        y, yf = yy[: x.shape[0]], yy[x.shape[0]:]
    else:
        y, yf = model(x), None
    return _capi.metrics(y, gt, res, factor, actions, n_actions, pred_flip=yf, sums=sums)
