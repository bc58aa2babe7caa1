"""CPU oracle for the evaluation epilogue (flip-TTA plumbing, de-normalisation, MPJPE family).

TEST INFRASTRUCTURE ONLY -- see the header of `oracle/kasf_oracle.py` for who may import this.
Parity status: PINNED against the reference's own numpy functions executed in the build container
(`oracle/make_golden.py` -> `tests/golden/metrics_*.npz`).

Follows, in float64 numpy like the reference does after `.cpu().numpy()`:
  * joint_flip ................. reference utils/utilities.py:128-135
  * de-normalise / root-relative  reference train_and_evaluate_sp.py:55-72
  * mpjpe / jpe / accel / p-mpjpe reference utils/error_calc.py:5-48
  * per-action aggregation ...... reference train_and_evaluate_sp.py:85-127
"""
from __future__ import annotations

import numpy as np

FLIP_LEFT = [1, 2, 3, 14, 15, 16]      # reference utils/utilities.py:128
FLIP_RIGHT = [4, 5, 6, 11, 12, 13]


def joint_flip(a: np.ndarray) -> np.ndarray:
    out = np.array(a, copy=True)
    out[..., 0] *= -1
    out[..., FLIP_LEFT + FLIP_RIGHT, :] = out[..., FLIP_RIGHT + FLIP_LEFT, :]
    return out


def denormalise(pred: np.ndarray, res: np.ndarray, factor: np.ndarray, gt: np.ndarray):
    """pred [B,T,17,3] normalised model output; res [B,2]=(w,h); factor [B,T]; gt [B,T,17,3] mm.

    Returns root-relative (pred_mm, gt_mm). reference train_and_evaluate_sp.py:55-72."""
    p = np.array(pred, dtype=np.float64, copy=True)
    p[:, :, 0, :] = 0
    w = res[:, 0].astype(np.float64)[:, None, None]
    h = res[:, 1].astype(np.float64)[:, None, None]
    p[..., 0] = (p[..., 0] + 1.0) * w / 2
    p[..., 1] = (p[..., 1] + h / w) * w / 2
    p[..., 2] = p[..., 2] * w / 2
    p = p * factor.astype(np.float64)[:, :, None, None]
    p = p - p[:, :, 0:1, :]
    g = gt.astype(np.float64)
    g = g - g[:, :, 0:1, :]
    return p, g


def jpe(pred, gt):               # [..,17,3] -> [..,17]      error_calc.py:10-12
    return np.linalg.norm(pred - gt, axis=-1)


def mpjpe(pred, gt):             # -> [..]                   error_calc.py:5-7
    return jpe(pred, gt).mean(-1)


def accel_error(pred, gt):       # [B,T,17,3] -> [B,T-2]     error_calc.py:15-19
    ap = pred[:, :-2] - 2 * pred[:, 1:-1] + pred[:, 2:]
    ag = gt[:, :-2] - 2 * gt[:, 1:-1] + gt[:, 2:]
    return np.linalg.norm(ap - ag, axis=-1).mean(-1)


def p_mpjpe(pred, gt):           # [N,17,3] x2 -> [N]        error_calc.py:21-48
    muX = gt.mean(1, keepdims=True)
    muY = pred.mean(1, keepdims=True)
    X0, Y0 = gt - muX, pred - muY
    nX = np.sqrt((X0 ** 2).sum((1, 2), keepdims=True))
    nY = np.sqrt((Y0 ** 2).sum((1, 2), keepdims=True))
    X0, Y0 = X0 / nX, Y0 / nY
    H = X0.transpose(0, 2, 1) @ Y0
    U, s, Vt = np.linalg.svd(H)
    V = Vt.transpose(0, 2, 1)
    R = V @ U.transpose(0, 2, 1)
    sg = np.sign(np.linalg.det(R))
    V[:, :, -1] *= sg[:, None]
    s[:, -1] *= sg
    R = V @ U.transpose(0, 2, 1)
    a = s.sum(1)[:, None, None] * nX / nY
    t = muX - a * (muY @ R)
    return np.linalg.norm(a * (pred @ R) + t - gt, axis=-1).mean(-1)


def evaluate(pred, res, factor, gt, actions=None):
    """Whole protocol for a batch. Returns dict of scalars + per-joint table, aggregated per
    action then over actions as the reference does (train_and_evaluate_sp.py:105-127)."""
    B, T = pred.shape[:2]
    p, g = denormalise(pred, res, factor, gt)
    e1 = mpjpe(p, g)                                            # [B,T]
    ej = jpe(p, g)                                              # [B,T,17]
    ea = accel_error(p, g)                                      # [B,T-2]
    e2 = p_mpjpe(p.reshape(B * T, 17, 3), g.reshape(B * T, 17, 3)).reshape(B, T)
    actions = np.zeros(B, dtype=np.int64) if actions is None else np.asarray(actions)
    names = sorted(set(actions.tolist()))
    per = {k: [] for k in ("mpjpe", "p_mpjpe", "accel")}
    perj = []
    for a in names:
        m = actions == a
        per["mpjpe"].append(e1[m].mean())
        per["p_mpjpe"].append(e2[m].mean())
        per["accel"].append(ea[m].mean())
        perj.append(ej[m].reshape(-1, 17).mean(0))
    return {
        "mpjpe": float(np.mean(per["mpjpe"])),
        "p_mpjpe": float(np.mean(per["p_mpjpe"])),
        "accel": float(np.mean(per["accel"])),
        "mpjpe_joint": np.mean(np.stack(perj), axis=0),
        "per_frame": {"mpjpe": e1, "p_mpjpe": e2, "accel": ea, "jpe": ej},
    }
