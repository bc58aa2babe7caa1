"""CPU restatement of the reference's evaluation loop AROUND the model -- the caller of the hot path.

TEST INFRASTRUCTURE (like everything under oracle/): only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.  Follows train_and_evaluate_sp.py:27-127 (`evaluate_one_epoch_new`) statement by statement: how the
reference drives `model(...)` (eval mode, no_grad, two forwards for the flip TTA, in-place zeroing of joint 0 on the
model's output, `.cpu().numpy()`), the per-clip de-normalisation in float32 (:63-72), the numpy metric functions of
utils/error_calc.py (restated in metrics_oracle.py) and the per-action aggregation (:105-127).

Pinned: tests/test_oracle_golden.py replays the predictions recorded in tests/golden/metrics.npz through this loop and
compares with the numbers the UNMODIFIED reference loop produced from them (oracle/make_golden.py), and
tests/golden/eval_loop_real_model.npz holds the reference loop's results with the real reference model.
"""
from __future__ import annotations

import numpy as np
import torch

from . import metrics_oracle as MO


def joint_flip(x: torch.Tensor) -> torch.Tensor:
    """utils/utilities.py:128-135 (torch version of the reference helper: clone, negate x, swap left / right joints)."""
    left, right = [1, 2, 3, 14, 15, 16], [4, 5, 6, 11, 12, 13]
    y = x.clone()
    y[..., 0] *= -1
    y[..., left + right, :] = y[..., right + left, :]
    return y


def evaluate_loop(model, loader, device, flip: bool, num_joints: int = 17):
    """train_and_evaluate_sp.py:27-127.  loader yields (joint_input, joint_label_scaled, joint_factor, joint_action,
    joint_res) batches like the reference's DataLoader."""
    model.eval()                                                              # :30
    per_action = {}
    with torch.no_grad():                                                     # :39
        for joint_input, label, factor, actions, res in loader:               # :40
            joint_input = joint_input.to(device)                              # :42
            if flip:                                                          # :46-51
                pred = (model(joint_input) + joint_flip(model(joint_flip(joint_input)))) / 2
            else:
                pred = model(joint_input)                                     # :53
            pred[:, :, 0, :] = 0                                              # :55 (in place, on the model's output)
            pred = pred.cpu().numpy()                                         # :57-60
            label, factor, res = label.cpu().numpy(), factor.cpu().numpy(), res.cpu().numpy()
            for i, d in enumerate(pred):                                      # :62-72 (float32, like the reference)
                w, h = res[i]
                d[:, :, :2] = (d[:, :, :2] + np.array([1, h / w])) * w / 2
                d[:, :, 2:] = d[:, :, 2:] * w / 2
                d *= factor[i][:, None, None]
                d = d - d[:, 0:1, :]
                g = label[i] - label[i][:, 0:1, :]
                a = per_action.setdefault(actions[i], {"mpjpe": [], "p": [], "acc": [], "jpe": []})
                a["mpjpe"].extend(MO.mpjpe(d[None], g[None])[0])              # :74-81, utils/error_calc.py
                a["jpe"].append(MO.jpe(d[None], g[None])[0])
                a["acc"].extend(MO.accel_error(d[None], g[None])[0])
                a["p"].extend(MO.p_mpjpe(d.astype(np.float64), g.astype(np.float64)))
    names = list(per_action)                                                  # :105-127
    return {"mpjpe": float(np.mean([np.mean(per_action[n]["mpjpe"]) for n in names])),
            "p_mpjpe": float(np.mean([np.mean(per_action[n]["p"]) for n in names])),
            "acceleration_error": float(np.mean([np.mean(per_action[n]["acc"]) for n in names])),
            "mpjpe_joint": np.mean([np.concatenate(per_action[n]["jpe"]).mean(0) for n in names], axis=0),
            "activity_name_sequence": names}
