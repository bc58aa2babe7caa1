"""Variable-length serving front end: lift a whole video's 2D keypoints to 3D.

Restates the clip handling of the reference demo (demo/demo.py:132-156 `resample` / `turn_into_clips`, :220-237 the
per-clip loop, demo/lib/utils.py:5-19 `flip_data` / `normalize_screen_coordinates`) with two deliberate differences:

* execution: all clips of the video and their mirrored copies go through ONE forward as a [2 * n_clips, T, 17, 3]
  batch instead of two forwards per clip;
* results: the test-time augmentation is the INTENDED one, (f(x) + flip(f(flip(x)))) / 2, as in the evaluation scripts
  (train_and_evaluate_sp.py:46-51).  The demo as it actually runs differs: its `flip_data` (demo/lib/utils.py:5-13)
  mutates its argument in place, so at demo.py:221-222 `input_2D` and `input_2D_aug` are the SAME flipped array and
  the demo computes (f(flip(x)) + flip(f(flip(x)))) / 2.  That aliasing bug is not reproduced; pass
  `reference_demo_aliasing=True` to `lift_video` to get the demo's literal behaviour (tests/test_gpu_io.py covers both).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _capi


def resample(n_frames: int, target_frame: int) -> np.ndarray:
    """Indices that stretch `n_frames` frames to `target_frame` (demo/demo.py:132-136)."""
    even = np.linspace(0, n_frames, num=target_frame, endpoint=False)
    return np.clip(np.floor(even), a_min=0, a_max=n_frames - 1).astype(np.uint32)


def turn_into_clips(keypoints: np.ndarray, target_frame_length: int) -> Tuple[List[np.ndarray], Optional[np.ndarray]]:
    """[P, n_frames, 17, C] -> list of [P, T, 17, C] clips; a short last clip (or a short video) is stretched by
    frame repetition and `downsample` holds the positions of its distinct frames (demo/demo.py:139-156).
    `downsample` is None when every clip is full (the reference leaves the name unbound in that case)."""
    clips, downsample = [], None
    n_frames = keypoints.shape[1]
    if n_frames <= target_frame_length:
        idx = resample(n_frames, target_frame_length)
        clips.append(keypoints[:, idx, ...])
        downsample = np.unique(idx, return_index=True)[1]
    else:
        for start in range(0, n_frames, target_frame_length):
            clip = keypoints[:, start:start + target_frame_length, ...]
            if clip.shape[1] != target_frame_length:
                idx = resample(clip.shape[1], target_frame_length)
                clips.append(clip[:, idx, ...])
                downsample = np.unique(idx, return_index=True)[1]
            else:
                clips.append(clip)
    return clips, downsample


def normalize_screen_coordinates(X: np.ndarray, w: float, h: float) -> np.ndarray:
    """Pixels -> [-1, 1] x [-h/w, h/w] (demo/lib/utils.py:15-19); a confidence channel passes through."""
    assert X.shape[-1] in (2, 3)
    out = np.array(X, dtype=np.float32, copy=True)
    out[..., :2] = X[..., :2] / w * 2 - np.asarray([1, h / w], np.float32)
    return out


@torch.no_grad()
def lift_video(model, keypoints: np.ndarray, width: int, height: int, flip: bool = True, return_rep: bool = False,
               max_batch: int = 4096, reference_demo_aliasing: bool = False) -> np.ndarray:
    """2D keypoints of one video [n_frames, 17, 2|3] (pixels [, confidence]) -> root-relative 3D poses
    [n_frames, 17, 3] in the model's normalised units, following demo/demo.py:220-244: split into T-frame clips,
    flip test-time augmentation, average, un-stretch the last clip, zero the root joint.

    With return_rep=True returns the 512-d motion representation [n_frames, 17, 512] instead
    (model/KASportsFormer.py:342-343; averaged over the two flips after mirroring the joints back).
    reference_demo_aliasing: reproduce demo.py:221-233 literally -- its in-place `flip_data` makes BOTH forwards see
    the flipped clip (module docstring); default False = the augmentation the evaluation scripts use."""
    kp = np.asarray(keypoints, np.float32)
    if kp.ndim != 3 or kp.shape[1] != 17 or kp.shape[2] not in (2, 3):
        raise ValueError("keypoints must be [n_frames, 17, 2|3]")
    if kp.shape[2] == 2:
        kp = np.concatenate([kp, np.ones(kp.shape[:2] + (1,), np.float32)], axis=-1)
    T = int(model.cfg["n_frames"])
    clips, downsample = turn_into_clips(kp[None], T)
    x = np.concatenate([normalize_screen_coordinates(c, width, height) for c in clips], axis=0)   # [n_clips,T,17,3]
    dev = next(model.parameters()).device
    outs = []
    for i0 in range(0, x.shape[0], max_batch):
        xb = torch.from_numpy(x[i0:i0 + max_batch]).to(dev)
        n = xb.shape[0]
        if flip:
            xf = _capi.joint_flip(xb)
            yy = model(torch.cat([xf if reference_demo_aliasing else xb, xf], dim=0), return_rep=return_rep)
            y, yf = yy[:n], yy[n:]
            if return_rep:   # mirror the joints back; the representation has no x axis to negate
                perm = torch.tensor(_capi.table(5), device=dev, dtype=torch.long)
                y = (y + yf[:, :, perm]) * 0.5
            else:
                y = (y + _capi.joint_flip(yf)) * 0.5
        else:
            y = model(xb, return_rep=return_rep)
        outs.append(y)
    y = torch.cat(outs, dim=0)
    frames = [y[i] for i in range(y.shape[0])]
    if downsample is not None:
        frames[-1] = frames[-1][torch.from_numpy(np.asarray(downsample, np.int64)).to(dev)]
    out = torch.cat(frames, dim=0)
    if not return_rep:
        out[:, 0, :] = 0
    return out.cpu().numpy()
