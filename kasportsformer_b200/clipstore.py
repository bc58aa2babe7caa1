"""Packed clip store + feeder for evaluation at B200 rates.

The reference keeps every test clip in its own pickle (data/preprocessor/clip_generate_sp.py:48-79: keys
`data_input` [T,17,3], `data_label_scaled` [T,17,3], `data_factor` [T], `data_res` (w, h), `data_action`,
`data_env`) and reads them with 19 DataLoader workers (data/reader/sp_dataset.py:30-92).  Unpickling 2 KB files
cannot feed 10^4 clips/s per GPU, so this module

  * packs a clip directory once into a *shard*: a directory of plain `.npy` arrays (`input` f32 [N,T,17,3],
    `gt` f32 [N,T,17,3], `factor` f32 [N,T], `res` f32 [N,2], `action` i32 [N]) + `meta.json` (action names in
    first-appearance order, T, N) -- memory-mappable, no parsing at read time (`pack_clip_dir`);
  * serves contiguous rank-local ranges of a shard (`ClipStore.shard`, the batch sharding of DESIGN.md section 6);
  * streams batches to the device through two pinned staging buffers and a copy stream, so the H2D copy of batch
    i+1 overlaps the forward of batch i (`ClipFeeder`);
  * runs the reference's evaluation protocol over a store (`evaluate_store`): flip-TTA forward, de-normalisation,
    MPJPE / P-MPJPE / acceleration error per action on the device, one all_gather of the per-action sums.

Batches keep the reference's order (sorted file names, sp_dataset.py:24), so results are reproducible per clip.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch

ARRAYS = ("input", "gt", "factor", "res", "action")


def pack_clip_dir(clip_dir: str, out_dir: str, action_names: Optional[List[str]] = None) -> Dict[str, object]:
    """Pack the reference's one-pickle-per-clip test directory into a shard under `out_dir`.

    Clips are taken in sorted file-name order like sp_dataset.py:21-27.  `action_names` fixes the action ->
    index mapping (default: first appearance)."""
    files = sorted(f for f in os.listdir(clip_dir) if f.endswith(".pkl"))
    if not files:
        raise ValueError(f"no .pkl clips under {clip_dir}")
    names = list(action_names) if action_names is not None else []
    xs, gts, fs, rs, acts = [], [], [], [], []
    for f in files:
        with open(os.path.join(clip_dir, f), "rb") as fh:
            d = pickle.load(fh)
        a = str(d["data_action"])
        if a not in names:
            if action_names is not None:
                raise ValueError(f"{f}: action {a!r} not in the given action_names")
            names.append(a)
        xs.append(np.asarray(d["data_input"], np.float32))
        gts.append(np.asarray(d["data_label_scaled"], np.float32))
        fs.append(np.asarray(d["data_factor"], np.float32).reshape(-1))
        rs.append(np.asarray(d["data_res"], np.float32).reshape(2))
        acts.append(names.index(a))
    x = np.stack(xs)
    if x.ndim != 4 or x.shape[2] != 17 or x.shape[3] not in (2, 3):
        raise ValueError(f"clips must be [T,17,2|3], got {x.shape[1:]}")
    if x.shape[3] == 2:     # gt configs store x,y only on some dumps: confidence 1.0 (data/reader/sp_reader.py:52-55)
        x = np.concatenate([x, np.ones(x.shape[:3] + (1,), np.float32)], axis=-1)
    os.makedirs(out_dir, exist_ok=True)
    arrs = {"input": x, "gt": np.stack(gts), "factor": np.stack(fs), "res": np.stack(rs),
            "action": np.asarray(acts, np.int32)}
    for k, v in arrs.items():
        np.save(os.path.join(out_dir, k + ".npy"), np.ascontiguousarray(v))
    meta = {"n_clips": int(x.shape[0]), "n_frames": int(x.shape[1]), "action_names": names, "source": clip_dir}
    with open(os.path.join(out_dir, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    return meta


class ClipStore:
    """Memory-mapped shard written by `pack_clip_dir` (or built from arrays with `from_arrays`)."""

    def __init__(self, path: Optional[str] = None, arrays: Optional[Dict[str, np.ndarray]] = None,
                 action_names: Optional[List[str]] = None, lo: int = 0, hi: Optional[int] = None):
        if path is not None:
            with open(os.path.join(path, "meta.json")) as fh:
                meta = json.load(fh)
            arrays = {k: np.load(os.path.join(path, k + ".npy"), mmap_mode="r") for k in ARRAYS}
            action_names = meta["action_names"]
        assert arrays is not None and action_names is not None
        n = arrays["input"].shape[0]
        for k in ARRAYS:
            if arrays[k].shape[0] != n:
                raise ValueError(f"array {k!r} has {arrays[k].shape[0]} clips, expected {n}")
        self.arrays, self.action_names = arrays, list(action_names)
        self.lo, self.hi = lo, n if hi is None else hi

    @classmethod
    def from_arrays(cls, input, gt, factor, res, action, action_names) -> "ClipStore":
        return cls(arrays={"input": np.asarray(input, np.float32), "gt": np.asarray(gt, np.float32),
                           "factor": np.asarray(factor, np.float32), "res": np.asarray(res, np.float32),
                           "action": np.asarray(action, np.int32)}, action_names=action_names)

    def __len__(self) -> int:
        return self.hi - self.lo

    @property
    def n_frames(self) -> int:
        return int(self.arrays["input"].shape[1])

    @property
    def n_actions(self) -> int:
        return len(self.action_names)

    def shard(self, rank: int, world: int) -> "ClipStore":
        """Contiguous range of this store owned by `rank` of `world` (sizes differ by at most one clip)."""
        n = len(self)
        a = self.lo + (n * rank) // world
        b = self.lo + (n * (rank + 1)) // world
        return ClipStore(arrays=self.arrays, action_names=self.action_names, lo=a, hi=b)

    def batch(self, i0: int, i1: int) -> Tuple[np.ndarray, ...]:
        """Clips [i0, i1) of this (sharded) store as numpy views, in ARRAYS order."""
        a, b = self.lo + i0, min(self.lo + i1, self.hi)
        return tuple(self.arrays[k][a:b] for k in ARRAYS)


class ClipFeeder:
    """Iterates a ClipStore in batches of device tensors (x, gt, res, factor, action).

    On a CUDA device every batch is copied from the memory map into one of two pinned staging sets and sent with
    non-blocking copies on a private stream; the consumer's stream waits on the copy's event, and a staging set is
    reused only after the copies that read it have completed.  The copy of batch i+1 is issued before batch i is
    handed out, so it overlaps the consumer's kernels.  On a CPU device (tests) batches are plain tensors."""

    def __init__(self, store: ClipStore, batch_size: int, device: torch.device):
        self.store, self.bs, self.dev = store, int(batch_size), torch.device(device)
        self.cuda = self.dev.type == "cuda"
        if self.cuda:
            self.stream = torch.cuda.Stream(self.dev)
            T = store.n_frames
            shapes = {"input": (self.bs, T, 17, 3), "gt": (self.bs, T, 17, 3), "factor": (self.bs, T),
                      "res": (self.bs, 2), "action": (self.bs,)}
            self.stage = [{k: torch.empty(shapes[k], dtype=torch.int32 if k == "action" else torch.float32,
                                          pin_memory=True) for k in ARRAYS} for _ in range(2)]
            self.stage_free = [None, None]      # event: the H2D copies out of staging set s have completed

    def __len__(self) -> int:
        return (len(self.store) + self.bs - 1) // self.bs

    def _issue(self, b: int, slot: int):
        arrs = self.store.batch(b * self.bs, (b + 1) * self.bs)
        n = arrs[0].shape[0]
        if not self.cuda:
            t = [torch.from_numpy(np.ascontiguousarray(a)) for a in arrs]
            return (t[0], t[1], t[3], t[2], t[4]), None
        if self.stage_free[slot] is not None:
            self.stage_free[slot].synchronize()
        out = {}
        with torch.cuda.stream(self.stream):
            for k, a in zip(ARRAYS, arrs):
                self.stage[slot][k][:n].numpy()[...] = a          # page-in + memcpy into pinned memory
                out[k] = self.stage[slot][k][:n].to(self.dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.stage_free[slot] = ev
        return (out["input"], out["gt"], out["res"], out["factor"], out["action"]), ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        nb = len(self)
        if nb == 0:
            return
        nxt = self._issue(0, 0)
        for b in range(nb):
            cur = nxt
            if b + 1 < nb:
                nxt = self._issue(b + 1, (b + 1) & 1)
            tensors, ev = cur
            if ev is not None:
                torch.cuda.current_stream(self.dev).wait_event(ev)
                for t in tensors:
                    t.record_stream(torch.cuda.current_stream(self.dev))
            yield tensors


@torch.no_grad()
def evaluate_store(model, store: ClipStore, batch_size: int = 1024, flip: bool = True, device=None, group=None):
    """The reference's `evaluate` protocol (train_and_evaluate_sp.py:30-127) over this rank's part of a store.

    Call with `store.shard(rank, world)` under torch.distributed: the per-action sums of all ranks are combined with
    one all_gather (evaluate.gather_sums) and every rank returns the global result dict."""
    from . import evaluate as E
    device = torch.device(device) if device is not None else next(model.parameters()).device
    sums = torch.zeros(store.n_actions, E.COLS, dtype=torch.float64, device=device)
    for x, gt, res, factor, action in ClipFeeder(store, batch_size, device):
        sums = E.evaluate_batch(model, x, gt, res, factor, action, store.n_actions, flip=flip, sums=sums)
    total = E.gather_sums(sums, group)
    out = E.finalize_metrics(total.cpu().numpy())
    out["activity_names"] = [store.action_names[i] for i in out["activity_index"]]
    return out
