"""The path's only collective -- the all_gather of per-rank metric partial sums -- exercised with two
gloo ranks on CPU (the host-side logic of kasportsformer_b200.evaluate; the sums themselves come from
the oracle here since no GPU is present)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kasportsformer_b200.evaluate import finalize_metrics, gather_sums
from kasportsformer_b200 import synthetic
from oracle import metrics_oracle as MO


def _sums_from_oracle(pred, res, factor, gt, actions, n_actions):
    o = MO.evaluate(pred, res, factor, gt, actions=actions)
    pf = o["per_frame"]
    s = np.zeros((n_actions, 22))
    for b, a in enumerate(actions):
        s[a, 0] += pf["mpjpe"][b].sum(); s[a, 1] += pf["p_mpjpe"][b].sum(); s[a, 2] += pf["accel"][b].sum()
        s[a, 3] += pf["mpjpe"].shape[1]; s[a, 4] += pf["accel"].shape[1]; s[a, 5:] += pf["jpe"][b].sum(0)
    return s


def _data():
    B, T = 12, 9
    pred = (synthetic.make_clips(B, T, 4, "gt") * 0.3).numpy().astype(np.float64)
    gt, factor, res, actions = synthetic.make_labels(B, T, 6, n_actions=3)
    return pred, res.numpy(), factor.numpy(), gt.numpy(), actions.numpy()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred, res, factor, gt, actions = _data()
    sl = slice(rank * 6, rank * 6 + 6)                      # contiguous batch shard per rank
    local = _sums_from_oracle(pred[sl], res[sl], factor[sl], gt[sl], actions[sl], 3)
    total = gather_sums(torch.from_numpy(local))
    out[rank] = finalize_metrics(total.numpy())["mpjpe"]
    dist.destroy_process_group()


def test_two_rank_metric_reduction_matches_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    pred, res, factor, gt, actions = _data()
    single = finalize_metrics(_sums_from_oracle(pred, res, factor, gt, actions, 3))
    ref = MO.evaluate(pred, res, factor, gt, actions=actions)
    assert abs(single["mpjpe"] - ref["mpjpe"]) < 1e-9
    assert abs(out[0] - single["mpjpe"]) < 1e-9 and abs(out[1] - single["mpjpe"]) < 1e-9
