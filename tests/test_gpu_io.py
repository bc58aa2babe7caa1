"""Rows f2 / f3 on the device: evaluation over a packed clip store through the pinned double-buffered feeder, and
the variable-length serving front end against a per-clip restatement of the reference demo loop."""
import numpy as np
from conftest import load_golden
import pytest
import torch

from kasportsformer_b200 import KASportsFormer, _capi, clipstore, serving, synthetic
from kasportsformer_b200.evaluate import evaluate_batch, finalize_metrics

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(T=27, L=2, seed=3, regime="stress"):
    m = KASportsFormer(n_layers=L, n_frames=T, num_heads=8).eval()
    cfg = dict(m.cfg)
    m.load_state_dict(synthetic.make_state(cfg, seed, regime), strict=True)
    return m.to(DEV)


def test_evaluate_store_matches_single_batch_and_shards():
    B, T, A = 70, 27, 4
    m = _model(T)
    x = synthetic.make_clips(B, T, 5, "det")
    gt, factor, res, actions = synthetic.make_labels(B, T, 6, n_actions=A)
    st = clipstore.ClipStore.from_arrays(x.numpy(), gt.numpy(), factor.numpy(), res.numpy(), actions.numpy(),
                                         [f"act{i}" for i in range(A)])
    whole = finalize_metrics(evaluate_batch(m, x.to(DEV), gt.to(DEV), res.to(DEV), factor.to(DEV), actions.to(DEV),
                                            A, flip=True).cpu().numpy())
    fed = clipstore.evaluate_store(m, st, batch_size=16, flip=True)        # 5 batches, last one ragged
    for k in ("mpjpe", "p_mpjpe", "acceleration_error"):
        assert abs(fed[k] - whole[k]) <= 1e-9 * abs(whole[k]), k
    assert fed["activity_names"] == [f"act{i}" for i in range(A)]
    # two shards evaluated separately add up to the same table (what the all_gather combines)
    from kasportsformer_b200 import evaluate as E
    tot = None
    for r in range(2):
        sums = torch.zeros(A, E.COLS, dtype=torch.float64, device=DEV)
        for xb, gb, rb, fb, ab in clipstore.ClipFeeder(st.shard(r, 2), 16, torch.device(DEV)):
            sums = evaluate_batch(m, xb, gb, rb, fb, ab, A, flip=True, sums=sums)
        tot = sums if tot is None else tot + sums
    two = finalize_metrics(tot.cpu().numpy())
    assert abs(two["mpjpe"] - whole["mpjpe"]) <= 1e-9 * whole["mpjpe"]


@pytest.mark.parametrize("n_frames", [11, 27, 40, 81])
def test_lift_video_matches_per_clip_demo_loop(n_frames):
    T = 27
    m = _model(T)
    rng = np.random.default_rng(n_frames)
    kp = (rng.random((n_frames, 17, 3)) * np.asarray([1920, 1080, 1.0])).astype(np.float32)
    out = serving.lift_video(m, kp, 1920, 1080, flip=True)
    assert out.shape == (n_frames, 17, 3) and np.all(out[:, 0] == 0)
    # the reference demo loop (demo/demo.py:220-244), clip by clip, two forwards per clip
    clips, down = serving.turn_into_clips(kp[None], T)
    ref = []
    for i, c in enumerate(clips):
        x = torch.from_numpy(serving.normalize_screen_coordinates(c, 1920, 1080)).to(DEV)
        y = (m(x) + _capi.joint_flip(m(_capi.joint_flip(x)))) / 2
        if i == len(clips) - 1 and down is not None:
            y = y[:, torch.from_numpy(np.asarray(down, np.int64)).to(DEV)]
        y[:, :, 0, :] = 0
        ref.append(y[0].cpu().numpy())
    ref = np.concatenate(ref, axis=0)
    assert ref.shape == out.shape
    assert np.abs(out - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    rep = serving.lift_video(m, kp, 1920, 1080, flip=False, return_rep=True)
    assert rep.shape == (n_frames, 17, 512) and np.isfinite(rep).all()
    # the demo as it actually runs: flip_data (demo/lib/utils.py:5-13) flips in place, so demo.py:232-233 feed the
    # flipped clip to both forwards; reproduced only on request
    z, _ = load_golden("serving.npz")
    lit = serving.lift_video(m, kp, 1920, 1080, flip=True, reference_demo_aliasing=True)
    ref2 = []
    for i, c in enumerate(clips):
        x = serving.normalize_screen_coordinates(c, 1920, 1080)
        xa = x.copy()
        xa[..., 0] *= -1                                   # flip_data, restated: negate x, swap left / right joints
        xa[:, :, [1, 2, 3, 14, 15, 16] + [4, 5, 6, 11, 12, 13]] = xa[:, :, [4, 5, 6, 11, 12, 13] + [1, 2, 3, 14, 15, 16]]
        x = xa                                             # the aliasing: `input_2D` IS `input_2D_aug` after the call
        xt = torch.from_numpy(x).to(DEV)
        y = (m(xt) + _capi.joint_flip(m(torch.from_numpy(xa).to(DEV)))) / 2
        if i == len(clips) - 1 and down is not None:
            y = y[:, torch.from_numpy(np.asarray(down, np.int64)).to(DEV)]
        y[:, :, 0, :] = 0
        ref2.append(y[0].cpu().numpy())
    ref2 = np.concatenate(ref2, axis=0)
    assert np.abs(lit - ref2).max() <= 1e-6 * max(1.0, np.abs(ref2).max())
    assert np.abs(lit - out).max() > 1e-4                 # and it is a different result
    x0 = z["flip_in"]
    assert np.array_equal(_capi.joint_flip(torch.from_numpy(x0).to(DEV)).cpu().numpy(), z["flip_out"])
