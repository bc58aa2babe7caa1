"""K8 (GPU evaluation epilogue) vs the metrics oracle and the reference eval-loop golden."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from kasportsformer_b200 import _capi, synthetic
from kasportsformer_b200.evaluate import finalize_metrics
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_joint_flip():
    z, _ = load_golden("metrics.npz")
    out = _capi.joint_flip(torch.from_numpy(z["pred"]).to(DEV)).cpu().numpy()
    assert np.array_equal(out, z["flip_of_pred"])


@pytest.mark.parametrize("flip", [False, True])
def test_metrics_vs_reference_eval_loop(flip):
    """mm-scale metrics within 0.01 mm of the unmodified reference loop (which works in float32)."""
    z, _ = load_golden("metrics.npz")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    sums, per = _capi.metrics(t(z["pred"]), t(z["gt"]), t(z["res"]), t(z["factor"]), t(z["actions"]), 3,
                              pred_flip=t(z["pred_flip"]) if flip else None, want_per_frame=True)
    r = finalize_metrics(sums.cpu().numpy())
    ref = z["eval_flip" if flip else "eval_noflip"]
    assert abs(r["mpjpe"] - ref[0]) <= 0.01 and abs(r["p_mpjpe"] - ref[1]) <= 0.01
    assert abs(r["acceleration_error"] - ref[2]) <= 0.01
    np.testing.assert_allclose(r["mpjpe_joint"], z[("eval_flip" if flip else "eval_noflip") + "_joint"], atol=0.01)
    # per-frame values against the float64 oracle: tight
    pred = z["pred"].astype(np.float64)
    if flip:
        pred = ((z["pred"] + MO.joint_flip(z["pred_flip"])) / np.float32(2)).astype(np.float64)
    o = MO.evaluate(pred, z["res"], z["factor"], z["gt"], actions=z["actions"])
    pf = per.cpu().numpy()
    np.testing.assert_allclose(pf[..., 0], o["per_frame"]["mpjpe"], rtol=1e-9)
    np.testing.assert_allclose(pf[..., 1], o["per_frame"]["p_mpjpe"], rtol=1e-7)
    np.testing.assert_allclose(pf[:, :-2, 2], o["per_frame"]["accel"], rtol=1e-9)


def test_metrics_large_batch_sums():
    B, T = 257, 27
    pred = synthetic.make_clips(B, T, 1, "gt") * 0.3
    gt, factor, res, actions = synthetic.make_labels(B, T, 2, n_actions=5)
    sums = _capi.metrics(pred.to(DEV), gt.to(DEV), res.to(DEV), factor.to(DEV), actions.to(DEV), 5)
    r = finalize_metrics(sums.cpu().numpy())
    o = MO.evaluate(pred.numpy().astype(np.float64), res.numpy(), factor.numpy(), gt.numpy(), actions=actions.numpy())
    assert abs(r["mpjpe"] - o["mpjpe"]) <= 1e-6 * o["mpjpe"]
    assert abs(r["p_mpjpe"] - o["p_mpjpe"]) <= 1e-6 * o["p_mpjpe"]
    assert abs(r["acceleration_error"] - o["accel"]) <= 1e-6 * o["accel"]


def test_out_of_range_action_is_skipped_not_folded_into_action_0():
    B, T = 6, 27
    pred = synthetic.make_clips(B, T, 3, "gt") * 0.3
    gt, factor, res, actions = synthetic.make_labels(B, T, 4, n_actions=3)
    bad = actions.clone()
    bad[1], bad[4] = 7, -1
    good = _capi.metrics(pred.to(DEV), gt.to(DEV), res.to(DEV), factor.to(DEV), actions.to(DEV), 3).cpu().numpy()
    got = _capi.metrics(pred.to(DEV), gt.to(DEV), res.to(DEV), factor.to(DEV), bad.to(DEV), 3).cpu().numpy()
    keep = [i for i in range(B) if i not in (1, 4)]
    want = _capi.metrics(pred[keep].to(DEV), gt[keep].to(DEV), res[keep].to(DEV), factor[keep].to(DEV),
                         actions[keep].to(DEV), 3).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert got[:, 3].sum() == (B - 2) * T and good[:, 3].sum() == B * T
    with pytest.raises(ValueError):
        finalize_metrics(got, expect_frames=B * T)
    finalize_metrics(good, expect_frames=B * T)
