"""What the two precision modes cost against the REFERENCE (not against an oracle that follows the kernel's roundings),
in millimetres, on weights with trained-like magnitudes.  Default-init weights scale every block by 1e-5 and hide
block-level error, so the acceptance numbers of README / DESIGN come from here:

  * precision="exact" (fp32 FMA everywhere, erf GELU) reproduces the reference goldens to <= 1e-2 mm max |dy|
    (BASELINE.json's bar; 1e-2 mm == 1.04e-5 normalised units at res_w = 1920, factor 1) and its MPJPE to <= 0.01 mm;
  * precision="fast" (bf16 tensor-core operands, fp16 hidden tile, tanh-form GELU) is bounded at the values measured on
    B200 (written next to each assert) -- two orders of magnitude above the bar once the layer scales are ~0.1.

Every run appends its measured numbers to gpurun_out/precision_report.jsonl (bench.py reports the same two figures)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, ROOT
from kasportsformer_b200 import KASportsFormer, synthetic
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MM = 1920.0 / 2.0            # normalised units -> mm at res_w = 1920, factor 1


def _report(**kw):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision_report.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")


def _model(meta):
    cfg = meta["cfg"]
    m = KASportsFormer(n_layers=cfg["n_layers"], num_heads=8, n_frames=cfg["n_frames"])
    m.load_state_dict(synthetic.make_state(cfg, meta["seed"], meta["regime"]), strict=True)
    return m.to(DEV).eval()


# measured on B200 (round 2, profiles/r02_precision_report.jsonl): fast-mode max |dy| against the reference golden is
# 0.75 / 2.71 / 0.42 mm (mean 0.13 / 0.08 / 0.09 mm), exact-mode 0.0005 / 0.0004 / 0.0004 mm; the bound is ~2x the measurement.
# (T = 81 runs on the split path: its maximum sits on one token behind a top-k edge that the bf16 operands of the preceding
#  blocks moved -- 1.33 mm with the fused kernel's similarity, 2.71 mm with the split path's; the mean is unchanged.)
FAST_BOUND_MM = {"stress_L2_T27.npz": 1.6, "stress_L1_T81.npz": 5.0, "stress_L1_T9.npz": 1.0}


@pytest.mark.parametrize("name", ["stress_L2_T27.npz", "stress_L1_T81.npz", "stress_L1_T9.npz"])
def test_stress_goldens_both_modes(name):
    z, meta = load_golden(name)
    m = _model(meta)
    x = synthetic.make_clips(meta["B"], meta["cfg"]["n_frames"], meta["clip_seed"], meta["kind"]).to(DEV)
    ref = z["t:y"]
    y_fast = m(x).cpu().numpy()
    m.precision = "exact"
    y_exact = m(x).cpu().numpy()
    e_fast, e_exact = np.abs(y_fast - ref).max() * MM, np.abs(y_exact - ref).max() * MM
    _report(test="stress_golden", golden=name, fast_max_mm=float(e_fast), exact_max_mm=float(e_exact),
            fast_mean_mm=float(np.abs(y_fast - ref).mean() * MM), exact_mean_mm=float(np.abs(y_exact - ref).mean() * MM))
    assert e_exact <= 1e-2, f"exact mode: max |dy| = {e_exact} mm"
    assert e_fast <= FAST_BOUND_MM[name], f"fast mode: max |dy| = {e_fast} mm"


def test_trained_like_26_layers_vs_reference():
    """The full-depth model with trained-like magnitudes: final joints and the eval protocol's MPJPE / P-MPJPE
    (reference utils/error_calc.py on the reference's output, recorded by oracle/make_golden.py)."""
    z, meta = load_golden("trained_like_L26_T27.npz")
    m = _model(meta)
    B = meta["B"]
    x = synthetic.make_clips(B, 27, meta["clip_seed"], meta["kind"]).to(DEV)
    gt, factor, res, _ = synthetic.make_labels(B, 27, seed=meta["label_seed"], n_actions=1)
    out = {}
    for mode in ("fast", "exact"):
        m.precision = mode
        y = m(x).cpu().numpy()
        r = MO.evaluate(y, res.numpy().astype(np.float64), factor.numpy(), gt.numpy())
        out[mode] = dict(max_mm=float(np.abs(y - z["y"]).max() * MM), mean_mm=float(np.abs(y - z["y"]).mean() * MM),
                         d_mpjpe_mm=float(abs(r["mpjpe"] - float(z["mpjpe"]))),
                         d_p_mpjpe_mm=float(abs(r["p_mpjpe"] - float(z["p_mpjpe"]))))
    _report(test="trained_like_L26_T27", **out)
    # exact: the reference's arithmetic; what is left is fp32 summation order (and, rarely, a flipped top-k near-tie)
    assert out["exact"]["max_mm"] <= 1e-2, out
    assert out["exact"]["d_mpjpe_mm"] <= 0.01 and out["exact"]["d_p_mpjpe_mm"] <= 0.01, out
    # (measured on B200: max 0.0007 mm, MPJPE / P-MPJPE equal to 6e-8 / 1.5e-6 mm)
    # fast: measured on B200 max 11.6 mm (a handful of tokens behind a flipped top-k edge), mean 0.31 mm, MPJPE within
    # 0.27 mm, P-MPJPE within 0.002 mm of the reference's 1422.6 / 368.5 mm (bounds ~2x)
    assert out["fast"]["max_mm"] <= 25.0 and out["fast"]["mean_mm"] <= 0.7, out
    assert out["fast"]["d_mpjpe_mm"] <= 0.6 and out["fast"]["d_p_mpjpe_mm"] <= 0.05, out


def test_exact_mode_contract():
    """exact mode: same module, same output contract (fresh tensor, return_rep, batch invariance)."""
    cfg = dict(n_layers=1, n_frames=27, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17,
               neighbour_num=4)
    m = KASportsFormer(n_layers=1, num_heads=8, n_frames=27)
    m.load_state_dict(synthetic.make_state(cfg, 7, "stress"))
    m = m.to(DEV).eval()
    m.precision = "exact"
    x = synthetic.make_clips(5, 27, 3, "det").to(DEV)
    y = m(x)
    assert y.shape == (5, 27, 17, 3) and torch.isfinite(y).all()
    assert torch.equal(torch.cat([m(x[:2]), m(x[2:])]), y)
    assert m(x, return_rep=True).shape == (5, 27, 17, 512)
    m.precision = "fast"
    assert (m(x) - y).abs().max().item() < 5e-2


def test_reference_ctor_default_num_heads_4_runs_exact_only():
    """`KASportsFormer()` with the reference constructor's own default num_heads=4 (model/KASportsFormer.py:291-295; head_dim
    32, no shipped YAML) constructs and runs -- in the fp32 path -- and matches the oracle; the tensor-core path refuses it."""
    from oracle import kasf_oracle as O
    cfg = dict(n_layers=2, n_frames=27, dim_feat=128, dim_rep=512, num_heads=4, mlp_ratio=4, num_joints=17,
               neighbour_num=4)
    state = synthetic.make_state(cfg, 12, "stress")
    m = KASportsFormer(n_layers=2, n_frames=27)              # num_heads defaults to 4
    m.load_state_dict(state)
    m = m.to(DEV).eval()
    assert m.precision == "exact"
    x = synthetic.make_clips(3, 27, 6, "det")
    y = m(x.to(DEV)).cpu()
    ref = O.forward(state, x, O.default_config(n_layers=2, n_frames=27, num_heads=4))
    assert (y - ref).abs().max().item() * MM <= 1e-2
    m.precision = "fast"
    with pytest.raises(NotImplementedError):
        m(x.to(DEV))
