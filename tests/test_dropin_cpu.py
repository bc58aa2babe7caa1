"""Drop-in boundary, host side (no GPU): the torch custom ops are registered with shape-only fake implementations, the
`model` import shadow swaps the B200 module into the reference's UNMODIFIED script, and the restatement of the
reference's evaluation loop that the GPU test runs is pinned against the real loop's recorded results."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import kasportsformer_b200  # noqa: F401  (registers torch.ops.kasf.*)
from conftest import load_golden, ROOT
from oracle import eval_loop_oracle as ELO
from oracle import ref_shim


def test_custom_ops_registered_with_fake_impls():
    x = torch.empty(3, 27, 17, 3, device="meta")
    blob = torch.empty(16, dtype=torch.uint8, device="meta")
    assert torch.ops.kasf.forward(x, blob, None, 26, 27, False, 0, 0).shape == (3, 27, 17, 3)
    assert torch.ops.kasf.forward(x, blob, None, 26, 27, True, 0, 0).shape == (3, 27, 17, 512)
    s = torch.ops.kasf.metrics(x, None, x, torch.empty(3, 2, device="meta"), torch.empty(3, 27, device="meta"),
                               torch.empty(3, dtype=torch.int32, device="meta"), 5)
    assert s.shape == (5, 22) and s.dtype == torch.float64
    with pytest.raises(NotImplementedError):          # there is no CPU kernel behind the op
        torch.ops.kasf.forward(torch.zeros(1, 27, 17, 3), torch.zeros(4, dtype=torch.uint8), None, 26, 27, False, 0, 0)


def test_eval_loop_restatement_replays_reference_golden():
    """oracle/eval_loop_oracle.py on the predictions recorded in metrics.npz == what the unmodified reference loop
    (train_and_evaluate_sp.py:27-149) returned for them."""
    z, _ = load_golden("metrics.npz")
    names = ["a", "b", "a", "c", "b", "a"]
    for flip in (False, True):
        class Fake(torch.nn.Module):
            calls = 0

            def forward(self, x):
                Fake.calls += 1
                return torch.from_numpy(z["pred"] if Fake.calls % 2 == 1 else z["pred_flip"]).clone()
        Fake.calls = 0
        loader = [(torch.zeros(6, 27, 17, 3), torch.from_numpy(z["gt"]), torch.from_numpy(z["factor"]), names,
                   torch.from_numpy(z["res"]))]
        r = ELO.evaluate_loop(Fake(), loader, "cpu", flip)
        ref = z["eval_flip" if flip else "eval_noflip"]
        assert abs(r["mpjpe"] - ref[0]) <= 1e-3 and abs(r["p_mpjpe"] - ref[1]) <= 1e-3
        assert abs(r["acceleration_error"] - ref[2]) <= 1e-3
        assert np.abs(r["mpjpe_joint"] - z[("eval_flip" if flip else "eval_noflip") + "_joint"]).max() <= 1e-3


_SWAP = r'''
import os, sys, types
shadow, ref = sys.argv[1], sys.argv[2]
sys.path[:0] = [shadow, ref]                      # the shadow `model` package wins over the reference's
class EasyDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__
ed = types.ModuleType("easydict"); ed.EasyDict = EasyDict; sys.modules["easydict"] = ed     # not installed here
os.environ["WANDB_MODE"] = "disabled"
import torch
import train_and_evaluate_sp as tes             # the reference's script, unmodified
import kasportsformer_b200
assert tes.load_model is kasportsformer_b200.load_model, "the script did not bind the drop-in factory"
from utils.utilities import yaml_config_reader    # the reference's own config reader
args = yaml_config_reader(os.path.join(ref, "configs", "sportspose-gt-kasportsformer.yaml"))
model = tes.load_model(args)                      # train_and_evaluate_sp.py:161
assert type(model) is kasportsformer_b200.KASportsFormer
assert tes.total_parameters_count(model) == 29365668
wrapped = torch.nn.DataParallel(model)            # :164-165
ckpt = {"module." + k: v for k, v in model.state_dict().items()}        # what utilities.py:115 saves
assert len(ckpt) == 2975
wrapped.load_state_dict(ckpt, strict=True)        # :174
try:
    wrapped.module.eval()(torch.zeros(1, 27, 17, 3))
except RuntimeError as e:
    assert "no CPU path" in str(e)
else:
    raise SystemExit("CPU forward did not raise")
print("SWAP-OK")
'''


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present (GPU box)")
def test_unmodified_reference_script_binds_the_dropin():
    """`PYTHONPATH=shadow:reference` and the unchanged train_and_evaluate_sp.py builds, wraps and loads OUR module the way
    its `evaluate()` does (:152-176): factory from the reference's YAML, nn.DataParallel, strict `module.`-prefixed load."""
    shadow = os.path.join(ROOT, "kasportsformer_b200", "shadow")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", _SWAP, shadow, ref_shim.REF], capture_output=True, text=True, env=env,
                       timeout=300)
    assert "SWAP-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
