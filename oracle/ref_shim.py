"""Import the UNMODIFIED reference (read-only checkout) on CPU.  Build-container only.

TEST INFRASTRUCTURE: used by `oracle/make_golden.py` to record golden vectors and by the optional
`tests/test_against_reference.py` (skipped when the checkout is absent, e.g. on the GPU box).
Nothing here copies reference sources; it only arranges `sys.modules` so that they import:
  * `timm.models.layers.DropPath` (model/KASportsFormer.py:12) is served by the reference's own
    identical implementation model/modules/drop.py:34-42 (never instantiated: drop_path == 0);
  * `easydict.EasyDict` (train_and_evaluate_sp.py:8, utils/utilities.py:9) -> an attribute dict;
  * `model.model_tools` (pulls timm.data / matplotlib / torchprofile at import) is replaced by a
    stub exposing the two names the eval script imports -- the script's own code is untouched.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("KASF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "model", "KASportsFormer.py"))


_done = False


def install():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF}")
    sys.path.insert(0, REF)
    from model.modules.drop import DropPath          # reference's own DropPath
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath = DropPath
    timm.models, timm_models.layers = timm_models, timm_layers
    sys.modules.setdefault("timm", timm)
    sys.modules.setdefault("timm.models", timm_models)
    sys.modules.setdefault("timm.models.layers", timm_layers)

    class EasyDict(dict):
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__
    ed = types.ModuleType("easydict")
    ed.EasyDict = EasyDict
    sys.modules.setdefault("easydict", ed)
    _done = True


def reference_model_class():
    install()
    from model.KASportsFormer import KASportsFormer
    return KASportsFormer


def reference_modules():
    """(bone_decomposer, error_calc module, joint_flip)."""
    install()
    from model.KASportsFormer import bone_decomposer
    import utils.error_calc as error_calc
    from utils.utilities import joint_flip
    return bone_decomposer, error_calc, joint_flip


def reference_eval_loop():
    """`evaluate_one_epoch_new` of train_and_evaluate_sp.py:27-149, imported unmodified."""
    install()
    if "model.model_tools" not in sys.modules:
        stub = types.ModuleType("model.model_tools")
        stub.load_model = lambda args: (_ for _ in ()).throw(RuntimeError("stub"))
        stub.total_parameters_count = lambda m: sum(p.numel() for p in m.parameters())
        sys.modules["model.model_tools"] = stub
    os.environ.setdefault("WANDB_MODE", "disabled")
    import train_and_evaluate_sp as tes
    return tes.evaluate_one_epoch_new, sys.modules["easydict"].EasyDict


def build_reference(cfg: dict, state=None):
    """Construct the real reference model from our cfg keys; optionally load a state dict."""
    import torch
    K = reference_model_class()
    m = K(n_layers=cfg["n_layers"], dim_in=3, dim_feat=cfg["dim_feat"], dim_rep=cfg["dim_rep"], dim_out=3,
          mlp_ratio=cfg["mlp_ratio"], act_layer=torch.nn.GELU, num_heads=cfg["num_heads"],
          num_joints=17, neighbour_num=cfg["neighbour_num"], n_frames=cfg["n_frames"],
          layer_scale_init_value=1e-5)
    if state is not None:
        m.load_state_dict(state, strict=True)
    return m.eval()
