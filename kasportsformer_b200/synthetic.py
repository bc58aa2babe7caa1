"""Deterministic synthetic weights and clips.

Everything is derived from numpy's PCG64 `random()` stream (uniform doubles produced by integer
arithmetic), so the same seed gives bit-identical float32 tensors in the build container and on the
GPU box -- the golden vectors in tests/golden/ were recorded from the real reference loaded with
exactly these weights.  Distributions follow the reference's initialisers (nn.Linear /
LayerNorm / BatchNorm1d defaults, reference model/modules/graph.py:46-50 for U,V,
model/KASportsFormer.py:264-266 fusion, :300-302 pos-embeds, :100-101 layer scale); the normal
initialisers are replaced by uniform ones of equal variance.

Two regimes:
  * "default": what a freshly constructed reference model looks like (layer_scale 1e-5 -> every block
    is nearly the identity; this is the acceptance regime of BASELINE.json);
  * "stress":  trained-like magnitudes (layer scales U(0.05,0.15), random LN/BN/fusion/pos-embeds,
    random BN running stats) which make per-stage parity meaningful.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .skeleton import LIMB_GROUPS, LIMB_CHANNEL_NAMES, LIMB_HIDDEN


class _Rng:
    def __init__(self, seed):
        self.g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, shape, lo, hi):
        u = self.g.random(size=shape)                       # float64 in [0,1)
        return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32))


def _linear(r: _Rng, out_f, in_f, bias=True, std=None):
    if std is None:
        bound = 1.0 / math.sqrt(in_f)                       # kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
    else:
        bound = std * math.sqrt(3.0)
    w = r.uniform((out_f, in_f), -bound, bound)
    b = r.uniform((out_f,), -1.0 / math.sqrt(in_f), 1.0 / math.sqrt(in_f)) if bias else None
    return w, b


def make_state(cfg: dict, seed: int = 0, regime: str = "default") -> Dict[str, torch.Tensor]:
    """Full reference-named state dict (float tensors + int64 num_batches_tracked)."""
    assert regime in ("default", "stress")
    r = _Rng(seed)
    stress = regime == "stress"
    D, T, J = cfg["dim_feat"], cfg["n_frames"], cfg["num_joints"]
    Hd = D * cfg["mlp_ratio"]
    s: Dict[str, torch.Tensor] = {}

    def put_linear(name, out_f, in_f, bias=True, std=None):
        w, b = _linear(r, out_f, in_f, bias, std)
        s[name + ".weight"] = w
        if bias:
            s[name + ".bias"] = b

    def put_ln(name):
        if stress:
            s[name + ".weight"] = r.uniform((D,), 0.7, 1.3)
            s[name + ".bias"] = r.uniform((D,), -0.2, 0.2)
        else:
            s[name + ".weight"] = torch.ones(D)
            s[name + ".bias"] = torch.zeros(D)

    for nm in ("pos_embed", "bone_pos_embed", "limb_pos_embed"):
        s[nm] = r.uniform((1, J, D), -0.3, 0.3) if stress else torch.zeros(1, J, D)
    for nm in ("joints_embed", "bone_embed", "limb_embed"):
        put_linear(nm, D, 3)
    put_ln("norm")
    for g, members in enumerate(LIMB_GROUPS):
        for ch in LIMB_CHANNEL_NAMES:
            put_linear(f"bone_refusion.mlp_layers.{g}.{ch}.fc1", LIMB_HIDDEN, len(members))
            put_linear(f"bone_refusion.mlp_layers.{g}.{ch}.fc2", 1, LIMB_HIDDEN)
    for l in range(cfg["n_layers"]):
        for br in ("att", "graph", "bone"):
            for mode in ("spatial", "temporal"):
                p = f"layers_with_bone.{l}.{br}_{mode}."
                for ls in ("layer_scale_1", "layer_scale_2"):
                    s[p + ls] = r.uniform((D,), 0.05, 0.15) if stress else torch.full((D,), 1e-5)
                put_ln(p + "norm1")
                put_ln(p + "norm1_limb")
                if br == "att":
                    put_linear(p + "mixer.proj", D, D)
                    put_linear(p + "mixer.qkv", 3 * D, D, bias=False)
                elif br == "bone":
                    put_linear(p + "mixer.proj", D, D)
                    put_linear(p + "mixer.qkv_q", D, D, bias=False)
                    put_linear(p + "mixer.qkv_kv", 2 * D, D, bias=False)
                else:
                    nodes = J if mode == "spatial" else T
                    put_linear(p + "mixer.U", D, D, std=math.sqrt(2.0 / D))
                    put_linear(p + "mixer.V", D, D, std=math.sqrt(2.0 / D))
                    if stress:
                        s[p + "mixer.batch_norm.weight"] = r.uniform((nodes,), 0.7, 1.3)
                        s[p + "mixer.batch_norm.bias"] = r.uniform((nodes,), -0.2, 0.2)
                        s[p + "mixer.batch_norm.running_mean"] = r.uniform((nodes,), -0.3, 0.3)
                        s[p + "mixer.batch_norm.running_var"] = r.uniform((nodes,), 0.5, 2.0)
                    else:
                        s[p + "mixer.batch_norm.weight"] = torch.ones(nodes)
                        s[p + "mixer.batch_norm.bias"] = torch.zeros(nodes)
                        s[p + "mixer.batch_norm.running_mean"] = torch.zeros(nodes)
                        s[p + "mixer.batch_norm.running_var"] = torch.ones(nodes)
                    s[p + "mixer.batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
                put_ln(p + "norm2")
                put_linear(p + "mlp.fc1", Hd, D)
                put_linear(p + "mlp.fc2", D, Hd)
        fp = f"layers_with_bone.{l}.fusion_three_channel"
        if stress:
            put_linear(fp, 3, 3 * D)
        else:
            s[fp + ".weight"] = torch.zeros(3, 3 * D)
            s[fp + ".bias"] = torch.full((3,), 1.0 / 3.0)
    put_linear("rep_logit.fc", cfg["dim_rep"], D)
    put_linear("head", 3, cfg["dim_rep"])
    return s


def make_clips(B: int, T: int, seed: int = 0, kind: str = "det"):
    """Synthetic 2D keypoint clips [B,T,17,3]: xy ~ 0.5*N(0,1)-like (sum of uniforms), confidence 1.0
    for "gt" configs (reference data/reader/sp_reader.py:52-55) or U(0.3,1) for "det" (:46-51)."""
    r = _Rng(seed + 7919)
    # Irwin-Hall(4) rescaled to unit variance: bit-reproducible, close to N(0,1)
    u = r.g.random(size=(B, T, 17, 2, 4)).sum(-1)
    xy = ((u - 2.0) * math.sqrt(3.0) * 0.5).astype(np.float32)
    conf = np.ones((B, T, 17, 1), np.float32) if kind == "gt" else \
        (0.3 + 0.7 * r.g.random(size=(B, T, 17, 1))).astype(np.float32)
    return torch.from_numpy(np.concatenate([xy, conf], axis=-1))


def make_labels(B: int, T: int, seed: int = 0, n_actions: int = 1, res=(1312.0, 1216.0)):
    """gt [B,T,17,3] mm ~ 250*N(0,1)-like, factor [B,T] ~ U(2,5), res [B,2], actions int32 [B]."""
    r = _Rng(seed + 104729)
    u = r.g.random(size=(B, T, 17, 3, 4)).sum(-1)
    gt = ((u - 2.0) * math.sqrt(3.0) * 250.0).astype(np.float32)
    factor = (2.0 + 3.0 * r.g.random(size=(B, T))).astype(np.float32)
    resa = np.tile(np.asarray(res, np.float32)[None], (B, 1))
    actions = (np.arange(B) % n_actions).astype(np.int32)
    return (torch.from_numpy(gt), torch.from_numpy(factor), torch.from_numpy(resa),
            torch.from_numpy(actions))


def state_digest(state: Dict[str, torch.Tensor]) -> str:
    """sha256 over the float tensors in sorted-name order (pins the generator across machines)."""
    import hashlib
    h = hashlib.sha256()
    for k in sorted(state):
        if state[k].is_floating_point():
            h.update(k.encode())
            h.update(state[k].contiguous().numpy().tobytes())
    return h.hexdigest()
