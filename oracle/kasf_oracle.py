"""CPU oracle for the KASportsFormer inference forward pass.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain tensor math on the CPU, of the algorithm the reference
implements in `model/KASportsFormer.py` and `model/modules/*.py`.  It is the checker that the CUDA
path is compared with; it is never the product path.  Only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity status: PINNED.  `oracle/make_golden.py` (run in the build container, where the unmodified
reference at /root/reference is importable on CPU) loads identical weights into the real reference
modules and records stage outputs into `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
this restatement against those vectors.  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), and its arithmetic lives in PyTorch (ATen), so the vectors recorded from the
reference executed here are the pin.

All functions take a flat ``state`` dict with the reference's state_dict names and work in the
dtype of ``x`` (float32 to mirror the reference, float64 to arbitrate summation-order noise and
top-k near ties: the reference cannot run in float64 because graph.py:111 hard-casts to float32).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

# --- constant tables, transcribed from the reference (they define results) -------------------
# reference model/KASportsFormer.py:46-47
BONE_CHILD = [0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15]
BONE_PARENT = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
# reference model/modules/bone_refusion.py:34-40
LIMB_GROUPS = [
    [0, 1, 2], [3, 4, 5], [6, 7], [8, 9], [10, 11, 12], [13, 14, 15],
    [6, 7, 1, 2], [6, 7, 4, 5], [6, 7, 11, 12], [6, 7, 14, 15], [6, 7, 9],
    [14, 15, 11, 12], [1, 2, 4, 5], [14, 15, 4, 5], [11, 12, 4, 5], [10, 0], [13, 3],
]
# reference model/modules/graph.py:16-17
CONNECTIONS = {10: [9], 9: [8, 10], 8: [7, 9, 11, 14], 14: [15, 8], 15: [16, 14], 11: [12, 8],
               12: [13, 11], 7: [0, 8], 0: [1, 7, 4], 1: [2, 0], 2: [3, 1], 4: [5, 0], 5: [6, 4],
               16: [15], 13: [12], 3: [2], 6: [5]}
LIMB_CH = ["mlp_dir_x", "mlp_dir_y", "mlp_len"]

Tensor = torch.Tensor
State = Dict[str, Tensor]
Hook = Optional[Callable[[str, Tensor], None]]


def default_config(**over):
    """Model keys of the shipped YAMLs (reference configs/*.yaml:66-92)."""
    cfg = dict(n_layers=26, dim_in=3, dim_feat=128, dim_rep=512, dim_out=3, mlp_ratio=4,
               num_heads=8, num_joints=17, neighbour_num=4, n_frames=27)
    cfg.update(over)
    return cfg


# --- kinematic anatomy features ------------------------------------------------------------------
def bone_features(x: Tensor) -> Tensor:
    """reference model/KASportsFormer.py:42-62 (bone_decomposer). x [B,T,17,3] -> [B,T,17,3]."""
    p = x[..., :2]
    d = p[:, :, BONE_CHILD] - p[:, :, BONE_PARENT]            # [B,T,16,2]
    length = torch.sqrt((d * d).sum(-1, keepdim=True))        # [B,T,16,1]
    length = torch.where(length == 0, torch.ones_like(length), length)
    u = d / length
    u = torch.cat([u, u.mean(dim=-2, keepdim=True)], dim=-2)
    length = torch.cat([length, length.mean(dim=-2, keepdim=True)], dim=-2)
    return torch.cat([u, length], dim=-1)


def gelu_erf(v: Tensor) -> Tensor:
    """exact (erf) GELU = 0.5 v (1 + erf(v / sqrt 2)); nn.GELU() default, model_tools.py:81."""
    return F.gelu(v)


# When True, the dense projections of the FormerModules round their operands to bfloat16 (fp32
# accumulation) at the points where the CUDA kernels do (tensor-core operands; K/V tiles), so the
# kernels can be compared with a tight tolerance.  Embeddings, head, LayerNorm, softmax, similarity
# stay fp32 in the kernels and here.  Default False = the reference's fp32 arithmetic.
EMULATE_BF16 = False


def _q(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(t.dtype) if EMULATE_BF16 else t


def linear(v: Tensor, w: Tensor, b: Optional[Tensor] = None, exact: bool = False) -> Tensor:
    """y = v W^T + b  (torch nn.Linear convention; W is [out, in])."""
    if EMULATE_BF16 and not exact:
        return F.linear(_q(v), _q(w), b)
    return F.linear(v, w, b)


def limb_features(state: State, x: Tensor) -> Tensor:
    """reference model/modules/bone_refusion.py:61-70 + bone_MLP.py:16-27.

    17 limb groups x 3 channels of tiny MLPs (n -> 16 -> 1) applied to the RAW joints."""
    outs = []
    for g, idx in enumerate(LIMB_GROUPS):
        chans = []
        for c, nm in enumerate(LIMB_CH):
            pre = f"bone_refusion.mlp_layers.{g}.{nm}."
            v = x[:, :, idx, c]                                         # [B,T,n]
            h = gelu_erf(linear(v, state[pre + "fc1.weight"], state[pre + "fc1.bias"], exact=True))
            chans.append(linear(h, state[pre + "fc2.weight"], state[pre + "fc2.bias"], exact=True))  # [B,T,1]
        outs.append(torch.cat(chans, dim=-1).unsqueeze(-2))             # [B,T,1,3]
    return torch.cat(outs, dim=-2)


def embed(state: State, x: Tensor, bone: Tensor, limb: Tensor):
    """reference model/KASportsFormer.py:325-330."""
    X = linear(x, state["joints_embed.weight"], state["joints_embed.bias"], exact=True) + state["pos_embed"]
    XB = linear(bone, state["bone_embed.weight"], state["bone_embed.bias"], exact=True) + state["bone_pos_embed"]
    XL = linear(limb, state["limb_embed.weight"], state["limb_embed.bias"], exact=True) + state["limb_pos_embed"]
    return X, XB, XL


# --- building blocks ---------------------------------------------------------------------------
def layer_norm(v: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    # (v - mean) / sqrt(biased_var + eps) * w + b over the last axis
    return F.layer_norm(v, (v.shape[-1],), w, b, eps)


def _attend(q: Tensor, k: Tensor, v: Tensor, mode: str, heads: int) -> Tensor:
    """q,k,v [B,T,J,C] -> [B,T,J,C]. reference selfattention.py:18-41 / bone_crossattention.py:19-41."""
    B, T, J, C = q.shape
    d = C // heads
    scale = d ** -0.5
    def split(t):  # [B,T,J,H,d] -> [B,H,T,J,d]
        return t.reshape(B, T, J, heads, d).permute(0, 3, 1, 2, 4)
    q, k, v = split(_q(q)), split(_q(k)), split(_q(v))   # Q/K/V tiles are kept in bf16 on chip
    if mode == "temporal":
        q, k, v = q.transpose(2, 3), k.transpose(2, 3), v.transpose(2, 3)   # [B,H,J,T,d]
    if EMULATE_BF16:
        # the kernel feeds the un-normalised probabilities to the P.V tensor-core MMA in bf16 and
        # divides by the fp32 row sum afterwards
        sc = (q @ k.transpose(-2, -1)) * scale
        p = torch.exp(sc - sc.amax(dim=-1, keepdim=True))
        o = (_q(p) @ v) / p.sum(dim=-1, keepdim=True)
    else:
        att = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
        o = att @ v
    if mode == "temporal":
        o = o.transpose(2, 3)                                               # [B,H,T,J,d]
    return o.permute(0, 2, 3, 1, 4).reshape(B, T, J, C)


def attention_mixer(state: State, pre: str, z: Tensor, mode: str, heads: int, norm: Optional[tuple] = None) -> Tensor:
    """reference model/modules/selfattention.py:44-60.

    norm = (x, gamma, beta) lets the bf16-emulating mode follow the kernels, which fold LN1's affine into the Q|K|V
    weights: operand = the normalised row rounded to bf16, weight = bf16(W diag(gamma)), explicit fp32 query bias
    W_q beta; the beta term of K cancels in the softmax, the one of V joins the projection bias."""
    C = z.shape[-1]
    if EMULATE_BF16 and norm is not None:
        x, gamma, beta = norm
        xhat = F.layer_norm(x, (C,), None, None, 1e-5)
        w = state[pre + "qkv.weight"]
        qkv = F.linear(_q(xhat), _q(w * gamma[None, :]))
        q = qkv[..., :C] + w[:C] @ beta
        o = _attend(q, qkv[..., C:2 * C], qkv[..., 2 * C:], mode, heads)
        wp = state[pre + "proj.weight"]
        return F.linear(_q(o), _q(wp), state[pre + "proj.bias"] + wp @ (w[2 * C:] @ beta))
    qkv = linear(z, state[pre + "qkv.weight"])
    o = _attend(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], mode, heads)
    return linear(o, state[pre + "proj.weight"], state[pre + "proj.bias"])


def bone_mixer(state: State, pre: str, z: Tensor, zl: Tensor, mode: str, heads: int,
               limb_norm: Optional[tuple] = None, norm: Optional[tuple] = None) -> Tensor:
    """reference model/modules/bone_crossattention.py:43-62.

    limb_norm = (xl, gamma, beta) lets the bf16-emulating mode follow the kernels, which fold the limb LayerNorm's
    affine into the K|V weights (operand = the normalised limb row, rounded to bf16; weight = bf16(W diag(gamma)));
    the beta term of K cancels in the softmax and the beta term of V is added, in fp32, to the projection bias."""
    C = z.shape[-1]
    if EMULATE_BF16 and norm is not None:      # LN1's affine folded into W_q, explicit query bias (see attention_mixer)
        x, gamma, beta = norm
        wq = state[pre + "qkv_q.weight"]
        q = F.linear(_q(F.layer_norm(x, (C,), None, None, 1e-5)), _q(wq * gamma[None, :])) + wq @ beta
    else:
        q = linear(z, state[pre + "qkv_q.weight"])
    if EMULATE_BF16 and limb_norm is not None:
        xl, gamma, beta = limb_norm
        xhat = F.layer_norm(xl, (C,), None, None, 1e-5)
        wkv = state[pre + "qkv_kv.weight"]
        kv = F.linear(_q(xhat), _q(wkv * gamma[None, :]))
        o = _attend(q, kv[..., :C], kv[..., C:], mode, heads)
        wp = state[pre + "proj.weight"]
        return F.linear(_q(o), _q(wp), state[pre + "proj.bias"] + wp @ (wkv[C:] @ beta))
    kv = linear(zl, state[pre + "qkv_kv.weight"])
    o = _attend(q, kv[..., :C], kv[..., C:], mode, heads)
    return linear(o, state[pre + "proj.weight"], state[pre + "proj.bias"])


def skeleton_adjacency(dtype=torch.float32) -> Tensor:
    """reference model/modules/graph.py:52-61."""
    a = torch.zeros(17, 17, dtype=dtype)
    for i, nb in CONNECTIONS.items():
        for j in nb:
            a[i, j] = 1
    return a


def temporal_adjacency(z: Tensor, k: int) -> Tensor:
    """z [G,T,C] -> 0/1 adjacency [G,T,T]; reference graph.py:108-111 (ties kept by >=)."""
    sim = z @ z.transpose(1, 2)
    thr = sim.topk(k=k, dim=-1, largest=True)[0][..., -1:]
    return (sim >= thr).to(z.dtype)


def gcn_mixer(state: State, pre: str, z: Tensor, mode: str, neighbour_num: int,
              adj_override: Optional[Tensor] = None) -> Tensor:
    """reference model/modules/graph.py:99-134. z [B,T,J,C] is the LayerNorm output."""
    B, T, J, C = z.shape
    if mode == "temporal":
        g = z.transpose(1, 2).reshape(B * J, T, C)
        adj = temporal_adjacency(g, neighbour_num) if adj_override is None else adj_override
    else:
        g = z.reshape(B * T, J, C)
        adj = skeleton_adjacency(z.dtype).to(z.device).unsqueeze(0)
    deg = adj.sum(-1)                                         # row sums, graph.py:81
    dis = deg ** -0.5
    norm_adj = dis.unsqueeze(-1) * adj * dis.unsqueeze(-2)   # D^-1/2 A D^-1/2, graph.py:86-88
    vz = linear(g, state[pre + "V.weight"], state[pre + "V.bias"])
    uz = linear(g, state[pre + "U.weight"], state[pre + "U.bias"])
    y = norm_adj @ vz + uz
    # BatchNorm1d(num_nodes) in eval mode: the channel axis is the NODE axis (graph.py:37,129)
    rm = state[pre + "batch_norm.running_mean"].view(1, -1, 1)
    rv = state[pre + "batch_norm.running_var"].view(1, -1, 1)
    bw = state[pre + "batch_norm.weight"].view(1, -1, 1)
    bb = state[pre + "batch_norm.bias"].view(1, -1, 1)
    y = (y - rm) / torch.sqrt(rv + 1e-5) * bw + bb
    out = torch.relu(g + y)                                   # inner residual adds the LN output
    if mode == "temporal":
        return out.reshape(B, J, T, C).transpose(1, 2)
    return out.reshape(B, T, J, C)


def mlp(state: State, pre: str, z: Tensor, norm: Optional[tuple] = None) -> Tensor:
    """reference model/modules/mlp.py:24-30.

    norm = (x, gamma, beta) lets the bf16-emulating mode follow the kernels, which fold LN2's affine into fc1
    (operand = the normalised row rounded to bf16, weight = bf16(W1 diag(gamma)), bias = b1 + W1 beta in fp32)."""
    if EMULATE_BF16 and norm is not None:
        x, gamma, beta = norm
        w1 = state[pre + "fc1.weight"]
        xhat = F.layer_norm(x, (x.shape[-1],), None, None, 1e-5)
        h = gelu_erf(F.linear(_q(xhat), _q(w1 * gamma[None, :]), state[pre + "fc1.bias"] + w1 @ beta))
    else:
        h = gelu_erf(linear(z, state[pre + "fc1.weight"], state[pre + "fc1.bias"]))
    if EMULATE_BF16:
        # the kernels keep the hidden activation and the fc2 weights in fp16 (fc2 is an f16 x f16 -> fp32 MMA)
        w2 = state[pre + "fc2.weight"]
        return F.linear(h.to(torch.float16).to(h.dtype), w2.to(torch.float16).to(w2.dtype), state[pre + "fc2.bias"])
    return linear(h, state[pre + "fc2.weight"], state[pre + "fc2.bias"])


BRANCHES = (("att", "attention"), ("graph", "graph"), ("bone", "bone"))


def former_module(state: State, pre: str, v: Tensor, xl: Optional[Tensor], kind: str, mode: str,
                  cfg: dict, hook: Hook = None) -> Tensor:
    """reference model/KASportsFormer.py:103-118 (use_layer_scale=True path)."""
    z = layer_norm(v, state[pre + "norm1.weight"], state[pre + "norm1.bias"])
    if kind == "attention":
        m = attention_mixer(state, pre + "mixer.", z, mode, cfg["num_heads"],
                            (v, state[pre + "norm1.weight"], state[pre + "norm1.bias"]))
    elif kind == "graph":
        m = gcn_mixer(state, pre + "mixer.", z, mode, cfg["neighbour_num"])
    else:
        zl = layer_norm(xl, state[pre + "norm1_limb.weight"], state[pre + "norm1_limb.bias"])
        m = bone_mixer(state, pre + "mixer.", z, zl, mode, cfg["num_heads"],
                       (xl, state[pre + "norm1_limb.weight"], state[pre + "norm1_limb.bias"]),
                       (v, state[pre + "norm1.weight"], state[pre + "norm1.bias"]))
    if hook:
        hook(pre + "mixer", m)
    v = v + state[pre + "layer_scale_1"] * m
    if hook:
        hook(pre + "mid", v)
    h = mlp(state, pre + "mlp.", layer_norm(v, state[pre + "norm2.weight"], state[pre + "norm2.bias"]),
            (v, state[pre + "norm2.weight"], state[pre + "norm2.bias"]))
    if hook:
        hook(pre + "mlp", h)
    v = v + state[pre + "layer_scale_2"] * h
    if hook:
        hook(pre + "out", v)
    return v


def fuse(state: State, pre: str, a: Tensor, g: Tensor, b: Tensor) -> Tensor:
    """reference model/KASportsFormer.py:279-282 (adaptive fusion)."""
    logits = linear(torch.cat([a, g, b], dim=-1), state[pre + "weight"], state[pre + "bias"], exact=True)
    alpha = torch.softmax(logits, dim=-1)
    return a * alpha[..., 0:1] + g * alpha[..., 1:2] + b * alpha[..., 2:3]


def layer(state: State, l: int, X: Tensor, XB: Optional[Tensor], XL: Tensor, cfg: dict,
          hook: Hook = None) -> Tensor:
    """reference model/KASportsFormer.py:268-286. XB is given for layer 0 only (:332-336)."""
    P = f"layers_with_bone.{l}."
    outs = []
    for br, kind in BRANCHES:
        src = XB if (br == "bone" and XB is not None) else X
        v = former_module(state, P + br + "_spatial.", src, XL, kind, "spatial", cfg, hook)
        v = former_module(state, P + br + "_temporal.", v, XL, kind, "temporal", cfg, hook)
        outs.append(v)
    y = fuse(state, P + "fusion_three_channel.", *outs)
    if hook:
        hook(P + "out", y)
    return y


def head(state: State, X: Tensor, return_rep: bool = False) -> Tensor:
    """reference model/KASportsFormer.py:339-345."""
    z = layer_norm(X, state["norm.weight"], state["norm.bias"])
    rep = torch.tanh(linear(z, state["rep_logit.fc.weight"], state["rep_logit.fc.bias"], exact=True))
    if return_rep:
        return rep
    return linear(rep, state["head.weight"], state["head.bias"], exact=True)


def forward(state: State, x: Tensor, cfg: Optional[dict] = None, return_rep: bool = False,
            hook: Hook = None) -> Tensor:
    """reference model/KASportsFormer.py:320-347. x [B,T,17,3] -> [B,T,17,3]."""
    cfg = cfg or default_config()
    if x.dtype != next(iter(state.values())).dtype:
        raise ValueError("state and input dtypes differ; use cast_state()")
    bone = bone_features(x)
    limb = limb_features(state, x)
    if hook:
        hook("bone", bone)
        hook("limb", limb)
    X, XB, XL = embed(state, x, bone, limb)
    if hook:
        hook("X", X), hook("XB", XB), hook("XL", XL)
    for l in range(cfg["n_layers"]):
        X = layer(state, l, X, XB if l == 0 else None, XL, cfg, hook)
    return head(state, X, return_rep)


def cast_state(state: State, dtype) -> State:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state.items()}
