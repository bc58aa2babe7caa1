"""GPU parity tests, stage by stage, through the C-ABI (ctypes -> libkasf.so) against the CPU oracle
and the golden vectors recorded from the reference.  Tolerances are written next to each check.

Weight regime "stress" (trained-like layer scales ~0.1) is used for stage tests because default init
(layer_scale 1e-5) hides block-level bugs."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from kasportsformer_b200 import _capi, synthetic
from oracle import kasf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cfg(n_layers=1, n_frames=27):
    return dict(n_layers=n_layers, n_frames=n_frames, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4,
                num_joints=17, neighbour_num=4)


_cache = {}


def _setup(n_layers, T, seed, regime):
    key = (n_layers, T, seed, regime)
    if key not in _cache:
        cfg = _cfg(n_layers, T)
        state = synthetic.make_state(cfg, seed, regime)
        blob = _capi.pack_state(cfg, {k: v for k, v in state.items() if v.is_floating_point()}, torch.device(DEV))
        _cache[key] = (cfg, state, blob)
    return _cache[key]


def test_tcgen05_gemm_primitive():
    """bf16 x bf16 -> fp32 tensor-core GEMM (tcgen05 + bulk copy + TMEM load) vs fp32 matmul of the
    bf16-rounded operands: relative error <= 1e-5 (only accumulation order differs)."""
    g = torch.Generator().manual_seed(0)
    for M, N in ((128, 128), (300, 384), (1000, 512)):
        a = torch.randn(M, 128, generator=g)
        w = torch.randn(N, 128, generator=g) * 0.1
        d = _capi.test_gemm(a.to(DEV), w.to(DEV)).cpu()
        ref = a.bfloat16().float() @ w.bfloat16().float().t()
        err = (d - ref).abs().max().item()
        assert err <= 1e-5 * ref.abs().max().item() + 1e-5, f"M={M} N={N}: {err}"


@pytest.mark.parametrize("T,B,kind", [(27, 5, "det"), (27, 3, "gt"), (81, 2, "det"), (9, 9, "det"), (243, 1, "gt")])
def test_kinematic_features(T, B, kind):
    """K1 vs oracle: bone / limb features and the three embeddings, fp32: <= 1e-5 relative."""
    cfg, state, blob = _setup(1, T, 11, "stress")
    x = synthetic.make_clips(B, T, 3, kind)
    x[0, 0, 1] = x[0, 0, 0]          # a zero-length bone (KASportsFormer.py:51)
    bone, limb, X, XB, XL = _capi.kinematic_features(cfg, blob, x.to(DEV))
    rb, rl = O.bone_features(x), O.limb_features(state, x)
    rX, rXB, rXL = O.embed(state, x, rb, rl)
    for name, got, ref in (("bone", bone, rb), ("limb", limb, rl), ("X", X, rX), ("XB", XB, rXB), ("XL", XL, rXL)):
        err = (got.cpu() - ref).abs().max().item()
        assert err <= 1e-5 * max(1.0, ref.abs().max().item()), f"{name}: {err}"


def test_kinematic_features_golden():
    z, meta = load_golden("stress_L2_T27.npz")
    cfg, state, blob = _setup(2, 27, meta["seed"], meta["regime"])
    x = synthetic.make_clips(meta["B"], 27, meta["clip_seed"], meta["kind"])
    bone, limb, X, XB, XL = _capi.kinematic_features(cfg, blob, x.to(DEV))
    cs = meta["ch_stride"]
    assert np.abs(bone.cpu().numpy() - z["t:bone"]).max() <= 1e-5
    assert np.abs(limb.cpu().numpy() - z["t:limb"]).max() <= 1e-5
    for nm, t in (("X", X), ("XB", XB), ("XL", XL)):
        assert np.abs(t.cpu().numpy()[..., ::cs] - z["t:" + nm]).max() <= 1e-5 * max(1.0, np.abs(z["t:" + nm]).max())


def _module_oracle(state, cfg, layer, br, kind, mode, v, XL, emulate):
    O.EMULATE_BF16 = emulate
    try:
        return O.former_module(state, f"layers_with_bone.{layer}.{br}_{mode}.", v, XL, kind, mode,
                               O.default_config(n_layers=cfg["n_layers"], n_frames=cfg["n_frames"]))
    finally:
        O.EMULATE_BF16 = False


MODULES = [("att", "attention"), ("graph", "graph"), ("bone", "bone")]


# temporal split path: T = 72 / 81 / 100 / 128 / 150 / 185 / 215 / 243 cover every instantiated key-step count of the
# attention core (5, 6, 7, 8, 10, 12, 14, 16 sixteen-key steps) and both M-tile counts of the GCN kernel
@pytest.mark.parametrize("T,B", [(27, 3), (27, 10), (9, 5), (81, 2), (128, 1), (100, 2), (243, 1), (243, 2), (150, 2),
                                 (72, 1), (185, 1), (215, 1)])
@pytest.mark.parametrize("mode", ["spatial", "temporal"])
@pytest.mark.parametrize("br,kind", MODULES)
def test_former_module(br, kind, mode, T, B):
    """One FormerModule (the fused kernel; for temporal modules with T > 64 the split path: projection kernel,
    per-sequence mixer-core kernel, fused tail) vs the oracle.

    (a) against the oracle emulating the kernel's bf16 operand rounding: update error <= 2e-3 of the
        update magnitude (accumulation order, bf16 re-rounding of aggregated rows);
    (b) against the fp32 oracle (= reference arithmetic): <= 3e-2 of the update magnitude (bf16 operands).
    Temporal GCN: the similarity top-k is discontinuous; a near-tie may flip one edge, so up to 0.5 % of
    rows may exceed (a)."""
    cfg, state, blob = _setup(1, T, 21, "stress")
    g = torch.Generator().manual_seed(100 + T + B)
    v = torch.randn(B, T, 17, 128, generator=g)
    XL = torch.randn(B, T, 17, 128, generator=g)
    out = _capi.former_module(cfg, blob, 0, kind, mode, v.to(DEV), XL.to(DEV)).cpu()
    ref_q = _module_oracle(state, cfg, 0, br, kind, mode, v, XL, True)
    ref = _module_oracle(state, cfg, 0, br, kind, mode, v, XL, False)
    upd = (ref - v).abs().max().item()
    assert upd > 1e-2
    err_rows = (out - ref_q).abs().amax(dim=-1).reshape(-1) / upd
    if kind == "graph" and mode == "temporal":
        assert (err_rows > 2e-3).float().mean().item() <= 0.005, f"{(err_rows > 2e-3).float().mean()}"
        assert err_rows.median().item() <= 5e-4
    else:
        assert err_rows.max().item() <= 4e-3, f"emulated-oracle error {err_rows.max().item()}"
    err_f = ((out - ref).abs().amax(dim=-1).reshape(-1) / upd)
    assert err_f.quantile(0.99).item() <= 3e-2, f"fp32-oracle error {err_f.quantile(0.99).item()}"
    if mode == "spatial" or T <= 32:
        # the opt-in two-tiles-in-flight kernel (KASF_FLAG_TWO_TILES; bone modules through limb tiles): same arithmetic
        out2 = _capi.former_module(cfg, blob, 0, kind, mode, v.to(DEV), XL.to(DEV), two_tiles=True).cpu()
        err2 = (out2 - ref_q).abs().amax(dim=-1).reshape(-1) / upd
        if kind == "graph" and mode == "temporal":
            assert (err2 > 2e-3).float().mean().item() <= 0.005 and err2.median().item() <= 5e-4
        else:
            assert err2.max().item() <= 4e-3, f"two-tiles kernel vs emulated oracle {err2.max().item()}"
    else:
        with pytest.raises(_capi.KasfError):
            _capi.former_module(cfg, blob, 0, kind, mode, v.to(DEV), XL.to(DEV), two_tiles=True)
    if kind == "bone":
        # the path kasf_forward takes: K|V operand from the pre-normalised bf16 limb tiles (one bulk copy per tile)
        out_lt = _capi.former_module(cfg, blob, 0, kind, mode, v.to(DEV), XL.to(DEV), use_limb_tiles=True).cpu()
        err_lt = (out_lt - ref_q).abs().amax(dim=-1).reshape(-1) / upd
        assert err_lt.max().item() <= 4e-3, f"limb-tile path vs emulated oracle {err_lt.max().item()}"
        assert ((out_lt - out).abs().max() / upd).item() <= 2e-3


def test_former_module_in_place_and_default_regime():
    """in == out aliasing is allowed; default-init weights (layer scale 1e-5) give a near-identity block."""
    cfg, state, blob = _setup(1, 27, 5, "default")
    v = torch.randn(4, 27, 17, 128, generator=torch.Generator().manual_seed(1))
    vd = v.to(DEV)
    _capi.former_module(cfg, blob, 0, "attention", "temporal", vd, None, out=vd)
    ref = _module_oracle(state, cfg, 0, "att", "attention", "temporal", v, None, False)
    assert (vd.cpu() - ref).abs().max().item() <= 2e-6


def test_fusion_and_head():
    """K6 / K7 are fp32: <= 1e-5 relative (fusion), <= 2e-6 absolute on y (head; |y| ~ 0.3)."""
    cfg, state, blob = _setup(1, 27, 21, "stress")
    g = torch.Generator().manual_seed(7)
    a, b, c = (torch.randn(3, 27, 17, 128, generator=g) for _ in range(3))
    out = _capi.fusion(cfg, blob, 0, a.to(DEV), b.to(DEV), c.to(DEV)).cpu()
    ref = O.fuse(state, "layers_with_bone.0.fusion_three_channel.", a, b, c)
    assert (out - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    y, rep = _capi.head(cfg, blob, a.to(DEV), return_rep=True)
    assert (y.cpu() - O.head(state, a)).abs().max().item() <= 2e-6
    assert (rep.cpu() - O.head(state, a, return_rep=True)).abs().max().item() <= 2e-6


def test_tables_match_host_copies():
    from kasportsformer_b200 import skeleton as S
    assert _capi.table(0) == list(S.BONE_CHILD) and _capi.table(1) == list(S.BONE_PARENT)
    assert _capi.table(2) == [len(g) for g in S.LIMB_GROUPS]
    mem = _capi.table(3)
    for i, g in enumerate(S.LIMB_GROUPS):
        assert [m for m in mem[4 * i:4 * i + 4] if m >= 0] == list(g)
    adj = np.array(_capi.table(4)).reshape(17, 17)
    assert np.array_equal(adj, O.skeleton_adjacency().numpy().astype(int))
    assert _capi.table(5) == list(S.flip_permutation())
