"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/kasf.h declares, agrees
with the host copies of the constant tables and the state_dict schema.  No compute call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from kasportsformer_b200 import KASportsFormer, _capi, build, load_model, AttrDict, total_parameters_count
from kasportsformer_b200 import skeleton as S
from oracle import kasf_oracle as O
from oracle import metrics_oracle as MO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(n_layers=26, n_frames=27, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17,
           neighbour_num=4)


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _capi.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "kasf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(kasf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.exported_symbols()), declared ^ set(_capi.exported_symbols())
    for name in declared:
        assert hasattr(lib, name)
    assert lib.kasf_version() == 2
    assert b"sm_100" in lib.kasf_strerror(-3)


def test_weight_image_matches_state_dict_schema(lib):
    ents = _capi.weight_entries(CFG)
    m = KASportsFormer(num_heads=8)
    sd = m.state_dict()
    floats = [(k, v.numel()) for k, v in sd.items() if v.is_floating_point()]
    assert [(n, c) for n, _, c in ents] == floats
    off = 0
    for _, o, c in ents:
        assert o == off
        off += c
    assert off == lib.kasf_weight_image_floats(C.byref(_capi.c_config(CFG))) == 29367956
    assert len(sd) == 2975 and total_parameters_count(m) == 29365668


def test_config_validation(lib):
    bad = dict(CFG, num_heads=6)
    assert lib.kasf_weight_entries(C.byref(_capi.c_config(bad))) == -2
    m4 = KASportsFormer()                      # the reference ctor's default num_heads=4 (no shipped YAML uses it)
    assert m4.cfg["num_heads"] == 4 and m4.precision == "exact"      # ... runs in the fp32 path only
    assert lib.kasf_weight_entries(C.byref(_capi.c_config(m4.cfg))) == lib.kasf_weight_entries(C.byref(_capi.c_config(CFG)))
    with pytest.raises(NotImplementedError):
        KASportsFormer(num_heads=6)
    with pytest.raises(NotImplementedError):
        KASportsFormer(num_heads=8, hierarchical=True)
    with pytest.raises(NotImplementedError):
        KASportsFormer(num_heads=8, n_frames=500)
    assert lib.kasf_packed_bytes(C.byref(_capi.c_config(CFG))) % 1024 == 0


def test_load_model_from_reference_yaml_keys():
    """All 24 model keys of the shipped YAMLs are accepted (reference model/model_tools.py:86-92)."""
    args = AttrDict(model_name="KASportsFormer", n_layers=2, dim_in=3, dim_feat=128, dim_rep=512, dim_out=3,
                    mlp_ratio=4, act_layer="gelu", attn_drop=0.0, drop=0.0, drop_path=0.0, use_layer_scale=True,
                    layer_scale_init_value=0.00001, use_adaptive_fusion=True, num_heads=8, qkv_bias=False,
                    qkv_scale=None, hierarchical=False, num_joints=17, use_temporal_similarity=True,
                    temporal_connection_len=1, use_tcn=False, graph_only=False, neighbour_num=4, n_frames=27)
    m = load_model(args)
    assert len(m.layers_with_bone) == 2
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(1, 27, 17, 3))    # no CPU path, fails loudly


def test_tables_match_host_copies(lib):
    assert _capi.table(0) == list(S.BONE_CHILD) == O.BONE_CHILD
    assert _capi.table(1) == list(S.BONE_PARENT) == O.BONE_PARENT
    assert _capi.table(2) == [len(g) for g in S.LIMB_GROUPS]
    mem = _capi.table(3)
    for i, g in enumerate(S.LIMB_GROUPS):
        assert [m for m in mem[4 * i:4 * i + 4] if m >= 0] == list(g) == O.LIMB_GROUPS[i]
    adj = np.array(_capi.table(4)).reshape(17, 17)
    assert np.array_equal(adj, O.skeleton_adjacency().numpy().astype(int))
    assert adj.sum() == 32 and np.array_equal(adj, adj.T)
    assert tuple(adj.sum(1)) == S.skeleton_degrees()
    assert _capi.table(5) == list(S.flip_permutation())
    x = np.arange(17 * 3, dtype=np.float64).reshape(1, 17, 3)
    assert np.array_equal(MO.joint_flip(x)[0, :, 1], x[0, list(S.flip_permutation()), 1])


def test_procrustes_routine_host(lib):
    """The metric kernel's Procrustes code (same source, run on the host) vs numpy SVD."""
    g = np.random.default_rng(0)
    for case in range(20):
        p = g.normal(size=(17, 3)) * 100
        t = g.normal(size=(17, 3)) * 100
        if case == 1:
            t = p.copy(); t[:, 0] *= -1            # pure reflection
        if case == 2:
            p[:, 2] = 0; t[:, 2] = 0               # planar
        if case == 3:
            t = 2.5 * p @ np.linalg.qr(g.normal(size=(3, 3)))[0] + 7
        a = lib.kasf_selftest_p_mpjpe_host(p.ctypes.data, t.ctypes.data)
        b = MO.p_mpjpe(p[None], t[None])[0]
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (case, a, b)


def test_launch_counts_and_passes(lib):
    """Host-side accounting of a forward (no GPU): 186 launches on the fused path; temporal modules of sequences
    longer than 64 frames take the split path (3 / 2 / 3 kernels instead of 1 per temporal module: 316 launches);
    a batch that does not fit the 3 GiB of streams is split into EQUAL passes (256 clips at T = 243: 2 x 128)."""
    def cfg(T):
        return dict(CFG, n_frames=T)
    assert _capi.forward_launches(cfg(27), 1024) == 1 + 2 + 26 * 7 + 1
    assert _capi.forward_launches(cfg(64), 16) == 1 + 2 + 26 * 7 + 1
    assert _capi.forward_launches(cfg(81), 256) == 1 + 2 + 26 * 12 + 1
    assert _capi.forward_launches(cfg(243), 128) == 1 + 2 + 26 * 12 + 1
    assert _capi.forward_launches(cfg(243), 256) == 2 * (1 + 2 + 26 * 12 + 1)
    # equal passes: the workspace of 256 clips at T = 243 is the one of 128 clips, not of 253
    assert _capi.workspace_bytes(cfg(243), 256) == _capi.workspace_bytes(cfg(243), 128)
    assert _capi.workspace_bytes(cfg(243), 128) > _capi.workspace_bytes(cfg(243), 64)
    # the split path needs scratch and limb tiles in (sequence, frame) order; the fused path no scratch
    c81, c27 = _capi.c_config(cfg(81)), _capi.c_config(cfg(27))
    assert lib.kasf_module_scratch_bytes(C.byref(c81), 4) > 0 and lib.kasf_module_scratch_bytes(C.byref(c27), 4) == 0
    assert lib.kasf_limb_tiles_bytes(C.byref(c81), 4, 1) == ((4 * 17 * 81 + 127) // 128) * 32768
