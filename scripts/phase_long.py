"""Phase cycles (thread 0's timeline, per sequence) of the split-path mixer-core kernels.  usage: phase_long.py B T"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import _capi, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 243
cfg = dict(n_layers=1, n_frames=T, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17, neighbour_num=4)
dev = torch.device("cuda:0")
state = synthetic.make_state(cfg, 0, "default")
blob = _capi.pack_state(cfg, {k: v for k, v in state.items() if v.is_floating_point()}, dev)
v = torch.randn(B, T, 17, 128, device=dev)
xl = torch.randn(B, T, 17, 128, device=dev)
names = {0: "ln_split", 1: "similarity_wait", 2: "threshold_bits", 3: "adjacency_store_sync", 4: "rescale_rowsum", 5: "aggregation_wait",
         6: "epilogue", 8: "load", 9: "attention", 10: "store", 11: "setup"}
for kind in ("graph", "attention"):
    out = torch.empty_like(v)
    prof = torch.zeros(24, dtype=torch.int64, device=dev)
    _capi._check(_capi.lib().kasf_former_module_profiled(C.byref(_capi.c_config(cfg)), _capi._ptr(blob), 0, _capi.KIND[kind], 1,
                 _capi._ptr(v), _capi._ptr(xl), _capi._ptr(out), B, _capi._stream(), _capi._ptr(prof)), "profiled")
    torch.cuda.synchronize()
    h = prof.cpu().tolist()
    print(kind, "T", T, "sequences", B * 17, {names[i]: round(h[i] / (B * 17)) for i in names if h[i]})
