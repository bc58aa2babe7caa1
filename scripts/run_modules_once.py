"""Launch each of the six FormerModule kernels (B clips, T frames) a few times: the target of ncu captures
   ncu --set full --clock-control none --import-source on -k regex:former_module -s 6 -c 6 -o out python scripts/run_modules_once.py"""
import sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import _capi, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 27
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = dict(n_layers=1, n_frames=T, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17, neighbour_num=4)
dev = torch.device("cuda:0")
state = synthetic.make_state(cfg, 0, "default")
blob = _capi.pack_state(cfg, {k: v for k, v in state.items() if v.is_floating_point()}, dev)
v = torch.randn(B, T, 17, 128, device=dev)
xl = torch.randn(B, T, 17, 128, device=dev)
lts = {m: _capi.limb_tiles(cfg, xl, m) for m in ("spatial", "temporal")}
torch.cuda.synchronize()
for _ in range(reps):
    for kind in ("attention", "graph", "bone"):
        for mode in ("spatial", "temporal"):
            _capi.former_module(cfg, blob, 0, kind, mode, v, xl, use_limb_tiles=True)
torch.cuda.synchronize()
print("done")
