"""Tiny forward + module calls for compute-sanitizer runs (memcheck / synccheck / racecheck)."""
import sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import KASportsFormer, _capi, synthetic
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 27
m = KASportsFormer(n_layers=1, num_heads=8, n_frames=T).eval()
m.load_state_dict(synthetic.make_state(dict(m.cfg), 1, "stress"), strict=True)
m = m.to(dev)
x = synthetic.make_clips(9, T, 3, "det").to(dev)
y = m(x)
torch.cuda.synchronize()
print("forward ok", float(y.abs().sum()))
