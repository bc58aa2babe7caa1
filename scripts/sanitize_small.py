"""Tiny forward + module calls for compute-sanitizer runs (memcheck / synccheck / racecheck).
usage: sanitize_small.py [T] [fast|exact|two_tiles]"""
import sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import KASportsFormer, _capi, synthetic
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 27
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
m = KASportsFormer(n_layers=1, num_heads=8, n_frames=T).eval()
m.load_state_dict(synthetic.make_state(dict(m.cfg), 1, "stress"), strict=True)
m = m.to(dev)
x = synthetic.make_clips(9, T, 3, "det").to(dev)
if mode == "two_tiles":
    y = _capi.forward(m.cfg, m.packed_weights(dev), x, two_tiles=True)
else:
    m.precision = mode
    y = m(x)
torch.cuda.synchronize()
print("forward ok", mode, float(y.abs().sum()))
