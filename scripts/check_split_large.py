"""Large-batch consistency check of the split path (many tiles / sequences per kernel): fast forward vs the exact
(fp32, unfused) forward of the same model on the same clips.  usage: check_split_large.py T B [layers]"""
import sys, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import KASportsFormer, synthetic
T = int(sys.argv[1]); B = int(sys.argv[2]); L = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
m = KASportsFormer(n_layers=L, num_heads=8, n_frames=T).eval()
m.load_state_dict(synthetic.make_state(dict(m.cfg), 3, "stress"), strict=True)
m = m.to(dev)
x = synthetic.make_clips(B, T, 5, "det").to(dev)
m.precision = "fast"
yf = m(x)
m.precision = "exact"
ye = m(x)
torch.cuda.synchronize()
d = (yf - ye).abs() * 960.0
per_clip = d.amax(dim=(1, 2, 3))
print(f"T={T} B={B} L={L}: fast vs exact |dy| mm: max {d.max().item():.3f} mean {d.mean().item():.4f}; per-clip max: median "
      f"{per_clip.median().item():.3f}, worst clip {per_clip.argmax().item()} ({per_clip.max().item():.3f}); finite {bool(torch.isfinite(yf).all())}")
# the error must not depend on where a clip sits in the batch (tiles / CTAs): compare the first and the last quarter
q = B // 4
print(f"   mean |dy| first quarter {d[:q].mean().item():.4f} mm, last quarter {d[-q:].mean().item():.4f} mm")
