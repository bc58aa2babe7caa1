"""Forward latency at small batches: stream launches vs CUDA-graph replay of the same kasf_forward call."""
import sys, time, torch
sys.path.insert(0, ".")
from kasportsformer_b200 import KASportsFormer, _capi, synthetic
dev = torch.device("cuda:0")
m = KASportsFormer(n_layers=26, num_heads=8, n_frames=27).eval().to(dev)
blob = m.packed_weights(dev)
for B in (1, 4, 16, 64, 256):
    x = synthetic.make_clips(B, 27, 1, "det").to(dev)
    for _ in range(3):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = m(x)
    e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        y = m(x)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 20 * 1e3
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        yg = _capi.forward(m.cfg, blob, x)          # warm on the capture stream (workspace allocation)
        torch.cuda.current_stream().synchronize()
        with torch.cuda.graph(g, stream=s):
            yg = _capi.forward(m.cfg, blob, x)
    torch.cuda.current_stream().wait_stream(s)
    g.replay(); torch.cuda.synchronize()
    ok = torch.equal(yg, y)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 20
    print(f"B={B:4d}  stream launches {plain:7.3f} ms (wall {wall:7.3f})   graph replay {graph:7.3f} ms   equal={ok}")
