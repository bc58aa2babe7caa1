#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: clips/s of the KASportsFormer inference forward
(T=27, J=17) on N B200s, plus the roofline of the dominant kernel and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--frames T]

One "step" = one forward over a batch of `--batch` (default 1024, BASELINE.json configs[1]) synthetic
clips per GPU.  N>1 is launched by torchrun (one rank per GPU); clips are batch-sharded, the forward has
no collective, and the step ends with the all_gather of the per-rank MPJPE partial sums (the only
exchange of the path).  Rank 0 prints ONE JSON line.

  value ...... whole-job clips/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e ........ the same through the drop-in nn.Module with HOST buffers: pinned x -> H2D -> forward -> D2H y
  roofline ... the fused FormerModule kernels (156 of the 186 launches per forward, >95 % of the time):
               algorithmic FLOPs / per-launch device time, against the measured sustained bf16 peak.
               The production forward runs the three branches of a layer on three streams, where "the duration of a
               launch" is not defined, so the same K steps are repeated right after the timed region with a CUDA
               event after every launch (branches serialised on one stream; `serialised_ms_per_step`) and the
               per-launch times come from that pass
  cpu_baseline oracle port of the reference forward timed on the host cores (bounded sample)
  gpu_eager_baseline (N = 1) the same eager PyTorch forward (oracle port: the ATen calls the reference issues -> cuBLAS /
               ATen kernels) on the SAME B200, same B x T input, in fp32, with TF32 allowed, and under bf16 autocast:
               the bar a hand-written path has to beat on this box
  sweep ...... clips/s of the forward at T = 81 and T = 243 (256 clips per GPU), same process, every N
               (BASELINE.json configs[3])
  parity ..... max |dy| and the MPJPE difference in mm against the reference golden with trained-like magnitudes
               (tests/golden/trained_like_L26_T27.npz) for precision "fast" (what `value` measures) and "exact",
               and the exact mode's clips/s

`--impl reference` times the reference's CPU implementation of the path: /root/reference does not travel to
the GPU box and is pure Python/PyTorch, so this arm runs the oracle port (oracle/kasf_oracle.py, the same
ATen calls the reference issues) with all host threads; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n_layers=26, n_frames=27, dim_feat=128, dim_rep=512, num_heads=8, mlp_ratio=4, num_joints=17,
           neighbour_num=4)
METRIC = "clips/sec KASportsFormer fwd (T=27,J=17) at 1/2/4/8 B200; MPJPE Δ vs ref"


# ------------------------------------------------------------------ algorithmic work
def module_macs_per_token(kind: str, mode: str, T: int) -> int:
    n = 17 if mode == "spatial" else T
    mlp = 2 * 128 * 512
    if kind == "graph":
        return 2 * 128 * 128 + mlp + n * 128 + (T * 128 if mode == "temporal" else 0)
    return 3 * 128 * 128 + 128 * 128 + mlp + 2 * n * 128           # qkv (or q + kv) + proj + mlp + QK^T + PV


def flops_per_clip(T: int, n_layers: int = 26) -> float:
    per_layer = sum(module_macs_per_token(k, m, T) for k in ("attention", "graph", "bone") for m in ("spatial", "temporal"))
    per_layer += 3 * 384                                             # fusion
    outside = 3 * 3 * 128 + 203 + 128 * 512 + 512 * 3                # embeds, limb MLPs, rep_logit, head
    return 2.0 * (n_layers * per_layer + outside) * 17 * T


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops_sustained"], src="measured (MEASURED_PEAKS.json, sustained)")
    return dict(hbm=6650.0, bf16=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------ clocks sampler
class Clocks(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append((float(f[0]), float(f[1]), f[2:]))
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm (oracle port of the reference)
def cpu_forward_rate(T: int, sample_B: int, repeats: int):
    from kasportsformer_b200 import synthetic
    from oracle import kasf_oracle as O
    cfg = dict(CFG, n_frames=T)
    torch.set_num_threads(os.cpu_count() or 1)
    state = synthetic.make_state(cfg, 0, "default")
    x = synthetic.make_clips(sample_B, T, 0, "det")
    ocfg = O.default_config(n_frames=T)
    with torch.no_grad():
        O.forward(state, x[:2], ocfg)                               # warm-up (thread pools, allocator)
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.forward(state, x, ocfg)
            best = min(best, time.perf_counter() - t0)
    return sample_B / best, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, sample_B = args.frames, 16
    torch.set_num_threads(os.cpu_count() or 1)
    from kasportsformer_b200 import synthetic
    from oracle import kasf_oracle as O
    cfg = dict(CFG, n_frames=T)
    state = synthetic.make_state(cfg, 0, "default")
    x = synthetic.make_clips(sample_B, T, 0, "det")
    ocfg = O.default_config(n_frames=T)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            O.forward(state, x[:4], ocfg)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.forward(state, x, ocfg)
        dt = time.perf_counter() - t0
    v = sample_B * args.steps / dt
    cores = os.cpu_count() or 1
    sample = f"{sample_B} clips/step of the B={args.batch} workload, {args.steps} steps, torch CPU fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"KASportsFormer sportspose-det T={T} J=17 forward, batch {args.batch} per GPU "
                               "(BASELINE.json configs[1]), default-init weights", "frames": T, "batch_per_gpu": args.batch,
                   "parallelism": f"dp{args.gpus} (batch-sharded clips, all_gather of MPJPE sums)",
                   "l2": f"per-step working set {6 * args.batch * T * 17 * 512 / 1e6:.0f} MB of fp32 streams > 126 MB L2 (no flush needed)"},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------ same-box eager baseline (oracle port on the GPU)
def gpu_eager_rates(T: int, B: int, dev):
    """Eager PyTorch forward of the reference's op sequence on the same GPU: clips/s in fp32 (TF32 off: torch's default
    for matmul), with TF32 allowed, and under bf16 autocast.  CUDA events, 1 warm-up + 2 timed forwards each."""
    from kasportsformer_b200 import synthetic
    from oracle import kasf_oracle as O
    cfg = dict(CFG, n_frames=T)
    state = {k: v.to(dev) for k, v in synthetic.make_state(cfg, 0, "default").items()}
    x = synthetic.make_clips(B, T, 0, "det").to(dev)
    ocfg = O.default_config(n_frames=T)
    out = {}

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            fn()
        b.record()
        torch.cuda.synchronize()
        return 2 * B / (a.elapsed_time(b) / 1e3)

    old = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.no_grad():
            torch.backends.cuda.matmul.allow_tf32 = False
            out["fp32"] = timed(lambda: O.forward(state, x, ocfg))
            torch.backends.cuda.matmul.allow_tf32 = True
            out["tf32"] = timed(lambda: O.forward(state, x, ocfg))
            torch.backends.cuda.matmul.allow_tf32 = False

            def ac():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    O.forward(state, x, ocfg)
            out["bf16_autocast"] = timed(ac)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return {"unit": "clips/s", "batch": B, "frames": T, **{k: round(v, 1) for k, v in out.items()},
            "what": "oracle port of the reference forward (eager ATen / cuBLAS) on the same B200, inputs resident, CUDA events"}


def parity_report(dev):
    """Both precision modes against the reference golden with trained-like magnitudes (26 layers, T = 27)."""
    import numpy as np
    from kasportsformer_b200 import KASportsFormer, synthetic
    from oracle import metrics_oracle as MO
    p = os.path.join(ROOT, "tests", "golden", "trained_like_L26_T27.npz")
    if not os.path.exists(p):
        return None
    z = np.load(p)
    meta = json.loads(str(z["meta"]))
    cfg = meta["cfg"]
    m = KASportsFormer(n_layers=cfg["n_layers"], num_heads=8, n_frames=cfg["n_frames"])
    m.load_state_dict(synthetic.make_state(cfg, meta["seed"], meta["regime"]))
    m = m.to(dev).eval()
    x = synthetic.make_clips(meta["B"], 27, meta["clip_seed"], meta["kind"]).to(dev)
    gt, factor, res, _ = synthetic.make_labels(meta["B"], 27, seed=meta["label_seed"], n_actions=1)
    mm = 1920.0 / 2.0
    rep = {"golden": "tests/golden/trained_like_L26_T27.npz (unmodified reference, layer scales 0.05-0.15)",
           "unit": "mm (normalised units x 960: res_w = 1920, factor 1)"}
    for mode in ("fast", "exact"):
        m.precision = mode
        y = m(x).cpu().numpy()
        r = MO.evaluate(y, res.numpy().astype(np.float64), factor.numpy(), gt.numpy())
        rep[mode] = {"max_abs_dy_mm": round(float(np.abs(y - z["y"]).max() * mm), 5),
                     "mean_abs_dy_mm": round(float(np.abs(y - z["y"]).mean() * mm), 5),
                     "d_mpjpe_mm": round(abs(r["mpjpe"] - float(z["mpjpe"])), 6),
                     "d_p_mpjpe_mm": round(abs(r["p_mpjpe"] - float(z["p_mpjpe"])), 6)}
    # exact-mode throughput (64 clips, default-init 26 layers)
    m.precision = "exact"
    xb = synthetic.make_clips(64, 27, 1, "det").to(dev)
    m(xb)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m(xb)
    b.record()
    torch.cuda.synchronize()
    rep["exact_clips_per_s"] = round(64 / (a.elapsed_time(b) / 1e3), 1)
    return rep


# ------------------------------------------------------------------ GPU arm
def sweep_point(T: int, B: int, dev, world, dist, steps: int = 3):
    """clips/s of the forward at another sequence length (inputs resident, CUDA events, max over ranks)."""
    from kasportsformer_b200 import KASportsFormer, _capi, synthetic
    cfg = dict(CFG, n_frames=T)
    model = KASportsFormer(n_layers=26, num_heads=8, n_frames=T)
    model.load_state_dict(synthetic.make_state(cfg, 0, "default"))
    model = model.to(dev).eval()
    blob = model.packed_weights(dev)
    rank = int(os.environ.get("RANK", "0"))
    x = synthetic.make_clips(B, T, seed=rank, kind="det").to(dev)
    y = torch.empty(B, T, 17, 3, device=dev)
    for _ in range(3):
        _capi.forward_into(cfg, blob, x, y, None)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        _capi.forward_into(cfg, blob, x, y, None)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    v = world * B / (ms / 1e3)
    pk = peaks()
    del model, blob, x, y
    torch.cuda.empty_cache()
    return {"frames": T, "batch_per_gpu": B, "value": round(v, 1), "unit": "clips/s", "ms_per_step": round(ms, 3),
            "steps": steps, "whole_forward_frac": round(v / world * flops_per_clip(T) / 1e12 / pk["bf16"], 4)}


def run_ours(args):
    import torch.distributed as dist
    from kasportsformer_b200 import KASportsFormer, _capi, synthetic
    from kasportsformer_b200.evaluate import gather_sums

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # keeps NCCL's banner off stdout (one JSON line)
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    T, B = args.frames, args.batch
    cfg = dict(CFG, n_frames=T)

    model = KASportsFormer(n_layers=26, num_heads=8, n_frames=T)
    model.load_state_dict(synthetic.make_state(cfg, 0, "default"))
    model = model.to(dev).eval()
    blob = model.packed_weights(dev)
    # every rank gets its own shard of a global synthetic clip set (weak scaling: B clips per GPU)
    x_host = synthetic.make_clips(B, T, seed=rank, kind="det").pin_memory()
    gt, factor, res, actions = [t.to(dev) for t in synthetic.make_labels(B, T, seed=rank, n_actions=4)]
    x = x_host.to(dev)
    y = torch.empty(B, T, 17, 3, device=dev)
    y_host = torch.empty(B, T, 17, 3).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timers = [_capi.LaunchTimer(cfg, B) for _ in range(args.steps)]
    n_launch = _capi.forward_launches(cfg, B)
    n_marks = _capi.forward_marks(cfg, B)

    def step(timer):
        _capi.forward_into(cfg, blob, x, y, timer)
        sums = _capi.metrics(y, gt, res, factor, actions, 4)
        return gather_sums(sums)

    for _ in range(max(args.warmup, 3)):
        step(None)
    barrier()
    clk = Clocks(local)
    clk.start()
    # ---- timed region: K steps of the production forward (the graph and bone branches of a layer run on side
    #      streams next to the attention branch) + metric reduction, inputs resident in HBM
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(None)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- the same K steps once more with an event after every launch (this serialises the branches on one stream,
    #      so a launch duration is well defined): per-launch device times for the roofline of the dominant kernel
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        step(timers[i])
    g1.record()
    barrier()
    ms_serial = g0.elapsed_time(g1)
    per_launch = [0.0] * n_marks
    for tm in timers:
        for i, v in enumerate(tm.launch_ms()):
            per_launch[i] += v / args.steps
    clocks = clk.summary()

    # ---- end-to-end through the public nn.Module with host buffers
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        yd = model(xd)
        y_host.copy_(yd, non_blocking=True)
    for _ in range(3):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, args.steps // 2)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- the other sequence lengths of BASELINE.json configs[3], every N (all ranks take part)
    sweep = []
    if T == 27 and not args.no_sweep:
        del model, x, y
        torch.cuda.empty_cache()
        for Ts, Bs in ((81, 256), (243, 256)):
            sweep.append(sweep_point(Ts, Bs, dev, world, dist if world > 1 else None))

    if rank == 0:
        pk = peaks()
        value = world * B * args.steps / (ms / 1e3)
        e2e = world * B * e2e_steps / (ms_e2e / 1e3)
        # launch order: features, 26 x (att_s, att_t, graph_s, graph_t, bone_s, bone_t, fusion), head
        tokens = B * T * 17
        names = ["att_s", "att_t", "graph_s", "graph_t", "bone_s", "bone_t", "fusion"]
        kinds = [("attention", "spatial"), ("attention", "temporal"), ("graph", "spatial"), ("graph", "temporal"),
                 ("bone", "spatial"), ("bone", "temporal")]
        per_kind_ms = {n: 0.0 for n in names}
        # kasf_forward micro-batches large B: every pass contributes 1 (features) + 26 x 7 + 1 (head) marks
        per_pass = 1 + 26 * 7 + 1
        passes = n_marks // per_pass
        feat_ms = sum(per_launch[ps * per_pass] for ps in range(passes))
        head_ms = sum(per_launch[ps * per_pass + per_pass - 1] for ps in range(passes))
        for ps in range(passes):
            for l in range(26):
                for i, n in enumerate(names):
                    per_kind_ms[n] += per_launch[ps * per_pass + 1 + l * 7 + i]
        mod_ms = sum(per_kind_ms[n] for n in names[:6])
        mod_flops = 26 * sum(2.0 * module_macs_per_token(k, m, T) * tokens for k, m in kinds)
        achieved = mod_flops / (mod_ms / 1e3) / 1e12
        step_ms = sum(per_launch)
        # DRAM traffic of the same kernels from the committed `ncu --set full` capture (per launch, mean of the six
        # instantiations); only valid for the configuration the capture was taken on
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj["batch"] == B and tj["frames"] == T:
                per = [v["dram_read_bytes"] + v["dram_write_bytes"] for v in tj["per_launch"].values()]
                traffic, traffic_src = sum(per) / len(per), tj["capture"]
        # algorithmic HBM bytes per launch: 512 B read + 512 B written per token (+ 256 B of the bf16 limb tile in the
        # two bone modules): mean over the six instantiations
        alg_bytes = tokens / passes * (512 * 2 * 6 + 256 * 2) / 6
        out = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"KASportsFormer sportspose-det T={T} J=17 forward, batch {B} per GPU "
                                   "(BASELINE.json configs[1]), default-init weights", "frames": T, "batch_per_gpu": B,
                       "parallelism": f"dp{world} (batch-sharded clips, all_gather of MPJPE sums)",
                       "l2": f"per-step working set {6 * tokens * 512 / 1e6:.0f} MB of fp32 streams > 126 MB L2 (no flush needed)"},
            "e2e": {"value": e2e, "unit": "clips/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4},
            "gpu_launches": (n_launch + 1) * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                         "frac": achieved / pk["bf16"], "traffic": traffic, "traffic_unit": "bytes/launch (dram read+write, ncu)",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_flops_per_launch": mod_flops / (156 * passes), "peak_source": pk["src"],
                         "kernel": f"former_module_kernel<KIND,MODE> (6 instantiations, {156 * passes} launches/forward)",
                         "share_of_step": mod_ms / step_ms,
                         "serialised_ms_per_step": ms_serial / args.steps,
                         "per_kind_ms_per_forward": {k: round(v, 4) for k, v in per_kind_ms.items()},
                         "other_ms": {"features": round(feat_ms, 4), "head": round(head_ms, 4)},
                         "whole_forward_frac": value / world * flops_per_clip(T) / 1e12 / pk["bf16"]},
        }
        if sweep:
            out["sweep"] = sweep
        if world == 1 and not args.no_extras:
            out["gpu_eager_baseline"] = gpu_eager_rates(T, B, dev)
            out["parity"] = parity_report(dev)
        if world == 1 and not args.no_cpu:
            v, secs = cpu_forward_rate(T, 32, 2)
            out["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"32 clips of the same workload, best of 2 forwards ({secs:.1f} s each), oracle port, torch CPU fp32"}
        print(json.dumps(out))
    for tm in timers:
        tm.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=27)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the T = 81 / 243 sweep points")
    ap.add_argument("--no-extras", action="store_true", help="skip the same-box eager baseline and the parity report")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
