"""ctypes binding of libkasf.so (declared in include/kasf.h).

PyTorch is used for device memory and streams only: every call passes raw device pointers and the
current CUDA stream through the C-ABI.  There is no fallback: if the shared library is missing or
the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# KASF_LIB selects another build of the same library (A/B measurements of kernel variants); never a fallback
LIB_PATH = os.environ.get("KASF_LIB") or os.path.join(_HERE, "libkasf.so")

KIND = {"attention": 0, "graph": 1, "bone": 2}
MODE = {"spatial": 0, "temporal": 1}
METRIC_COLS = 22


class KasfConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layers", "n_frames", "dim_feat", "dim_rep", "num_heads",
                                          "mlp_ratio", "num_joints", "neighbour_num")]


class KasfForwardOpts(C.Structure):
    _fields_ = [("precision", C.c_int32), ("flags", C.c_uint32), ("ctx", C.c_void_p), ("image_dev", C.c_void_p)]


PRECISION = {"fast": 0, "exact": 1}
FLAG_TWO_TILES = 1


class KasfError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

_SIGNATURES = {
    "kasf_version": (C.c_int, []),
    "kasf_strerror": (C.c_char_p, [C.c_int]),
    "kasf_device_supported": (C.c_int, []),
    "kasf_weight_entries": (C.c_int, [C.POINTER(KasfConfig)]),
    "kasf_weight_entry": (C.c_int, [C.POINTER(KasfConfig), C.c_int, C.c_char_p, C.c_size_t,
                                    C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "kasf_weight_image_floats": (C.c_size_t, [C.POINTER(KasfConfig)]),
    "kasf_packed_bytes": (C.c_size_t, [C.POINTER(KasfConfig)]),
    "kasf_pack_weights": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "kasf_workspace_bytes": (C.c_size_t, [C.POINTER(KasfConfig), C.c_int]),
    "kasf_forward": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kasf_forward_launches": (C.c_int, [C.POINTER(KasfConfig), C.c_int]),
    "kasf_forward_marks": (C.c_int, [C.POINTER(KasfConfig), C.c_int]),
    "kasf_module_scratch_bytes": (C.c_size_t, [C.POINTER(KasfConfig), C.c_int]),
    "kasf_former_module_ws": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "kasf_forward_timed": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]),
    "kasf_event_create": (C.c_void_p, []),
    "kasf_event_destroy": (None, [C.c_void_p]),
    "kasf_event_elapsed_ms": (C.c_float, [C.c_void_p, C.c_void_p]),
    "kasf_kinematic_features": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p]),
    "kasf_limb_tiles_bytes": (C.c_size_t, [C.POINTER(KasfConfig), C.c_int, C.c_int]),
    "kasf_limb_tiles": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "kasf_former_module_lt": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "kasf_former_module": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "kasf_former_module_profiled": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "kasf_ctx_create": (C.c_void_p, []),
    "kasf_ctx_destroy": (None, [C.c_void_p]),
    "kasf_workspace_bytes_ex": (C.c_size_t, [C.POINTER(KasfConfig), C.c_int, C.c_int]),
    "kasf_forward_ex": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "kasf_former_module_ex": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint32,
                                        C.c_void_p]),
    "kasf_former_module_profiled_lt": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_void_p]),
    "kasf_fusion": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "kasf_head": (C.c_int, [C.POINTER(KasfConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_int, C.c_void_p]),
    "kasf_metrics": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "kasf_joint_flip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "kasf_table": (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "kasf_test_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "kasf_selftest_p_mpjpe_host": (C.c_double, [C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load libkasf.so (once). Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise KasfError(f"{LIB_PATH} not found: build it with `python -m "
                                    "kasportsformer_b200.build` (nvcc, sm_100a). No CPU fallback exists.")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype, fn.argtypes = res, args
                _lib = l
    return _lib


def _check(code: int, what: str):
    if code != 0:
        msg = lib().kasf_strerror(code)
        raise KasfError(f"{what} failed: {msg.decode() if msg else code} ({code})")


def check_config_supported(cfg: dict):
    """Host-side mirror of the C validation (include/kasf.h, kasf_config)."""
    fixed = dict(dim_feat=128, dim_rep=512, mlp_ratio=4, num_joints=17, neighbour_num=4)
    bad = [f"{k}={cfg[k]} (built for {v})" for k, v in fixed.items() if cfg[k] != v]
    if cfg["num_heads"] not in (4, 8):
        bad.append(f"num_heads={cfg['num_heads']} (8: all kernels; 4: precision='exact' only)")
    if not (1 <= cfg["n_layers"] <= 1024):
        bad.append(f"n_layers={cfg['n_layers']}")
    if not (4 <= cfg["n_frames"] <= 243):
        bad.append(f"n_frames={cfg['n_frames']} (supported: 4..243)")
    if bad:
        raise NotImplementedError("kasportsformer_b200 kernels are specialised; unsupported: " + ", ".join(bad))


def c_config(cfg: dict) -> KasfConfig:
    return KasfConfig(**{k: int(cfg[k]) for k, _ in KasfConfig._fields_})


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_device(dev: torch.device):
    if dev.type != "cuda":
        raise KasfError("kasportsformer_b200 has no CPU path")
    with torch.cuda.device(dev):
        _check(lib().kasf_device_supported(), "kasf_device_supported")


# --- weights -----------------------------------------------------------------------------------
def weight_entries(cfg: dict):
    """[(name, offset_floats, numel)] of the canonical fp32 weight image (defined by the library)."""
    l, cc = lib(), c_config(cfg)
    n = l.kasf_weight_entries(C.byref(cc))
    if n < 0:
        _check(n, "kasf_weight_entries")
    out, buf = [], C.create_string_buffer(256)
    off, num = C.c_size_t(), C.c_size_t()
    for i in range(n):
        _check(l.kasf_weight_entry(C.byref(cc), i, buf, 256, C.byref(off), C.byref(num)), "kasf_weight_entry")
        out.append((buf.value.decode(), off.value, num.value))
    return out


def build_weight_image(cfg: dict, state: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Flatten a reference-named state dict into the canonical fp32 image (host tensor)."""
    entries = weight_entries(cfg)
    total = lib().kasf_weight_image_floats(C.byref(c_config(cfg)))
    img = torch.empty(total, dtype=torch.float32)
    names = set()
    for name, off, numel in entries:
        t = state[name]
        if t.numel() != numel:
            raise KasfError(f"{name}: expected {numel} elements, got {tuple(t.shape)}")
        img[off:off + numel] = t.detach().reshape(-1).to(dtype=torch.float32, device="cpu")
        names.add(name)
    extra = [k for k in state if k not in names and state[k].is_floating_point()]
    if extra:
        raise KasfError(f"state has tensors the library does not know: {extra[:5]}")
    return img


def pack_state(cfg: dict, state: Dict[str, torch.Tensor], device: torch.device, keep_image: bool = False):
    """The kernel-ready blob (and, with keep_image, the fp32 weight image on the device: precision="exact" reads it)."""
    _require_device(device)
    l, cc = lib(), c_config(cfg)
    # assemble on the device the tensors live on if they already are there (no host round trip)
    img = build_weight_image(cfg, state).to(device)
    nbytes = l.kasf_packed_bytes(C.byref(cc))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        _check(l.kasf_pack_weights(C.byref(cc), _ptr(img), _ptr(blob), nbytes, _stream()), "kasf_pack_weights")
        torch.cuda.current_stream().synchronize()   # img may be freed after return
    return (blob, img) if keep_image else blob


# --- forward + stages ----------------------------------------------------------------------------
# Per (device, host thread, stream) state of the callers of this module: the workspace and the forward context (two
# side streams + three events, kasf_ctx_create).  nn.DataParallel calls forward from one thread per device: the key
# keeps those callers apart, the lock protects the dictionaries themselves.
_ws_cache: Dict[tuple, torch.Tensor] = {}
_ctx_cache: Dict[tuple, "ForwardContext"] = {}
_state_lock = threading.Lock()


class ForwardContext:
    """Owner of a kasf_forward_ctx (side streams + events for the branch-parallel forward)."""

    def __init__(self, device: torch.device):
        with torch.cuda.device(device):
            self.handle = lib().kasf_ctx_create()
        if not self.handle:
            raise KasfError("kasf_ctx_create failed")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib().kasf_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _caller_key(device: torch.device):
    return (device.index, threading.get_ident(), torch.cuda.current_stream(device).cuda_stream)


def _workspace(cfg: dict, B: int, device: torch.device, precision: int = 0) -> torch.Tensor:
    n = lib().kasf_workspace_bytes_ex(C.byref(c_config(cfg)), B, precision)
    key = _caller_key(device)
    with _state_lock:
        ws = _ws_cache.get(key)
        if ws is None or ws.numel() < n:
            ws = torch.empty(n, dtype=torch.uint8, device=device)
            _ws_cache[key] = ws
    return ws


def _context(device: torch.device) -> ForwardContext:
    key = _caller_key(device)
    with _state_lock:
        ctx = _ctx_cache.get(key)
        if ctx is None:
            ctx = _ctx_cache[key] = ForwardContext(device)
    return ctx


def _launch_forward(cfg, blob, x, y, rep, ws, ctx, precision=0, flags=0, image=None):
    opts = KasfForwardOpts(precision, flags, ctx.handle if ctx is not None else None, _ptr(image))
    _check(lib().kasf_forward_ex(C.byref(c_config(cfg)), _ptr(blob), _ptr(x), _ptr(y), _ptr(rep), x.shape[0],
                                 _ptr(ws), ws.numel(), _stream(), C.byref(opts)), "kasf_forward_ex")


def forward(cfg: dict, blob: torch.Tensor, x: torch.Tensor, return_rep: bool = False, precision: str = "fast",
            image: Optional[torch.Tensor] = None, two_tiles: bool = False, ws: Optional[torch.Tensor] = None,
            ctx: Optional[ForwardContext] = None, branch_streams: bool = True) -> torch.Tensor:
    """kasf_forward_ex.  precision "exact" needs `image` (the fp32 weight image on the device, pack_state(...,
    keep_image=True)).  `ws` / `ctx`: caller-owned workspace and context (CUDA-graph capture); default: cached per
    (device, thread, stream)."""
    _require_device(x.device)
    B, T = x.shape[0], x.shape[1]
    y = torch.empty(B, T, 17, 3, dtype=torch.float32, device=x.device)
    rep = torch.empty(B, T, 17, cfg["dim_rep"], dtype=torch.float32, device=x.device) if return_rep else None
    if B == 0:
        return rep if return_rep else y
    prec = PRECISION[precision]
    if prec == 1 and image is None:
        raise KasfError('precision="exact" needs the fp32 weight image')
    with torch.cuda.device(x.device):
        if ws is None:
            ws = _workspace(cfg, B, x.device, prec)
        if ctx is None and branch_streams and prec == 0:
            ctx = _context(x.device)
        _launch_forward(cfg, blob, x, y, rep, ws, ctx, prec, FLAG_TWO_TILES if two_tiles else 0, image)
    return rep if return_rep else y


def workspace_bytes(cfg: dict, B: int, precision: str = "fast") -> int:
    return lib().kasf_workspace_bytes_ex(C.byref(c_config(cfg)), B, PRECISION[precision])


class LaunchTimer:
    """Per-stage CUDA events for `forward(..., timer=...)`: device time of every stage (features, each
    FormerModule, each fusion, head) of a forward; one stage = one kernel for n_frames <= 128."""

    def __init__(self, cfg: dict, B: int):
        self.n = lib().kasf_forward_marks(C.byref(c_config(cfg)), B) + 1
        self.events = (C.c_void_p * self.n)(*[lib().kasf_event_create() for _ in range(self.n)])

    def launch_ms(self):
        """[ms] per launch, in launch order (call after synchronising the stream)."""
        return [lib().kasf_event_elapsed_ms(self.events[i], self.events[i + 1]) for i in range(self.n - 1)]

    def close(self):
        for e in self.events:
            lib().kasf_event_destroy(e)


def forward_into(cfg: dict, blob: torch.Tensor, x: torch.Tensor, y: torch.Tensor, timer: "LaunchTimer" = None,
                 two_tiles: bool = False):
    """Forward into a preallocated output (bench loop: no allocation inside the timed region)."""
    B = x.shape[0]
    with torch.cuda.device(x.device):
        ws = _workspace(cfg, B, x.device)
        if timer is None:
            _launch_forward(cfg, blob, x, y, None, ws, _context(x.device), 0, FLAG_TWO_TILES if two_tiles else 0)
        else:
            _check(lib().kasf_forward_timed(C.byref(c_config(cfg)), _ptr(blob), _ptr(x), _ptr(y), None, B,
                                            _ptr(ws), ws.numel(), _stream(), timer.events, timer.n),
                   "kasf_forward_timed")
    return y


def forward_launches(cfg: dict, B: int) -> int:
    return lib().kasf_forward_launches(C.byref(c_config(cfg)), B)


def forward_marks(cfg: dict, B: int) -> int:
    return lib().kasf_forward_marks(C.byref(c_config(cfg)), B)


def kinematic_features(cfg, blob, x, want_raw=True):
    _require_device(x.device)
    B, T = x.shape[:2]
    f = lambda c: torch.empty(B, T, 17, c, dtype=torch.float32, device=x.device)
    bone, limb = (f(3), f(3)) if want_raw else (None, None)
    X, XB, XL = f(128), f(128), f(128)
    with torch.cuda.device(x.device):
        _check(lib().kasf_kinematic_features(C.byref(c_config(cfg)), _ptr(blob), _ptr(x), _ptr(bone),
                                             _ptr(limb), _ptr(X), _ptr(XB), _ptr(XL), B, _stream()),
               "kasf_kinematic_features")
    return bone, limb, X, XB, XL


def limb_tiles(cfg, XL, mode):
    """Normalised limb rows as bf16 operand tiles in the tile order of `mode` (None when the mode has no tile path)."""
    _require_device(XL.device)
    B = XL.shape[0]
    n = lib().kasf_limb_tiles_bytes(C.byref(c_config(cfg)), B, MODE[mode])
    if n == 0:
        return None
    tiles = torch.empty(n, dtype=torch.uint8, device=XL.device)
    with torch.cuda.device(XL.device):
        _check(lib().kasf_limb_tiles(C.byref(c_config(cfg)), _ptr(XL), _ptr(tiles), B, MODE[mode], _stream()),
               "kasf_limb_tiles")
    return tiles


def former_module(cfg, blob, layer, kind, mode, v, XL=None, out=None, use_limb_tiles=False, two_tiles=False):
    """One FormerModule.  use_limb_tiles: bone modules take their K|V operand from pre-normalised limb tiles (the
    path kasf_forward uses) instead of normalising XL inside the kernel.  two_tiles: the two-tiles-in-flight kernel
    (KASF_FLAG_TWO_TILES; bone modules then always through limb tiles)."""
    _require_device(v.device)
    B = v.shape[0]
    out = torch.empty_like(v) if out is None else out
    with torch.cuda.device(v.device):
        nscr = lib().kasf_module_scratch_bytes(C.byref(c_config(cfg)), B) if mode == "temporal" else 0
        scr = torch.empty(max(nscr, 256), dtype=torch.uint8, device=v.device)
        if two_tiles:
            lt = limb_tiles(cfg, XL, mode) if kind == "bone" else None
            _check(lib().kasf_former_module_ex(C.byref(c_config(cfg)), _ptr(blob), layer, KIND[kind], MODE[mode],
                                               _ptr(v), _ptr(XL), _ptr(lt), _ptr(out), B, _ptr(scr), nscr,
                                               FLAG_TWO_TILES, _stream()), "kasf_former_module_ex")
        elif use_limb_tiles and kind == "bone":
            lt = limb_tiles(cfg, XL, mode)
            _check(lib().kasf_former_module_lt(C.byref(c_config(cfg)), _ptr(blob), layer, KIND[kind], MODE[mode],
                                               _ptr(v), _ptr(XL), _ptr(lt), _ptr(out), B, _ptr(scr), nscr, _stream()),
                   "kasf_former_module_lt")
        else:
            _check(lib().kasf_former_module_ws(C.byref(c_config(cfg)), _ptr(blob), layer, KIND[kind], MODE[mode],
                                               _ptr(v), _ptr(XL), _ptr(out), B, _ptr(scr), nscr, _stream()),
                   "kasf_former_module_ws")
    return out


PHASES = ["limb_kv", "load_ln1", "qkv_mma_wait", "qkv_drain", "attention", "proj_mma_wait", "similarity",
          "aggregation", "v_mma_wait", "mixer_epilogue", "ln2", "mlp", "out_epilogue", "rows_wait",
          "mlp_wait", "mlp_tmem_ld", "last_fc2_wait", "first_fc1_wait"]


def former_module_phases(cfg, blob, layer, kind, mode, v, XL=None):
    """Run one module with the phase-cycle hook; returns {phase: mean cycles per tile per CTA}."""
    _require_device(v.device)
    out = torch.empty_like(v)
    prof = torch.zeros(24, dtype=torch.int64, device=v.device)
    with torch.cuda.device(v.device):
        _check(lib().kasf_former_module_profiled(C.byref(c_config(cfg)), _ptr(blob), layer, KIND[kind], MODE[mode],
                                                 _ptr(v), _ptr(XL), _ptr(out), v.shape[0], _stream(), _ptr(prof)),
               "kasf_former_module_profiled")
        torch.cuda.synchronize()
    B, T = v.shape[0], v.shape[1]
    tiles = (B * T + 6) // 7 if mode == "spatial" else (B * 17 + (128 // T) - 1) // (128 // T)
    return {n: prof[i].item() / tiles for i, n in enumerate(PHASES)}, tiles


PHASES_V2 = {0: "rows_wait", 1: "ln1", 2: "qkv_wait_drain", 3: "mixer_core", 4: "epilogue_waits", 5: "mixer_epilogue",
             8: "x1_wait", 9: "ln2", 10: "fc1_wait", 11: "gelu", 12: "gelu_buffer_wait", 13: "fc2_wait", 14: "out_epilogue"}


def former_module_phases_v2(cfg, blob, layer, kind, mode, v, XL=None):
    """Phase cycles of the two-tiles-in-flight kernel (bone modules through limb tiles): mean cycles per tile for
    the mixer group (slots 0-5) and the MLP group (slots 8-14)."""
    _require_device(v.device)
    out = torch.empty_like(v)
    prof = torch.zeros(32 + 2 * 6 * 600, dtype=torch.int64, device=v.device)
    with torch.cuda.device(v.device):
        lt = limb_tiles(cfg, XL, mode) if kind == "bone" else None
        _check(lib().kasf_former_module_profiled_lt(C.byref(c_config(cfg)), _ptr(blob), layer, KIND[kind], MODE[mode],
                                                    _ptr(v), _ptr(XL), _ptr(lt), _ptr(out), v.shape[0], _stream(),
                                                    _ptr(prof)), "kasf_former_module_profiled_lt")
        torch.cuda.synchronize()
    B, T = v.shape[0], v.shape[1]
    tiles = (B * T + 6) // 7 if mode == "spatial" else (B * 17 + (128 // T) - 1) // (128 // T)
    h = prof.cpu()
    former_module_phases_v2.trace = [e for r in range(6)
                                     for e in h[32 + 2 * 600 * r:32 + 2 * (600 * r + int(h[24 + r]))].reshape(-1, 2).tolist()]
    return {n: prof[i].item() / tiles for i, n in PHASES_V2.items()}, tiles


def fusion(cfg, blob, layer, a, g, b):
    _require_device(a.device)
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _check(lib().kasf_fusion(C.byref(c_config(cfg)), _ptr(blob), layer, _ptr(a), _ptr(g), _ptr(b),
                                 _ptr(out), a.shape[0], _stream()), "kasf_fusion")
    return out


def head(cfg, blob, X, return_rep=False):
    _require_device(X.device)
    B, T = X.shape[:2]
    y = torch.empty(B, T, 17, 3, dtype=torch.float32, device=X.device)
    rep = torch.empty(B, T, 17, 512, dtype=torch.float32, device=X.device) if return_rep else None
    with torch.cuda.device(X.device):
        _check(lib().kasf_head(C.byref(c_config(cfg)), _ptr(blob), _ptr(X), _ptr(y), _ptr(rep), B, _stream()),
               "kasf_head")
    return (y, rep) if return_rep else y


def joint_flip(x: torch.Tensor) -> torch.Tensor:
    _require_device(x.device)
    x = x.contiguous()
    out = torch.empty_like(x)
    frames = x.numel() // 51
    with torch.cuda.device(x.device):
        _check(lib().kasf_joint_flip(_ptr(x), _ptr(out), frames, _stream()), "kasf_joint_flip")
    return out


def metrics(pred, gt, res, factor, actions, n_actions, pred_flip=None, sums=None, want_per_frame=False):
    _require_device(pred.device)
    B, T = pred.shape[:2]
    dev = pred.device
    if sums is None:
        sums = torch.zeros(n_actions, METRIC_COLS, dtype=torch.float64, device=dev)
    per = torch.empty(B, T, 3, dtype=torch.float64, device=dev) if want_per_frame else None
    with torch.cuda.device(dev):
        _check(lib().kasf_metrics(T, _ptr(pred), _ptr(pred_flip), _ptr(gt), _ptr(res), _ptr(factor),
                                  _ptr(actions), n_actions, _ptr(sums), _ptr(per), B, _stream()),
               "kasf_metrics")
    return (sums, per) if want_per_frame else sums


def table(which: int):
    buf = (C.c_int32 * 512)()
    n = lib().kasf_table(which, buf, 512)
    if n < 0:
        _check(n, "kasf_table")
    return list(buf[:n])


def test_gemm(a: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    _require_device(a.device)
    M, N = a.shape[0], w.shape[0]
    d = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _check(lib().kasf_test_gemm(_ptr(a), _ptr(w), _ptr(d), M, N, _stream()), "kasf_test_gemm")
    return d
