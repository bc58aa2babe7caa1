// kasf_api.cu -- extern "C" entry points of libkasf.so (declared in include/kasf.h) and the
// orchestration of the forward pass (reference model/KASportsFormer.py:320-347).
#include <cstdlib>
#include <new>

#include "kasf_internal.h"

using namespace kasf;

namespace {

int device_ok() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return KASF_EARCH;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return KASF_EARCH;
    return major == 10 ? KASF_OK : KASF_EARCH;
}

// clips processed per pass of kasf_forward: bounds the workspace (6 fp32 streams of [chunk*T*17,128]).
int clip_chunk(const kasf_config* cfg, int B) {
    const long long per_clip = (long long)cfg->n_frames * J * D * 4 * 6;
    long long c = (3LL << 30) / per_clip;   // ~3 GiB of streams
    if (c < 1) c = 1;
    if (c >= B) return B > 0 ? B : 1;
    const long long passes = (B + c - 1) / c;            // equal passes: a 3-clip remainder pass costs a full set of launches
    return (int)((B + passes - 1) / passes);
}

struct Streams {
    float *X, *XB, *XL, *A, *G, *Bn;
};
Streams carve(void* ws, long long tokens) {
    const size_t n = ((size_t)tokens * D * 4 + 1023) / 1024 * 1024;
    uint8_t* p = static_cast<uint8_t*>(ws);
    Streams s;
    s.X = (float*)(p), s.XB = (float*)(p + n), s.XL = (float*)(p + 2 * n);
    s.A = (float*)(p + 3 * n), s.G = (float*)(p + 4 * n), s.Bn = (float*)(p + 5 * n);
    return s;
}

}  // namespace

// side streams + events of one caller (device, host thread): see kasf.h
struct kasf_forward_ctx {
    cudaStream_t s[2];
    cudaEvent_t e[3];
};

extern "C" {

kasf_forward_ctx* kasf_ctx_create(void) {
    kasf_forward_ctx* c = new (std::nothrow) kasf_forward_ctx();
    if (!c) return nullptr;
    bool ok = true;
    for (int i = 0; i < 2; ++i) c->s[i] = nullptr;
    for (int i = 0; i < 3; ++i) c->e[i] = nullptr;
    for (int i = 0; i < 2 && ok; ++i) ok = cudaStreamCreateWithFlags(&c->s[i], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i) ok = cudaEventCreateWithFlags(&c->e[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        kasf_ctx_destroy(c);
        return nullptr;
    }
    return c;
}

void kasf_ctx_destroy(kasf_forward_ctx* c) {
    if (!c) return;
    for (int i = 0; i < 3; ++i)
        if (c->e[i]) cudaEventDestroy(c->e[i]);
    for (int i = 0; i < 2; ++i)
        if (c->s[i]) cudaStreamDestroy(c->s[i]);
    delete c;
}

int kasf_version(void) { return KASF_VERSION; }

const char* kasf_strerror(int code) {
    switch (code) {
        case KASF_OK: return "ok";
        case KASF_EINVAL: return "invalid argument";
        case KASF_ESHAPE: return "shape/config not implemented by this build";
        case KASF_EARCH: return "device is not sm_100 (no fallback path exists)";
        case KASF_ENOMEM: return "caller-provided buffer too small";
        default:
            if (code <= -1000) return cudaGetErrorString((cudaError_t)(-code - 1000));
            return "unknown error";
    }
}

int kasf_device_supported(void) { return device_ok(); }

int kasf_weight_entries(const kasf_config* cfg) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    int n = 0;
    walk_image(cfg, nullptr, nullptr, [&](const char*, size_t, size_t) { ++n; });
    return n;
}

int kasf_weight_entry(const kasf_config* cfg, int index, char* name, size_t name_cap, size_t* offset_floats,
                      size_t* numel) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!name || !offset_floats || !numel || index < 0) return KASF_EINVAL;
    int n = 0;
    bool found = false;
    walk_image(cfg, nullptr, nullptr, [&](const char* nm, size_t off, size_t cnt) {
        if (n++ == index) {
            found = true;
            snprintf(name, name_cap, "%s", nm);
            *offset_floats = off;
            *numel = cnt;
        }
    });
    return found ? KASF_OK : KASF_EINVAL;
}

size_t kasf_weight_image_floats(const kasf_config* cfg) {
    if (config_ok(cfg)) return 0;
    return walk_image(cfg, nullptr, nullptr, [](const char*, size_t, size_t) {});
}

size_t kasf_packed_bytes(const kasf_config* cfg) { return config_ok(cfg) ? 0 : packed_bytes(cfg); }

int kasf_pack_weights(const kasf_config* cfg, const float* image_dev, void* packed_dev, size_t packed_cap,
                      void* stream) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!image_dev || !packed_dev) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return pack_weights(cfg, image_dev, packed_dev, packed_cap, (cudaStream_t)stream);
}

size_t kasf_workspace_bytes(const kasf_config* cfg, int B) {
    if (config_ok(cfg) || B <= 0) return 0;
    const int chunk = clip_chunk(cfg, B);
    const long long tokens = (long long)chunk * cfg->n_frames * J;
    return 6 * (((size_t)tokens * D * 4 + 1023) / 1024 * 1024) + 3 * module_scratch_bytes(chunk, cfg->n_frames) +
           limb_tiles_bytes(chunk, cfg->n_frames, KASF_MODE_SPATIAL) + limb_tiles_bytes(chunk, cfg->n_frames, KASF_MODE_TEMPORAL);
}

size_t kasf_workspace_bytes_ex(const kasf_config* cfg, int B, int precision) {
    if (precision == KASF_PRECISION_EXACT) return (config_ok(cfg) || B <= 0) ? 0 : exact_workspace_bytes(cfg, B);
    return kasf_workspace_bytes(cfg, B);
}

size_t kasf_module_scratch_bytes(const kasf_config* cfg, int B) {
    if (config_ok(cfg) || B <= 0) return 0;
    return module_scratch_bytes(B, cfg->n_frames);
}

int kasf_forward_marks(const kasf_config* cfg, int B) {
    if (config_ok(cfg) || B <= 0) return 0;
    const int chunk = clip_chunk(cfg, B);
    const int passes = (B + chunk - 1) / chunk;
    return passes * (1 + cfg->n_layers * 7 + 1);
}

int kasf_forward_launches(const kasf_config* cfg, int B) {
    if (config_ok(cfg) || B <= 0) return 0;
    const int chunk = clip_chunk(cfg, B);
    const int passes = (B + chunk - 1) / chunk;
    // T > KASF_SPLIT_T: a temporal module is 3 kernels (attention, bone) or 2 (graph) instead of 1
    const int per_layer = cfg->n_frames > KASF_SPLIT_T ? 7 + 2 + 1 + 2 : 7;
    return passes * (1 + 2 /* limb_tiles_kernel, spatial + temporal */ + cfg->n_layers * per_layer + 1);
}

int kasf_kinematic_features(const kasf_config* cfg, const void* packed_dev, const float* x_dev, float* bone_dev,
                            float* limb_dev, float* X_dev, float* XB_dev, float* XL_dev, int B, void* stream) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !x_dev || !X_dev || !XB_dev || !XL_dev || B < 0) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_features((const uint8_t*)packed_dev, x_dev, bone_dev, limb_dev, X_dev, XB_dev, XL_dev,
                           (long long)B * cfg->n_frames, (cudaStream_t)stream);
}

int kasf_former_module_ws(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                          const float* in_dev, const float* XL_dev, float* out_dev, int B, void* scratch_dev,
                          size_t scratch_bytes, void* stream) {
    int rc = fast_config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !in_dev || !out_dev || B < 0 || layer < 0 || layer >= cfg->n_layers) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_former_module((const uint8_t*)packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B,
                                cfg->n_frames, (cudaStream_t)stream, nullptr, scratch_dev, scratch_bytes);
}

size_t kasf_limb_tiles_bytes(const kasf_config* cfg, int B, int mode) {
    if (config_ok(cfg) || B <= 0 || mode < 0 || mode > 1) return 0;
    return limb_tiles_bytes(B, cfg->n_frames, mode);
}

int kasf_limb_tiles(const kasf_config* cfg, const float* XL_dev, void* limb_tiles_dev, int B, int mode, void* stream) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!XL_dev || !limb_tiles_dev || B < 0 || mode < 0 || mode > 1) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_limb_tiles(XL_dev, limb_tiles_dev, B, cfg->n_frames, mode, (cudaStream_t)stream);
}

int kasf_former_module_lt(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                          const float* in_dev, const float* XL_dev, const void* limb_tiles_dev, float* out_dev, int B,
                          void* scratch_dev, size_t scratch_bytes, void* stream) {
    int rc = fast_config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !in_dev || !out_dev || B < 0 || layer < 0 || layer >= cfg->n_layers) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_former_module((const uint8_t*)packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B,
                                cfg->n_frames, (cudaStream_t)stream, nullptr, scratch_dev, scratch_bytes, limb_tiles_dev);
}

int kasf_former_module_ex(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                          const float* in_dev, const float* XL_dev, const void* limb_tiles_dev, float* out_dev, int B,
                          void* scratch_dev, size_t scratch_bytes, uint32_t flags, void* stream) {
    int rc = fast_config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !in_dev || !out_dev || B < 0 || layer < 0 || layer >= cfg->n_layers) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_former_module((const uint8_t*)packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B,
                                cfg->n_frames, (cudaStream_t)stream, nullptr, scratch_dev, scratch_bytes, limb_tiles_dev, flags);
}

int kasf_former_module(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                       const float* in_dev, const float* XL_dev, float* out_dev, int B, void* stream) {
    return kasf_former_module_ws(cfg, packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B, nullptr, 0, stream);
}

int kasf_former_module_profiled(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                                const float* in_dev, const float* XL_dev, float* out_dev, int B, void* stream,
                                unsigned long long* phase_cycles_dev) {
    int rc = fast_config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !in_dev || !out_dev || !phase_cycles_dev || B < 0 || layer < 0 || layer >= cfg->n_layers)
        return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    // split path (debug hook only): the scratch area is allocated here for the duration of the call
    void* scr = nullptr;
    const size_t scr_bytes = mode == KASF_MODE_TEMPORAL ? module_scratch_bytes(B, cfg->n_frames) : 0;
    if (scr_bytes && cudaMalloc(&scr, scr_bytes) != cudaSuccess) return KASF_ENOMEM;
    rc = launch_former_module((const uint8_t*)packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B,
                              cfg->n_frames, (cudaStream_t)stream, phase_cycles_dev, scr, scr_bytes);
    if (scr) {
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(scr);
    }
    return rc;
}

int kasf_former_module_profiled_lt(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                                   const float* in_dev, const float* XL_dev, const void* limb_tiles_dev, float* out_dev,
                                   int B, void* stream, unsigned long long* phase_cycles_dev) {
    int rc = fast_config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !in_dev || !out_dev || !phase_cycles_dev || B < 0 || layer < 0 || layer >= cfg->n_layers)
        return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_former_module((const uint8_t*)packed_dev, layer, kind, mode, in_dev, XL_dev, out_dev, B,
                                cfg->n_frames, (cudaStream_t)stream, phase_cycles_dev, nullptr, 0, limb_tiles_dev);
}

int kasf_fusion(const kasf_config* cfg, const void* packed_dev, int layer, const float* att_dev,
                const float* graph_dev, const float* bone_dev, float* out_dev, int B, void* stream) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !att_dev || !graph_dev || !bone_dev || !out_dev || B < 0 || layer < 0 || layer >= cfg->n_layers)
        return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_fusion((const uint8_t*)packed_dev, layer, att_dev, graph_dev, bone_dev, out_dev,
                         (long long)B * cfg->n_frames * J, (cudaStream_t)stream);
}

int kasf_head(const kasf_config* cfg, const void* packed_dev, const float* X_dev, float* y_dev, float* rep_dev,
              int B, void* stream) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    if (!packed_dev || !X_dev || (!y_dev && !rep_dev) || B < 0) return KASF_EINVAL;
    if ((rc = device_ok())) return rc;
    return launch_head((const uint8_t*)packed_dev, X_dev, y_dev, rep_dev, (long long)B * cfg->n_frames * J,
                       (cudaStream_t)stream);
}

static int forward_impl(const kasf_config* cfg, const void* packed_dev, const float* x_dev, float* y_dev,
                        float* rep_dev, int B, void* ws_dev, size_t ws_bytes, void* stream, void** events,
                        const kasf_forward_opts* opts) {
    int rc = config_ok(cfg);
    if (rc) return rc;
    const int precision = opts ? opts->precision : KASF_PRECISION_FAST;
    if (precision != KASF_PRECISION_FAST && precision != KASF_PRECISION_EXACT) return KASF_EINVAL;
    if (!x_dev || (!y_dev && !rep_dev) || !ws_dev || B < 0) return KASF_EINVAL;
    if (((uintptr_t)ws_dev & 255) != 0) return KASF_EINVAL;
    if (ws_bytes < kasf_workspace_bytes_ex(cfg, B, precision)) return KASF_ENOMEM;
    if ((rc = device_ok())) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == KASF_PRECISION_EXACT) {
        if (!opts->image_dev || !packed_dev || events) return KASF_EINVAL;
        return exact_forward(cfg, opts->image_dev, (const uint8_t*)packed_dev, x_dev, y_dev, rep_dev, B, ws_dev, ws_bytes, st);
    }
    if (!packed_dev) return KASF_EINVAL;
    if ((rc = fast_config_ok(cfg))) return rc;
    const uint8_t* blob = (const uint8_t*)packed_dev;
    const int T = cfg->n_frames;
    const int chunk = clip_chunk(cfg, B);
    const unsigned flags = opts ? opts->flags : 0u;
    // the two-tiles kernel where it applies (spatial modules; temporal ones with short sequences)
    const unsigned fl_s = flags & KASF_FLAG_TWO_TILES, fl_t = T <= 32 ? fl_s : 0u;
    int ev = 0;
    // side streams for the graph / bone branches (see the layer loop): the caller's context, never under per-stage timing
    kasf_forward_ctx* ctx = (opts && !events) ? opts->ctx : nullptr;
    bool forked = false;
#define KASF_MARK() do { if (events) cudaEventRecord((cudaEvent_t)events[ev++], st); } while (0)
    KASF_MARK();
    for (int b0 = 0; b0 < B && !rc; b0 += chunk) {
        const int nb = B - b0 < chunk ? B - b0 : chunk;
        const long long tokens = (long long)nb * T * J;
        Streams s = carve(ws_dev, (long long)chunk * T * J);
        const size_t stream_bytes = ((size_t)chunk * T * J * D * 4 + 1023) / 1024 * 1024;
        void* scr = static_cast<uint8_t*>(ws_dev) + 6 * stream_bytes;     // temporal modules, T > KASF_SPLIT_T only
        const size_t scr_bytes = module_scratch_bytes(chunk, T);
        // normalised limb rows as bf16 operand tiles (spatial / temporal tile order), shared by all layers
        // (split path: one scratch area per branch, the branches run concurrently)
        void* scr_g = static_cast<uint8_t*>(scr) + scr_bytes;
        void* scr_b = static_cast<uint8_t*>(scr) + 2 * scr_bytes;
        uint8_t* lt_s = static_cast<uint8_t*>(scr) + 3 * scr_bytes;
        uint8_t* lt_t = lt_s + limb_tiles_bytes(chunk, T, KASF_MODE_SPATIAL);
        if (limb_tiles_bytes(chunk, T, KASF_MODE_TEMPORAL) == 0) lt_t = nullptr;
        const float* x = x_dev + (size_t)b0 * T * J * 3;
        if ((rc = launch_features(blob, x, nullptr, nullptr, s.X, s.XB, s.XL, (long long)nb * T, st))) break;
        if ((rc = launch_limb_tiles(s.XL, lt_s, nb, T, KASF_MODE_SPATIAL, st))) break;
        if ((rc = launch_limb_tiles(s.XL, lt_t, nb, T, KASF_MODE_TEMPORAL, st))) break;
        KASF_MARK();
        for (int l = 0; l < cfg->n_layers && !rc; ++l) {
            // three branches, each spatial module then temporal module (KASportsFormer.py:268-275).  The branches are
            // independent until the fusion, so with a context the graph and bone branches run on its two side
            // streams: the last, partial wave of one persistent kernel (26.7 tiles per SM at B = 1024) is filled by the
            // first CTAs of another branch's kernel instead of idling.
            const float* bone_src = l == 0 ? s.XB : s.X;
            cudaStream_t sg = st, sb = st;
            if (ctx && cudaEventRecord(ctx->e[0], st) == cudaSuccess &&               // X of this layer is final
                cudaStreamWaitEvent(ctx->s[0], ctx->e[0], 0) == cudaSuccess &&
                cudaStreamWaitEvent(ctx->s[1], ctx->e[0], 0) == cudaSuccess) {
                sg = ctx->s[0], sb = ctx->s[1];
                forked = true;
            }
            do {
                if ((rc = launch_former_module(blob, l, KASF_KIND_ATTENTION, KASF_MODE_SPATIAL, s.X, nullptr, s.A, nb, T, st, nullptr, nullptr, 0, nullptr, fl_s))) break;
                KASF_MARK();
                if ((rc = launch_former_module(blob, l, KASF_KIND_ATTENTION, KASF_MODE_TEMPORAL, s.A, nullptr, s.A, nb, T, st, nullptr, scr, scr_bytes, nullptr, fl_t))) break;
                KASF_MARK();
                if ((rc = launch_former_module(blob, l, KASF_KIND_GRAPH, KASF_MODE_SPATIAL, s.X, nullptr, s.G, nb, T, sg, nullptr, nullptr, 0, nullptr, fl_s))) break;
                KASF_MARK();
                if ((rc = launch_former_module(blob, l, KASF_KIND_GRAPH, KASF_MODE_TEMPORAL, s.G, nullptr, s.G, nb, T, sg, nullptr, scr_g, scr_bytes, nullptr, fl_t))) break;
                KASF_MARK();
                if ((rc = launch_former_module(blob, l, KASF_KIND_BONE, KASF_MODE_SPATIAL, bone_src, s.XL, s.Bn, nb, T, sb, nullptr, nullptr, 0, lt_s, fl_s))) break;
                KASF_MARK();
                if ((rc = launch_former_module(blob, l, KASF_KIND_BONE, KASF_MODE_TEMPORAL, s.Bn, s.XL, s.Bn, nb, T, sb, nullptr, scr_b, scr_bytes, lt_t, lt_t ? fl_t : 0u))) break;
                KASF_MARK();
            } while (0);
            if (forked) {
                // join -- also after a launch error, so that the caller's stream (or capture) is never left with
                // un-joined side-stream work
                bool ok = cudaEventRecord(ctx->e[1], sg) == cudaSuccess && cudaEventRecord(ctx->e[2], sb) == cudaSuccess &&
                          cudaStreamWaitEvent(st, ctx->e[1], 0) == cudaSuccess && cudaStreamWaitEvent(st, ctx->e[2], 0) == cudaSuccess;
                forked = false;
                if (!ok && !rc) rc = cuda_status() ? cuda_status() : KASF_EINVAL;
            }
            if (rc) break;
            if ((rc = launch_fusion(blob, l, s.A, s.G, s.Bn, s.X, tokens, st))) break;
            KASF_MARK();
        }
        if (rc) break;
        float* y = y_dev ? y_dev + (size_t)b0 * T * J * 3 : nullptr;
        float* rep = rep_dev ? rep_dev + (size_t)b0 * T * J * REP : nullptr;
        if ((rc = launch_head(blob, s.X, y, rep, tokens, st))) break;
        KASF_MARK();
    }
#undef KASF_MARK
    return rc;
}

int kasf_forward(const kasf_config* cfg, const void* packed_dev, const float* x_dev, float* y_dev, float* rep_dev,
                 int B, void* ws_dev, size_t ws_bytes, void* stream) {
    return forward_impl(cfg, packed_dev, x_dev, y_dev, rep_dev, B, ws_dev, ws_bytes, stream, nullptr, nullptr);
}

int kasf_forward_ex(const kasf_config* cfg, const void* packed_dev, const float* x_dev, float* y_dev, float* rep_dev,
                    int B, void* ws_dev, size_t ws_bytes, void* stream, const kasf_forward_opts* opts) {
    return forward_impl(cfg, packed_dev, x_dev, y_dev, rep_dev, B, ws_dev, ws_bytes, stream, nullptr, opts);
}

int kasf_forward_timed(const kasf_config* cfg, const void* packed_dev, const float* x_dev, float* y_dev,
                       float* rep_dev, int B, void* ws_dev, size_t ws_bytes, void* stream, void** events,
                       int n_events) {
    if (!events || n_events < kasf_forward_marks(cfg, B) + 1) return KASF_EINVAL;
    return forward_impl(cfg, packed_dev, x_dev, y_dev, rep_dev, B, ws_dev, ws_bytes, stream, events, nullptr);
}

void* kasf_event_create(void) {
    cudaEvent_t e = nullptr;
    return cudaEventCreate(&e) == cudaSuccess ? (void*)e : nullptr;
}
void kasf_event_destroy(void* e) {
    if (e) cudaEventDestroy((cudaEvent_t)e);
}
float kasf_event_elapsed_ms(void* a, void* b) {
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) return -1.f;
    return ms;
}

int kasf_metrics(int T, const float* pred_dev, const float* pred_flip_dev, const float* gt_dev, const float* res_dev,
                 const float* factor_dev, const int32_t* action_dev, int n_actions, double* sums_dev,
                 double* per_frame_dev, int B, void* stream) {
    if (!pred_dev || !gt_dev || !res_dev || !factor_dev || !sums_dev || B < 0) return KASF_EINVAL;
    int rc = device_ok();
    if (rc) return rc;
    return launch_metrics(T, pred_dev, pred_flip_dev, gt_dev, res_dev, factor_dev, action_dev, n_actions, sums_dev,
                          per_frame_dev, B, (cudaStream_t)stream);
}

int kasf_joint_flip(const float* in_dev, float* out_dev, int64_t n_frames_total, void* stream) {
    if (!in_dev || !out_dev || n_frames_total < 0) return KASF_EINVAL;
    int rc = device_ok();
    if (rc) return rc;
    return launch_flip(in_dev, out_dev, n_frames_total, (cudaStream_t)stream);
}

int kasf_table(int which, int32_t* out, int cap) {
    if (!out) return KASF_EINVAL;
    const int* src = nullptr;
    int n = 0;
    int adj[289];
    switch (which) {
        case 0: src = h_bone_child, n = 16; break;
        case 1: src = h_bone_parent, n = 16; break;
        case 2: src = h_limb_size, n = 17; break;
        case 3: src = h_limb_member, n = 68; break;
        case 4:
            for (int i = 0; i < 289; ++i) adj[i] = 0;
            for (int i = 0; i < 17; ++i)
                for (int k = 0; k < 4; ++k)
                    if (h_nbr[i * 4 + k] >= 0) adj[i * 17 + h_nbr[i * 4 + k]] = 1;
            src = adj, n = 289;
            break;
        case 5: src = h_flip, n = 17; break;
        default: return KASF_EINVAL;
    }
    if (cap < n) return KASF_ENOMEM;
    for (int i = 0; i < n; ++i) out[i] = src[i];
    return n;
}

int kasf_test_gemm(const float* a_dev, const float* w_dev, float* d_dev, int M, int N, void* stream) {
    if (!a_dev || !w_dev || !d_dev) return KASF_EINVAL;
    int rc = device_ok();
    if (rc) return rc;
    return launch_test_gemm(a_dev, w_dev, d_dev, M, N, (cudaStream_t)stream);
}

// self-test hook: the Procrustes routine of the metric kernel executed on the host (same source)
double kasf_selftest_p_mpjpe_host(const double* pred_17x3, const double* gt_17x3) {
    return host_p_mpjpe(pred_17x3, gt_17x3);
}

}  // extern "C"
