/*
 * kasf.h -- C-ABI of libkasf.so: the B200 (sm_100a) implementation of the KASportsFormer
 * inference forward pass (2D keypoint clips [B,T,17,3] -> 3D poses [B,T,17,3]) and of the
 * MPJPE-family evaluation reduction.
 *
 * The reference (jw0r1n/KASportsFormer) has no native boundary: the path sits behind a Python
 * class.  Each entry point below names the reference interface it replaces (file:line relative to
 * the reference checkout).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is a CUDA device pointer owned by the caller
 *     (the library never allocates or frees device memory and keeps no global mutable state);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and not synchronised;
 *   - the caller selects the device (cudaSetDevice) before calling;
 *   - return value: 0 on success, a negative KASF_E* code otherwise (kasf_strerror() names it);
 *     CUDA launch errors are returned as -(1000 + cudaError_t);
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every compute
 *     entry point returns KASF_EARCH.
 */
#ifndef KASF_H_
#define KASF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KASF_VERSION 2

#define KASF_OK        0
#define KASF_EINVAL   -1   /* null pointer / bad enum / misaligned buffer                     */
#define KASF_ESHAPE   -2   /* a shape this build does not implement (see kasf_config)         */
#define KASF_EARCH    -3   /* current device is not sm_100                                     */
#define KASF_ENOMEM   -4   /* caller-provided buffer too small                                 */

/* Model hyper-parameters: the YAML model keys of the reference (configs/*.yaml:66-92) that reach
 * KASportsFormer.__init__ (model/KASportsFormer.py:291-295).  This build implements
 * dim_feat=128, num_heads=8, mlp_ratio=4, dim_rep=512, num_joints=17, dim_in=dim_out=3,
 * neighbour_num=4, n_frames in [4,243]; anything else -> KASF_ESHAPE. */
typedef struct kasf_config {
    int32_t n_layers;
    int32_t n_frames;
    int32_t dim_feat;
    int32_t dim_rep;
    int32_t num_heads;
    int32_t mlp_ratio;
    int32_t num_joints;
    int32_t neighbour_num;
} kasf_config;

/* FormerModule selectors (model/KASportsFormer.py:65-118) */
#define KASF_KIND_ATTENTION 0   /* model/modules/selfattention.py      */
#define KASF_KIND_GRAPH     1   /* model/modules/graph.py              */
#define KASF_KIND_BONE      2   /* model/modules/bone_crossattention.py */
#define KASF_MODE_SPATIAL   0
#define KASF_MODE_TEMPORAL  1

int         kasf_version(void);
const char* kasf_strerror(int code);
/* 0 if the current device can run this library, KASF_EARCH otherwise. */
int         kasf_device_supported(void);

/* ---- weights -------------------------------------------------------------------------------
 * The "weight image" is one contiguous float32 array holding every floating-point tensor of the
 * reference state_dict (names exactly as `KASportsFormer.state_dict()` produces them; the int64
 * `num_batches_tracked` buffers are not part of it).  The library defines the order:
 * kasf_weight_entry() enumerates (name, offset, numel).  Replaces: the nn.Module parameter tree
 * built by model/KASportsFormer.py:291-318. */
int    kasf_weight_entries(const kasf_config* cfg);
int    kasf_weight_entry(const kasf_config* cfg, int index, char* name, size_t name_cap,
                         size_t* offset_floats, size_t* numel);
size_t kasf_weight_image_floats(const kasf_config* cfg);

/* Pack the fp32 image (dev) into the kernel-ready blob (dev): bf16 tensor-core operand tiles in
 * the 128B-swizzled K-major shared-memory layout, fp32 vectors, BatchNorm folded to scale/shift. */
size_t kasf_packed_bytes(const kasf_config* cfg);
int    kasf_pack_weights(const kasf_config* cfg, const float* image_dev, void* packed_dev,
                         size_t packed_cap, void* stream);

/* ---- forward -------------------------------------------------------------------------------
 * Replaces KASportsFormer.forward (model/KASportsFormer.py:320-347).
 *   x_dev   float32 [B, T, 17, 3] contiguous (x, y, confidence)
 *   y_dev   float32 [B, T, 17, 3]            (3D pose, normalised units)
 *   rep_dev float32 [B, T, 17, 512] or NULL  (`return_rep=True` output, :342-343)
 *   ws_dev  scratch of at least kasf_workspace_bytes(cfg, B) bytes, 256-byte aligned         */
size_t kasf_workspace_bytes(const kasf_config* cfg, int B);
/* kasf_forward enqueues all its work on `stream`, one kernel after the other, fast precision: the stateless form. */
int    kasf_forward(const kasf_config* cfg, const void* packed_dev, const float* x_dev,
                    float* y_dev, float* rep_dev, int B, void* ws_dev, size_t ws_bytes,
                    void* stream);

/* Forward with options.
 *   precision  KASF_PRECISION_FAST: bf16 tensor-core operands with fp32 accumulation, fp16 MLP hidden tile, tanh-form
 *              GELU (max deviation 5.6e-5 from erf); fp32 residual stream, statistics, softmax, similarity, head.
 *              KASF_PRECISION_EXACT: the reference's arithmetic -- fp32 FMA everywhere, erf GELU -- on CUDA cores
 *              (no tensor cores): for checking a checkpoint's accuracy, an order of magnitude slower.  Needs the
 *              fp32 weight image (`image_dev`, the array kasf_pack_weights consumed) instead of the packed blob and
 *              kasf_workspace_bytes_ex(cfg, B, KASF_PRECISION_EXACT) bytes of workspace.
 *   flags      KASF_FLAG_TWO_TILES: the two-tiles-in-flight FormerModule kernel where it applies (spatial modules,
 *              temporal ones with n_frames <= 32) -- parity-tested, currently slower than the default kernel.
 *   ctx        NULL, or a kasf_forward_ctx: two side streams and three events, created once by the caller for a
 *              (device, host thread) pair.  With a context the graph and bone branches of every layer run on the side
 *              streams (forked after the previous fusion, joined before the next one): the last, partial wave of one
 *              persistent kernel is filled by the first CTAs of another branch's kernel.  Also valid under stream
 *              capture (the fork / join becomes part of the captured graph; keep the context alive with the graph).
 *              The library itself keeps no state and reads no environment variables. */
#define KASF_PRECISION_FAST  0
#define KASF_PRECISION_EXACT 1
#define KASF_FLAG_TWO_TILES  1u
typedef struct kasf_forward_ctx kasf_forward_ctx;
kasf_forward_ctx* kasf_ctx_create(void);                 /* on the current device; NULL on failure */
void              kasf_ctx_destroy(kasf_forward_ctx* ctx);
typedef struct kasf_forward_opts {
    int32_t precision;
    uint32_t flags;
    kasf_forward_ctx* ctx;
    const float* image_dev;      /* KASF_PRECISION_EXACT only */
} kasf_forward_opts;
size_t kasf_workspace_bytes_ex(const kasf_config* cfg, int B, int precision);
int    kasf_forward_ex(const kasf_config* cfg, const void* packed_dev, const float* x_dev,
                       float* y_dev, float* rep_dev, int B, void* ws_dev, size_t ws_bytes,
                       void* stream, const kasf_forward_opts* opts /* NULL: defaults */);
/* Number of kernel launches one kasf_forward(B) enqueues (for bench accounting). */
int    kasf_forward_launches(const kasf_config* cfg, int B);
/* Number of timing marks of kasf_forward_timed: one per stage (features, every FormerModule, every fusion,
 * head).  Equal to the launch count for n_frames <= 64; for longer sequences a temporal module is 2-3
 * kernels behind one mark. */
int    kasf_forward_marks(const kasf_config* cfg, int B);
/* Same forward, additionally recording events[0] before the first stage and events[i] after the
 * i-th stage on `stream` (n_events >= marks + 1), so a caller can read per-stage device times
 * of the very launches it is timing.  Stage order per pass: features, then per layer att_s, att_t,
 * graph_s, graph_t, bone_s, bone_t, fusion, and finally head.  Events come from kasf_event_create. */
int    kasf_forward_timed(const kasf_config* cfg, const void* packed_dev, const float* x_dev,
                          float* y_dev, float* rep_dev, int B, void* ws_dev, size_t ws_bytes,
                          void* stream, void** events, int n_events);
void*  kasf_event_create(void);
void   kasf_event_destroy(void* event);
float  kasf_event_elapsed_ms(void* start, void* stop);   /* after the stream has been synchronised */

/* ---- per-stage entry points (used by the parity tests; same kernels kasf_forward runs) -------
 * Kinematic anatomy features + the three embeddings.  Replaces bone_decomposer
 * (model/KASportsFormer.py:42-62), BoneRefusion.forward (model/modules/bone_refusion.py:61-70,
 * bone_MLP.py:16-27) and the embeddings (model/KASportsFormer.py:325-330).
 *   bone_dev / limb_dev  float32 [B,T,17,3] or NULL (raw features, for tests)
 *   X/XB/XL              float32 [B,T,17,128]                                                  */
int kasf_kinematic_features(const kasf_config* cfg, const void* packed_dev, const float* x_dev,
                            float* bone_dev, float* limb_dev, float* X_dev, float* XB_dev,
                            float* XL_dev, int B, void* stream);

/* One FormerModule (model/KASportsFormer.py:103-118): out = v + ls1*mixer(LN1(v)[,LN1_limb(XL)]);
 * out = out + ls2*MLP(LN2(out)).  `layer` in [0,n_layers), kind/mode as above.
 * in/out float32 [B,T,17,128] (may alias); XL_dev needed for KASF_KIND_BONE only. */
int kasf_former_module(const kasf_config* cfg, const void* packed_dev, int layer, int kind,
                       int mode, const float* in_dev, const float* XL_dev, float* out_dev, int B,
                       void* stream);
/* Same with caller-provided scratch.  Temporal modules of sequences that do not pack into a 128-row tile
 * (n_frames > 64: the T=81 and T=243 configs) run as a projection kernel, a per-sequence mixer-core kernel
 * and the fused tail, exchanging bf16 Q/K/V (or A_hat z) through `scratch_dev`
 * (>= kasf_module_scratch_bytes(cfg, B) bytes, 256-byte aligned; 0 bytes needed for n_frames <= 64,
 * then identical to kasf_former_module).  kasf_former_module returns KASF_ENOMEM in that case. */
size_t kasf_module_scratch_bytes(const kasf_config* cfg, int B);
int kasf_former_module_ws(const kasf_config* cfg, const void* packed_dev, int layer, int kind,
                          int mode, const float* in_dev, const float* XL_dev, float* out_dev, int B,
                          void* scratch_dev, size_t scratch_bytes, void* stream);

/* Bone modules, fast path: the limb stream is the same for every layer and the limb LayerNorm's affine is folded
 * into the packed K|V weights, so the K|V operand of all bone modules is the normalised limb row.  kasf_limb_tiles
 * normalises XL once (fp32 two-pass statistics, model/modules/bone_crossattention.py:47-51) and writes it as bf16
 * [128 x 128] operand tiles in the tile order of `mode`; a bone module of the same mode given `limb_tiles_dev`
 * (128-byte aligned, kasf_limb_tiles_bytes bytes) fetches its K|V operand with one bulk copy per tile and never
 * reads XL_dev (temporal tiles of n_frames > 64 are in the split path's (sequence, frame) row order).  Without
 * limb tiles (NULL) the module normalises the fp32 limb rows itself.  kasf_forward uses this path internally. */
size_t kasf_limb_tiles_bytes(const kasf_config* cfg, int B, int mode);
int kasf_limb_tiles(const kasf_config* cfg, const float* XL_dev, void* limb_tiles_dev, int B, int mode,
                    void* stream);
int kasf_former_module_lt(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                          const float* in_dev, const float* XL_dev, const void* limb_tiles_dev,
                          float* out_dev, int B, void* scratch_dev, size_t scratch_bytes, void* stream);
/* The same with `flags` (KASF_FLAG_TWO_TILES: returns KASF_ESHAPE where that kernel does not apply -- temporal
 * modules with n_frames > 32, bone modules without limb tiles). */
int kasf_former_module_ex(const kasf_config* cfg, const void* packed_dev, int layer, int kind, int mode,
                          const float* in_dev, const float* XL_dev, const void* limb_tiles_dev,
                          float* out_dev, int B, void* scratch_dev, size_t scratch_bytes, uint32_t flags,
                          void* stream);

/* Profiling hook (own kernel instantiations, n_frames <= 32 and fused path only; otherwise the counters stay zero):
 * same launch, additionally accumulating (atomicAdd by thread 0 of every CTA) the SM cycles
 * spent in each phase of the kernel into phase_cycles_dev[24] (caller zeroes it): 0 limb K/V, 1 load+LN1,
 * 2 QKV MMA wait, 3 Q/K/V drain, 4 attention core, 5 projection MMA wait, 6 similarity/top-k, 7 aggregation,
 * 8 V MMA wait, 9 mixer epilogue, 10 LN2, 11 MLP epilogue compute, 12 output epilogue, 13 wait for the gathered
 * rows, 14 MLP waits for MMAs inside the chunk loop, 15 MLP tensor-memory loads, 16 wait for the last fc2 chunk,
 * 17 wait for the first fc1 chunk (18..23 spare). */
int kasf_former_module_profiled(const kasf_config* cfg, const void* packed_dev, int layer, int kind,
                                int mode, const float* in_dev, const float* XL_dev, float* out_dev,
                                int B, void* stream, unsigned long long* phase_cycles_dev);
/* The same with the bone modules' limb operand from kasf_limb_tiles (limb_tiles_dev, may be null).  When the
 * two-tiles-in-flight kernel serves the call (spatial modules, temporal ones with n_frames <= 32; bone modules only
 * with limb tiles) the slots are: mixer group 0 wait for the gathered rows, 1 LN1, 2 Q|K|V waits and drains (graph:
 * z hand-over), 3 attention core / adjacency + aggregation, 4 waits before the epilogue (x re-read, previous output
 * epilogue, projection), 5 mixer epilogue; MLP group 8 wait for x1, 9 LN2, 10 fc1 waits, 11 GELU epilogues, 12 wait
 * for the fc2 that last read the GELU buffer, 13 wait for the last fc2, 14 output epilogue. */
int kasf_former_module_profiled_lt(const kasf_config* cfg, const void* packed_dev, int layer, int kind,
                                   int mode, const float* in_dev, const float* XL_dev,
                                   const void* limb_tiles_dev, float* out_dev, int B, void* stream,
                                   unsigned long long* phase_cycles_dev);

/* Adaptive fusion of the three branches (model/KASportsFormer.py:279-282). */
int kasf_fusion(const kasf_config* cfg, const void* packed_dev, int layer, const float* att_dev,
                const float* graph_dev, const float* bone_dev, float* out_dev, int B, void* stream);

/* Final LayerNorm -> Linear(128,512) -> tanh -> Linear(512,3) (model/KASportsFormer.py:339-345). */
int kasf_head(const kasf_config* cfg, const void* packed_dev, const float* X_dev, float* y_dev,
              float* rep_dev, int B, void* stream);

/* ---- evaluation epilogue ---------------------------------------------------------------------
 * Replaces the per-clip numpy loop of train_and_evaluate_sp.py:55-103 and utils/error_calc.py:5-48.
 *   pred_dev     float32 [B,T,17,3] model output (normalised); joint 0 is treated as zero (:55)
 *   pred_flip_dev float32 [B,T,17,3] or NULL: output for the left/right flipped input; when given
 *                the two are un-flipped and averaged first (:46-51, utils/utilities.py:128-135)
 *   gt_dev       float32 [B,T,17,3] ground truth in mm
 *   res_dev      float32 [B,2] (res_w,res_h);  factor_dev float32 [B,T];  action_dev int32 [B]
 *   sums_dev     float64 [n_actions, KASF_METRIC_COLS]: ACCUMULATED (atomicAdd) per-action sums:
 *                col 0 sum MPJPE(frames), 1 sum P-MPJPE, 2 sum accel-error, 3 #frames, 4 #accel
 *                frames, 5..21 sum per-joint error.  Caller zeroes it and reduces across ranks.
 *                A clip whose action index is outside [0, n_actions) is a caller error: it is SKIPPED (never folded
 *                into another action), so column 3 then sums to fewer than B*T frames.
 *   per_frame_dev float64 [B,T,3] or NULL: per-frame (mpjpe, p_mpjpe, accel) for tests.        */
#define KASF_METRIC_COLS 22
int kasf_metrics(int T, const float* pred_dev, const float* pred_flip_dev, const float* gt_dev,
                 const float* res_dev, const float* factor_dev, const int32_t* action_dev,
                 int n_actions, double* sums_dev, double* per_frame_dev, int B, void* stream);

/* Left/right flip of a clip batch (utils/utilities.py:128-135): x -> -x, swap joint pairs. */
int kasf_joint_flip(const float* in_dev, float* out_dev, int64_t n_frames_total, void* stream);

/* ---- constant tables baked into the kernels (for cross-checking host copies) ---------------
 * which: 0 bone child[16], 1 bone parent[16], 2 limb group sizes[17], 3 limb members[17*4]
 * (-1 padded), 4 skeleton adjacency[17*17], 5 flip permutation[17].  Returns count written. */
int kasf_table(int which, int32_t* out, int cap);

/* ---- test hook: plain tcgen05 GEMM  D[M,N] = A[M,128] * W[N,128]^T (bf16 in, fp32 out) ------- */
int kasf_test_gemm(const float* a_dev, const float* w_dev, float* d_dev, int M, int N,
                   void* stream);

/* ---- self-test hook: the metric kernel's Procrustes routine (same source) executed on the host. */
double kasf_selftest_p_mpjpe_host(const double* pred_17x3, const double* gt_17x3);

#ifdef __cplusplus
}
#endif
#endif /* KASF_H_ */
