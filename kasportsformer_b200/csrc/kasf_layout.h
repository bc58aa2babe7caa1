// kasf_layout.h -- (1) the canonical fp32 weight image (reference state_dict names, in the order
// the reference's state_dict() yields them) and (2) the packed, kernel-ready blob.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/kasf.h"

namespace kasf {

constexpr int D = 128;        // dim_feat
constexpr int HID = 512;      // mlp hidden
constexpr int REP = 512;      // dim_rep
constexpr int J = 17;         // joints
constexpr int HEADS = 8;
constexpr int DH = 16;
constexpr int LIMB_HID = 16;
constexpr int MAX_T = 243;

inline int config_ok(const kasf_config* c) {
    if (!c) return KASF_EINVAL;
    // num_heads: 8 (every shipped YAML; the tensor-core kernels are specialised to 8 x 16) or 4 (the reference
    // constructor's default, head_dim 32: KASF_PRECISION_EXACT only)
    if (c->dim_feat != D || c->dim_rep != REP || (c->num_heads != HEADS && c->num_heads != 4) || c->mlp_ratio != 4 ||
        c->num_joints != J || c->neighbour_num != 4)
        return KASF_ESHAPE;
    if (c->n_layers < 1 || c->n_layers > 1024) return KASF_ESHAPE;
    if (c->n_frames < 4 || c->n_frames > MAX_T) return KASF_ESHAPE;
    return KASF_OK;
}

// the tensor-core ("fast") kernels
inline int fast_config_ok(const kasf_config* c) {
    const int rc = config_ok(c);
    return rc ? rc : (c->num_heads == HEADS ? KASF_OK : KASF_ESHAPE);
}

static const int kLimbSize[17] = {3, 3, 2, 2, 3, 3, 4, 4, 4, 4, 3, 4, 4, 4, 4, 2, 2};

// ------------------------------------------------------------------ fp32 image offsets (floats)
struct ModuleImg {          // one FormerModule
    size_t ls1, ls2, n1w, n1b, nlw, nlb, n2w, n2b, fc1w, fc1b, fc2w, fc2b;
    size_t projw, projb, qkvw, qw, kvw;            // attention / bone
    size_t Uw, Ub, Vw, Vb, bnw, bnb, bnm, bnv;     // graph
};
struct LayerImg {
    ModuleImg m[6];         // att_s, att_t, graph_s, graph_t, bone_s, bone_t
    size_t fw, fb;
};
struct GlobalImg {
    size_t pos, bpos, lpos, jw, jb, bw, bb, lw, lb, nw, nb;
    size_t limb[17][3][4];  // fc1w, fc1b, fc2w, fc2b
    size_t repw, repb, hw, hb;
};

static const char* const kBranch[3] = {"att", "graph", "bone"};
static const char* const kMode[2] = {"spatial", "temporal"};
static const char* const kLimbCh[3] = {"mlp_dir_x", "mlp_dir_y", "mlp_len"};

// Walks the image in canonical order.  `emit(name, offset, numel)` is called per tensor when given.
template <typename Emit>
inline size_t walk_image(const kasf_config* c, GlobalImg* G, LayerImg* L /* [n_layers] or null */, Emit emit) {
    size_t off = 0;
    char nm[160];
    auto put = [&](const char* name, size_t n) {
        size_t o = off;
        emit(name, o, n);
        off += n;
        return o;
    };
    GlobalImg g;
    g.pos = put("pos_embed", J * D);
    g.bpos = put("bone_pos_embed", J * D);
    g.lpos = put("limb_pos_embed", J * D);
    g.jw = put("joints_embed.weight", D * 3);
    g.jb = put("joints_embed.bias", D);
    g.bw = put("bone_embed.weight", D * 3);
    g.bb = put("bone_embed.bias", D);
    g.lw = put("limb_embed.weight", D * 3);
    g.lb = put("limb_embed.bias", D);
    g.nw = put("norm.weight", D);
    g.nb = put("norm.bias", D);
    for (int i = 0; i < 17; ++i)
        for (int ch = 0; ch < 3; ++ch) {
            snprintf(nm, sizeof nm, "bone_refusion.mlp_layers.%d.%s.fc1.weight", i, kLimbCh[ch]);
            g.limb[i][ch][0] = put(nm, LIMB_HID * kLimbSize[i]);
            snprintf(nm, sizeof nm, "bone_refusion.mlp_layers.%d.%s.fc1.bias", i, kLimbCh[ch]);
            g.limb[i][ch][1] = put(nm, LIMB_HID);
            snprintf(nm, sizeof nm, "bone_refusion.mlp_layers.%d.%s.fc2.weight", i, kLimbCh[ch]);
            g.limb[i][ch][2] = put(nm, LIMB_HID);
            snprintf(nm, sizeof nm, "bone_refusion.mlp_layers.%d.%s.fc2.bias", i, kLimbCh[ch]);
            g.limb[i][ch][3] = put(nm, 1);
        }
    for (int l = 0; l < c->n_layers; ++l) {
        LayerImg li;
        memset(&li, 0, sizeof li);
        for (int br = 0; br < 3; ++br)
            for (int md = 0; md < 2; ++md) {
                ModuleImg& m = li.m[br * 2 + md];
                char p[96];
                snprintf(p, sizeof p, "layers_with_bone.%d.%s_%s.", l, kBranch[br], kMode[md]);
                auto P = [&](const char* s, size_t n) {
                    snprintf(nm, sizeof nm, "%s%s", p, s);
                    return put(nm, n);
                };
                m.ls1 = P("layer_scale_1", D);
                m.ls2 = P("layer_scale_2", D);
                m.n1w = P("norm1.weight", D);
                m.n1b = P("norm1.bias", D);
                m.nlw = P("norm1_limb.weight", D);
                m.nlb = P("norm1_limb.bias", D);
                if (br == 0) {
                    m.projw = P("mixer.proj.weight", D * D);
                    m.projb = P("mixer.proj.bias", D);
                    m.qkvw = P("mixer.qkv.weight", 3 * D * D);
                } else if (br == 1) {
                    const size_t nodes = md == 0 ? J : (size_t)c->n_frames;
                    m.Uw = P("mixer.U.weight", D * D);
                    m.Ub = P("mixer.U.bias", D);
                    m.Vw = P("mixer.V.weight", D * D);
                    m.Vb = P("mixer.V.bias", D);
                    m.bnw = P("mixer.batch_norm.weight", nodes);
                    m.bnb = P("mixer.batch_norm.bias", nodes);
                    m.bnm = P("mixer.batch_norm.running_mean", nodes);
                    m.bnv = P("mixer.batch_norm.running_var", nodes);
                } else {
                    m.projw = P("mixer.proj.weight", D * D);
                    m.projb = P("mixer.proj.bias", D);
                    m.qw = P("mixer.qkv_q.weight", D * D);
                    m.kvw = P("mixer.qkv_kv.weight", 2 * D * D);
                }
                m.n2w = P("norm2.weight", D);
                m.n2b = P("norm2.bias", D);
                m.fc1w = P("mlp.fc1.weight", HID * D);
                m.fc1b = P("mlp.fc1.bias", HID);
                m.fc2w = P("mlp.fc2.weight", D * HID);
                m.fc2b = P("mlp.fc2.bias", D);
            }
        snprintf(nm, sizeof nm, "layers_with_bone.%d.fusion_three_channel.weight", l);
        li.fw = put(nm, 3 * 3 * D);
        snprintf(nm, sizeof nm, "layers_with_bone.%d.fusion_three_channel.bias", l);
        li.fb = put(nm, 3);
        if (L) L[l] = li;
    }
    g.repw = put("rep_logit.fc.weight", REP * D);
    g.repb = put("rep_logit.fc.bias", REP);
    g.hw = put("head.weight", 3 * REP);
    g.hb = put("head.bias", 3);
    if (G) *G = g;
    return off;
}

// ------------------------------------------------------------------ packed blob (bytes)
// [ global block ][ layer 0 ][ layer 1 ] ...
//   layer   = 6 x module + fusion block
//   module  = vector block (fp32) + 12 weight chunks ([128 x 128] 16-bit operand tiles, 32 KB each: bf16, except
//             the fc2 chunks 8..11, which are fp16 like the hidden activation they multiply)
// 1 (default): the MLP hidden activation and the fc2 weights are fp16 and the GELU epilogue runs in packed half
// precision; 0: bf16 hidden tile, fp32 epilogue (kept for A/B measurements)
#ifndef KASF_HALF_GELU
#define KASF_HALF_GELU 1
#endif
constexpr size_t CHUNK_BYTES = 32768;
constexpr int MOD_CHUNKS = 12;
// vector block, offsets in floats
constexpr int V_LS1 = 0, V_LS2 = 128, V_N1W = 256, V_N1B = 384, V_NLW = 512, V_NLB = 640, V_N2W = 768,
              V_N2B = 896, V_BMIX = 1024 /* proj bias | U bias */, V_BV = 1152, V_B1 = 1280, V_B2 = 1792,
              V_BNS = 1920 /* [256] */, V_BNT = 2176 /* [256] */,
              V_B1H = 2560 /* fc1 bias as 512 fp16 values (256 floats) for the packed-half GELU epilogue */,
              V_BQ = 2816 /* [128] query bias W_q beta_1 (attention / bone: LN1's affine is folded into the weights) */,
              V_FLOATS = 2944;
constexpr size_t MOD_VEC_BYTES = V_FLOATS * 4;                                  // 11776
constexpr size_t MOD_BYTES = MOD_VEC_BYTES + MOD_CHUNKS * CHUNK_BYTES;          // 404992
constexpr size_t FUSION_BYTES = 5120;                                            // W[3][384], b[3] fp32
constexpr size_t LAYER_BYTES = 6 * MOD_BYTES + FUSION_BYTES;
// chunk order inside a module (natural order; the kernel's producer walks its own sequence):
//   attention: 0..2 Wq,Wk,Wv   3 Wproj        4..7 W1[0..3]  8..11 W2[0..3]
//   bone:      0 Wq  1..2 Wk,Wv 3 Wproj       4..7 W1        8..11 W2
//   graph:     0 Wu  1 Wv       (2,3 unused)  4..7 W1        8..11 W2

// global block, offsets in floats
constexpr int G_LIMB = 0;                       // [17][3][100]: W1[16][4] b1[16] W2[16] b2 (pad to 100)
constexpr int G_LIMB_STRIDE = 100;
constexpr int G_EMB_W = G_LIMB + 17 * 3 * G_LIMB_STRIDE;   // 5100: [3 embeds][3 in][128]  (in-major)
constexpr int G_EMB_B = G_EMB_W + 3 * 3 * D;               // [3][128]
constexpr int G_POS = G_EMB_B + 3 * D;                     // [3][17][128]
constexpr int G_NORM = G_POS + 3 * J * D;                  // gamma[128] beta[128]
constexpr int G_REPW = G_NORM + 2 * D;                     // Wrep^T [128 k][512 n]
constexpr int G_REPB = G_REPW + D * REP;                   // [512]
constexpr int G_HEADW = G_REPB + REP;                      // [3][512]
constexpr int G_HEADB = G_HEADW + 3 * REP;                 // [3] (+pad)
constexpr int G_FLOATS = G_HEADB + 4;
constexpr size_t GLOBAL_F_BYTES = ((size_t)G_FLOATS * 4 + 1023) / 1024 * 1024;
// ... followed by rep_logit's weight as bf16 TRIPLES for the tensor-core head (kasf_head.cu): W = hi + mid + lo exactly
// (3 x 8 mantissa bits), 8 pieces of 64 output columns, each piece = three [64 n x 128 k] operand images (hi | mid | lo,
// 16 KB each: two K sub-tiles of [64 rows x 128 B], 128-byte swizzle) = 48 KB that one bulk copy lands MMA-ready
constexpr size_t G_REP3_OFF = GLOBAL_F_BYTES;
constexpr size_t REP3_IMG = 16384, REP3_PIECE = 3 * REP3_IMG;
constexpr size_t GLOBAL_BYTES = GLOBAL_F_BYTES + 8 * REP3_PIECE;
// byte offset of element (n, k) of a [64 x 128] image
__host__ __device__ constexpr uint32_t img64_off(uint32_t n, uint32_t k) {
    return (k >> 6) * 8192u + n * 128u + ((((k & 63u) >> 3) ^ (n & 7u)) << 4) + ((k & 7u) << 1);
}

inline size_t packed_bytes(const kasf_config* c) { return GLOBAL_BYTES + (size_t)c->n_layers * LAYER_BYTES; }
__host__ __device__ inline size_t module_off(int layer, int mod) {
    return GLOBAL_BYTES + (size_t)layer * LAYER_BYTES + (size_t)mod * MOD_BYTES;
}
__host__ __device__ inline size_t fusion_off(int layer) {
    return GLOBAL_BYTES + (size_t)layer * LAYER_BYTES + 6 * MOD_BYTES;
}

}  // namespace kasf
