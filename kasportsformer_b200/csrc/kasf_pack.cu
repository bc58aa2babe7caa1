// kasf_pack.cu -- fp32 weight image -> kernel-ready blob (see kasf_layout.h).
//   * dense projection weights -> bf16 [128 x 128] operand tiles in the 128B-swizzled K-major
//     shared-memory image, so that a plain bulk copy (UBLKCP) lands them MMA-ready;
//   * LayerNorm / layer-scale / bias vectors copied to a per-module fp32 vector block;
//   * BatchNorm1d(eval) folded to per-node scale/shift (reference model/modules/graph.py:37,129);
//   * features/head weights re-laid for coalesced access.
#include "kasf_internal.h"

namespace kasf {

struct PackModuleArgs {
    ModuleImg img;
    int kind;       // 0 att, 1 graph, 2 bone
    int nodes;      // BN nodes (graph only)
    size_t dst;     // byte offset of the module in the blob
};
struct PackLayerArgs {
    PackModuleArgs m[6];
    size_t fw, fb, fdst;
};

// one block per (module, chunk); 256 threads; chunk = [128 n][128 k] bf16
__global__ void pack_chunks_kernel(const float* __restrict__ img, uint8_t* __restrict__ blob, PackLayerArgs a) {
    const int mod = blockIdx.x / MOD_CHUNKS, c = blockIdx.x % MOD_CHUNKS;
    const PackModuleArgs& pm = a.m[mod];
    const float* src = nullptr;   // element (n, k) at src[n * ld + k]
    int ld = D;
    if (c >= 4 && c < 8) {                    // W1 rows [128c', 128c'+128)
        src = img + pm.img.fc1w + (size_t)(c - 4) * 128 * D;
    } else if (c >= 8) {                      // W2[:, 128c' : 128c'+128]
        src = img + pm.img.fc2w + (size_t)(c - 8) * 128;
        ld = HID;
    } else if (pm.kind == 0) {
        src = c < 3 ? img + pm.img.qkvw + (size_t)c * 128 * D : img + pm.img.projw;
    } else if (pm.kind == 2) {
        src = c == 0 ? img + pm.img.qw : (c < 3 ? img + pm.img.kvw + (size_t)(c - 1) * 128 * D : img + pm.img.projw);
    } else {
        src = c == 0 ? img + pm.img.Uw : (c == 1 ? img + pm.img.Vw : nullptr);
    }
    uint8_t* dst = blob + pm.dst + MOD_VEC_BYTES + (size_t)c * CHUNK_BYTES;
    // bone K|V projections: the limb LayerNorm's gamma is folded into the weight columns, so that the operand is the
    // layer-independent normalised limb row (computed once per forward, see limb_tiles_kernel); beta goes to the
    // output-projection bias (pack_vectors_kernel)
    // likewise LN1's gamma is folded into the Q|K|V (attention) / Q (bone) weight columns: their operand is the
    // normalised stream row, which a producer kernel can hand over ready-made (see xhat tiles in kasf_module.cu)
    // ... and LN2's gamma into the fc1 weight columns of every module (beta: fc1 bias, pack_vectors_kernel)
    const float* kscale = (pm.kind == 2 && (c == 1 || c == 2)) ? img + pm.img.nlw
                          : ((pm.kind == 0 && c < 3) || (pm.kind == 2 && c == 0)) ? img + pm.img.n1w
                          : (c >= 4 && c < 8) ? img + pm.img.n2w : nullptr;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {   // pairs of k
        const int n = i >> 6, k = (i & 63) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (src) {
            v0 = src[(size_t)n * ld + k];
            v1 = src[(size_t)n * ld + k + 1];
        }
        if (kscale) v0 *= kscale[k], v1 *= kscale[k + 1];
        if (c >= 8 && !KASF_HALF_GELU) {
            *reinterpret_cast<uint32_t*>(dst + tile_off_bf16(n, k)) = pack_bf16(0.5f * v0, 0.5f * v1);
        } else if (c >= 8) {
            // the MLP epilogue produces 2*GELU (exact power-of-two rescaling) as fp16: fc2 is an f16 x f16 MMA
            *reinterpret_cast<uint32_t*>(dst + tile_off_bf16(n, k)) = pack_f16(0.5f * v0, 0.5f * v1);
        } else {
            *reinterpret_cast<uint32_t*>(dst + tile_off_bf16(n, k)) = pack_bf16(v0, v1);
        }
    }
}

// one block per module (+1 for fusion): vector block
__global__ void pack_vectors_kernel(const float* __restrict__ img, uint8_t* __restrict__ blob, PackLayerArgs a) {
    if (blockIdx.x == 6) {
        float* f = reinterpret_cast<float*>(blob + a.fdst);
        for (int i = threadIdx.x; i < 3 * 3 * D; i += blockDim.x) f[i] = img[a.fw + i];
        if (threadIdx.x < 3) f[3 * 3 * D + threadIdx.x] = img[a.fb + threadIdx.x];
        return;
    }
    const PackModuleArgs& pm = a.m[blockIdx.x];
    const ModuleImg& m = pm.img;
    float* v = reinterpret_cast<float*>(blob + pm.dst);
    for (int i = threadIdx.x; i < V_FLOATS; i += blockDim.x) v[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        v[V_LS1 + i] = img[m.ls1 + i];
        v[V_LS2 + i] = img[m.ls2 + i];
        v[V_N1W + i] = img[m.n1w + i];
        v[V_N1B + i] = img[m.n1b + i];
        v[V_NLW + i] = img[m.nlw + i];
        v[V_NLB + i] = img[m.nlb + i];
        v[V_N2W + i] = img[m.n2w + i];
        v[V_N2B + i] = img[m.n2b + i];
        v[V_B2 + i] = img[m.fc2b + i];
        if (pm.kind == 1) {
            v[V_BMIX + i] = img[m.Ub + i];
            v[V_BV + i] = img[m.Vb + i];
        } else {
            v[V_BMIX + i] = img[m.projb + i];
        }
    }
    if (pm.kind == 0 || pm.kind == 2) {
        // LN1's affine (g1, b1) folded into the projections of the stream rows: q = (Wq g1) xhat + Wq b1 keeps its
        // offset as an explicit query bias (V_BQ, added when Q is drained); the K offset cancels in the softmax; the V
        // offset (self-attention only) joins the projection bias below.  LN1 itself is left without affine.
        const float* wq = pm.kind == 0 ? img + m.qkvw : img + m.qw;
        for (int o = threadIdx.x; o < D; o += blockDim.x) {
            float acc = 0.f;
            for (int k = 0; k < D; ++k) acc = fmaf(wq[(size_t)o * D + k], img[m.n1b + k], acc);
            v[V_BQ + o] = acc;
            v[V_N1W + o] = 1.f;
            v[V_N1B + o] = 0.f;
        }
    }
    if (pm.kind == 0) {
        __shared__ float wvb0[D];
        __syncthreads();
        for (int o = threadIdx.x; o < D; o += blockDim.x) {
            const float* wv = img + m.qkvw + (size_t)(2 * D + o) * D;
            float acc = 0.f;
            for (int k = 0; k < D; ++k) acc = fmaf(wv[k], img[m.n1b + k], acc);
            wvb0[o] = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < D; i += blockDim.x) {
            const float* wp = img + m.projw + (size_t)i * D;
            float acc = 0.f;
            for (int o = 0; o < D; ++o) acc = fmaf(wp[o], wvb0[o], acc);
            v[V_BMIX + i] = img[m.projb + i] + acc;
        }
    }
    if (pm.kind == 2) {
        // Bone cross-attention with K = (Wk g) xhat + Wk b, V = (Wv g) xhat + Wv b (g, b: limb LayerNorm affine):
        // the K offset adds the same q . (Wk b) to every key of a query and cancels in the softmax; the V offset passes
        // through the attention unchanged (the probabilities sum to 1) and becomes part of the projection bias:
        // b_proj' = b_proj + W_proj (Wv b).  The limb LayerNorm itself is left without affine (gamma 1, beta 0).
        __shared__ float wvb[D];
        __syncthreads();
        for (int o = threadIdx.x; o < D; o += blockDim.x) {
            const float* wv = img + m.kvw + (size_t)(D + o) * D;
            float acc = 0.f;
            for (int k = 0; k < D; ++k) acc = fmaf(wv[k], img[m.nlb + k], acc);
            wvb[o] = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < D; i += blockDim.x) {
            const float* wp = img + m.projw + (size_t)i * D;
            float acc = 0.f;
            for (int o = 0; o < D; ++o) acc = fmaf(wp[o], wvb[o], acc);
            v[V_BMIX + i] = img[m.projb + i] + acc;
            v[V_NLW + i] = 1.f;
            v[V_NLB + i] = 0.f;
        }
    }
    for (int i = threadIdx.x; i < HID; i += blockDim.x) {
        // fc1 bias + W1 beta_2: LN2's affine is folded into fc1 (the kernels write the plain normalised row)
        const float* w1 = img + m.fc1w + (size_t)i * D;
        float acc = img[m.fc1b + i];
        for (int k = 0; k < D; ++k) acc = fmaf(w1[k], img[m.n2b + k], acc);
        v[V_B1 + i] = acc;
        reinterpret_cast<__half*>(v + V_B1H)[i] = __float2half_rn(acc);
    }
    if (pm.kind == 1)
        for (int i = threadIdx.x; i < pm.nodes; i += blockDim.x) {
            // eval BatchNorm: y*s + t,  s = w / sqrt(var + eps),  t = b - mean * s
            const float s = img[m.bnw + i] / sqrtf(img[m.bnv + i] + 1e-5f);
            v[V_BNS + i] = s;
            v[V_BNT + i] = img[m.bnb + i] - img[m.bnm + i] * s;
        }
}

struct PackGlobalArgs {
    GlobalImg g;
};
__global__ void pack_global_kernel(const float* __restrict__ img, uint8_t* __restrict__ blob, PackGlobalArgs a) {
    float* o = reinterpret_cast<float*>(blob);
    const GlobalImg& g = a.g;
    const int limb_size[17] = KASF_LIMB_SIZE;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < 17 * 3 * G_LIMB_STRIDE; i += nth) {
        const int gi = i / (3 * G_LIMB_STRIDE), ch = (i / G_LIMB_STRIDE) % 3, e = i % G_LIMB_STRIDE;
        const int n = limb_size[gi];
        float v = 0.f;
        if (e < 64) {            // W1[h][m] padded to 4 members
            const int h = e >> 2, mm = e & 3;
            if (mm < n) v = img[g.limb[gi][ch][0] + h * n + mm];
        } else if (e < 80) v = img[g.limb[gi][ch][1] + (e - 64)];
        else if (e < 96) v = img[g.limb[gi][ch][2] + (e - 80)];
        else if (e == 96) v = img[g.limb[gi][ch][3]];
        o[G_LIMB + i] = v;
    }
    const size_t ew[3] = {g.jw, g.bw, g.lw}, eb[3] = {g.jb, g.bb, g.lb}, ep[3] = {g.pos, g.bpos, g.lpos};
    for (int i = tid; i < 3 * 3 * D; i += nth) {     // [e][in][c] <- weight[c][in]
        const int e = i / (3 * D), in = (i / D) % 3, c = i % D;
        o[G_EMB_W + i] = img[ew[e] + c * 3 + in];
    }
    for (int i = tid; i < 3 * D; i += nth) o[G_EMB_B + i] = img[eb[i / D] + i % D];
    for (int i = tid; i < 3 * J * D; i += nth) o[G_POS + i] = img[ep[i / (J * D)] + i % (J * D)];
    for (int i = tid; i < D; i += nth) {
        o[G_NORM + i] = img[g.nw + i];
        o[G_NORM + D + i] = img[g.nb + i];
    }
    for (int i = tid; i < D * REP; i += nth) {       // Wrep^T[k][n] <- W[n][k]
        const int k = i / REP, n = i % REP;
        o[G_REPW + i] = img[g.repw + (size_t)n * D + k];
    }
    for (int i = tid; i < REP; i += nth) o[G_REPB + i] = img[g.repb + i];
    // rep_logit's weight as bf16 triples (hi + mid + lo == the fp32 value): operand images of the tensor-core head
    for (int i = tid; i < REP * D; i += nth) {
        const int n = i / D, k = i % D;
        const float w = img[g.repw + i];
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(h);
        const __nv_bfloat16 m = __float2bfloat16_rn(r1);
        const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
        uint8_t* piece = blob + G_REP3_OFF + (size_t)(n >> 6) * REP3_PIECE + img64_off(n & 63, k);
        *reinterpret_cast<__nv_bfloat16*>(piece) = h;
        *reinterpret_cast<__nv_bfloat16*>(piece + REP3_IMG) = m;
        *reinterpret_cast<__nv_bfloat16*>(piece + 2 * REP3_IMG) = l;
    }
    for (int i = tid; i < 3 * REP; i += nth) o[G_HEADW + i] = img[g.hw + i];
    for (int i = tid; i < 4; i += nth) o[G_HEADB + i] = i < 3 ? img[g.hb + i] : 0.f;
}

int pack_weights(const kasf_config* cfg, const float* image, void* packed, size_t cap, cudaStream_t st) {
    if (cap < packed_bytes(cfg)) return KASF_ENOMEM;
    GlobalImg G;
    LayerImg* L = new LayerImg[cfg->n_layers];
    walk_image(cfg, &G, L, [](const char*, size_t, size_t) {});
    uint8_t* blob = static_cast<uint8_t*>(packed);
    PackGlobalArgs ga;
    ga.g = G;
    pack_global_kernel<<<64, 256, 0, st>>>(image, blob, ga);
    for (int l = 0; l < cfg->n_layers; ++l) {
        PackLayerArgs a;
        for (int m = 0; m < 6; ++m) {
            a.m[m].img = L[l].m[m];
            a.m[m].kind = m / 2;
            a.m[m].nodes = (m % 2 == 0) ? J : cfg->n_frames;
            a.m[m].dst = module_off(l, m);
        }
        a.fw = L[l].fw;
        a.fb = L[l].fb;
        a.fdst = fusion_off(l);
        pack_chunks_kernel<<<6 * MOD_CHUNKS, 256, 0, st>>>(image, blob, a);
        pack_vectors_kernel<<<7, 256, 0, st>>>(image, blob, a);
    }
    delete[] L;
    return cuda_status();
}

}  // namespace kasf
