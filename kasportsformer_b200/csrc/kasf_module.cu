// kasf_module.cu -- K2/K3/K4/K5: one fused kernel per FormerModule
//
//     v   = v + ls1 * MIXER(LN1(v) [, LN1_limb(XL)])          reference model/KASportsFormer.py:103-110
//     out = v + ls2 * fc2(GELU(fc1(LN2(v))))                   reference model/KASportsFormer.py:111, mlp.py:24-30
//
// MIXER = Attention (selfattention.py:44-60), BoneCrossAttention (bone_crossattention.py:43-62) or
// GCN (graph.py:99-134), each in its spatial (17 joints of a frame) or temporal (T frames of a joint)
// grouping.  A persistent CTA owns a tile of <=128 group-aligned tokens for the whole module: the fp32
// residual rows are read from HBM once and written once; everything in between stays on chip.
//
//   dense projections ....... tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands in 128B-swizzled smem;
//                             weights arrive pre-swizzled by bulk async copies (TMA engine) through a
//                             2-slot mbarrier ring fed by a dedicated producer warp
//   LayerNorm ................ fp32, warp per row, two-pass statistics
//   attention core ........... thread per query row, K/V (bf16) broadcast from smem, online softmax (fp32)
//   temporal-GCN adjacency ... fp32 similarity on CUDA cores, exact 4th-largest selection with ">=" ties,
//                             row-sum degrees, D^-1/2 A D^-1/2 applied as a sparse gather
//   epilogues ................ thread per row straight out of TMEM (tcgen05.ld 32x32b)
//
// Shared memory map (bytes):   STASH  fp32 residual tile [128][128]             65536
//                              AUX    K|V bf16 / z fp32 / MLP hidden tiles      65536
//                              ATILE  bf16 A operand [128 x 128]                 32768
//                              RING   2 x weight chunk [128 x 128] bf16          65536
//                              barriers + adjacency scratch
#include "kasf_internal.h"

namespace kasf {

__constant__ int c_nbr[68] = KASF_NBR;
__constant__ int c_deg[17] = KASF_DEG;

constexpr int CW = 8;                         // compute warps
constexpr int MOD_THREADS = (CW + 1) * 32;    // + producer warp
constexpr uint32_t SM_STASH = 0;
constexpr uint32_t SM_AUX = 65536;
constexpr uint32_t SM_ATILE = 131072;
constexpr uint32_t SM_RING = 163840;
constexpr uint32_t SM_BARS = 229376;
constexpr uint32_t SM_ADJ = SM_BARS + 256;        // u32 [128][4]
constexpr uint32_t SM_ROWSUM = SM_ADJ + 2048;     // f32 [128]
constexpr uint32_t SM_DEG = SM_ROWSUM + 512;      // u8  [128]
constexpr uint32_t SM_TOTAL = SM_DEG + 128;       // 232320 <= 232448
static_assert(SM_TOTAL <= 232448, "shared memory budget");

// TMEM columns
constexpr uint32_t TM_Q = 0, TM_K = 128, TM_V = 256, TM_MIX = 384;   // mixer phase
constexpr uint32_t TM_H0 = 0, TM_H1 = 128, TM_OUT = 256;            // MLP phase

enum { B_FULL0 = 0, B_FULL1, B_EMPTY0, B_EMPTY1, B_MMA, B_HFULL0, B_HFULL1, B_HSFREE0, B_HSFREE1, B_OUT, B_COUNT };

struct ModParams {
    const uint8_t* mod;      // packed module (vector block + chunks)
    const float* in;         // residual stream in   [B,T,17,128]
    const float* xl;         // limb stream (bone modules)
    float* out;              // residual stream out  (may alias in)
    int B, T;
    int ntiles;
    int groups_per_tile;     // temporal: sequences per tile
};

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// erf-GELU through tanh: 0.5 v (1 + erf(v / sqrt 2)) = 0.5 v (1 + tanh(v (a + b v^2 + c v^4))) with a minimax
// fit of (a, b, c) (max abs deviation from the erf form 5.6e-5) and the hardware tanh (MUFU, rel. error
// 2^-11).  The result is rounded to bf16 (2^-9) right after, so this is below the operand rounding; the
// exact erff form costs ~5x the instructions and made the MLP epilogue the bottleneck of the kernel.
__device__ __forceinline__ float gelu_erf(float v) {
#ifdef KASF_EXACT_GELU
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
#else
    const float v2 = v * v;
    const float pl = fmaf(v2, fmaf(v2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v * pl));
    const float hv = 0.5f * v;
    return fmaf(hv, t, hv);
#endif
}

// fp32 [128][128] tile with XOR-swizzled 16-byte chunks: conflict-free both for "warp per row" and
// "thread per row" access.
__device__ __forceinline__ uint32_t f32_off(uint32_t r, uint32_t chunk) { return r * 512u + ((chunk ^ (r & 7u)) << 4); }

struct WaitBar {
    uint64_t* bar;
    uint32_t phase;
    __device__ __forceinline__ void wait() {
        mbar_wait(bar, phase);
        phase ^= 1;
    }
};

// Token index of tile row r, or -1 (padding).  Spatial: 7 whole frames per tile (rows = frame*17+joint).
// Temporal: `gpt` whole (clip, joint) sequences per tile (rows = seq*T + t).
template <int MODE>
__device__ __forceinline__ long long row_token(const ModParams& p, int tile, int r) {
    if (MODE == KASF_MODE_SPATIAL) {
        if (r >= 119) return -1;
        const long long tok = (long long)tile * 119 + r;
        return tok < (long long)p.B * p.T * J ? tok : -1;
    } else {
        const int g = r / p.T;
        if (g >= p.groups_per_tile) return -1;
        const long long seq = (long long)tile * p.groups_per_tile + g;
        if (seq >= (long long)p.B * J) return -1;
        const long long b = seq / J;
        const int j = (int)(seq % J), t = r - g * p.T;
        return (b * p.T + t) * J + j;
    }
}

// LayerNorm of the 128 rows of a tile.  SRC_GLOBAL: rows gathered from `src` by row_token (and stashed
// in STASH when `stash` is set); otherwise rows come from STASH.  Writes the bf16 A operand tile and,
// optionally, the fp32 normalised rows to AUX.
template <int MODE, bool SRC_GLOBAL, bool STASH_IT, bool Z_TO_AUX>
__device__ __forceinline__ void ln_tile(const ModParams& p, uint8_t* sm, int tile, const float* src,
                                        const float* gamma, const float* beta, int warp, int lane) {
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(beta + lane * 4);
#pragma unroll 1
    for (int rb = 0; rb < 16; rb += 4) {
        float4 v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = warp + CW * (rb + u);
            if (SRC_GLOBAL) {
                const long long tok = row_token<MODE>(p, tile, r);
                ok[u] = tok >= 0;
                v[u] = ok[u] ? *reinterpret_cast<const float4*>(src + tok * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                ok[u] = true;
                v[u] = *reinterpret_cast<const float4*>(sm + SM_STASH + f32_off(r, lane));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = warp + CW * (rb + u);
            const float mean = wsum(v[u].x + v[u].y + v[u].z + v[u].w) * (1.0f / D);
            const float dx = v[u].x - mean, dy = v[u].y - mean, dz = v[u].z - mean, dw = v[u].w - mean;
            const float var = wsum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
            const float rstd = 1.0f / sqrtf(var + 1e-5f);
            float4 z;
            z.x = ok[u] ? dx * rstd * g4.x + b4.x : 0.f;
            z.y = ok[u] ? dy * rstd * g4.y + b4.y : 0.f;
            z.z = ok[u] ? dz * rstd * g4.z + b4.z : 0.f;
            z.w = ok[u] ? dw * rstd * g4.w + b4.w : 0.f;
            if (STASH_IT) *reinterpret_cast<float4*>(sm + SM_STASH + f32_off(r, lane)) = v[u];
            if (Z_TO_AUX) *reinterpret_cast<float4*>(sm + SM_AUX + f32_off(r, lane)) = z;
            uint2 pk;
            pk.x = pack_bf16(z.x, z.y);
            pk.y = pack_bf16(z.z, z.w);
            *reinterpret_cast<uint2*>(sm + SM_ATILE + tile_off_bf16(r, lane * 4)) = pk;
        }
    }
}

// thread <-> (row, column half) mapping of the TMEM epilogues
struct EpiMap {
    int row;          // tile row == TMEM lane
    int half;         // columns [64*half, 64*half+64)
    uint32_t tbase;   // tmem base + lane offset
};

// x1 = stash + ls1 * mix   (written back to STASH in place), mix read from TMEM cols TM_MIX..+127
template <int KIND, int MODE>
__device__ __forceinline__ void epilogue_mixer(const ModParams& p, uint8_t* sm, const float* vec, const EpiMap& e,
                                               int tile) {
    // GCN: mix = relu(z + BN_node(acc + bU + rowsum*bV)); others: mix = acc + bproj
    float bn_s = 1.f, bn_t = 0.f, rs = 0.f;
    if (KIND == KASF_KIND_GRAPH) {
        int node;
        if (MODE == KASF_MODE_SPATIAL) node = e.row % J;
        else node = e.row % p.T;
        bn_s = vec[V_BNS + node];
        bn_t = vec[V_BNT + node];
        rs = *reinterpret_cast<const float*>(sm + SM_ROWSUM + e.row * 4);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        uint32_t acc[32];
        tmem_ld32(e.tbase + TM_MIX + e.half * 64 + b * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            const int col = e.half * 64 + b * 32 + c4 * 4;
            float4* xs = reinterpret_cast<float4*>(sm + SM_STASH + f32_off(e.row, col >> 2));
            float4 x = *xs;
            const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS1 + col);
            const float4 bm = *reinterpret_cast<const float4*>(vec + V_BMIX + col);
            float m0 = __uint_as_float(acc[c4 * 4 + 0]) + bm.x, m1 = __uint_as_float(acc[c4 * 4 + 1]) + bm.y,
                  m2 = __uint_as_float(acc[c4 * 4 + 2]) + bm.z, m3 = __uint_as_float(acc[c4 * 4 + 3]) + bm.w;
            if (KIND == KASF_KIND_GRAPH) {
                const float4 bv = *reinterpret_cast<const float4*>(vec + V_BV + col);
                const float4 z = *reinterpret_cast<const float4*>(sm + SM_AUX + f32_off(e.row, col >> 2));
                m0 = fmaxf(z.x + ((m0 + rs * bv.x) * bn_s + bn_t), 0.f);
                m1 = fmaxf(z.y + ((m1 + rs * bv.y) * bn_s + bn_t), 0.f);
                m2 = fmaxf(z.z + ((m2 + rs * bv.z) * bn_s + bn_t), 0.f);
                m3 = fmaxf(z.w + ((m3 + rs * bv.w) * bn_s + bn_t), 0.f);
            }
            x.x = fmaf(ls.x, m0, x.x);
            x.y = fmaf(ls.y, m1, x.y);
            x.z = fmaf(ls.z, m2, x.z);
            x.w = fmaf(ls.w, m3, x.w);
            *xs = x;
        }
    }
}

// ---------------------------------------------------------------------------------------------- attention core
// softmax(q k^T / 4) v for every (group, head) of the tile on warp-level tensor-core MMAs
// (mma.sync.m16n8k16 bf16 -> fp32): the problems are 17x17 / TxT with head_dim 16 -- far too small and
// too many for tcgen05 tiles.  Q [128 x 128] bf16 sits in the A tile (operand layout), K|V in AUX.  A work
// item is (group, head, 16-query block); its output overwrites the Q block it consumed.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MAXNT>   // key tiles of 8 this instantiation can hold in registers (gsize <= 8 * MAXNT)
__device__ __forceinline__ void attention_core_impl(uint8_t* sm, int warp, int lane, int gsize, int nrows) {
    const uint32_t q_base = smem_u32(sm + SM_ATILE), kv_base = smem_u32(sm + SM_AUX);
    const int ngroups = nrows / gsize;
    const int mtiles = (gsize + 15) >> 4;
    const int nkt = ((gsize + 15) >> 4) << 1;                     // key tiles, rounded up to pairs
    const int items = ngroups * HEADS * mtiles;
    const int g8 = lane >> 2, t4 = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const float scale = 0.25f * 1.4426950408889634f;              // head_dim^-1/2 * log2(e)
#pragma unroll 1
    for (int item = warp; item < items; item += CW) {
        const int mt = item % mtiles, h = (item / mtiles) % HEADS, g = item / (mtiles * HEADS);
        const int gr0 = g * gsize, m0 = gr0 + mt * 16;
        uint32_t qa[4];
        {
            const int row = min(m0 + (mi & 1) * 8 + r8, 127);
            ldsm_x4(q_base + tile_off_bf16(row, h * DH + (mi >> 1) * 8), qa);
        }
        float s[MAXNT][4];
#pragma unroll
        for (int nt = 0; nt < MAXNT; nt += 2) {
            if (nt < nkt) {
                uint32_t kb[4];
                const int row = min(gr0 + 8 * (nt + (mi >> 1)) + r8, 127);
                ldsm_x4(kv_base + f32_off(row, 2 * h + (mi & 1)), kb);
#pragma unroll
                for (int i = 0; i < 4; ++i) s[nt][i] = 0.f, s[nt + 1][i] = 0.f;
                mma_bf16_16816(s[nt], qa, kb[0], kb[1]);
                mma_bf16_16816(s[nt + 1], qa, kb[2], kb[3]);
            }
        }
        // ---- softmax over the keys of the group (rows g8 and g8+8 of the block), fp32
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt) {
            if (nt < nkt) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int key = nt * 8 + t4 * 2 + (i & 1);
                    s[nt][i] = key < gsize ? s[nt][i] * scale : -INFINITY;
                }
                mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
                mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt) {
            if (nt < nkt) {
                s[nt][0] = ex2_approx(s[nt][0] - mx0);
                s[nt][1] = ex2_approx(s[nt][1] - mx0);
                s[nt][2] = ex2_approx(s[nt][2] - mx1);
                s[nt][3] = ex2_approx(s[nt][3] - mx1);
                l0 += s[nt][0] + s[nt][1];
                l1 += s[nt][2] + s[nt][3];
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        // ---- O = P V  (P re-used from the score registers as the A operand)
        float o[2][4];
#pragma unroll
        for (int dn = 0; dn < 2; ++dn)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[dn][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < MAXNT / 2; ++ks) {
            if (2 * ks < nkt) {
                uint32_t pa[4], vb[4];
                pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
                pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
                pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
                pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
                const int row = min(gr0 + 16 * ks + (mi & 1) * 8 + r8, 127);
                ldsm_x4_t(kv_base + f32_off(row, 16 + 2 * h + (mi >> 1)), vb);
                mma_bf16_16816(o[0], pa, vb[0], vb[1]);
                mma_bf16_16816(o[1], pa, vb[2], vb[3]);
            }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int qr0 = mt * 16 + g8, qr1 = qr0 + 8;               // query index inside the group
#pragma unroll
        for (int dn = 0; dn < 2; ++dn) {
            const int col = h * DH + dn * 8 + t4 * 2;
            if (qr0 < gsize)
                *reinterpret_cast<uint32_t*>(sm + SM_ATILE + tile_off_bf16(gr0 + qr0, col)) = pack_bf16(o[dn][0] * i0, o[dn][1] * i0);
            if (qr1 < gsize)
                *reinterpret_cast<uint32_t*>(sm + SM_ATILE + tile_off_bf16(gr0 + qr1, col)) = pack_bf16(o[dn][2] * i1, o[dn][3] * i1);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void attention_core(uint8_t* sm, int warp, int lane, int gsize, int nrows) {
    if (MODE == KASF_MODE_SPATIAL || gsize <= 32) attention_core_impl<4>(sm, warp, lane, gsize, nrows);
    else if (gsize <= 64) attention_core_impl<8>(sm, warp, lane, gsize, nrows);
    else attention_core_impl<16>(sm, warp, lane, gsize, nrows);
}

// ----------------------------------------------------------------------------------------------
template <int KIND, int MODE>
__global__ void __launch_bounds__(MOD_THREADS, 1) former_module_kernel(const ModParams p) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BARS + B_COUNT * 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* vec = reinterpret_cast<const float*>(p.mod);
    const uint8_t* chunks = p.mod + MOD_VEC_BYTES;

    if (tid == 0) {
        if ((smem_u32(sm) & 1023u) != 0) __trap();
        for (int i = 0; i < B_COUNT; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // chunk consumption order (indices into the module's 12 chunks; see kasf_layout.h)
    constexpr int NCH = KIND == KASF_KIND_GRAPH ? 10 : 12;
    //                         mixer chunks                         MLP: W1_0 W1_1 W2_0 W1_2 W2_1 W1_3 W2_2 W2_3
    constexpr int ORD_ATT[12] = {0, 1, 2, 3, 4, 5, 8, 6, 9, 7, 10, 11};
    constexpr int ORD_BONE[12] = {1, 2, 0, 3, 4, 5, 8, 6, 9, 7, 10, 11};
    constexpr int ORD_GCN[12] = {0, 1, 4, 5, 8, 6, 9, 7, 10, 11, 0, 0};

    if (warp == CW) {
        // ===================== producer warp: stream weight chunks through the ring =====================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
#pragma unroll 1
                for (int i = 0; i < NCH; ++i) {
                    const int ci = KIND == KASF_KIND_ATTENTION ? ORD_ATT[i] : (KIND == KASF_KIND_BONE ? ORD_BONE[i] : ORD_GCN[i]);
                    const uint32_t slot = cnt & 1, ph = (cnt >> 1) & 1;
                    while (!mbar_try_wait(&bars[B_EMPTY0 + slot], ph ^ 1)) __nanosleep(128);
                    mbar_arrive_expect_tx(&bars[B_FULL0 + slot], CHUNK_BYTES);
                    bulk_g2s(sm + SM_RING + slot * CHUNK_BYTES, chunks + (size_t)ci * CHUNK_BYTES, CHUNK_BYTES,
                             &bars[B_FULL0 + slot]);
                    ++cnt;
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== compute warps =====================
        EpiMap e;
        e.row = (warp & 3) * 32 + lane;
        e.half = warp >> 2;
        e.tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t ccnt = 0;   // chunks consumed (meaningful in thread 0)
        WaitBar mma{&bars[B_MMA], 0}, hfull0{&bars[B_HFULL0], 0}, hfull1{&bars[B_HFULL1], 0},
            hsfree0{&bars[B_HSFREE0], 0}, hsfree1{&bars[B_HSFREE1], 0}, outb{&bars[B_OUT], 0};
        const uint32_t a_addr = smem_u32(sm + SM_ATILE);
        const uint32_t ring_addr = smem_u32(sm + SM_RING);
        const uint32_t hs_addr = smem_u32(sm + SM_AUX);

        // issue one weight chunk's MMA: D[tmem col] (+)= A(a_smem) * ring[slot]^T ; frees the slot when done
        auto mma_chunk = [&](uint32_t tcol, uint32_t a_smem, bool acc) {
            const uint32_t slot = ccnt & 1, ph = (ccnt >> 1) & 1;
            mbar_wait(&bars[B_FULL0 + slot], ph);
            tc_fence_after();
            umma_tile_k128(tmem + tcol, a_smem, ring_addr + slot * CHUNK_BYTES, 128, acc);
            tc_commit(&bars[B_EMPTY0 + slot]);
            ++ccnt;
        };

        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            // rows of this tile that carry tokens, and the group geometry
            int gsize, nrows;
            if (MODE == KASF_MODE_SPATIAL) {
                gsize = J;
                const long long left = (long long)p.B * p.T * J - (long long)tile * 119;
                nrows = (int)(left < 119 ? left : 119);
            } else {
                gsize = p.T;
                const long long left = (long long)p.B * J - (long long)tile * p.groups_per_tile;
                nrows = (int)(left < p.groups_per_tile ? left : p.groups_per_tile) * p.T;
            }

            if (KIND == KASF_KIND_BONE) {
                // ---- K,V from the limb stream: LN_limb(XL) Wkv^T
                ln_tile<MODE, true, false, false>(p, sm, tile, p.xl, vec + V_NLW, vec + V_NLB, warp, lane);
                fence_proxy_async();
                tc_fence_before();
                csync();
                if (tid == 0) {
                    mma_chunk(TM_K, a_addr, false);
                    mma_chunk(TM_V, a_addr, false);
                    tc_commit(&bars[B_MMA]);
                }
                mma.wait();   // A tile free again (and K,V complete)
                tc_fence_after();
            }
            // ---- load residual rows, stash them, LN1 -> A operand
            ln_tile<MODE, true, true, KIND == KASF_KIND_GRAPH>(p, sm, tile, p.in, vec + V_N1W, vec + V_N1B, warp, lane);
            fence_proxy_async();
            tc_fence_before();
            csync();

            if (KIND != KASF_KIND_GRAPH) {
                if (tid == 0) {
                    mma_chunk(TM_Q, a_addr, false);
                    if (KIND == KASF_KIND_ATTENTION) {
                        mma_chunk(TM_K, a_addr, false);
                        mma_chunk(TM_V, a_addr, false);
                    }
                    tc_commit(&bars[B_MMA]);
                }
                mma.wait();
                tc_fence_after();
                // ---- Q,K,V: TMEM -> bf16 smem.  Q goes to the (now free) A tile in operand layout, where the
                //      attention output later replaces it block by block; K|V go to AUX (row pitch 512 B).
#pragma unroll
                for (int qkv = 0; qkv < 3; ++qkv)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        uint32_t acc[32];
                        tmem_ld32(e.tbase + qkv * 128 + e.half * 64 + b * 32, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint4 pk;
                            pk.x = pack_bf16(__uint_as_float(acc[c * 8 + 0]), __uint_as_float(acc[c * 8 + 1]));
                            pk.y = pack_bf16(__uint_as_float(acc[c * 8 + 2]), __uint_as_float(acc[c * 8 + 3]));
                            pk.z = pack_bf16(__uint_as_float(acc[c * 8 + 4]), __uint_as_float(acc[c * 8 + 5]));
                            pk.w = pack_bf16(__uint_as_float(acc[c * 8 + 6]), __uint_as_float(acc[c * 8 + 7]));
                            if (qkv == 0) {
                                *reinterpret_cast<uint4*>(sm + SM_ATILE + tile_off_bf16(e.row, e.half * 64 + b * 32 + c * 8)) = pk;
                            } else {
                                const uint32_t chunk = (qkv - 1) * 16 + e.half * 8 + b * 4 + c;
                                *reinterpret_cast<uint4*>(sm + SM_AUX + f32_off(e.row, chunk)) = pk;
                            }
                        }
                    }
                tc_fence_before();
                csync();
                attention_core<MODE>(sm, warp, lane, gsize, nrows);
                fence_proxy_async();
                tc_fence_before();
                csync();
                if (tid == 0) {
                    mma_chunk(TM_MIX, a_addr, false);   // output projection
                    tc_commit(&bars[B_MMA]);
                }
                mma.wait();
                tc_fence_after();
            } else {
                // ================= GCN mixer =================
                if (tid == 0) {
                    mma_chunk(TM_MIX, a_addr, false);   // U z
                    tc_commit(&bars[B_MMA]);
                }
                float* rowsum = reinterpret_cast<float*>(sm + SM_ROWSUM);
                if (MODE == KASF_MODE_TEMPORAL) {
                    // ---- similarity S = z z^T per sequence (fp32), 4th-largest threshold, adjacency bits
                    uint32_t* adj = reinterpret_cast<uint32_t*>(sm + SM_ADJ);
                    uint8_t* degs = sm + SM_DEG;
                    const int T = p.T;
                    const int nib = (T + 3) >> 2;
                    const int ngroups = nrows / T;
#pragma unroll 1
                    for (int item = warp; item < ngroups * nib; item += CW) {
                        const int g = item / nib, ib = item - g * nib;
                        const int gr0 = g * T, i0 = gr0 + ib * 4, gend = gr0 + T;
                        float s[4][4];
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int q = 0; q < 4; ++q) s[a][q] = 0.f;
                        int ri[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) ri[a] = min(i0 + a, gend - 1);
#pragma unroll 2
                        for (int kc = 0; kc < 32; ++kc) {
                            float4 zi[4];
#pragma unroll
                            for (int a = 0; a < 4; ++a)
                                zi[a] = *reinterpret_cast<const float4*>(sm + SM_AUX + f32_off(ri[a], kc));
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int j = lane + 32 * q;
                                if (32 * q < T) {
                                    const int jr = gr0 + min(j, T - 1);
                                    const float4 zj = *reinterpret_cast<const float4*>(sm + SM_AUX + f32_off(jr, kc));
#pragma unroll
                                    for (int a = 0; a < 4; ++a)
                                        s[a][q] = fmaf(zi[a].w, zj.w, fmaf(zi[a].z, zj.z, fmaf(zi[a].y, zj.y, fmaf(zi[a].x, zj.x, s[a][q]))));
                                }
                            }
                        }
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            if (i0 + a >= gend) break;   // warp-uniform
                            float v[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) v[q] = (lane + 32 * q < T) ? s[a][q] : -INFINITY;
                            float thr = 0.f;
                            // k-th largest with multiplicity (torch.topk semantics), k = 4
#pragma unroll 1
                            for (int it = 0; it < 4; ++it) {
                                const float lm = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
                                thr = wmax(lm);
                                const unsigned owners = __ballot_sync(0xffffffffu, lm == thr);
                                if (lane == __ffs(owners) - 1) {   // remove exactly one instance
                                    if (v[0] == thr) v[0] = -INFINITY;
                                    else if (v[1] == thr) v[1] = -INFINITY;
                                    else if (v[2] == thr) v[2] = -INFINITY;
                                    else v[3] = -INFINITY;
                                }
                            }
                            int deg = 0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const unsigned bits = __ballot_sync(0xffffffffu, (lane + 32 * q < T) && s[a][q] >= thr);
                                deg += __popc(bits);
                                if (lane == 0) adj[(i0 + a) * 4 + q] = bits;
                            }
                            if (lane == 0) degs[i0 + a] = (uint8_t)deg;
                        }
                    }
                    csync();
                }
                mma.wait();   // U z done: the A tile may be overwritten (temporal: hidden behind the similarity)
                tc_fence_after();
                // ---- aggregation  agg_i = sum_j A_ij / sqrt(d_i d_j) * z_j   (warp per row) -> bf16 A tile
#pragma unroll 1
                for (int rr = 0; rr < 16; ++rr) {
                    const int r = warp + CW * rr;
                    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float rs = 0.f;
                    if (r < nrows) {
                        if (MODE == KASF_MODE_SPATIAL) {
                            const int j = r % J, base = r - j;
                            const float di = 1.0f / sqrtf((float)c_deg[j]);
#pragma unroll
                            for (int n = 0; n < 4; ++n) {
                                const int nb = c_nbr[j * 4 + n];
                                if (nb >= 0) {
                                    const float cf = di * (1.0f / sqrtf((float)c_deg[nb]));
                                    const float4 z = *reinterpret_cast<const float4*>(sm + SM_AUX + f32_off(base + nb, lane));
                                    a4.x = fmaf(cf, z.x, a4.x), a4.y = fmaf(cf, z.y, a4.y);
                                    a4.z = fmaf(cf, z.z, a4.z), a4.w = fmaf(cf, z.w, a4.w);
                                    rs += cf;
                                }
                            }
                        } else {
                            const uint32_t* adj = reinterpret_cast<const uint32_t*>(sm + SM_ADJ);
                            const uint8_t* degs = sm + SM_DEG;
                            const int gr0 = (r / p.T) * p.T;
                            const float di = 1.0f / sqrtf((float)degs[r]);
#pragma unroll 1
                            for (int q = 0; q < 4; ++q) {
                                unsigned bits = (32 * q < p.T) ? adj[r * 4 + q] : 0u;
                                while (bits) {
                                    const int jb = __ffs(bits) - 1;
                                    bits &= bits - 1;
                                    const int jr = gr0 + 32 * q + jb;
                                    const float cf = di * (1.0f / sqrtf((float)degs[jr]));
                                    const float4 z = *reinterpret_cast<const float4*>(sm + SM_AUX + f32_off(jr, lane));
                                    a4.x = fmaf(cf, z.x, a4.x), a4.y = fmaf(cf, z.y, a4.y);
                                    a4.z = fmaf(cf, z.z, a4.z), a4.w = fmaf(cf, z.w, a4.w);
                                    rs += cf;
                                }
                            }
                        }
                    }
                    if (lane == 0) rowsum[r] = rs;
                    uint2 pk;
                    pk.x = pack_bf16(a4.x, a4.y);
                    pk.y = pack_bf16(a4.z, a4.w);
                    *reinterpret_cast<uint2*>(sm + SM_ATILE + tile_off_bf16(r, lane * 4)) = pk;
                }
                fence_proxy_async();
                tc_fence_before();
                csync();
                if (tid == 0) {
                    mma_chunk(TM_MIX, a_addr, true);    // += (A_hat z) V^T
                    tc_commit(&bars[B_MMA]);
                }
                mma.wait();
                tc_fence_after();
            }

            // ---- x1 = x + ls1 * mixer ; LN2 -> A operand
            epilogue_mixer<KIND, MODE>(p, sm, vec, e, tile);
            tc_fence_before();
            csync();
            ln_tile<MODE, false, false, false>(p, sm, tile, nullptr, vec + V_N2W, vec + V_N2B, warp, lane);
            fence_proxy_async();
            csync();

            // ---- MLP: 4 hidden chunks of 128, software-pipelined over two TMEM / smem buffers
            if (tid == 0) {
                tc_fence_after();
                mma_chunk(TM_H0, a_addr, false);
                tc_commit(&bars[B_HFULL0]);
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int buf = c & 1;
                if (tid == 0 && c + 1 < 4) {
                    mma_chunk(buf ? TM_H0 : TM_H1, a_addr, false);
                    tc_commit(&bars[buf ? B_HFULL0 : B_HFULL1]);
                }
                if (buf) hfull1.wait(); else hfull0.wait();
                tc_fence_after();
                if (c >= 2) { if (buf) hsfree1.wait(); else hsfree0.wait(); }
                // GELU epilogue: TMEM hidden chunk -> bf16 A operand tile in AUX[buf]
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    uint32_t acc[32];
                    tmem_ld32(e.tbase + (buf ? TM_H1 : TM_H0) + e.half * 64 + b * 32, acc);
                    tmem_ld_wait();
                    const float* b1 = vec + V_B1 + c * 128 + e.half * 64 + b * 32;
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const float4 ba = *reinterpret_cast<const float4*>(b1 + c8 * 8);
                        const float4 bb = *reinterpret_cast<const float4*>(b1 + c8 * 8 + 4);
                        uint4 pk;
                        pk.x = pack_bf16(gelu_erf(__uint_as_float(acc[c8 * 8 + 0]) + ba.x), gelu_erf(__uint_as_float(acc[c8 * 8 + 1]) + ba.y));
                        pk.y = pack_bf16(gelu_erf(__uint_as_float(acc[c8 * 8 + 2]) + ba.z), gelu_erf(__uint_as_float(acc[c8 * 8 + 3]) + ba.w));
                        pk.z = pack_bf16(gelu_erf(__uint_as_float(acc[c8 * 8 + 4]) + bb.x), gelu_erf(__uint_as_float(acc[c8 * 8 + 5]) + bb.y));
                        pk.w = pack_bf16(gelu_erf(__uint_as_float(acc[c8 * 8 + 6]) + bb.z), gelu_erf(__uint_as_float(acc[c8 * 8 + 7]) + bb.w));
                        *reinterpret_cast<uint4*>(sm + SM_AUX + buf * TILE_BYTES + tile_off_bf16(e.row, e.half * 64 + b * 32 + c8 * 8)) = pk;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                csync();
                if (tid == 0) {
                    tc_fence_after();
                    mma_chunk(TM_OUT, hs_addr + buf * TILE_BYTES, c > 0);
                    if (c < 2) tc_commit(&bars[buf ? B_HSFREE1 : B_HSFREE0]);
                    if (c == 3) tc_commit(&bars[B_OUT]);
                }
            }
            outb.wait();
            tc_fence_after();
            // ---- out = x1 + ls2 * (acc + b2) -> STASH, then coalesced store
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                uint32_t acc[32];
                tmem_ld32(e.tbase + TM_OUT + e.half * 64 + b * 32, acc);
                tmem_ld_wait();
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const int col = e.half * 64 + b * 32 + c4 * 4;
                    float4* xs = reinterpret_cast<float4*>(sm + SM_STASH + f32_off(e.row, col >> 2));
                    float4 x = *xs;
                    const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS2 + col);
                    const float4 b2 = *reinterpret_cast<const float4*>(vec + V_B2 + col);
                    x.x = fmaf(ls.x, __uint_as_float(acc[c4 * 4 + 0]) + b2.x, x.x);
                    x.y = fmaf(ls.y, __uint_as_float(acc[c4 * 4 + 1]) + b2.y, x.y);
                    x.z = fmaf(ls.z, __uint_as_float(acc[c4 * 4 + 2]) + b2.z, x.z);
                    x.w = fmaf(ls.w, __uint_as_float(acc[c4 * 4 + 3]) + b2.w, x.w);
                    *xs = x;
                }
            }
            tc_fence_before();
            csync();
#pragma unroll 4
            for (int rr = 0; rr < 16; ++rr) {
                const int r = warp + CW * rr;
                const long long tok = row_token<MODE>(p, tile, r);
                if (tok >= 0)
                    *reinterpret_cast<float4*>(p.out + tok * D + lane * 4) =
                        *reinterpret_cast<const float4*>(sm + SM_STASH + f32_off(r, lane));
            }
            csync();   // STASH / AUX / ATILE are reused by the next tile
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KIND, int MODE>
static int launch_one(const ModParams& p, cudaStream_t st) {
    cudaFuncSetAttribute(former_module_kernel<KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    const int grid = p.ntiles < 148 ? p.ntiles : 148;
    former_module_kernel<KIND, MODE><<<grid, MOD_THREADS, SM_TOTAL, st>>>(p);
    return cuda_status();
}

int launch_former_module(const uint8_t* blob, int layer, int kind, int mode, const float* in, const float* XL,
                         float* out, int B, int T, cudaStream_t st) {
    if (B <= 0) return KASF_OK;
    if (kind < 0 || kind > 2 || mode < 0 || mode > 1) return KASF_EINVAL;
    if (kind == KASF_KIND_BONE && !XL) return KASF_EINVAL;
    if (mode == KASF_MODE_TEMPORAL && T > 128) return KASF_ESHAPE;   // TODO: two-tile sequences (T=243)
    ModParams p;
    // module order in the blob: att_s, att_t, graph_s, graph_t, bone_s, bone_t
    p.mod = blob + module_off(layer, kind * 2 + mode);
    p.in = in;
    p.xl = XL;
    p.out = out;
    p.B = B;
    p.T = T;
    if (mode == KASF_MODE_SPATIAL) {
        p.groups_per_tile = 7;
        p.ntiles = (int)(((long long)B * T + 6) / 7);
    } else {
        p.groups_per_tile = 128 / T;
        p.ntiles = (int)(((long long)B * J + p.groups_per_tile - 1) / p.groups_per_tile);
    }
#define KASF_CASE(K, M) \
    if (kind == K && mode == M) return launch_one<K, M>(p, st);
    KASF_CASE(0, 0) KASF_CASE(0, 1) KASF_CASE(1, 0) KASF_CASE(1, 1) KASF_CASE(2, 0) KASF_CASE(2, 1)
#undef KASF_CASE
    return KASF_EINVAL;
}

// ------------------------------------------------------------------ test hook: plain tcgen05 GEMM
// D[M,N] = A[M,128] W[N,128]^T with bf16-rounded operands; exercises the operand layout, the bulk
// copy + mbarrier ring, UMMA descriptors, tcgen05.ld epilogue.  One CTA per 128-row tile.
__global__ void __launch_bounds__(128, 1)
test_gemm_kernel(const float* __restrict__ a, const uint8_t* __restrict__ wchunks, float* __restrict__ d, int M, int N) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 65536);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + 65536 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int row0 = blockIdx.x * 128;
    for (int rr = 0; rr < 32; ++rr) {
        const int r = warp + 4 * rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M) v = *reinterpret_cast<const float4*>(a + (size_t)(row0 + r) * D + lane * 4);
        uint2 pk;
        pk.x = pack_bf16(v.x, v.y);
        pk.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(sm + tile_off_bf16(r, lane * 4)) = pk;
    }
    fence_proxy_async();
    __syncthreads();
    uint32_t ph = 0;
    for (int nc = 0; nc < N / 128; ++nc) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[0], CHUNK_BYTES);
            bulk_g2s(sm + 32768, wchunks + (size_t)nc * CHUNK_BYTES, CHUNK_BYTES, &bars[0]);
            mbar_wait(&bars[0], ph);
            tc_fence_after();
            umma_tile_k128(tmem, smem_u32(sm), smem_u32(sm + 32768), 128, false);
            tc_commit(&bars[1]);
        }
        mbar_wait(&bars[1], ph);
        tc_fence_after();
        ph ^= 1;
        const int r = warp * 32 + lane;
        for (int b = 0; b < 4; ++b) {
            uint32_t acc[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + b * 32, acc);
            tmem_ld_wait();
            if (row0 + r < M)
                for (int i = 0; i < 32; ++i) d[(size_t)(row0 + r) * N + nc * 128 + b * 32 + i] = __uint_as_float(acc[i]);
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 128);
}

__global__ void test_pack_kernel(const float* __restrict__ w, uint8_t* __restrict__ chunks, int N) {
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int n = i >> 6, k = (i & 63) * 2;
        const float* src = w + (size_t)(c * 128 + n) * D + k;
        *reinterpret_cast<uint32_t*>(chunks + (size_t)c * CHUNK_BYTES + tile_off_bf16(n, k)) = pack_bf16(src[0], src[1]);
    }
}

int launch_test_gemm(const float* a, const float* w, float* d, int M, int N, cudaStream_t st) {
    if (M <= 0 || N <= 0 || N % 128) return KASF_ESHAPE;
    uint8_t* chunks = nullptr;   // test hook only: scratch owned for the duration of the call
    if (cudaMalloc(&chunks, (size_t)(N / 128) * CHUNK_BYTES) != cudaSuccess) return KASF_ENOMEM;
    test_pack_kernel<<<N / 128, 256, 0, st>>>(w, chunks, N);
    const int smem = 65536 + 128;
    cudaFuncSetAttribute(test_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    test_gemm_kernel<<<(M + 127) / 128, 128, smem, st>>>(a, chunks, d, M, N);
    int rc = cuda_status();
    cudaStreamSynchronize(st);
    cudaFree(chunks);
    return rc;
}

}  // namespace kasf
