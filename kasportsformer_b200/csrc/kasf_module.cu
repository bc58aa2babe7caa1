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
//   residual rows ........... each thread owns half a row (64 columns): 256-bit global loads issued ONE TILE
//                             AHEAD into registers (the gather latency hides behind the previous tile's MLP),
//                             then resident in TENSOR MEMORY (128 of the 512 TMEM columns) for the whole
//                             module; updated in place by the epilogues; 256-bit stores straight from the
//                             output epilogue (no staging, full 32-byte sectors)
//   dense projections ....... tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands in 128B-swizzled smem;
//                             weights arrive pre-swizzled by bulk async copies (TMA engine) through a
//                             3-slot mbarrier ring fed by a dedicated producer warp
//   LayerNorm ............... fp32, two threads per row (thread = TMEM lane), exact two-pass statistics
//   attention core .......... warp-level tensor-core MMAs (mma.sync m16n8k16 bf16) on Q/K/V bf16 tiles in
//                             smem, fp32 softmax on the accumulator fragments
//   temporal-GCN adjacency .. similarity z z^T to fp32 accuracy on tensor cores (3xTF32 split:
//                             hi*hi + hi*lo + lo*hi), exact 4th-largest selection with multiplicity and ">="
//                             ties, row-sum degrees, D^-1/2 A D^-1/2 applied as a sparse fp32 gather
//   epilogues ............... thread per row straight out of TMEM (tcgen05.ld / tcgen05.st 32x32b)
//
// Shared memory map (bytes):   AUX    K|V bf16 / z fp32 / MLP hidden tiles                 65536
//                              ATILE  bf16 A operand [128 x 128] (also Q, attention output)  32768
//                              RING   3 x weight chunk [128 x 128] bf16                      98304
//                              VEC    the module's fp32 vectors (LN, layer scale, biases)     11776
//                              LN partials, adjacency bit masks, degrees, barriers
// Tensor memory columns:       0..127 Q -> mixer output -> hidden chunk 0 | 128..255 K -> hidden chunk 1
//                              256..383 V -> fc2 accumulator             | 384..511 residual rows X
#include <cstdlib>
#include <cstring>

#include "kasf_internal.h"

namespace kasf {

__constant__ int c_nbr[68] = KASF_NBR;
__constant__ int c_deg[17] = KASF_DEG;
// degree^-1/2 for the skeleton degrees 1..4 (index = degree)
__constant__ float c_rsd[5] = {0.f, 1.0f, 0.70710678118654752440f, 0.57735026918962576451f, 0.5f};

constexpr int CW = 8;                         // compute warps (TMEM epilogues, LayerNorm, mixer cores)
constexpr int MOD_THREADS = (CW + 4) * 32;    // + service warpgroup: weight producer, MMA issuer (+ 2 register donors)
constexpr int W_PRODUCER = CW, W_MMA = CW + 1;
constexpr int RING = 3;
// Three 32 KB buffers:
//   B0 ....... A operand tile: LN1 / LN_limb / A_hat z, then Q -> attention output, then LN2 (fc1)
//   B1|B2 .... staged fp32 rows (gather), then bf16 K|V or fp32 z (GCN), then the two hidden tiles of the MLP
constexpr uint32_t SM_B0 = 0, SM_B1 = 32768, SM_B2 = 65536;
constexpr uint32_t SM_STAGE = SM_B1;                  // fp32 [128][128], XOR-swizzled 16-byte chunks (f32_off)
constexpr uint32_t SM_Z = SM_B1;                      // same layout (GCN: z = LN1(x) in fp32)
constexpr uint32_t SM_KV = SM_B1;                     // bf16 K|V, row pitch 512 B (f32_off)
constexpr uint32_t SM_HS = SM_B1;                     // hidden tiles 0, 1 (operand layout)
constexpr uint32_t SM_A0 = SM_B0, SM_A1 = SM_B0;      // A operand tile (one buffer; the two names mark the two uses)
#ifndef KASF_RING_SHIFT
#define KASF_RING_SHIFT 0
#endif
constexpr uint32_t SM_RING = 98304 + KASF_RING_SHIFT;
constexpr uint32_t SM_VEC = SM_RING + RING * 32768;
constexpr uint32_t SM_PART = SM_VEC + (uint32_t)MOD_VEC_BYTES;   // float2 [128][2]
constexpr uint32_t SM_PART2 = SM_PART + 2048;         // second float2 [128][2]: ln_stats_merge alternates between the two
constexpr uint32_t SM_ADJ = SM_PART2 + 2048;          // u32 [128][4]
constexpr uint32_t SM_ROWSUM = SM_ADJ + 2048;         // f32 [128]
constexpr uint32_t SM_RSD = SM_ROWSUM + 512;          // f32 [128]  degree^-1/2 of the temporal adjacency rows
constexpr uint32_t SM_BARS = SM_RSD + 512;
constexpr uint32_t SM_TOTAL = SM_BARS + 256;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(MOD_VEC_BYTES == 11776, "vector block size");

// Where the small per-tile arrays live and which named barriers pair the two warps of a row: the helpers below are
// shared by this kernel (LayV1) and by the two-tiles-in-flight kernel of kasf_module_v2.cuh (its own layouts)
struct LayV1 {
    static constexpr uint32_t PART = SM_PART, ADJ = SM_ADJ, ROWSUM = SM_ROWSUM, RSD = SM_RSD;
    static constexpr int PAIR_BAR = 2;
};

constexpr uint32_t TM_MIX = 0, TM_K = 128, TM_V = 256, TM_X = 384;   // mixer phase (Q lives at TM_MIX)
constexpr uint32_t TM_H0 = 0, TM_H1 = 128, TM_OUT = 256;            // MLP phase
// MLP phase: the two GELU output chunks as the fc2 A OPERAND IN TENSOR MEMORY (packed 16-bit pairs, lane = row, 64
// columns per 128-wide chunk) in the columns the residual rows occupy during the mixer phase -- x1 stays in registers
constexpr uint32_t TM_HS = 384;

// mbarriers.  Ring: FULL (bulk-copy bytes) / EMPTY (tcgen05.commit).  Compute warps -> MMA warp: AREADY (the A
// operand tile is written), HSREADY (hidden tile c written, hidden accumulator drained).  MMA warp -> compute
// warps: MMA (mixer projections done), HFULL (fc1 chunk in TMEM), HSFREE (fc2 has read the hidden tile),
// OUT (fc2 complete).  ROWS: the cp.async row gather of a tile has landed (one arrival per compute thread).
// Bone modules fed from pre-normalised limb tiles: LIMBFULL (the tile's bf16 limb rows have landed in the A tile by
// bulk copy), A0FREE (tcgen05.commit: the last MMA reading the A tile has completed, the producer may overwrite it),
// OUTDONE (every compute warp has drained the fc2 accumulator of the previous tile: V may be projected into it).
// MMAK / MMAV (self-attention): K / V are in tensor memory -- they are drained to shared memory while the next
// projection runs (one barrier each: a waiter may lag at most one phase behind an mbarrier).
enum { B_FULL0 = 0, B_EMPTY0 = RING, B_AREADY = 2 * RING, B_MMA, B_HFULL0, B_HFULL1, B_HSREADY0, B_HSREADY1,
       B_HSFREE0, B_HSFREE1, B_OUT, B_ROWS, B_MMAK, B_MMAV, B_LIMBFULL, B_A0FREE, B_OUTDONE, B_HFREE0, B_HFREE1, B_COUNT };
static_assert(B_COUNT * 8 + 8 <= 256, "barrier block");

struct ModParams {
    const uint8_t* mod;      // packed module (vector block + chunks)
    const float* in;         // residual stream in   [B,T,17,128]
    const float* xl;         // limb stream (bone modules)
    float* out;              // residual stream out  (may alias in)
    int B, T;
    int ntiles;
    int groups_per_tile;     // temporal: sequences per tile
    unsigned long long* prof;   // optional [24] per-phase cycle counters (debug/profiling hook), else null
    // long sequences (T > KASF_SPLIT_T, split path): scratch in (sequence, frame) row order, row = seq * T + t
    __nv_bfloat16* sq;       // [B*17*T, 128] Q, then the attention output O (in place) | GCN: A_hat z
    __nv_bfloat16* sk;       // [B*17*T, 128] K
    __nv_bfloat16* sv;       // [B*17*T, 128] V
    float* srow;             // [B*17*T] GCN: row sums of A_hat
    // bone modules: optional pre-normalised limb rows, one bf16 [128 x 128] operand-tile image per tile of this mode
    // (limb_tiles_kernel); null: the kernel normalises the fp32 limb rows itself
    const uint8_t* xlt;
};
// internal row mapping of the split path: a tile is 128 consecutive rows of the (sequence, frame) order
#define KASF_MODE_LONG 2

__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// the two warps (w, w+4) that share the rows of one TMEM lane quarter
template <class L = LayV1>
__device__ __forceinline__ void pair_sync(int warp) { asm volatile("bar.sync %0, 64;" ::"r"(L::PAIR_BAR + (warp & 3)) : "memory"); }

// TWICE the erf-GELU, through tanh: v (1 + erf(v / sqrt 2)) = v (1 + tanh(v (a + b v^2 + c v^4))) with a minimax
// fit of (a, b, c) (max abs deviation from the erf form 5.6e-5 on GELU) and the hardware tanh (MUFU, rel. error
// 2^-11).  The result is rounded to bf16 (2^-9) right after, so this is below the operand rounding; the
// exact erff form costs ~5x the instructions and made the MLP epilogue the bottleneck of the kernel.  The
// factor 1/2 lives in the packed fc2 weights (kasf_pack.cu), where it is exact.
// 2*GELU(v) = fma(v, gelu2_tanh(gelu2_arg(v)), v): split so that the epilogue can software-pipeline the MUFU
__device__ __forceinline__ float gelu2_arg(float v) {
#ifdef KASF_EXACT_GELU
    return v;
#else
    // v^2 is clamped where the fitted polynomial peaks (1.70 at v^2 = 48.6): beyond |v| = 7 the tanh is saturated
    // anyway, and the unclamped quartic would turn negative past |v| = 10.7
    const float v2 = fminf(v * v, 48.5f);
    return v * fmaf(v2, fmaf(v2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
#endif
}
__device__ __forceinline__ float gelu2_tanh(float w) {
#ifdef KASF_EXACT_GELU
    return erff(w * 0.70710678118654752440f);
#else
    float t;
    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(w));
    return t;
#endif
}

// The same in packed half precision (default): the fc1 accumulator is rounded to fp16 (2^-11; the hidden tile
// that feeds fc2 is fp16 as well, so nothing is lost against the former bf16 tile, 2^-9), and bias, polynomial,
// tanh (MUFU.TANH.F16) and the final FMA run on two columns per instruction: 5.5 issue slots per element instead
// of 9-11, which leaves the epilogue bound by the MUFU pipe alone (8 cycles per warp instruction).  |fc1 output|
// saturates at 65504 and 2*GELU overflows beyond 32752 -- three orders of magnitude above anything LayerNorm-fed
// projections produce.
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 gelu2_arg_h2(__half2 v) {
    const __half2 s = __hmin2(__hmul2(v, v), __float2half2_rn(48.5f));
    const __half2 p = __hfma2(s, __hfma2(s, __float2half2_rn(-0.0003828259195935171f), __float2half2_rn(0.03722352208203997f)),
                              __float2half2_rn(0.7972238404651819f));
    return __hmul2(v, p);
}
__device__ __forceinline__ __half2 tanh_h2(__half2 w) {
    uint32_t t;
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h2u(w)));
    return u2h(t);
}

// fp32 [128][128] tile with XOR-swizzled 16-byte chunks: conflict-free both for "warp per row" and
// "thread per row" access.
__device__ __forceinline__ uint32_t f32_off(uint32_t r, uint32_t chunk) { return r * 512u + ((chunk ^ (r & 7u)) << 4); }

// Waiter side of the mbarriers of a compute thread: one shared-memory base and ONE register of phase bits (bit i =
// parity the thread expects next on barrier i), instead of a pointer and a phase per barrier
struct Waiter {
    uint32_t base;     // shared-space address of bars[0]
    uint32_t phases;
    __device__ __forceinline__ void wait(int idx) {
        const uint32_t addr = base + idx * 8, parity = (phases >> idx) & 1u;
        uint32_t ok;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
#ifdef KASF_SUSPEND_WAITS
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
#else
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#endif
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(addr), "r"(parity), "r"(0x989680u)
                : "memory");
        } while (!ok);
        phases ^= 1u << idx;
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
    } else if (MODE == KASF_MODE_LONG) {
        const long long R = (long long)tile * 128 + r;
        if (R >= (long long)p.B * J * p.T) return -1;
        const long long seq = R / p.T;
        const int t = (int)(R - seq * p.T);
        return ((seq / J) * p.T + t) * J + seq % J;
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

// thread <-> (row, column half) mapping of everything that touches tensor memory
struct EpiMap {
    int row;          // tile row == TMEM lane
    int half;         // columns [64*half, 64*half+64)
    int warp;
    uint32_t tbase;   // tmem base + lane offset
};

// LayerNorm statistics of a row whose two halves live in two threads (exact two-pass, fp32)
template <class L = LayV1>
__device__ __forceinline__ void ln_stats(uint8_t* sm, const EpiMap& e, const float (&xv)[64], float& mean, float& rstd) {
    float2* part = reinterpret_cast<float2*>(sm + L::PART);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) s0 += xv[i], s1 += xv[i + 1], s2 += xv[i + 2], s3 += xv[i + 3];
    part[e.row * 2 + e.half].x = (s0 + s1) + (s2 + s3);
    pair_sync<L>(e.warp);
    mean = (part[e.row * 2].x + part[e.row * 2 + 1].x) * (1.0f / D);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float d0 = xv[i] - mean, d1 = xv[i + 1] - mean, d2 = xv[i + 2] - mean, d3 = xv[i + 3] - mean;
        q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1), q2 = fmaf(d2, d2, q2), q3 = fmaf(d3, d3, q3);
    }
    part[e.row * 2 + e.half].y = (q0 + q1) + (q2 + q3);
    pair_sync<L>(e.warp);
    const float var = (part[e.row * 2].y + part[e.row * 2 + 1].y) * (1.0f / D);
    rstd = 1.0f / sqrtf(var + 1e-5f);
}

// The same statistics with ONE exchange between the two threads of a row instead of two: each half computes its own
// mean and sum of squared deviations about a pivot (its first value: no cancellation), and the halves are merged with the
// pairwise-variance formula  M2 = M2_a + M2_b + (mean_a - mean_b)^2 * n/2.  Same instruction count as the two-pass form,
// one named barrier and one shared-memory round trip less per LayerNorm; agrees with it to fp32 rounding.  `flip`
// alternates between two partial buffers (a thread may be one LayerNorm ahead of its partner's reads).
__device__ __forceinline__ void ln_stats_merge(uint8_t* sm, const EpiMap& e, const float (&xv)[64], float& mean, float& rstd,
                                               uint32_t& flip) {
    float2* part = reinterpret_cast<float2*>(sm + (flip ? SM_PART2 : SM_PART));
    flip ^= 1u;
    const float pv = xv[0];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        const float d0 = xv[i] - pv, d1 = xv[i + 1] - pv, d2 = xv[i + 2] - pv, d3 = xv[i + 3] - pv;
        s0 += d0, s1 += d1, s2 += d2, s3 += d3;
        q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1), q2 = fmaf(d2, d2, q2), q3 = fmaf(d3, d3, q3);
    }
    const float s = (s0 + s1) + (s2 + s3), q = (q0 + q1) + (q2 + q3);
    const float mh = fmaf(s, 1.0f / 64, pv), m2h = fmaf(-s * (1.0f / 64), s, q);
    part[e.row * 2 + e.half] = make_float2(mh, m2h);
    pair_sync(e.warp);
    const float2 a = part[e.row * 2], b = part[e.row * 2 + 1];
    const float dm = a.x - b.x;
    mean = 0.5f * (a.x + b.x);
    const float var = (a.y + b.y + 32.0f * dm * dm) * (1.0f / D);
    rstd = 1.0f / sqrtf(var + 1e-5f);
}

// z = (x - mean) * rstd [* gamma + beta] for this thread's 64 columns -> bf16 A operand tile (+ fp32 copy in AUX).
// AFFINE = false wherever the LayerNorm's affine is folded into the weights it feeds (kasf_pack.cu): LN1 / LN_limb of
// the attention and bone modules, LN2 of every module; only the GCN's LN1 keeps it (z itself is used in fp32).
template <bool Z_TO_AUX, bool AFFINE>
__device__ __forceinline__ void ln_write(uint8_t* sm, uint32_t a_tile, const EpiMap& e, const float (&xv)[64], float mean,
                                         float rstd, const float* gamma, const float* beta, bool ok) {
    const float nm = -mean * rstd;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int col = e.half * 64 + c * 8;
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaf(xv[c * 8 + i], rstd, nm);
        if (AFFINE) {
            const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(beta + col), b1 = *reinterpret_cast<const float4*>(beta + col + 4);
            z[0] = fmaf(z[0], g0.x, b0.x), z[1] = fmaf(z[1], g0.y, b0.y), z[2] = fmaf(z[2], g0.z, b0.z), z[3] = fmaf(z[3], g0.w, b0.w);
            z[4] = fmaf(z[4], g1.x, b1.x), z[5] = fmaf(z[5], g1.y, b1.y), z[6] = fmaf(z[6], g1.z, b1.z), z[7] = fmaf(z[7], g1.w, b1.w);
        }
        if (!ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] = 0.f;
        }
        uint4 pk;
        pk.x = pack_bf16(z[0], z[1]), pk.y = pack_bf16(z[2], z[3]), pk.z = pack_bf16(z[4], z[5]), pk.w = pack_bf16(z[6], z[7]);
        *reinterpret_cast<uint4*>(sm + a_tile + tile_off_bf16(e.row, col)) = pk;
        if (Z_TO_AUX) {
            *reinterpret_cast<float4*>(sm + SM_Z + f32_off(e.row, col >> 2)) = make_float4(z[0], z[1], z[2], z[3]);
            *reinterpret_cast<float4*>(sm + SM_Z + f32_off(e.row, (col >> 2) + 1)) = make_float4(z[4], z[5], z[6], z[7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- warp-level MMAs
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
__device__ __forceinline__ void mma_tf32_1688(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------- attention core
// softmax(q k^T / 4) v for every (group, head) of the tile on warp-level tensor-core MMAs: the problems are
// 17x17 / TxT with head_dim 16 -- far too small and too many for tcgen05 tiles.  Q [128 x 128] bf16 sits in
// the A tile (operand layout), K|V in AUX.  A work item is (group, head, 16-query block); its output
// overwrites the Q block it consumed.  U independent items are interleaved per warp to hide the
// ldmatrix -> mma -> shuffle -> ex2 -> mma dependency chain.  GS > 0: the group size is a compile-time
// constant (17 joints), so key tiles past the group and the key masks fold away.
// The 17th joint of a spatial group.  17 query rows are one full 16-row MMA block plus ONE row, and a second block for
// that row costs as much as the first (half of the spatial attention core, 3.4k cycles per tile).  Here the 56
// (group, head) leftover queries of a tile go to the CUDA cores instead, four lanes each (eight pairs per warp and
// round, seven rounds per tile): a lane takes the keys s, s+4, .. of the group (dot products over the 16 head
// dimensions, bf16 operands from shared memory, fp32 accumulation), maximum and sum are reduced over the quad, P is
// rounded to bf16 as in the MMA path, P V is reduced with a transpose-reduce (8 + 4 shuffles) that leaves every lane
// with four finished output dimensions.  (Eight lanes per pair, to balance the warps in half-rounds, was slower: 7.0k
// instead of 6.5k cycles for attention core + projection wait; the second MMA block cost 8.0k.)
__device__ __forceinline__ void attention_row16(uint8_t* sm, int round, int lane, int nrows) {
    const int pair = round * 8 + (lane >> 2), s4 = lane & 3;
    const int g = pair >> 3, h = pair & 7;
    const bool live = (g + 1) * J <= nrows;
    const int gr0 = live ? g * J : 0, qrow = gr0 + J - 1;
    const float scale = 0.25f * 1.4426950408889634f;
    auto unpack16 = [](const uint4& a, const uint4& b, float (&f)[16]) {
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) f[2 * i] = __uint_as_float(w[i] << 16), f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    };
    float q[16];
    unpack16(*reinterpret_cast<const uint4*>(sm + SM_A0 + tile_off_bf16(qrow, h * DH)),
             *reinterpret_cast<const uint4*>(sm + SM_A0 + tile_off_bf16(qrow, h * DH + 8)), q);
    float sc[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int key = s4 + 4 * i;                        // (i == 4: key 16, lane 0 of the quad only)
        const int krow = gr0 + min(key, J - 1);
        float k[16];
        unpack16(*reinterpret_cast<const uint4*>(sm + SM_KV + f32_off(krow, 2 * h)),
                 *reinterpret_cast<const uint4*>(sm + SM_KV + f32_off(krow, 2 * h + 1)), k);
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < 16; c += 2) d0 = fmaf(q[c], k[c], d0), d1 = fmaf(q[c + 1], k[c + 1], d1);
        sc[i] = key < J ? d0 + d1 : -INFINITY;
    }
    float mx = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), sc[4]);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float nm = -mx * scale;
    float l = 0.f, o[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) o[c] = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int key = s4 + 4 * i;
        const int krow = gr0 + min(key, J - 1);
        const float pf = ex2_approx(fmaf(sc[i], scale, nm));          // (masked key: ex2(-inf) = 0)
        l += pf;
        const float pb = __uint_as_float(pack_bf16(pf, 0.f) << 16);   // P rounded to bf16, as the MMA operand is
        float v[16];
        unpack16(*reinterpret_cast<const uint4*>(sm + SM_KV + f32_off(krow, 16 + 2 * h)),
                 *reinterpret_cast<const uint4*>(sm + SM_KV + f32_off(krow, 16 + 2 * h + 1)), v);
#pragma unroll
        for (int c = 0; c < 16; ++c) o[c] = fmaf(pb, v[c], o[c]);
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    // transpose-reduce over the quad: after xor 2 a lane holds 8 dimensions (summed over two lanes), after xor 1 four
    float r8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float send = (s4 & 2) ? o[c] : o[c + 8];
        const float keep = (s4 & 2) ? o[c + 8] : o[c];
        r8[c] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    float r4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float send = (s4 & 1) ? r8[c] : r8[c + 4];
        const float keep = (s4 & 1) ? r8[c + 4] : r8[c];
        r4[c] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    const int d0 = ((s4 >> 1) << 3) + ((s4 & 1) << 2);    // this lane's four dimensions
    const float inv = rcp_approx(l);
    __syncwarp();                                          // every lane of the quad has read q before it is overwritten
    if (live) {
        uint2 pk;
        pk.x = pack_bf16(r4[0] * inv, r4[1] * inv), pk.y = pack_bf16(r4[2] * inv, r4[3] * inv);
        *reinterpret_cast<uint2*>(sm + SM_A0 + tile_off_bf16(qrow, h * DH + d0)) = pk;
    }
}

template <int MAXNT, int U, int GS>   // MAXNT: key tiles of 8 held in registers (gsize <= 8 * MAXNT)
__device__ __forceinline__ void attention_core_impl(uint8_t* sm, int warp, int lane, int gsize_rt, int nrows) {
    const uint32_t q_base = smem_u32(sm + SM_A0), kv_base = smem_u32(sm + SM_KV);
    const int gsize = GS ? GS : gsize_rt;
    const int ngroups = nrows / gsize;
    // 17-joint groups: one 16-query block on the tensor cores, the 17th query on the CUDA cores (attention_row16)
    const int mtiles = GS == J ? 1 : (gsize + 15) >> 4;
    const int nkt = (gsize + 7) >> 3;                             // key tiles that hold keys of the group
    const int items = ngroups * HEADS * mtiles;
    // item / mtiles by multiplication (items < 1024, mtiles <= 8: exact); a runtime integer division per item and
    // lane cost 8 % of this phase's issue slots
    const uint32_t inv_mtiles = (65536u + (uint32_t)mtiles - 1u) / (uint32_t)mtiles;
    const int g8 = lane >> 2, t4 = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const float scale = 0.25f * 1.4426950408889634f;              // head_dim^-1/2 * log2(e)
    // every warp takes a contiguous, equally long range of items (spatial: 56 items = 7 per warp, i.e. a round of four
    // and a round of three; a round-robin over rounds of U gave six warps eight items and two warps four)
    const int ipw = (items + CW - 1) / CW, it_end = min(items, (warp + 1) * ipw);
#pragma unroll 1
    for (int it0 = warp * ipw; it0 < it_end; it0 += U) {
        int h[U], gr0[U], mt[U];
        bool live[U];
        uint32_t qa[U][4];
        float s[U][MAXNT][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            live[u] = it0 + u < it_end;
            const int item = live[u] ? it0 + u : it0;
            const int gh = GS ? item / mtiles : (int)(((uint32_t)item * inv_mtiles) >> 16);
            mt[u] = item - gh * mtiles, h[u] = gh & (HEADS - 1);
            gr0[u] = (gh >> 3) * gsize;
            // (query rows past the end of the group repeat its last row: their results are dropped, and no item ever
            //  reads the Q block another item may be overwriting with its output)
            const int row = min(gr0[u] + mt[u] * 16 + (mi & 1) * 8 + r8, gr0[u] + gsize - 1);
            ldsm_x4(q_base + tile_off_bf16(row, h[u] * DH + (mi >> 1) * 8), qa[u]);
        }
#pragma unroll
        for (int nt = 0; nt < MAXNT; nt += 2) {
            if (nt < nkt) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t kb[4];
                    const int row = min(gr0[u] + 8 * (nt + (mi >> 1)) + r8, 127);
                    ldsm_x4(kv_base + f32_off(row, 2 * h[u] + (mi & 1)), kb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) s[u][nt][i] = 0.f, s[u][nt + 1][i] = 0.f;
                    mma_bf16_16816(s[u][nt], qa[u], kb[0], kb[1]);
                    if (nt + 1 < nkt) mma_bf16_16816(s[u][nt + 1], qa[u], kb[2], kb[3]);
                }
            }
        }
        // ---- softmax over the keys of the group (rows g8 and g8+8 of the block), fp32; the 1/4 * log2(e)
        //      scale rides in the FFMA that feeds ex2
        float l0[U], l1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt) {
                if (nt < nkt) {
                    if (nt * 8 + 8 > gsize) {          // only the last key tile of the group can hold keys past its end
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int key = nt * 8 + t4 * 2 + (i & 1);
                            if (!(key < gsize)) s[u][nt][i] = -INFINITY;
                        }
                    }
                    mx0 = fmaxf(mx0, fmaxf(s[u][nt][0], s[u][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[u][nt][2], s[u][nt][3]));
                }
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            const float nm0 = -mx0 * scale, nm1 = -mx1 * scale;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt) {
                if (nt < nkt) {
                    s[u][nt][0] = ex2_approx(fmaf(s[u][nt][0], scale, nm0));
                    s[u][nt][1] = ex2_approx(fmaf(s[u][nt][1], scale, nm0));
                    s[u][nt][2] = ex2_approx(fmaf(s[u][nt][2], scale, nm1));
                    s[u][nt][3] = ex2_approx(fmaf(s[u][nt][3], scale, nm1));
                    a0 += s[u][nt][0] + s[u][nt][1];
                    a1 += s[u][nt][2] + s[u][nt][3];
                }
            }
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            l0[u] = a0, l1[u] = a1;
        }
        // ---- O = P V  (P re-used from the score registers as the A operand)
        float o[U][2][4];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int dn = 0; dn < 2; ++dn)
#pragma unroll
                for (int i = 0; i < 4; ++i) o[u][dn][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < MAXNT / 2; ++ks) {
            if (2 * ks < nkt) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t pa[4], vb[4];
                    pa[0] = pack_bf16(s[u][2 * ks][0], s[u][2 * ks][1]);
                    pa[1] = pack_bf16(s[u][2 * ks][2], s[u][2 * ks][3]);
                    if (2 * ks + 1 < nkt) {
                        pa[2] = pack_bf16(s[u][2 * ks + 1][0], s[u][2 * ks + 1][1]);
                        pa[3] = pack_bf16(s[u][2 * ks + 1][2], s[u][2 * ks + 1][3]);
                    } else {
                        pa[2] = 0u, pa[3] = 0u;
                    }
                    const int row = min(gr0[u] + 16 * ks + (mi & 1) * 8 + r8, 127);
                    ldsm_x4_t(kv_base + f32_off(row, 16 + 2 * h[u] + (mi >> 1)), vb);
                    mma_bf16_16816(o[u][0], pa, vb[0], vb[1]);
                    mma_bf16_16816(o[u][1], pa, vb[2], vb[3]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float i0 = rcp_approx(l0[u]), i1 = rcp_approx(l1[u]);
            const int qr0 = mt[u] * 16 + g8, qr1 = qr0 + 8;           // query index inside the group
            const uint32_t o0 = tile_off_bf16(gr0[u] + qr0, h[u] * DH + t4 * 2), o1 = tile_off_bf16(gr0[u] + qr1, h[u] * DH + t4 * 2);
#pragma unroll
            for (int dn = 0; dn < 2; ++dn) {
                // columns h*16 + dn*8 + ..: the next 16-byte chunk of the row, i.e. chunk index ^ 1 after swizzling
                const uint32_t x0 = dn ? (o0 ^ 16u) : o0, x1 = dn ? (o1 ^ 16u) : o1;
                if (live[u] && qr0 < gsize)
                    *reinterpret_cast<uint32_t*>(sm + SM_A0 + x0) = pack_bf16(o[u][dn][0] * i0, o[u][dn][1] * i0);
                if (live[u] && qr1 < gsize)
                    *reinterpret_cast<uint32_t*>(sm + SM_A0 + x1) = pack_bf16(o[u][dn][2] * i1, o[u][dn][3] * i1);
            }
        }
    }
}

// TC = sequence-length class of the instantiation: 0: group size <= 32, 1: <= 64, 2: <= 128 (key tiles held in
// registers).  One class per kernel keeps the cold variants out of the register allocation.
// work items interleaved per warp (measured, profiles/r01z_attention_interleave.txt): two for short temporal groups,
// four for the 17-joint spatial groups
#ifndef KASF_ATT_U_T0
#define KASF_ATT_U_T0 2
#endif
#ifndef KASF_ATT_U_S
#define KASF_ATT_U_S 4
#endif
template <int MODE, int TC>
__device__ __forceinline__ void attention_core(uint8_t* sm, int warp, int lane, int gsize, int nrows) {
    if (MODE == KASF_MODE_SPATIAL) {
        attention_core_impl<4, KASF_ATT_U_S, J>(sm, warp, lane, gsize, nrows);
        // seven rounds of 8 (group, head) pairs, one per warp 1..7.  (Measured: an MMA round of four items costs 2.1k
        // cycles, a round here 2.3k; warps 0-5 have two MMA rounds, warps 6, 7 one.  Two rounds for warps 6, 7 and one for
        // warps 3-5 was slower: 6.9k instead of 6.5k cycles for attention core + projection wait.)
        if (warp >= 1) attention_row16(sm, 7 - warp, lane, nrows);
    }
    else if (TC == 0) attention_core_impl<4, KASF_ATT_U_T0, 0>(sm, warp, lane, gsize, nrows);
    else if (TC == 1) attention_core_impl<8, 2, 0>(sm, warp, lane, gsize, nrows);
    else attention_core_impl<16, 1, 0>(sm, warp, lane, gsize, nrows);
}

// ---------------------------------------------------------------------------------------------- top-4 networks
__device__ __forceinline__ void cmpx(float& a, float& b) {   // a >= b afterwards
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    a = hi, b = lo;
}
__device__ __forceinline__ void sort4_desc(float (&c)[4]) {
    cmpx(c[0], c[1]), cmpx(c[2], c[3]), cmpx(c[0], c[2]), cmpx(c[1], c[3]), cmpx(c[1], c[2]);
}
// a <- the four largest of (a U b), sorted descending; a and b sorted descending (duplicates kept)
__device__ __forceinline__ void merge_top4(float (&a)[4], const float (&b)[4]) {
    float c[4] = {fmaxf(a[0], b[3]), fmaxf(a[1], b[2]), fmaxf(a[2], b[1]), fmaxf(a[3], b[0])};   // bitonic
    cmpx(c[0], c[2]), cmpx(c[1], c[3]), cmpx(c[0], c[1]), cmpx(c[2], c[3]);
    a[0] = c[0], a[1] = c[1], a[2] = c[2], a[3] = c[3];
}

// ---------------------------------------------------------------------------------------------- temporal adjacency
// S = z z^T per sequence to fp32 accuracy on tensor cores (3xTF32: the operand is split into the 19 bits the
// tensor core reads and the exact fp32 remainder; hi*hi + hi*lo + lo*hi), then per row the 4th largest value
// with multiplicity (torch.topk, graph.py:109) and A_ij = S_ij >= thr (:111).  Work item = (sequence,
// 16-row block); the row's bit mask and degree go to smem.
template <int MAXNT, class L = LayV1>
__device__ __forceinline__ void similarity_topk_impl(uint8_t* sm, int warp, int lane, int T, int nrows) {
    uint32_t* adj = reinterpret_cast<uint32_t*>(sm + L::ADJ);
    float* rsd = reinterpret_cast<float*>(sm + L::RSD);
    const int ngroups = nrows / T, mtiles = (T + 15) >> 4, nkt = (T + 7) >> 3;
    const int g8 = lane >> 2, t4 = lane & 3;
#pragma unroll 1
    for (int item = warp; item < ngroups * mtiles; item += CW) {
        const int g = item / mtiles, mt = item - g * mtiles;
        const int gr0 = g * T, m0 = gr0 + mt * 16;
        const int ra = min(m0 + g8, 127), rb = min(m0 + g8 + 8, 127);
        // hi*hi and the two cross terms accumulate in separate registers (short sequences): two independent MMA
        // dependency chains per key tile instead of one three times as long
#ifndef KASF_SIM_SPLIT_ACC
#define KASF_SIM_SPLIT_ACC 1
#endif
        constexpr bool SPLIT_ACC = KASF_SIM_SPLIT_ACC && MAXNT <= 8;
        float s[MAXNT][4], sx[SPLIT_ACC ? MAXNT : 1][4];
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
        if (SPLIT_ACC) {
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) sx[nt][i] = 0.f;
        }
        // K = 128 in eight double steps of 16 columns.  The k index of an MMA is a free permutation as long as A and B
        // agree: lane t4 of a quad takes the four consecutive columns 16 d + 4 t4 .. + 3 of its rows with ONE 128-bit
        // load (first k-step: columns +0 / +1 in slots t4 / t4+4, second k-step: +2 / +3), and every address is a
        // per-row base plus an immediate (scalar loads with per-load swizzle arithmetic made this phase
        // instruction-bound: 3.9k instructions per work item).
        const uint32_t zbase = SM_Z;   // offsets from sm: plain shared-memory loads the compiler may schedule freely
        const uint32_t zrow_a = zbase + ra * 512 + ((t4 ^ (ra & 7)) << 4);
        const uint32_t zrow_b = zbase + rb * 512 + ((t4 ^ (rb & 7)) << 4);
        uint32_t zrow_j[MAXNT];
#pragma unroll
        for (int nt = 0; nt < MAXNT; ++nt) {
            const int rj = min(gr0 + 8 * nt + g8, 127);
            zrow_j[nt] = zbase + rj * 512 + ((t4 ^ (rj & 7)) << 4);
        }
        auto split = [](float v, uint32_t& hi, uint32_t& lo) {
            hi = __float_as_uint(v) & 0xffffe000u;
            lo = __float_as_uint(v - __uint_as_float(hi));
        };
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            // chunk 4 d + t4: 128-byte segment d >> 1, swizzled slot (t4 ^ x) [^ 4 for odd d]
            const int off = (d >> 1) * 128;
            const float4 fa = *reinterpret_cast<const float4*>(sm + ((zrow_a + off) ^ ((d & 1) << 6)));
            const float4 fb = *reinterpret_cast<const float4*>(sm + ((zrow_b + off) ^ ((d & 1) << 6)));
            uint32_t ah[2][4], al[2][4];
            split(fa.x, ah[0][0], al[0][0]), split(fb.x, ah[0][1], al[0][1]), split(fa.y, ah[0][2], al[0][2]), split(fb.y, ah[0][3], al[0][3]);
            split(fa.z, ah[1][0], al[1][0]), split(fb.z, ah[1][1], al[1][1]), split(fa.w, ah[1][2], al[1][2]), split(fb.w, ah[1][3], al[1][3]);
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt) {
                if (nt < nkt) {
                    const float4 fj = *reinterpret_cast<const float4*>(sm + ((zrow_j[nt] + off) ^ ((d & 1) << 6)));
                    uint32_t bh[4], bl[4];
                    split(fj.x, bh[0], bl[0]), split(fj.y, bh[1], bl[1]), split(fj.z, bh[2], bl[2]), split(fj.w, bh[3], bl[3]);
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        if (SPLIT_ACC) {
                            mma_tf32_1688(s[nt], ah[k2], bh[2 * k2], bh[2 * k2 + 1]);
                            mma_tf32_1688(sx[nt], al[k2], bh[2 * k2], bh[2 * k2 + 1]);
                            mma_tf32_1688(sx[nt], ah[k2], bl[2 * k2], bl[2 * k2 + 1]);
                        } else {
                            mma_tf32_1688(s[nt], al[k2], bh[2 * k2], bh[2 * k2 + 1]);
                            mma_tf32_1688(s[nt], ah[k2], bl[2 * k2], bl[2 * k2 + 1]);
                            mma_tf32_1688(s[nt], ah[k2], bh[2 * k2], bh[2 * k2 + 1]);
                        }
                    }
                }
            }
        }
        if (SPLIT_ACC) {
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) s[nt][i] += sx[nt][i];
        }
        // ---- rows g8 (values s[nt][0..1]) and g8+8 (s[nt][2..3]); a row is spread over the 4 lanes of a quad
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            float v[MAXNT][2];
#pragma unroll
            for (int nt = 0; nt < MAXNT; ++nt)
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    v[nt][i] = (nt < nkt && nt * 8 + t4 * 2 + i < T) ? s[nt][hrow * 2 + i] : -INFINITY;
            // 4th largest value of the row with multiplicity (torch.topk + ">=", graph.py:109-111), by sorting
            // networks: every lane reduces its 2 MAXNT values to a sorted top-4 (groups of four: 5 compare-exchanges,
            // then a bitonic top-4 merge per further group), the quad merges its four lists with two shuffle rounds.
            // (Four rounds of "quad maximum, remove one instance" cost 4.7k cycles per tile.)
            float best[4];
#pragma unroll
            for (int g4 = 0; g4 < MAXNT / 2; ++g4) {
                float c[4] = {v[2 * g4][0], v[2 * g4][1], v[2 * g4 + 1][0], v[2 * g4 + 1][1]};
                sort4_desc(c);
                if (g4 == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) best[i] = c[i];
                } else {
                    merge_top4(best, c);
                }
            }
            {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = __shfl_xor_sync(0xffffffffu, best[i], 1);
                merge_top4(best, o);
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = __shfl_xor_sync(0xffffffffu, best[i], 2);
                // the 4th largest of the union of two sorted top-4 lists: min_i max(a_i, b_{3-i})
                best[0] = fminf(fminf(fmaxf(best[0], o[3]), fmaxf(best[1], o[2])), fminf(fmaxf(best[2], o[1]), fmaxf(best[3], o[0])));
            }
            const float thr = best[0];
            // adjacency bits of this row: key 8 nt + 2 t4 + i
            const int row = m0 + g8 + hrow * 8;
            int deg = 0;
#pragma unroll
            for (int w = 0; w < (MAXNT + 3) / 4; ++w) {
                uint32_t bits = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int nt = w * 4 + q;
                    if (nt < MAXNT && nt < nkt) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            if (nt * 8 + t4 * 2 + i < T && s[nt][hrow * 2 + i] >= thr) bits |= 1u << (q * 8 + t4 * 2 + i);
                    }
                }
                bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
                bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
                deg += __popc(bits);
                if (t4 == 0 && row < gr0 + T) adj[row * 4 + w] = bits;
            }
            if (t4 == 0 && row < gr0 + T) rsd[row] = 1.0f / sqrtf((float)deg);
        }
    }
}

template <int TC, class L = LayV1>
__device__ __forceinline__ void similarity_topk(uint8_t* sm, int warp, int lane, int T, int nrows) {
    similarity_topk_impl<TC == 0 ? 4 : (TC == 1 ? 8 : 16), L>(sm, warp, lane, T, nrows);
}

// ---------------------------------------------------------------------------------------------- tile geometry
// rows of tile `tile` that carry tokens (rows [0, n) are valid, the rest is padding)
template <int MODE>
__device__ __forceinline__ int tile_rows(const ModParams& p, int tile) {
    if (MODE == KASF_MODE_SPATIAL) {
        const long long left = (long long)p.B * p.T * J - (long long)tile * 119;
        return (int)(left < 119 ? left : 119);
    } else if (MODE == KASF_MODE_LONG) {
        const long long left = (long long)p.B * J * p.T - (long long)tile * 128;
        return (int)(left < 128 ? left : 128);
    } else {
        const long long left = (long long)p.B * J - (long long)tile * p.groups_per_tile;
        return (int)(left < p.groups_per_tile ? left : p.groups_per_tile) * p.T;
    }
}

// Row gather of a tile, issued by the compute warps themselves (a service warp gets an LDGSTS through every ~75
// cycles next to busy compute warps -- measured -- while 8 warps x 16 instructions are gone in a few hundred):
// warp w copies rows w, w+8, ..., one coalesced 512-byte cp.async (16 B per lane) per row into the staging buffer
// (swizzled like every fp32 tile, so the thread-per-row reads are conflict-free); every lane then attaches the
// completion of its copies to `bar`.  Issued right after fc2 completes: the rows land behind the output epilogue.
template <int MODE>
__device__ __forceinline__ void gather_rows(const ModParams& p, uint8_t* sm, int tile, const float* src, uint64_t* bar,
                                            int warp, int lane, int part) {
    // part 0: rows 0..63 (they land in B1, free as soon as fc2 of hidden chunk 2 has read it), part 1: rows 64..127
    // (B2, free after the last fc2 chunk).  Every lane attaches its copies to `bar` once per part.
    const int n = tile_rows<MODE>(p, tile);
    const int lo = part * 64, hi = min(n, lo + 64);
    if (MODE == KASF_MODE_SPATIAL) {
        const float* g = src + (long long)tile * 119 * D + lane * 4;
#pragma unroll 4
        for (int r = lo + warp; r < hi; r += CW) cp_async16(sm + SM_STAGE + f32_off(r, lane), g + (size_t)r * D, 16u);
    } else {
        // lane l (< 8) knows the token of row lo + warp + 8 l; broadcast row by row
        const int myrow = lo + warp + CW * (lane & 7);
        const int mytok = myrow < hi ? (int)row_token<MODE>(p, tile, myrow) : 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int tok = __shfl_sync(0xffffffffu, mytok, k);
            const int r = lo + warp + CW * k;
            if (r < hi) cp_async16(sm + SM_STAGE + f32_off(r, lane), src + (size_t)tok * D + lane * 4, 16u);
        }
    }
    cp_async_mbar_arrive(bar);
}

// The same for the 16 rows that this warp's lane quarter and column half own: rows (warp & 3) * 32 + (warp >> 2) * 16 + k.
// Arrives twice per lane (the barrier counts two half gathers per tile).
template <int MODE>
__device__ __forceinline__ void gather_rows_owned(const ModParams& p, uint8_t* sm, int tile, const float* src, uint64_t* bar,
                                                  int warp, int lane) {
    const int n = tile_rows<MODE>(p, tile);
    const int r0 = (warp & 3) * 32 + (warp >> 2) * 16;
    if (MODE == KASF_MODE_SPATIAL) {
        const float* g = src + (long long)tile * 119 * D + lane * 4;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
            const int r = r0 + k;
            if (r < n) cp_async16(sm + SM_STAGE + f32_off(r, lane), g + (size_t)r * D, 16u);
        }
    } else {
        const int myrow = r0 + (lane & 15);
        const int mytok = myrow < n ? (int)row_token<MODE>(p, tile, myrow) : 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int tok = __shfl_sync(0xffffffffu, mytok, k);
            const int r = r0 + k;
            if (r < n) cp_async16(sm + SM_STAGE + f32_off(r, lane), src + (size_t)tok * D + lane * 4, 16u);
        }
    }
    cp_async_mbar_arrive(bar);
    cp_async_mbar_arrive(bar);
}

// this thread's 64 staged values of its row (zeros for padding rows)
__device__ __forceinline__ void read_staged(const uint8_t* sm, const EpiMap& e, float (&xv)[64], bool ok) {
    if (ok) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(sm + SM_STAGE + f32_off(e.row, e.half * 16 + c));
            xv[c * 4] = v.x, xv[c * 4 + 1] = v.y, xv[c * 4 + 2] = v.z, xv[c * 4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) xv[i] = 0.f;
    }
}

// compute warp -> MMA warp: this warp's part of a shared-memory operand tile is written (generic proxy) and
// its tensor-memory reads are complete
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// ----------------------------------------------------------------------------------------------
// Warp roles: 8 compute warps | weight producer (1 lane) | MMA issuer (1 lane) | row gatherer (1 warp).
// The compute warps never issue an MMA and (apart from the mixer cores, which exchange rows through shared
// memory) never meet at a CTA barrier: every hand-over is an mbarrier, so warps drift apart and the MUFU-bound
// GELU epilogues of one warp overlap the tensor-memory loads, stores and MMAs triggered by the others.
// MODE == KASF_MODE_LONG is the tail of the split path for sequences that do not pack into a tile (T > KASF_SPLIT_T): the mixer
// core ran in its own kernels (kasf_long section below) and left the attention output / A_hat z in scratch;
// this kernel then does projection -> residual -> LN2 -> MLP -> residual on plain 128-row tiles.
// PROF: the per-phase cycle counters of kasf_former_module_profiled are compiled in (own instantiations, so that the
// production kernels carry neither the clock registers nor the sixteen checks per tile)
template <int KIND, int MODE, int TC, bool PROF = false>
__global__ void __launch_bounds__(MOD_THREADS, 1) former_module_kernel(const ModParams p) {
    constexpr bool POST = MODE == KASF_MODE_LONG;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BARS + B_COUNT * 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* vec = reinterpret_cast<const float*>(sm + SM_VEC);
    const uint8_t* chunks = p.mod + MOD_VEC_BYTES;

    if (tid == 0) {
        if ((smem_u32(sm) & 1023u) != 0) __trap();
        for (int i = 0; i < B_COUNT; ++i) {
            const bool by_warps = i == B_AREADY || i == B_HSREADY0 || i == B_HSREADY1 || i == B_HFREE0 || i == B_HFREE1;
            mbar_init(&bars[i], (by_warps || i == B_OUTDONE) ? CW : (i == B_ROWS ? CW * 32 * 2 : 1));   // ROWS: two half gathers
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = tid; i < V_FLOATS / 4; i += MOD_THREADS)   // the module's fp32 vectors, once per CTA
        reinterpret_cast<float4*>(sm + SM_VEC)[i] = reinterpret_cast<const float4*>(p.mod)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // chunk consumption order (indices into the module's 12 chunks; see kasf_layout.h)
    constexpr int NCH = KIND == KASF_KIND_GRAPH ? 10 : (POST ? 9 : 12);
    constexpr int ORD_POST[12] = {3, 4, 5, 6, 8, 7, 9, 10, 11, 0, 0, 0};   // attention / bone tail: projection + MLP
    //                         mixer chunks                         MLP: W1_0 W1_1 W1_2 W2_0 W1_3 W2_1 W2_2 W2_3
    constexpr int ORD_ATT[12] = {1, 2, 0, 3, 4, 5, 6, 8, 7, 9, 10, 11};   // K, V, Q: K|V drain while Q runs
    constexpr int ORD_BONE[12] = {1, 2, 0, 3, 4, 5, 6, 8, 7, 9, 10, 11};
    constexpr int ORD_GCN[12] = {0, 1, 4, 5, 6, 8, 7, 9, 10, 11, 0, 0};
    const bool limb_tiles = KIND == KASF_KIND_BONE && !POST && p.xlt != nullptr;
    const float* first_src = (KIND == KASF_KIND_BONE && !POST && !limb_tiles) ? p.xl : p.in;

    if (warp >= CW) {
        // ===================== service warpgroup (hands its registers to the compute warpgroups) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == W_PRODUCER) {
            // ---- weight chunks: L2 -> ring slots (bulk copies, MMA-ready swizzled images)
            if (lane == 0) {
                uint32_t slot = 0, ph = 0;
                uint32_t ph_free = 0;
                for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                    if (limb_tiles) {
                        // the tile's normalised limb rows: one 32 KB bulk copy straight into the A operand tile, as soon
                        // as the previous tile's last fc1 chunk has read it
                        if (tile != (int)blockIdx.x) {
                            mbar_wait_suspend(&bars[B_A0FREE], ph_free);
                            ph_free ^= 1;
                        }
                        mbar_arrive_expect_tx(&bars[B_LIMBFULL], TILE_BYTES);
                        bulk_g2s(sm + SM_A1, p.xlt + (size_t)tile * TILE_BYTES, TILE_BYTES, &bars[B_LIMBFULL]);
                    }
#pragma unroll 1
                    for (int i = 0; i < NCH; ++i) {
                        const int ci = KIND == KASF_KIND_GRAPH ? ORD_GCN[i]
                                       : (POST ? ORD_POST[i] : (KIND == KASF_KIND_ATTENTION ? ORD_ATT[i] : ORD_BONE[i]));
                        mbar_wait_suspend(&bars[B_EMPTY0 + slot], ph ^ 1);
                        mbar_arrive_expect_tx(&bars[B_FULL0 + slot], CHUNK_BYTES);
                        bulk_g2s(sm + SM_RING + slot * CHUNK_BYTES, chunks + (size_t)ci * CHUNK_BYTES, CHUNK_BYTES,
                                 &bars[B_FULL0 + slot]);
                        if (++slot == RING) slot = 0, ph ^= 1;
                    }
                }
            }
        } else if (warp == W_MMA && lane == 0) {
            // ---- the only thread that issues tcgen05.mma
            const uint32_t a0_addr = smem_u32(sm + SM_A0), a1_addr = smem_u32(sm + SM_A1), ring_addr = smem_u32(sm + SM_RING);
            uint32_t cslot = 0, cph = 0, ph_a = 0, ph_hs0 = 0, ph_hs1 = 0, ph_limb = 0, ph_hf = 0;
            auto chunk = [&](uint32_t tcol, uint32_t a_smem, bool acc, bool fp16_operands = false) {
                mbar_wait(&bars[B_FULL0 + cslot], cph);
                tc_fence_after();
                umma_tile_k128(tmem + tcol, a_smem, ring_addr + cslot * CHUNK_BYTES, 128, acc, fp16_operands);
                tc_commit(&bars[B_EMPTY0 + cslot]);
                if (++cslot == RING) cslot = 0, cph ^= 1;
            };
            // fc2 chunk: A = GELU chunk in tensor memory (8 columns per K-step of 16), B = ring slot
            auto chunk_ts = [&](uint32_t tcol, uint32_t a_tmem, bool acc) {
                mbar_wait(&bars[B_FULL0 + cslot], cph);
                tc_fence_after();
                const uint32_t idesc = KASF_HALF_GELU ? umma_idesc_f16(128, 128) : umma_idesc_bf16(128, 128);
                const uint64_t db = umma_desc_sw128(ring_addr + cslot * CHUNK_BYTES);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_ts(tmem + tcol, a_tmem + ks * 8, db + (uint64_t)(((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4), idesc,
                            (acc || ks > 0) ? 1u : 0u);
                tc_commit(&bars[B_EMPTY0 + cslot]);
                if (++cslot == RING) cslot = 0, cph ^= 1;
            };
            auto wait_a = [&]() {
                mbar_wait(&bars[B_AREADY], ph_a);
                ph_a ^= 1;
                tc_fence_after();
            };
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                if (POST && KIND != KASF_KIND_GRAPH) {
                    // Q, K, V were projected by the pre kernel
                } else if (KIND == KASF_KIND_ATTENTION) {
                    wait_a();                                  // LN1(x)
                    chunk(TM_K, a1_addr, false);
                    tc_commit(&bars[B_MMAK]);
                    chunk(TM_V, a1_addr, false);
                    tc_commit(&bars[B_MMAV]);
                    chunk(TM_MIX, a1_addr, false);             // Q last: it is drained into the A tile itself
                    tc_commit(&bars[B_MMA]);
                } else if (KIND == KASF_KIND_BONE) {
                    if (limb_tiles) {
                        // K|V of this tile are projected while the compute warps still run the previous tile's output
                        // epilogue; V lands in the fc2 accumulator's columns and waits for them to be drained
                        mbar_wait(&bars[B_LIMBFULL], ph_limb);
                        tc_fence_after();
                        chunk(TM_K, a1_addr, false);
                        mbar_wait(&bars[B_OUTDONE], ph_limb);
                        ph_limb ^= 1;
                        tc_fence_after();
                        chunk(TM_V, a1_addr, false);
                    } else {
                        wait_a();                              // LN_limb(XL)
                        chunk(TM_K, a1_addr, false);
                        chunk(TM_V, a1_addr, false);
                    }
                    tc_commit(&bars[B_MMA]);
                    wait_a();                                  // LN1(x)
                    chunk(TM_MIX, a1_addr, false);             // Q
                    tc_commit(&bars[B_MMA]);
                } else {
                    wait_a();                                  // z = LN1(x)
                    chunk(TM_MIX, a1_addr, false);             // U z
                    tc_commit(&bars[B_MMA]);
                }
                wait_a();                                      // attention output (B0) | A_hat z (B2)
                chunk(TM_MIX, KIND == KASF_KIND_GRAPH ? a1_addr : a0_addr, KIND == KASF_KIND_GRAPH);   // proj | += (A_hat z) V^T
                tc_commit(&bars[B_MMA]);
                // ---- MLP: fc1 chunks run two ahead of the GELU epilogues, fc2 accumulates behind them
                wait_a();                                      // LN2(x1)
                chunk(TM_H0, a0_addr, false);
                tc_commit(&bars[B_HFULL0]);
                chunk(TM_H1, a0_addr, false);
                tc_commit(&bars[B_HFULL1]);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int buf = c & 1;
                    if (c < 2) {
                        // fc1 chunk c+2 as soon as every warp has READ accumulator c (a few hundred cycles into its GELU
                        // epilogue), not when the epilogue is finished: fc2 of chunk c then follows its HSREADY directly,
                        // and the hidden-operand buffer it frees is what the epilogue of chunk c+2 waits for
                        mbar_wait(&bars[buf ? B_HFREE1 : B_HFREE0], ph_hf);
                        tc_fence_after();
                        chunk(buf ? TM_H1 : TM_H0, a0_addr, false);
                        tc_commit(&bars[buf ? B_HFULL1 : B_HFULL0]);
                        if (c == 1 && limb_tiles) tc_commit(&bars[B_A0FREE]);   // last reader of the A tile
                    }
                    if (buf) { mbar_wait(&bars[B_HSREADY1], ph_hs1); ph_hs1 ^= 1; }
                    else { mbar_wait(&bars[B_HSREADY0], ph_hs0); ph_hs0 ^= 1; }
                    tc_fence_after();
                    if (c == 1) ph_hf ^= 1;
                    chunk_ts(TM_OUT, tmem + TM_HS + buf * 64, c > 0);                          // fc2: fp16 x fp16
                    if (c < 3) tc_commit(&bars[buf ? B_HSFREE1 : B_HSFREE0]);   // (c == 2: B1 is free for the next tile's rows)
                    if (c == 3) tc_commit(&bars[B_OUT]);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== compute warps =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        EpiMap e;
        e.row = (warp & 3) * 32 + lane;
        e.half = warp >> 2;
        e.warp = warp;
        e.tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        Waiter wt{smem_u32(bars), 0u};
        uint32_t ln_flip = 0;

        long long pt0 = PROF ? clock64() : 0;
#define PMARK(k)                                                      \
    do {                                                              \
        if (PROF && tid == 0) {                                       \
            const long long pt1 = clock64();                          \
            atomicAdd(p.prof + (k), (unsigned long long)(pt1 - pt0)); \
            pt0 = pt1;                                                \
        }                                                             \
    } while (0)

#ifdef KASF_STAGGER
        {   // experiment: de-synchronise the SMs (equal tiles keep all 148 in the same phase: bursts on L2 / HBM)
            const long long t0 = clock64();
            const long long d = (long long)(blockIdx.x & 7) * KASF_STAGGER;
            while (clock64() - t0 < d) {}
            pt0 = PROF ? clock64() : 0;
        }
#endif
        gather_rows<MODE>(p, sm, blockIdx.x, first_src, &bars[B_ROWS], warp, lane, 0);
        gather_rows<MODE>(p, sm, blockIdx.x, first_src, &bars[B_ROWS], warp, lane, 1);
        if (limb_tiles && lane == 0) mbar_arrive(&bars[B_OUTDONE]);   // no previous tile: the accumulator columns are free
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int nrows = tile_rows<MODE>(p, tile);
            const int gsize = MODE == KASF_MODE_SPATIAL ? J : p.T;
            const bool row_ok = e.row < nrows;
            const long long tok = row_ok ? row_token<MODE>(p, tile, e.row) : -1;
            float xv[64];
            float mean, rstd;
            // The next tile's row gather, issued right after LN2 has been handed to the tensor cores -- where the warps
            // would otherwise just wait for the first fc1 chunk -- instead of behind the last fc2 chunks.  The staging
            // buffer (= K|V / z / the former shared-memory hidden tiles) is dead by then: the projection MMA, which this
            // thread has seen complete, waited for the AREADY arrival of all eight warps, so every warp has left the mixer
            // core, and the MLP no longer touches shared memory (its hidden activation lives in tensor memory).  Only a
            // graph module still reads z in its mixer epilogue -- each thread its own row -- so a warp gathers exactly
            // the rows that it and its partner warp (same rows, other column half) own, after meeting that partner.
            auto next_rows = [&]() {
                if (tile + (int)gridDim.x < p.ntiles) {
                    if (KIND == KASF_KIND_GRAPH) pair_sync(e.warp);
                    gather_rows_owned<MODE>(p, sm, tile + (int)gridDim.x, first_src, &bars[B_ROWS], warp, lane);
                }
            };

            if (KIND == KASF_KIND_BONE && !POST && !limb_tiles) {
                // ---- K,V from the limb stream: LN_limb(XL) Wkv^T
                wt.wait(B_ROWS);
                read_staged(sm, e, xv, row_ok);
                csync();                                   // every limb row is in registers: the staging buffer
                gather_rows<MODE>(p, sm, tile, p.in, &bars[B_ROWS], warp, lane, 0);   // receives the residual rows
                gather_rows<MODE>(p, sm, tile, p.in, &bars[B_ROWS], warp, lane, 1);
                ln_stats_merge(sm, e, xv, mean, rstd, ln_flip);
                ln_write<false, false>(sm, SM_A1, e, xv, mean, rstd, nullptr, nullptr, row_ok);
                warp_arrive(&bars[B_AREADY], lane);
                PMARK(0);
            }
            // split path: this row's mixer-core output (bf16, scratch) is requested first, its latency runs under the
            // staged-row reads and the tensor-memory stores
            uint4 pre8[8];
            if (POST) {
                const uint4* src = reinterpret_cast<const uint4*>(p.sq + ((long long)tile * 128 + e.row) * D + e.half * 64);
#pragma unroll
                for (int c = 0; c < 8; ++c) pre8[c] = row_ok ? __ldg(src + c) : make_uint4(0u, 0u, 0u, 0u);
            }
            // ---- residual rows: staging -> registers -> tensor memory (resident); LN1 -> A operand
            wt.wait(B_ROWS);
            PMARK(13);
            read_staged(sm, e, xv, row_ok);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                uint32_t xr[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) xr[i] = __float_as_uint(xv[b * 32 + i]);
                tmem_st32(e.tbase + TM_X + e.half * 64 + b * 32, xr);
            }
            if (POST && KIND != KASF_KIND_GRAPH) {
                // ---- split path: the attention output of this tile's rows (bf16, scratch) is the A operand
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(sm + SM_A0 + tile_off_bf16(e.row, e.half * 64 + c * 8)) = pre8[c];
            } else {
                ln_stats_merge(sm, e, xv, mean, rstd, ln_flip);
                if (KIND == KASF_KIND_BONE) {
                    wt.wait(B_MMA);                            // K,V complete: the A tile may be overwritten
                    tc_fence_after();
                }
                if (KIND == KASF_KIND_GRAPH) csync();      // z (fp32) overwrites the staging rows of other threads
                ln_write<KIND == KASF_KIND_GRAPH, KIND == KASF_KIND_GRAPH>(sm, SM_A1, e, xv, mean, rstd, vec + V_N1W, vec + V_N1B, row_ok);
            }
            tmem_st_wait();
            warp_arrive(&bars[B_AREADY], lane);
            PMARK(1);

            if (KIND != KASF_KIND_GRAPH && POST) {
                wt.wait(B_MMA);                                // output projection
                tc_fence_after();
                PMARK(5);
            } else if (KIND != KASF_KIND_GRAPH) {
                // ---- Q,K,V: TMEM -> bf16 smem.  K|V go to AUX (row pitch 512 B) as soon as they are complete -- while
                //      the next projection still runs on the tensor cores; Q goes last, into the (then free) A tile
                //      in operand layout, where the attention output later replaces it block by block.
                auto drain = [&](int qkv) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        uint32_t acc[32];
                        tmem_ld32(e.tbase + qkv * 128 + e.half * 64 + b * 32, acc);
                        tmem_ld_wait();
                        if (qkv == 0) {                    // query bias W_q beta_1 (LN1's affine lives in the weights)
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4) {
                                const float4 bq = *reinterpret_cast<const float4*>(vec + V_BQ + e.half * 64 + b * 32 + c4 * 4);
                                acc[c4 * 4 + 0] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 0]) + bq.x);
                                acc[c4 * 4 + 1] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 1]) + bq.y);
                                acc[c4 * 4 + 2] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 2]) + bq.z);
                                acc[c4 * 4 + 3] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 3]) + bq.w);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint4 pk;
                            pk.x = pack_bf16(__uint_as_float(acc[c * 8 + 0]), __uint_as_float(acc[c * 8 + 1]));
                            pk.y = pack_bf16(__uint_as_float(acc[c * 8 + 2]), __uint_as_float(acc[c * 8 + 3]));
                            pk.z = pack_bf16(__uint_as_float(acc[c * 8 + 4]), __uint_as_float(acc[c * 8 + 5]));
                            pk.w = pack_bf16(__uint_as_float(acc[c * 8 + 6]), __uint_as_float(acc[c * 8 + 7]));
                            if (qkv == 0) {
                                *reinterpret_cast<uint4*>(sm + SM_A0 + tile_off_bf16(e.row, e.half * 64 + b * 32 + c * 8)) = pk;
                            } else {
                                const uint32_t chunk = (qkv - 1) * 16 + e.half * 8 + b * 4 + c;
                                *reinterpret_cast<uint4*>(sm + SM_KV + f32_off(e.row, chunk)) = pk;
                            }
                        }
                    }
                };
                if (KIND == KASF_KIND_ATTENTION) {
                    // K complete => every warp has arrived on AREADY, i.e. has read its staged rows: AUX is free
                    wt.wait(B_MMAK);
                    tc_fence_after();
                    drain(1);
                    wt.wait(B_MMAV);
                    tc_fence_after();
                    drain(2);
                } else {
                    csync();                               // every residual row has left the staging buffer (= AUX)
                    drain(1);                              // K,V were complete before LN1(x) was written
                    drain(2);
                }
                wt.wait(B_MMA);                                // Q in tensor memory
                tc_fence_after();
                PMARK(2);
                drain(0);
                csync();                                   // every warp reads the K|V rows of the others
                PMARK(3);
                attention_core<MODE, TC>(sm, warp, lane, gsize, nrows);
                warp_arrive(&bars[B_AREADY], lane);
                PMARK(4);
                wt.wait(B_MMA);                                // output projection
                tc_fence_after();
                PMARK(5);
            } else {
                // ================= GCN mixer =================
                if (!POST) csync();                        // z of the whole tile is in shared memory
                if (MODE == KASF_MODE_TEMPORAL) {
                    similarity_topk<TC>(sm, warp, lane, p.T, nrows);
                    csync();
                    PMARK(6);
                }
                // ---- aggregation  agg_i = sum_j A_ij / sqrt(d_i d_j) * z_j : this thread's 64 columns of its row
                //      (measured slower: all four neighbour slots branch-free with zero weights -- twice the loads, 5.8k vs
                //       4.4k cycles per tile; warp-per-row with conflict-free row loads -- the rows of a warp serialise
                //       their adjacency -> degree -> row dependency chains, 7.5k / 18k)
                float rs = 0.f;
                uint4 agg8[8];                             // split path: A_hat z of this row (bf16) from scratch
                if (POST) {
                    const long long R = (long long)tile * 128 + e.row;
#pragma unroll
                    for (int c = 0; c < 8; ++c) agg8[c] = pre8[c];
                    rs = row_ok ? __ldg(p.srow + R) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 64; ++i) xv[i] = 0.f;
                if (row_ok && !POST) {
                    if (MODE == KASF_MODE_SPATIAL) {
                        const int j = e.row % J, base = e.row - j;
                        const float di = c_rsd[c_deg[j]];
#pragma unroll 1
                        for (int n = 0; n < 4; ++n) {
                            const int nb = c_nbr[j * 4 + n];
                            if (nb < 0) break;
                            const float cf = di * c_rsd[c_deg[nb]];
                            rs += cf;
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const float4 z = *reinterpret_cast<const float4*>(sm + SM_Z + f32_off(base + nb, e.half * 16 + c));
                                xv[c * 4] = fmaf(cf, z.x, xv[c * 4]), xv[c * 4 + 1] = fmaf(cf, z.y, xv[c * 4 + 1]);
                                xv[c * 4 + 2] = fmaf(cf, z.z, xv[c * 4 + 2]), xv[c * 4 + 3] = fmaf(cf, z.w, xv[c * 4 + 3]);
                            }
                        }
                    } else {
                        const uint32_t* adj = reinterpret_cast<const uint32_t*>(sm + SM_ADJ);
                        const float* rsd = reinterpret_cast<const float*>(sm + SM_RSD);
                        const int gr0 = (e.row / p.T) * p.T;
                        const float di = rsd[e.row];
#pragma unroll 1
                        for (int q = 0; q < 4; ++q) {
                            unsigned bits = (32 * q < p.T) ? adj[e.row * 4 + q] : 0u;
#pragma unroll 1
                            while (bits) {
                                const int jb = __ffs(bits) - 1;
                                bits &= bits - 1;
                                const int jr = gr0 + 32 * q + jb;
                                const float cf = di * rsd[jr];
                                rs += cf;
#pragma unroll
                                for (int c = 0; c < 16; ++c) {
                                    const float4 z = *reinterpret_cast<const float4*>(sm + SM_Z + f32_off(jr, e.half * 16 + c));
                                    xv[c * 4] = fmaf(cf, z.x, xv[c * 4]), xv[c * 4 + 1] = fmaf(cf, z.y, xv[c * 4 + 1]);
                                    xv[c * 4 + 2] = fmaf(cf, z.z, xv[c * 4 + 2]), xv[c * 4 + 3] = fmaf(cf, z.w, xv[c * 4 + 3]);
                                }
                            }
                        }
                    }
                }
                if (e.half == 0) *reinterpret_cast<float*>(sm + SM_ROWSUM + e.row * 4) = rs;
                wt.wait(B_MMA);   // U z done: the A tile may be overwritten
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 pk;
                    if (POST) {
                        pk = agg8[c];
                    } else {
                        pk.x = pack_bf16(xv[c * 8 + 0], xv[c * 8 + 1]), pk.y = pack_bf16(xv[c * 8 + 2], xv[c * 8 + 3]);
                        pk.z = pack_bf16(xv[c * 8 + 4], xv[c * 8 + 5]), pk.w = pack_bf16(xv[c * 8 + 6], xv[c * 8 + 7]);
                    }
                    *reinterpret_cast<uint4*>(sm + SM_A1 + tile_off_bf16(e.row, e.half * 64 + c * 8)) = pk;
                }
                pair_sync(e.warp);                         // the row sum written by the partner thread (half 0)
                warp_arrive(&bars[B_AREADY], lane);
                PMARK(7);
                wt.wait(B_MMA);                                // += (A_hat z) V^T
                tc_fence_after();
                PMARK(8);
            }

            // ---- x1 = x + ls1 * mixer  (TMEM -> registers; x1 STAYS in registers until the output epilogue: its
            //      tensor-memory columns hold the GELU chunks during the MLP), LN2 -> A operand
            {
                // GCN: mix = relu(z + BN_node(acc + bU + rowsum*bV)); others: mix = acc + bproj
                float bn_s = 1.f, bn_t = 0.f, rs = 0.f;
                if (KIND == KASF_KIND_GRAPH) {
                    const int node = MODE == KASF_MODE_SPATIAL ? e.row % J
                                     : (POST ? (int)(((long long)tile * 128 + e.row) % p.T) : e.row % p.T);
                    bn_s = vec[V_BNS + node];
                    bn_t = vec[V_BNT + node];
                    rs = *reinterpret_cast<const float*>(sm + SM_ROWSUM + e.row * 4);
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    uint32_t acc[32], xr[32];
                    tmem_ld32(e.tbase + TM_MIX + e.half * 64 + b * 32, acc);
                    tmem_ld32(e.tbase + TM_X + e.half * 64 + b * 32, xr);
                    tmem_ld_wait();
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const int col = e.half * 64 + b * 32 + c4 * 4;
                        const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS1 + col);
                        const float4 bm = *reinterpret_cast<const float4*>(vec + V_BMIX + col);
                        float m0 = __uint_as_float(acc[c4 * 4 + 0]) + bm.x, m1 = __uint_as_float(acc[c4 * 4 + 1]) + bm.y,
                              m2 = __uint_as_float(acc[c4 * 4 + 2]) + bm.z, m3 = __uint_as_float(acc[c4 * 4 + 3]) + bm.w;
                        if (KIND == KASF_KIND_GRAPH) {
                            const float4 bv = *reinterpret_cast<const float4*>(vec + V_BV + col);
                            const float4 z = *reinterpret_cast<const float4*>(sm + SM_Z + f32_off(e.row, col >> 2));
                            m0 = fmaxf(z.x + ((m0 + rs * bv.x) * bn_s + bn_t), 0.f);
                            m1 = fmaxf(z.y + ((m1 + rs * bv.y) * bn_s + bn_t), 0.f);
                            m2 = fmaxf(z.z + ((m2 + rs * bv.z) * bn_s + bn_t), 0.f);
                            m3 = fmaxf(z.w + ((m3 + rs * bv.w) * bn_s + bn_t), 0.f);
                        }
                        xv[b * 32 + c4 * 4 + 0] = fmaf(ls.x, m0, __uint_as_float(xr[c4 * 4 + 0]));
                        xv[b * 32 + c4 * 4 + 1] = fmaf(ls.y, m1, __uint_as_float(xr[c4 * 4 + 1]));
                        xv[b * 32 + c4 * 4 + 2] = fmaf(ls.z, m2, __uint_as_float(xr[c4 * 4 + 2]));
                        xv[b * 32 + c4 * 4 + 3] = fmaf(ls.w, m3, __uint_as_float(xr[c4 * 4 + 3]));
                    }
                }
            }
            PMARK(9);
            ln_stats_merge(sm, e, xv, mean, rstd, ln_flip);
            ln_write<false, false>(sm, SM_A0, e, xv, mean, rstd, nullptr, nullptr, row_ok);
            warp_arrive(&bars[B_AREADY], lane);
            PMARK(10);
            next_rows();

            // ---- MLP epilogues: hidden chunk c (fc1 accumulator in TMEM) -> 2*GELU -> 16-bit A operand tile AUX[c & 1]
            //      (requesting the next 32 accumulator columns while the current ones are computed was measured
            //       slower: tcgen05.ld next to running MMAs has ~250 cycles of latency either way, and the shorter
            //       software pipelines lose more than the overlap wins)
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int buf = c & 1;
                wt.wait(buf ? B_HFULL1 : B_HFULL0);
                tc_fence_after();
                if (c >= 2) { wt.wait(buf ? B_HSFREE1 : B_HSFREE0); }
                PMARK(c == 0 ? 17 : 14);
                uint32_t acc[2][32];
                tmem_ld32(e.tbase + (buf ? TM_H1 : TM_H0) + e.half * 64, acc[0]);
                tmem_ld32(e.tbase + (buf ? TM_H1 : TM_H0) + e.half * 64 + 32, acc[1]);
                tmem_ld_wait();
                if (c < 2) {                                   // accumulator c is in registers: fc1 of chunk c+2 may overwrite it
                    tc_fence_before();
                    warp_arrive(&bars[buf ? B_HFREE1 : B_HFREE0], lane);
                }
                PMARK(15);
                // (bias + polynomial of two of the eight groups on the fp32 FMA pipe instead of in packed half, to relieve
                //  the half-rate fp16 pipe, was measured slower: 5.7k instead of 5.4k cycles per tile, 14.2k vs 14.5k clips/s
                //  -- the loop is bound by issue slots and dependent chains, not by that pipe)
                // two-stage software pipeline over groups of 8 columns: the tanh arguments of group g+1 are computed
                // while the MUFU results of group g are in flight (the compiler's own schedule consumed each result
                // a few instructions after issuing it: 1960 vs 1440 cycles per chunk, scripts/micro/gelu_epi.cu)
                uint32_t hsw[32];         // this thread's 64 hidden values of the chunk as 16-bit pairs: 32 TMEM columns
#if KASF_HALF_GELU
                const uint4* b1h = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(vec + V_B1H) + c * 128 + e.half * 64);
                __half2 v[2][4], w[2][4];
                auto stage1 = [&](int g, int s2) {
                    const uint4 bh = b1h[g];
                    const uint32_t* a8 = &acc[g >> 2][(g & 3) * 8];
                    v[s2][0] = __hadd2(u2h(pack_f16(__uint_as_float(a8[0]), __uint_as_float(a8[1]))), u2h(bh.x));
                    v[s2][1] = __hadd2(u2h(pack_f16(__uint_as_float(a8[2]), __uint_as_float(a8[3]))), u2h(bh.y));
                    v[s2][2] = __hadd2(u2h(pack_f16(__uint_as_float(a8[4]), __uint_as_float(a8[5]))), u2h(bh.z));
                    v[s2][3] = __hadd2(u2h(pack_f16(__uint_as_float(a8[6]), __uint_as_float(a8[7]))), u2h(bh.w));
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[s2][i] = gelu2_arg_h2(v[s2][i]);
                };
                stage1(0, 0);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int s2 = g & 1;
                    __half2 t[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) t[i] = tanh_h2(w[s2][i]);
                    if (g + 1 < 8) stage1(g + 1, s2 ^ 1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) hsw[g * 4 + i] = h2u(__hfma2(v[s2][i], t[i], v[s2][i]));
                }
#else
                const float* b1 = vec + V_B1 + c * 128 + e.half * 64;
                float v[2][8], w[2][8];
                auto stage1 = [&](int g, int s2) {
                    const float4 ba = *reinterpret_cast<const float4*>(b1 + g * 8);
                    const float4 bb = *reinterpret_cast<const float4*>(b1 + g * 8 + 4);
                    const uint32_t* a8 = &acc[g >> 2][(g & 3) * 8];
                    v[s2][0] = __uint_as_float(a8[0]) + ba.x, v[s2][1] = __uint_as_float(a8[1]) + ba.y;
                    v[s2][2] = __uint_as_float(a8[2]) + ba.z, v[s2][3] = __uint_as_float(a8[3]) + ba.w;
                    v[s2][4] = __uint_as_float(a8[4]) + bb.x, v[s2][5] = __uint_as_float(a8[5]) + bb.y;
                    v[s2][6] = __uint_as_float(a8[6]) + bb.z, v[s2][7] = __uint_as_float(a8[7]) + bb.w;
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[s2][i] = gelu2_arg(v[s2][i]);
                };
                stage1(0, 0);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int s2 = g & 1;
                    float t[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) t[i] = gelu2_tanh(w[s2][i]);
                    if (g + 1 < 8) stage1(g + 1, s2 ^ 1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) hsw[g * 4 + i] = pack_bf16(fmaf(v[s2][2 * i], t[2 * i], v[s2][2 * i]), fmaf(v[s2][2 * i + 1], t[2 * i + 1], v[s2][2 * i + 1]));
                }
#endif
                // the chunk as the fc2 A operand in tensor memory: lane = row, column c = hidden (2c, 2c+1); one store
                // instead of eight swizzled 16-byte shared-memory stores, and fc2 reads only its weights from shared memory
                tmem_st32(e.tbase + TM_HS + buf * 64 + e.half * 32, hsw);
                tmem_st_wait();
                warp_arrive(&bars[buf ? B_HSREADY1 : B_HSREADY0], lane);
                PMARK(11);
            }
            // (reading the next tile's staged rows and taking their LN1 statistics here, under the last fc2 chunks, was
            //  measured slower -- 13.2k vs 13.9k clips/s: 64 more live registers through the out epilogue cost more
            //  than the ~0.9k cycles of wait they fill)
            wt.wait(B_HSFREE0);                            // (fc2 of chunk 2: keeps the barrier's phases in step)
            wt.wait(B_OUT);
            tc_fence_after();
            PMARK(16);
            // ---- out = x1 + ls2 * (acc + b2).  The four lanes of a quad own four rows; a 4 x 4 transpose of their 8-float
            //      pieces (two xor-shuffle stages) lets the quad write ONE 128-byte line per store instruction (8 lines of
            //      4 sectors per warp instruction) instead of 32 lines of one sector each.  Measured (scripts/micro/
            //      stg_patterns.cu): 3.94k cycles per 64 KB tile for the thread-per-row pattern on a single SM, 2.04k for
            //      line-per-quad, 0.39k for the shuffles.  (Staging the rows in shared memory for coalesced stores was
            //      slower than either: the extra CTA barriers and the second pass over the data cost more than they save.
            //      Keeping the finished rows in registers and issuing these stores in the NEXT tile, behind its LN1
            //      hand-over where the warps wait 2.2k cycles for the Q|K|V MMAs, moved 2.0k cycles of stores under
            //      that wait and won nothing: the wait shrank by 0.75k and the LN / drain / attention phases that
            //      followed grew by as much -- the stores drain through the same load/store pipeline as the
            //      shared-memory traffic of those phases.  25.3k vs 25.2k cycles per tile, bone modules 4 % slower.)
            {
                const int s4 = lane & 3;
                long long tokr[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) tokr[r] = __shfl_sync(0xffffffffu, tok, (lane & ~3) + r);
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    uint32_t acc[32];
                    tmem_ld32(e.tbase + TM_OUT + e.half * 64 + b * 32, acc);
                    tmem_ld_wait();
                    const float* xr = xv + b * 32;         // x1, in registers since the mixer epilogue
                    float o[4][8];                         // piece c8 = columns 8 c8 .. 8 c8 + 7 of this block
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const int col = e.half * 64 + b * 32 + c8 * 8;
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS2 + col + h4 * 4);
                            const float4 b2 = *reinterpret_cast<const float4*>(vec + V_B2 + col + h4 * 4);
                            const int i = c8 * 8 + h4 * 4;
                            o[c8][h4 * 4 + 0] = fmaf(ls.x, __uint_as_float(acc[i + 0]) + b2.x, xr[i + 0]);
                            o[c8][h4 * 4 + 1] = fmaf(ls.y, __uint_as_float(acc[i + 1]) + b2.y, xr[i + 1]);
                            o[c8][h4 * 4 + 2] = fmaf(ls.z, __uint_as_float(acc[i + 2]) + b2.z, xr[i + 2]);
                            o[c8][h4 * 4 + 3] = fmaf(ls.w, __uint_as_float(acc[i + 3]) + b2.w, xr[i + 3]);
                        }
                    }
                    // transpose among the quad: afterwards o[r] = piece s4 of the row owned by lane (quad base + r)
#pragma unroll
                    for (int i = 0; i < 2; ++i)            // stage xor 2: o[i + 2] of lanes 0,1 <-> o[i] of lanes 2,3
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float send = (s4 & 2) ? o[i][k] : o[i + 2][k];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
                            if (s4 & 2) o[i][k] = recv; else o[i + 2][k] = recv;
                        }
#pragma unroll
                    for (int i = 0; i < 4; i += 2)         // stage xor 1: o[i + 1] of even lanes <-> o[i] of odd lanes
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float send = (s4 & 1) ? o[i][k] : o[i + 1][k];
                            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                            if (s4 & 1) o[i][k] = recv; else o[i + 1][k] = recv;
                        }
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (tokr[r] >= 0) stg256(p.out + tokr[r] * D + e.half * 64 + b * 32 + s4 * 8, o[r]);
                }
            }
            if (limb_tiles) warp_arrive(&bars[B_OUTDONE], lane);   // fc2 accumulator drained: V of the next tile may land
            // (no CTA barrier here: the next tile's first shared-memory writes touch buffers whose last readers
            //  were MMAs already observed complete by every thread, and its MMAs wait for AREADY)
            PMARK(12);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KIND, int MODE, int TC>
static int launch_one(const ModParams& p, cudaStream_t st) {
    const int sms = sm_count();
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    if constexpr (TC == 0 && MODE != KASF_MODE_LONG) {     // profiling hook: short sequences only
        if (p.prof) {
            cudaFuncSetAttribute(former_module_kernel<KIND, MODE, TC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
            former_module_kernel<KIND, MODE, TC, true><<<grid, MOD_THREADS, SM_TOTAL, st>>>(p);
            return cuda_status();
        }
    }
    cudaFuncSetAttribute(former_module_kernel<KIND, MODE, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    former_module_kernel<KIND, MODE, TC><<<grid, MOD_THREADS, SM_TOTAL, st>>>(p);
    return cuda_status();
}

#include "kasf_module_v2.cuh"

// ============================================================================================== split path (T > KASF_SPLIT_T)
// A temporal sequence that does not pack into a 128-row tile -- longer than the tile (T = 243), or so long that one
// sequence per tile would leave a third of the MMA rows empty (T = 81) -- runs as two or three kernels over bf16
// scratch in (sequence, frame) row order:
//   attention / bone:  long_pre_kernel (LN1 [, limb operand] -> Q, K, V on tcgen05 -> bf16 scratch)
//                      long_attention_kernel (one CTA per sequence and four heads: softmax(q k^T / 4) v)
//                      former_module_kernel<KIND, KASF_MODE_LONG> (projection, residual, LN2, MLP, residual)
//   graph:             lgt::long_gcn_tc_kernel (kasf_long_gcn.cuh, one CTA per sequence: z = LN1(x) as three bf16
//                      pieces, fp32-accurate similarity, exact 4th-largest threshold, degrees, A_hat z and row sums
//                      on the tcgen05 tensor cores -> scratch)
//                      former_module_kernel<GRAPH, KASF_MODE_LONG> (U z + (A_hat z) V^T, BN, residuals, MLP)
// Same arithmetic as the fused kernels (bf16 operands, fp32 accumulation / statistics / softmax / similarity).

// ---- Q, K, V projection of 128 consecutive (sequence, frame) rows.  Persistent: one CTA per SM walks the tiles
//      with the three weight chunks (96 KB), the vector block, the barriers and the tensor-memory allocation set
//      up once; 8 warps, thread = half a row.  The fp32 rows of the NEXT tile are gathered with cp.async into the
//      staging buffer while the MMAs of the current tile run and its accumulators are drained (256-bit stores: one
//      full sector per thread and instruction), and each of Q, K, V is drained as soon as its own MMAs complete.
//      A bone module takes the limb operand of a tile ready-made (limb_tiles_kernel<LONG>: one 32 KB bulk copy
//      straight into the A tile, requested as soon as the tile's last MMA has read the previous operand); without
//      limb tiles it normalises the fp32 limb rows itself.
//      (The first version was one CTA per tile and paid the 96 KB weight load, the allocation and every latency in
//      series: 23.7k cycles per tile at T = 81, more than the tail kernel's whole MLP.)
template <int KIND>
__global__ void __launch_bounds__(256, 1) long_pre_kernel(const ModParams p) {
    extern __shared__ __align__(1024) uint8_t sm[];
    enum { BW = 0, BR = 1, BL = 2, BQ = 3, BK = 4, BV = 5 };
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BARS + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* vec = reinterpret_cast<const float*>(sm + SM_VEC);
    const uint8_t* chunks = p.mod + MOD_VEC_BYTES;
    const bool lt = KIND == KASF_KIND_BONE && p.xlt != nullptr;
    const float* first_src = (KIND == KASF_KIND_BONE && !lt) ? p.xl : p.in;
    if (tid == 0) {
        mbar_init(&bars[BW], 1);
        mbar_init(&bars[BR], 512);                         // every lane, once per 64-row part
        mbar_init(&bars[BL], 1);
        mbar_init(&bars[BQ], 1);
        mbar_init(&bars[BK], 1);
        mbar_init(&bars[BV], 1);
        fence_mbar_init();
        // ring slot i <- chunk ORD[i]: attention Q,K,V = 0,1,2; bone K,V,Q = 1,2,0
        mbar_arrive_expect_tx(&bars[BW], 3 * CHUNK_BYTES);
        for (int i = 0; i < 3; ++i) {
            const int ci = KIND == KASF_KIND_ATTENTION ? i : (i + 1) % 3;
            bulk_g2s(sm + SM_RING + i * CHUNK_BYTES, chunks + (size_t)ci * CHUNK_BYTES, CHUNK_BYTES, &bars[BW]);
        }
        if (lt) {
            mbar_arrive_expect_tx(&bars[BL], TILE_BYTES);
            bulk_g2s(sm + SM_A0, p.xlt + (size_t)blockIdx.x * TILE_BYTES, TILE_BYTES, &bars[BL]);
        }
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = tid; i < V_FLOATS / 4; i += 256) reinterpret_cast<float4*>(sm + SM_VEC)[i] = reinterpret_cast<const float4*>(p.mod)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gather_rows<KASF_MODE_LONG>(p, sm, blockIdx.x, first_src, &bars[BR], warp, lane, 0);
    gather_rows<KASF_MODE_LONG>(p, sm, blockIdx.x, first_src, &bars[BR], warp, lane, 1);
    const uint32_t tmem = *tmem_slot;
    EpiMap e;
    e.row = (warp & 3) * 32 + lane, e.half = warp >> 2, e.warp = warp;
    e.tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_addr = smem_u32(sm + SM_A0), ring = smem_u32(sm + SM_RING);
    float xv[64], mean, rstd;
    uint32_t ph_r = 0, ph_l = 0, ph_m = 0;
    if (tid == 0) mbar_wait(&bars[BW], 0);                 // weights resident (the issuing thread is their only reader)
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const bool row_ok = e.row < tile_rows<KASF_MODE_LONG>(p, tile);
        const int next = tile + (int)gridDim.x;
        const long long R = (long long)tile * 128 + e.row;
        auto drain = [&](int qkv) {
            // this thread's 64 columns as four 32-byte pieces of bf16 pairs; after the quad transpose the four lanes of
            // a quad write one 128-byte line (the half row of ONE row) per instruction
            __nv_bfloat16* base = (qkv == 0 ? p.sq : (qkv == 1 ? p.sk : p.sv)) + ((long long)tile * 128 + (e.row & ~3)) * D + e.half * 64;
            uint32_t w[4][8];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                uint32_t acc[32];
                tmem_ld32(e.tbase + qkv * 128 + e.half * 64 + b * 32, acc);
                tmem_ld_wait();
                if (qkv == 0) {                                // query bias W_q beta_1
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        acc[i] = __float_as_uint(__uint_as_float(acc[i]) + vec[V_BQ + e.half * 64 + b * 32 + i]);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        w[b * 2 + c][i] = pack_bf16(__uint_as_float(acc[c * 16 + 2 * i]), __uint_as_float(acc[c * 16 + 2 * i + 1]));
            }
            quad_transpose8(w, lane);
            const int nrows_t = tile_rows<KASF_MODE_LONG>(p, tile);
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if ((e.row & ~3) + r < nrows_t) stg256u(base + r * D + (lane & 3) * 16, w[r]);
        };
        if (KIND == KASF_KIND_BONE) {
            if (lt) {
                tc_fence_before();
                __syncthreads();                           // every thread has drained K, V of the previous tile
                if (tid == 0) {
                    mbar_wait(&bars[BL], ph_l);
                    tc_fence_after();
                    umma_tile_k128(tmem + TM_K, a_addr, ring, 128, false);
                    umma_tile_k128(tmem + TM_V, a_addr, ring + CHUNK_BYTES, 128, false);
                    tc_commit(&bars[BV]);                  // (one barrier for both: nobody waits for K alone)
                }
                ph_l ^= 1;
            } else {
                mbar_wait(&bars[BR], ph_r);
                ph_r ^= 1;
                read_staged(sm, e, xv, row_ok);
                ln_stats(sm, e, xv, mean, rstd);
                ln_write<false, false>(sm, SM_A0, e, xv, mean, rstd, nullptr, nullptr, row_ok);
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    umma_tile_k128(tmem + TM_K, a_addr, ring, 128, false);
                    umma_tile_k128(tmem + TM_V, a_addr, ring + CHUNK_BYTES, 128, false);
                    tc_commit(&bars[BV]);                  // (one barrier for both: nobody waits for K alone)
                }
                gather_rows<KASF_MODE_LONG>(p, sm, tile, p.in, &bars[BR], warp, lane, 0);
                gather_rows<KASF_MODE_LONG>(p, sm, tile, p.in, &bars[BR], warp, lane, 1);
            }
        }
        mbar_wait(&bars[BR], ph_r);
        ph_r ^= 1;
        read_staged(sm, e, xv, row_ok);
        ln_stats(sm, e, xv, mean, rstd);
        if (KIND == KASF_KIND_BONE) {
            mbar_wait(&bars[BV], ph_m);                    // K, V MMAs have read the limb operand
            tc_fence_after();
        }
        ln_write<false, false>(sm, SM_A0, e, xv, mean, rstd, nullptr, nullptr, row_ok);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            if (KIND == KASF_KIND_ATTENTION) {
                umma_tile_k128(tmem + TM_MIX, a_addr, ring, 128, false);
                tc_commit(&bars[BQ]);
                umma_tile_k128(tmem + TM_K, a_addr, ring + CHUNK_BYTES, 128, false);
                tc_commit(&bars[BK]);
                umma_tile_k128(tmem + TM_V, a_addr, ring + 2 * CHUNK_BYTES, 128, false);
                tc_commit(&bars[BV]);
            } else {
                umma_tile_k128(tmem + TM_MIX, a_addr, ring + 2 * CHUNK_BYTES, 128, false);
                tc_commit(&bars[BQ]);
            }
        }
        if (next < p.ntiles) {                              // staging is free: every thread has its rows in registers
            gather_rows<KASF_MODE_LONG>(p, sm, next, first_src, &bars[BR], warp, lane, 0);
            gather_rows<KASF_MODE_LONG>(p, sm, next, first_src, &bars[BR], warp, lane, 1);
        }
        if (KIND == KASF_KIND_BONE) {
            drain(1);
            drain(2);
            mbar_wait(&bars[BQ], ph_m);
            tc_fence_after();
            if (lt && tid == 0 && next < p.ntiles) {        // the A tile is free: next tile's limb operand
                mbar_arrive_expect_tx(&bars[BL], TILE_BYTES);
                bulk_g2s(sm + SM_A0, p.xlt + (size_t)next * TILE_BYTES, TILE_BYTES, &bars[BL]);
            }
            drain(0);
        } else {
            mbar_wait(&bars[BQ], ph_m);
            tc_fence_after();
            drain(0);
            mbar_wait(&bars[BK], ph_m);
            tc_fence_after();
            drain(1);
            mbar_wait(&bars[BV], ph_m);
            tc_fence_after();
            drain(2);
        }
        ph_m ^= 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- attention core of one (clip, joint) sequence, T <= 256: Q | K | V bf16 [256 x 128] resident in shared
//      memory (rows >= T zero), warp = head, 16-query blocks, two passes over the keys (row maximum, then
//      probabilities and P V) so that the rounding is the one of the fused kernel; the output replaces Q.
//      The tiles hold the sequence padded to whole 64-key blocks (la_rows), so sequences of up to 128 frames take
//      half the shared memory and two CTAs share an SM (MINB = 2).
__host__ __device__ constexpr uint32_t la_rows(int T) { return (uint32_t)((T + 63) & ~63); }
//      A CTA owns FOUR heads of a sequence (64 of the 128 columns; grid = 2 x sequences, 4 warps): the tiles are half
//      as large, so two CTAs (four for T <= 128) share an SM and the load / store phases of one -- 17 % of the
//      run time of the first, one-CTA-per-SM version -- run under the arithmetic of the others.
__device__ __forceinline__ uint32_t la_off(uint32_t r, uint32_t c16) { return r * 128u + ((c16 ^ (r & 7u)) << 4); }

template <int MINB, int NT16>
__global__ void __launch_bounds__(128, MINB) long_attention_kernel(const ModParams p) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = p.T;
    const long long seq = blockIdx.x >> 1;
    const int hg = blockIdx.x & 1;                              // head group: columns 64 hg .. 64 hg + 63
    __nv_bfloat16* gq = p.sq + seq * T * D + hg * 64;
    const __nv_bfloat16* gk = p.sk + seq * T * D + hg * 64;
    const __nv_bfloat16* gv = p.sv + seq * T * D + hg * 64;
    const uint32_t rows = la_rows(T), LA_TILE = rows * 128u;   // one [rows x 64 bf16] tile
    long long pt0 = p.prof ? clock64() : 0;                    // phase-cycle hook: thread 0's timeline, slots 8..10
    auto mark = [&](int k) {
        if (p.prof && tid == 0) {
            const long long t1 = clock64();
            atomicAdd(p.prof + k, (unsigned long long)(t1 - pt0));
            pt0 = t1;
        }
    };
    // ---- load (16-byte chunks, coalesced), zero the tail rows
    for (uint32_t i = tid; i < 3 * rows * 8; i += 128) {
        const uint32_t m = i / (rows * 8), r = (i >> 3) % rows, c = i & 7;
        uint8_t* dst = sm + m * LA_TILE + la_off(r, c);
        if (r < T) {
            const __nv_bfloat16* src = (m == 0 ? gq : (m == 1 ? gk : gv)) + (size_t)r * D + c * 8;
            cp_async16(dst, src, 16u);
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    mark(8);
    const uint32_t qb = smem_u32(sm), kb = smem_u32(sm + LA_TILE), vb = smem_u32(sm + 2 * LA_TILE);
    const int h = warp;
    const int g8 = lane >> 2, t4 = lane & 3, mi = lane >> 3, r8 = lane & 7;
    const float scale = 0.25f * 1.4426950408889634f;
    const int mtiles = (T + 15) >> 4, nkb = (T + 63) >> 6;        // 16-query blocks, 64-key blocks
    if constexpr (NT16 > 0) {
        // ---- single pass over NT16 >= T / 16 sixteen-key steps: ALL scores of a query block stay in registers (8 NT16
        //      per thread and block), so Q K^T is computed once -- the two-pass form below computes it twice, a third of
        //      the kernel's mma.sync work -- and only for the keys that exist (T = 81: 96 instead of 128).  Same
        //      arithmetic and rounding (global row maximum first).  With one CTA per SM this form was slower (all warps
        //      in their mma-only and MUFU-only phases at the same time); with two to four independent CTAs per SM the
        //      phases of different CTAs overlap.
        constexpr int U = NT16 <= 8 ? 2 : 1, NT = NT16 * 2;
#pragma unroll 1
        for (int mt0 = 0; mt0 < mtiles; mt0 += U) {
            uint32_t qa[U][4];
            float s[U][NT][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int mt = min(mt0 + u, mtiles - 1);             // (odd count: the last round repeats a block)
                ldsm_x4(qb + la_off(mt * 16 + (mi & 1) * 8 + r8, 2 * h + (mi >> 1)), qa[u]);
            }
#pragma unroll
            for (int nt = 0; nt < NT; nt += 2) {
                uint32_t kf[4];
                ldsm_x4(kb + la_off(8 * (nt + (mi >> 1)) + r8, 2 * h + (mi & 1)), kf);
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s[u][nt][i] = 0.f, s[u][nt + 1][i] = 0.f;
                    mma_bf16_16816(s[u][nt], qa[u], kf[0], kf[1]);
                    mma_bf16_16816(s[u][nt + 1], qa[u], kf[2], kf[3]);
                }
            }
            float nm0[U], nm1[U], l0[U], l1[U], o[U][2][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int nt = NT - 4; nt < NT; ++nt)                  // only the last 32 keys can lie past the end
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (nt * 8 + t4 * 2 + (i & 1) >= T) s[u][nt][i] = -INFINITY;
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mx0 = fmaxf(mx0, fmaxf(s[u][nt][0], s[u][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[u][nt][2], s[u][nt][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                nm0[u] = -mx0 * scale, nm1[u] = -mx1 * scale;
                l0[u] = 0.f, l1[u] = 0.f;
#pragma unroll
                for (int dn = 0; dn < 2; ++dn)
#pragma unroll
                    for (int i = 0; i < 4; ++i) o[u][dn][i] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < NT16; ++ks) {
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int nt = 2 * ks; nt < 2 * ks + 2; ++nt) {
                        s[u][nt][0] = ex2_approx(fmaf(s[u][nt][0], scale, nm0[u])), s[u][nt][1] = ex2_approx(fmaf(s[u][nt][1], scale, nm0[u]));
                        s[u][nt][2] = ex2_approx(fmaf(s[u][nt][2], scale, nm1[u])), s[u][nt][3] = ex2_approx(fmaf(s[u][nt][3], scale, nm1[u]));
                        l0[u] += s[u][nt][0] + s[u][nt][1];
                        l1[u] += s[u][nt][2] + s[u][nt][3];
                    }
                uint32_t vf[4];
                ldsm_x4_t(vb + la_off(16 * ks + (mi & 1) * 8 + r8, 2 * h + (mi >> 1)), vf);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int n0 = 2 * ks;
                    uint32_t pa[4];
                    pa[0] = pack_bf16(s[u][n0][0], s[u][n0][1]), pa[1] = pack_bf16(s[u][n0][2], s[u][n0][3]);
                    pa[2] = pack_bf16(s[u][n0 + 1][0], s[u][n0 + 1][1]), pa[3] = pack_bf16(s[u][n0 + 1][2], s[u][n0 + 1][3]);
                    mma_bf16_16816(o[u][0], pa, vf[0], vf[1]);
                    mma_bf16_16816(o[u][1], pa, vf[2], vf[3]);
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (mt0 + u < mtiles) {
                    l0[u] += __shfl_xor_sync(0xffffffffu, l0[u], 1), l1[u] += __shfl_xor_sync(0xffffffffu, l1[u], 1);
                    l0[u] += __shfl_xor_sync(0xffffffffu, l0[u], 2), l1[u] += __shfl_xor_sync(0xffffffffu, l1[u], 2);
                    const float i0 = rcp_approx(l0[u]), i1 = rcp_approx(l1[u]);
#pragma unroll
                    for (int dn = 0; dn < 2; ++dn) {
                        const int r0 = (mt0 + u) * 16 + g8, r1 = r0 + 8;
                        *reinterpret_cast<uint32_t*>(sm + la_off(r0, 2 * h + dn) + t4 * 4) = pack_bf16(o[u][dn][0] * i0, o[u][dn][1] * i0);
                        *reinterpret_cast<uint32_t*>(sm + la_off(r1, 2 * h + dn) + t4 * 4) = pack_bf16(o[u][dn][2] * i1, o[u][dn][3] * i1);
                    }
                }
            }
        }
    } else {
    // Two query blocks per round (U = 2): their ldmatrix -> mma -> ex2 -> mma chains are independent, which is the
    // only instruction-level parallelism a warp has here (one block at a time left both pipes mostly idle; keeping
    // all scores of a block in registers for a single pass was slower still: the eight warps then run their
    // mma-only and MUFU-only phases in lock-step).
    constexpr int U = 2;
#pragma unroll 1
    for (int mt0 = 0; mt0 < mtiles; mt0 += U) {
        uint32_t qa[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int mt = min(mt0 + u, mtiles - 1);             // (odd count: the last round repeats a block)
            ldsm_x4(qb + la_off(mt * 16 + (mi & 1) * 8 + r8, 2 * h + (mi >> 1)), qa[u]);
        }
        float mx0[U], mx1[U];
#pragma unroll
        for (int u = 0; u < U; ++u) mx0[u] = -INFINITY, mx1[u] = -INFINITY;
        auto scores = [&](int kblk, float (&s)[U][8][4]) {
#pragma unroll
            for (int nt = 0; nt < 8; nt += 2) {
                uint32_t kf[4];
                ldsm_x4(kb + la_off(kblk * 64 + 8 * (nt + (mi >> 1)) + r8, 2 * h + (mi & 1)), kf);
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) s[u][nt][i] = 0.f, s[u][nt + 1][i] = 0.f;
                    mma_bf16_16816(s[u][nt], qa[u], kf[0], kf[1]);
                    mma_bf16_16816(s[u][nt + 1], qa[u], kf[2], kf[3]);
                }
            }
            if (kblk * 64 + 64 > T) {                             // only the last key block holds keys past the end
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (kblk * 64 + nt * 8 + t4 * 2 + (i & 1) >= T) s[u][nt][i] = -INFINITY;
            }
        };
#pragma unroll 1
        for (int kblk = 0; kblk < nkb; ++kblk) {                 // pass 1: row maxima
            float s[U][8][4];
            scores(kblk, s);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    mx0[u] = fmaxf(mx0[u], fmaxf(s[u][nt][0], s[u][nt][1]));
                    mx1[u] = fmaxf(mx1[u], fmaxf(s[u][nt][2], s[u][nt][3]));
                }
        }
        float nm0[U], nm1[U], l0[U], l1[U], o[U][2][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            mx0[u] = fmaxf(mx0[u], __shfl_xor_sync(0xffffffffu, mx0[u], 1));
            mx1[u] = fmaxf(mx1[u], __shfl_xor_sync(0xffffffffu, mx1[u], 1));
            mx0[u] = fmaxf(mx0[u], __shfl_xor_sync(0xffffffffu, mx0[u], 2));
            mx1[u] = fmaxf(mx1[u], __shfl_xor_sync(0xffffffffu, mx1[u], 2));
            nm0[u] = -mx0[u] * scale, nm1[u] = -mx1[u] * scale;
            l0[u] = 0.f, l1[u] = 0.f;
#pragma unroll
            for (int dn = 0; dn < 2; ++dn)
#pragma unroll
                for (int i = 0; i < 4; ++i) o[u][dn][i] = 0.f;
        }
#pragma unroll 1
        for (int kblk = 0; kblk < nkb; ++kblk) {                 // pass 2: probabilities, P V
            float s[U][8][4];
            scores(kblk, s);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    s[u][nt][0] = ex2_approx(fmaf(s[u][nt][0], scale, nm0[u])), s[u][nt][1] = ex2_approx(fmaf(s[u][nt][1], scale, nm0[u]));
                    s[u][nt][2] = ex2_approx(fmaf(s[u][nt][2], scale, nm1[u])), s[u][nt][3] = ex2_approx(fmaf(s[u][nt][3], scale, nm1[u]));
                    l0[u] += s[u][nt][0] + s[u][nt][1];
                    l1[u] += s[u][nt][2] + s[u][nt][3];
                }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t vf[4];
                ldsm_x4_t(vb + la_off(kblk * 64 + 16 * ks + (mi & 1) * 8 + r8, 2 * h + (mi >> 1)), vf);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t pa[4];
                    pa[0] = pack_bf16(s[u][2 * ks][0], s[u][2 * ks][1]), pa[1] = pack_bf16(s[u][2 * ks][2], s[u][2 * ks][3]);
                    pa[2] = pack_bf16(s[u][2 * ks + 1][0], s[u][2 * ks + 1][1]), pa[3] = pack_bf16(s[u][2 * ks + 1][2], s[u][2 * ks + 1][3]);
                    mma_bf16_16816(o[u][0], pa, vf[0], vf[1]);
                    mma_bf16_16816(o[u][1], pa, vf[2], vf[3]);
                }
            }
        }
        // the output of this (head, query block) replaces the Q block this warp alone reads
        __syncwarp();
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (mt0 + u < mtiles) {
                l0[u] += __shfl_xor_sync(0xffffffffu, l0[u], 1), l1[u] += __shfl_xor_sync(0xffffffffu, l1[u], 1);
                l0[u] += __shfl_xor_sync(0xffffffffu, l0[u], 2), l1[u] += __shfl_xor_sync(0xffffffffu, l1[u], 2);
                const float i0 = rcp_approx(l0[u]), i1 = rcp_approx(l1[u]);
#pragma unroll
                for (int dn = 0; dn < 2; ++dn) {
                    const int r0 = (mt0 + u) * 16 + g8, r1 = r0 + 8;
                    *reinterpret_cast<uint32_t*>(sm + la_off(r0, 2 * h + dn) + t4 * 4) = pack_bf16(o[u][dn][0] * i0, o[u][dn][1] * i0);
                    *reinterpret_cast<uint32_t*>(sm + la_off(r1, 2 * h + dn) + t4 * 4) = pack_bf16(o[u][dn][2] * i1, o[u][dn][3] * i1);
                }
            }
        }
    }
    }
    __syncthreads();
    mark(9);
    for (int i = tid; i < T * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(gq + (size_t)r * D + c * 8) = *reinterpret_cast<const uint4*>(sm + la_off(r, c));
    }
    mark(10);
}

#include "kasf_long_gcn.cuh"

size_t module_scratch_bytes(int B, int T) {
    if (T <= KASF_SPLIT_T || B <= 0) return 0;
    const size_t rows = (size_t)B * J * T;
    return 3 * rows * D * 2 + ((rows * 4 + 255) / 256) * 256;
}

static int launch_long(ModParams p, int kind, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (p.T > 256) return KASF_ESHAPE;
    if (!scratch || scratch_bytes < module_scratch_bytes(p.B, p.T) || ((uintptr_t)scratch & 255) != 0) return KASF_ENOMEM;
    const size_t rows = (size_t)p.B * J * p.T;
    p.sq = static_cast<__nv_bfloat16*>(scratch);
    p.sk = p.sq + rows * D;
    p.sv = p.sk + rows * D;
    p.srow = reinterpret_cast<float*>(p.sv + rows * D);
    p.groups_per_tile = 0;
    p.ntiles = (int)((rows + 127) / 128);
    const int seqs = p.B * J;
    int rc;
    if (kind == KASF_KIND_GRAPH) {
        if ((rc = p.T <= 128 ? lgt::launch_gcn_tc<1>(p, seqs, st) : lgt::launch_gcn_tc<2>(p, seqs, st))) return rc;
        return launch_one<KASF_KIND_GRAPH, KASF_MODE_LONG, 0>(p, st);
    }
    const int pre_grid = p.ntiles < sm_count() ? p.ntiles : sm_count();
    if (kind == KASF_KIND_ATTENTION) {
        cudaFuncSetAttribute(long_pre_kernel<KASF_KIND_ATTENTION>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        long_pre_kernel<KASF_KIND_ATTENTION><<<pre_grid, 256, SM_TOTAL, st>>>(p);
    } else {
        cudaFuncSetAttribute(long_pre_kernel<KASF_KIND_BONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        long_pre_kernel<KASF_KIND_BONE><<<pre_grid, 256, SM_TOTAL, st>>>(p);
    }
    if ((rc = cuda_status())) return rc;
    const int la_bytes = 3 * (int)la_rows(p.T) * 128;
#define KASF_LA(MINB, NT16)                                                                                    \
    do {                                                                                                       \
        cudaFuncSetAttribute(long_attention_kernel<MINB, NT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, la_bytes); \
        long_attention_kernel<MINB, NT16><<<2 * seqs, 128, la_bytes, st>>>(p);                                \
    } while (0)
    const int nt16 = (p.T + 15) >> 4;                    // sixteen-key steps; instantiated: 5, 6, 7, 8, 10, 12, 14, 16
    if (nt16 <= 4) KASF_LA(4, 0);                        // (T <= 64 never takes the split path: two-pass fallback)
    else if (nt16 == 5) KASF_LA(3, 5);
    else if (nt16 == 6) KASF_LA(3, 6);
    else if (nt16 == 7) KASF_LA(3, 7);
    else if (nt16 == 8) KASF_LA(3, 8);
    else if (nt16 <= 10) KASF_LA(2, 10);
    else if (nt16 <= 12) KASF_LA(2, 12);
    else if (nt16 <= 14) KASF_LA(2, 14);
    else KASF_LA(2, 16);
#undef KASF_LA
    if ((rc = cuda_status())) return rc;
    return kind == KASF_KIND_ATTENTION ? launch_one<KASF_KIND_ATTENTION, KASF_MODE_LONG, 0>(p, st)
                                       : launch_one<KASF_KIND_BONE, KASF_MODE_LONG, 0>(p, st);
}

// ---------------------------------------------------------------------------------------------- limb tiles
// The limb stream XL is the same for all 26 layers, and the limb LayerNorm's affine is folded into the K|V weights
// (kasf_pack.cu), so the operand of every bone module's K|V projection is the normalised limb row itself.  This
// kernel normalises XL once per forward and writes it, rounded to bf16, as ready-made [128 x 128] operand-tile
// images (128-byte swizzle) in the tile order of the spatial (MODE 0) or temporal (MODE 1) modules: a bone module
// then fetches its tile's limb operand with one 32 KB bulk copy and never touches the fp32 limb rows.
// Warp per row, lane = 4 columns, exact two-pass fp32 statistics (eps 1e-5, bone_crossattention.py:47-51).
template <int MODE>
__global__ void __launch_bounds__(256) limb_tiles_kernel(const ModParams p, uint8_t* __restrict__ tiles) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        uint8_t* dst = tiles + (size_t)tile * TILE_BYTES;
        const int n = tile_rows<MODE>(p, tile);
        for (int r = warp; r < 128; r += 8) {
            uint2 pk = make_uint2(0u, 0u);
            if (r < n) {
                const long long tok = row_token<MODE>(p, tile, r);
                const float4 v = *reinterpret_cast<const float4*>(p.xl + tok * D + lane * 4);
                float s = (v.x + v.y) + (v.z + v.w);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s * (1.0f / D);
                const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
                float q = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                const float rstd = 1.0f / sqrtf(q * (1.0f / D) + 1e-5f);
                pk.x = pack_bf16(d0 * rstd, d1 * rstd);
                pk.y = pack_bf16(d2 * rstd, d3 * rstd);
            }
            *reinterpret_cast<uint2*>(dst + tile_off_bf16(r, lane * 4)) = pk;
        }
    }
}

static void mode_tiling(ModParams& p, int mode) {
    if (mode == KASF_MODE_SPATIAL) {
        p.groups_per_tile = 7;
        p.ntiles = (int)(((long long)p.B * p.T + 6) / 7);
    } else if (p.T > KASF_SPLIT_T) {                       // split path: 128 consecutive (sequence, frame) rows
        p.groups_per_tile = 0;
        p.ntiles = (int)(((long long)p.B * J * p.T + 127) / 128);
    } else {
        p.groups_per_tile = 128 / p.T;
        p.ntiles = (int)(((long long)p.B * J + p.groups_per_tile - 1) / p.groups_per_tile);
    }
}

size_t limb_tiles_bytes(int B, int T, int mode) {
    ModParams p;
    p.B = B, p.T = T;
    mode_tiling(p, mode);
    return (size_t)p.ntiles * TILE_BYTES;
}

int launch_limb_tiles(const float* XL, void* tiles, int B, int T, int mode, cudaStream_t st) {
    if (B <= 0 || limb_tiles_bytes(B, T, mode) == 0) return KASF_OK;
    if (!XL || !tiles || ((uintptr_t)tiles & 127) != 0 || ((uintptr_t)XL & 15) != 0) return KASF_EINVAL;
    ModParams p;
    memset(&p, 0, sizeof p);
    p.xl = XL;
    p.B = B, p.T = T;
    mode_tiling(p, mode);
    const int grid = p.ntiles < sm_count() * 8 ? p.ntiles : sm_count() * 8;
    if (mode == KASF_MODE_SPATIAL) limb_tiles_kernel<KASF_MODE_SPATIAL><<<grid, 256, 0, st>>>(p, static_cast<uint8_t*>(tiles));
    else if (T > KASF_SPLIT_T) limb_tiles_kernel<KASF_MODE_LONG><<<grid, 256, 0, st>>>(p, static_cast<uint8_t*>(tiles));
    else limb_tiles_kernel<KASF_MODE_TEMPORAL><<<grid, 256, 0, st>>>(p, static_cast<uint8_t*>(tiles));
    return cuda_status();
}

int launch_former_module(const uint8_t* blob, int layer, int kind, int mode, const float* in, const float* XL,
                         float* out, int B, int T, cudaStream_t st, unsigned long long* prof, void* scratch,
                         size_t scratch_bytes, const void* limb_tiles, unsigned flags) {
    if (B <= 0) return KASF_OK;
    if (kind < 0 || kind > 2 || mode < 0 || mode > 1) return KASF_EINVAL;
    if (kind == KASF_KIND_BONE && !XL) return KASF_EINVAL;
    if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)XL) & 31) != 0) return KASF_EINVAL;   // 256-bit row accesses
    if ((long long)B * T * J >= (1LL << 31)) return KASF_ESHAPE;                             // 32-bit token indices
    ModParams p;
    // module order in the blob: att_s, att_t, graph_s, graph_t, bone_s, bone_t
    p.mod = blob + module_off(layer, kind * 2 + mode);
    p.in = in;
    p.xl = XL;
    p.out = out;
    p.B = B;
    p.T = T;
    p.prof = prof;
    p.sq = p.sk = p.sv = nullptr;
    p.srow = nullptr;
    p.xlt = kind == KASF_KIND_BONE ? static_cast<const uint8_t*>(limb_tiles) : nullptr;
    if (((uintptr_t)p.xlt & 127) != 0) return KASF_EINVAL;
    if (mode == KASF_MODE_TEMPORAL && T > KASF_SPLIT_T)
        return (flags & KASF_FLAG_TWO_TILES) ? KASF_ESHAPE : launch_long(p, kind, scratch, scratch_bytes, st);
    int tc = 0;
    if (mode == KASF_MODE_SPATIAL) {
        p.groups_per_tile = 7;
        p.ntiles = (int)(((long long)B * T + 6) / 7);
    } else {
        p.groups_per_tile = 128 / T;
        p.ntiles = (int)(((long long)B * J + p.groups_per_tile - 1) / p.groups_per_tile);
        tc = T <= 32 ? 0 : (T <= 64 ? 1 : 2);
    }
    // Two tiles in flight per SM (kasf_module_v2.cuh) -- opt-in (KASF_FLAG_TWO_TILES): parity-green, but measured SLOWER
    // than this file's one-tile kernel on B200 (10.4k vs 13.3k clips/s at B = 1024, T = 27; DESIGN.md section 3 has the
    // phase cycles, the event trace and the ncu counters that explain why).  The stage tests cover both kernels.
    if (flags & KASF_FLAG_TWO_TILES) {
        if (tc != 0 || (kind == KASF_KIND_BONE && !p.xlt)) return KASF_ESHAPE;
        const int sms = sm_count();
#define KASF_CASE2(K, M) \
    if (kind == K && mode == M) return v2::launch_v2<K, M>(p, st, sms);
        KASF_CASE2(0, 0) KASF_CASE2(0, 1) KASF_CASE2(1, 0) KASF_CASE2(1, 1) KASF_CASE2(2, 0) KASF_CASE2(2, 1)
#undef KASF_CASE2
    }
#define KASF_CASE(K, M, C) \
    if (kind == K && mode == M && tc == C) return launch_one<K, M, C>(p, st);
    KASF_CASE(0, 0, 0) KASF_CASE(1, 0, 0) KASF_CASE(2, 0, 0)
    KASF_CASE(0, 1, 0) KASF_CASE(0, 1, 1) KASF_CASE(0, 1, 2)
    KASF_CASE(1, 1, 0) KASF_CASE(1, 1, 1) KASF_CASE(1, 1, 2)
    KASF_CASE(2, 1, 0) KASF_CASE(2, 1, 1) KASF_CASE(2, 1, 2)
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
