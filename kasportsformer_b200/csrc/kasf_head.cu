// kasf_head.cu -- K6 adaptive fusion and K7 output head.  Both fp32 on CUDA cores.
//
// K6 replaces the fusion of reference model/KASportsFormer.py:279-282:
//     alpha = softmax(Linear_384->3(cat(att, graph, bone)));  x = sum_i alpha_i * branch_i
//   HBM-bound: 3 x 512 B read + 512 B written per token; one warp per token, float4 per lane.
//
// K7 replaces reference model/KASportsFormer.py:339-345:
//     y = Linear_512->3(tanh(Linear_128->512(LayerNorm(x))))
//   Precision-critical: tf32/bf16 here breaks the 1e-2 mm bar.  Two kernels:
//   head_kernel (plain fp32 FMA, accurate tanhf: the exact-precision path) and head_tc_kernel (the 128 -> 512 projection
//   on tcgen05 at fp32 accuracy through a bf16 triple split of both operands: the fast path).
#include "kasf_internal.h"

namespace kasf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------ K6
__global__ void __launch_bounds__(256)
fusion_kernel(const float* __restrict__ fw /* W[3][384], b[3] */, const float* __restrict__ a,
              const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ out, long long tokens) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    float4 w[3][3];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int br = 0; br < 3; ++br) w[o][br] = *reinterpret_cast<const float4*>(fw + o * 384 + br * 128 + lane * 4);
    const float b0 = fw[1152], b1 = fw[1153], b2 = fw[1154];
#ifndef KASF_FUSION_ILP
#define KASF_FUSION_ILP 4
#endif
    // KASF_FUSION_ILP tokens per warp and round: their 3 x 16-byte loads per lane are all issued before the first
    // logit reduction (one token per round left the kernel at 78 % of the measured copy bandwidth)
    for (long long t0 = warp; t0 < tokens; t0 += nwarps * KASF_FUSION_ILP) {
        float4 va[KASF_FUSION_ILP], vg[KASF_FUSION_ILP], vb[KASF_FUSION_ILP];
#pragma unroll
        for (int u = 0; u < KASF_FUSION_ILP; ++u) {
            const long long t = t0 + (long long)u * nwarps;
            if (t < tokens) {
                va[u] = *reinterpret_cast<const float4*>(a + t * D + lane * 4);
                vg[u] = *reinterpret_cast<const float4*>(g + t * D + lane * 4);
                vb[u] = *reinterpret_cast<const float4*>(b + t * D + lane * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < KASF_FUSION_ILP; ++u) {
            const long long t = t0 + (long long)u * nwarps;
            if (t >= tokens) break;
            float l[3];
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                float s = va[u].x * w[o][0].x + va[u].y * w[o][0].y + va[u].z * w[o][0].z + va[u].w * w[o][0].w;
                s += vg[u].x * w[o][1].x + vg[u].y * w[o][1].y + vg[u].z * w[o][1].z + vg[u].w * w[o][1].w;
                s += vb[u].x * w[o][2].x + vb[u].y * w[o][2].y + vb[u].z * w[o][2].z + vb[u].w * w[o][2].w;
                l[o] = warp_sum(s);
            }
            l[0] += b0, l[1] += b1, l[2] += b2;
            const float m = fmaxf(l[0], fmaxf(l[1], l[2]));
            const float e0 = expf(l[0] - m), e1 = expf(l[1] - m), e2 = expf(l[2] - m);
            const float inv = 1.0f / (e0 + e1 + e2);
            const float a0 = e0 * inv, a1 = e1 * inv, a2 = e2 * inv;
            float4 r;
            r.x = va[u].x * a0 + vg[u].x * a1 + vb[u].x * a2;
            r.y = va[u].y * a0 + vg[u].y * a1 + vb[u].y * a2;
            r.z = va[u].z * a0 + vg[u].z * a1 + vb[u].z * a2;
            r.w = va[u].w * a0 + vg[u].w * a1 + vb[u].w * a2;
            *reinterpret_cast<float4*>(out + t * D + lane * 4) = r;
        }
    }
}

int launch_fusion(const uint8_t* blob, int layer, const float* a, const float* g, const float* b, float* out,
                  long long tokens, cudaStream_t st) {
    if (tokens <= 0) return KASF_OK;
    const float* fw = reinterpret_cast<const float*>(blob + fusion_off(layer));
    const int grid = (int)min((tokens + 7) / 8, (long long)sm_count() * 16);
    fusion_kernel<<<grid, 256, 0, st>>>(fw, a, g, b, out, tokens);
    return cuda_status();
}

// ------------------------------------------------------------------------------------ K7
constexpr int HT = 32;                          // tokens per tile
constexpr int HEAD_SMEM = HT * D * 4 + HT * REP * 4;

__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ gw, const float* __restrict__ X, float* __restrict__ y,
            float* __restrict__ rep_out, long long tokens) {
    extern __shared__ float4 hsm4[];
    float* s_z = reinterpret_cast<float*>(hsm4);          // [HT][128]
    float* s_rep = s_z + HT * D;                          // [HT][512]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* gamma = gw + G_NORM;
    const float* beta = gw + G_NORM + D;
    const float* Wt = gw + G_REPW;                        // [128 k][512 n]
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(beta + lane * 4);
    const float br0 = gw[G_REPB + tid], br1 = gw[G_REPB + tid + 256];

    const long long ntiles = (tokens + HT - 1) / HT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long t0 = tile * HT;
        __syncthreads();
        // ---- LayerNorm (eps 1e-5, biased variance), warp per token
        for (int r = warp; r < HT; r += 8) {
            const long long t = t0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < tokens) v = *reinterpret_cast<const float4*>(X + t * D + lane * 4);
            const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
            const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
            const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
            const float rstd = 1.0f / sqrtf(var + 1e-5f);
            float4 z;
            z.x = dx * rstd * g4.x + b4.x;
            z.y = dy * rstd * g4.y + b4.y;
            z.z = dz * rstd * g4.z + b4.z;
            z.w = dw * rstd * g4.w + b4.w;
            *reinterpret_cast<float4*>(s_z + r * D + lane * 4) = z;
        }
        __syncthreads();
        // ---- rep = tanh(z Wrep^T + b): thread owns columns tid and tid+256 for all HT tokens
        float acc0[HT], acc1[HT];
#pragma unroll
        for (int r = 0; r < HT; ++r) acc0[r] = br0, acc1[r] = br1;
        // weights of the next four k are requested before the current four are consumed (L2 latency off the FMA chain)
        float w0[4], w1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            w0[q] = Wt[q * REP + tid];
            w1[q] = Wt[q * REP + tid + 256];
        }
#pragma unroll 1
        for (int k = 0; k < D; k += 4) {
            float n0[4], n1[4];
            const int kn = k + 4 < D ? k + 4 : k;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                n0[q] = Wt[(kn + q) * REP + tid];
                n1[q] = Wt[(kn + q) * REP + tid + 256];
            }
#pragma unroll
            for (int r = 0; r < HT; ++r) {
                const float4 z = *reinterpret_cast<const float4*>(s_z + r * D + k);
                acc0[r] = fmaf(z.x, w0[0], acc0[r]);
                acc1[r] = fmaf(z.x, w1[0], acc1[r]);
                acc0[r] = fmaf(z.y, w0[1], acc0[r]);
                acc1[r] = fmaf(z.y, w1[1], acc1[r]);
                acc0[r] = fmaf(z.z, w0[2], acc0[r]);
                acc1[r] = fmaf(z.z, w1[2], acc1[r]);
                acc0[r] = fmaf(z.w, w0[3], acc0[r]);
                acc1[r] = fmaf(z.w, w1[3], acc1[r]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) w0[q] = n0[q], w1[q] = n1[q];
        }
#pragma unroll
        for (int r = 0; r < HT; ++r) {
            const float r0 = tanhf(acc0[r]), r1 = tanhf(acc1[r]);
            s_rep[r * REP + tid] = r0;
            s_rep[r * REP + tid + 256] = r1;
            if (rep_out && t0 + r < tokens) {
                rep_out[(t0 + r) * REP + tid] = r0;
                rep_out[(t0 + r) * REP + tid + 256] = r1;
            }
        }
        __syncthreads();
        // ---- y = rep Whead^T + b: 96 (token, out) dot products of length 512, warp-cooperative
        if (y) {
            for (int p = warp; p < HT * 3; p += 8) {
                const int r = p / 3, o = p % 3;
                const float* wh = gw + G_HEADW + o * REP;
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < REP / 32; ++q) s = fmaf(s_rep[r * REP + q * 32 + lane], wh[q * 32 + lane], s);
                s = warp_sum(s);
                if (lane == 0 && t0 + r < tokens) y[(t0 + r) * 3 + o] = s + gw[G_HEADB + o];
            }
        }
    }
}

// ------------------------------------------------------------------------------------ K7 on tensor cores
// The same head with the 128 -> 512 projection (97 % of its FLOPs) on tcgen05 at fp32 accuracy: both operands are split
// into bf16 TRIPLES (x = hi + mid + lo exactly: 3 x 8 mantissa bits) and the product is accumulated in fp32 from the six
// partial products whose weight is above 2^-24: hi*hi + hi*mid + mid*hi + hi*lo + lo*hi + mid*mid (the three dropped
// ones are below 2^-24 of the result).  The fp32 FMA kernel above spends 2.0 ms per 1,024 clips on this (49 % of the
// FMA pipe: it is FMA-bound, profiles/r02_ncu_small_kernels_summary.md); six bf16 MMAs per output cost as much tensor
// time as 3xTF32 and reuse the bf16 operand layout of the FormerModule kernels.
//   tile = 128 tokens; LayerNorm (fp32, warp per row) -> three A operand tiles; rep_logit's weight arrives as 8 pieces
//   of 64 output columns, three pre-split operand images each (kasf_pack.cu), through a 2-stage ring; D = 128 x 512 fp32
//   fills tensor memory; the epilogue of a piece (bias, tanh, optional store of the representation, the three 512 -> 3
//   dot products) runs while the tensor cores work on the next pieces.
// Warp roles: 8 compute warps | weight producer (1 lane) | MMA issuer (1 lane).
constexpr uint32_t HT_A = 0;                        // 3 x [128 x 128] bf16 (hi, mid, lo)
constexpr uint32_t HT_W = 3 * 32768;                // 2 stages x 48 KB
constexpr uint32_t HT_HW = HT_W + 2 * (uint32_t)REP3_PIECE;   // Whead fp32 [3][512]
constexpr uint32_t HT_BR = HT_HW + 3 * REP * 4;     // brep [512]
constexpr uint32_t HT_YP = HT_BR + REP * 4;         // partial y [128][2][4]
constexpr uint32_t HT_BARS = HT_YP + 128 * 2 * 4 * 4;
constexpr uint32_t HT_TOTAL = HT_BARS + 256;
enum { HB_FULL0 = 0, HB_EMPTY0 = 2, HB_AREADY = 4, HB_ADONE, HB_PDONE0, HB_COUNT = HB_PDONE0 + 8 };

// tanh to ~2e-7 absolute: 1 - 2 / (1 + e^2x) with ex2.approx (2^-22) and an approximate reciprocal
__device__ __forceinline__ float tanh_e2(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    return 1.0f - __fdividef(2.0f, 1.0f + e);
}

__global__ void __launch_bounds__(320, 1)
head_tc_kernel(const uint8_t* __restrict__ blob, const float* __restrict__ X, float* __restrict__ y,
               float* __restrict__ rep_out, long long tokens) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + HT_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + HT_BARS + HB_COUNT * 8);
    const float* gw = reinterpret_cast<const float*>(blob);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles = (tokens + 127) / 128;
    if (tid == 0) {
        if ((smem_u32(sm) & 1023u) != 0) __trap();
        for (int i = 0; i < HB_COUNT; ++i) mbar_init(&bars[i], i == HB_AREADY ? 8 : 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = tid; i < 3 * REP; i += 320) reinterpret_cast<float*>(sm + HT_HW)[i] = gw[G_HEADW + i];
    for (int i = tid; i < REP; i += 320) reinterpret_cast<float*>(sm + HT_BR)[i] = gw[G_REPB + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {          // ---- weight pieces: L2 -> ring (one 48 KB bulk copy each)
            uint32_t st = 0, ph = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int p = 0; p < 8; ++p) {
                    mbar_wait_suspend(&bars[HB_EMPTY0 + st], ph ^ 1);
                    mbar_arrive_expect_tx(&bars[HB_FULL0 + st], (uint32_t)REP3_PIECE);
                    bulk_g2s(sm + HT_W + st * REP3_PIECE, blob + G_REP3_OFF + (size_t)p * REP3_PIECE, (uint32_t)REP3_PIECE,
                             &bars[HB_FULL0 + st]);
                    if (++st == 2) st = 0, ph ^= 1;
                }
        }
    } else if (warp == 9) {
        if (lane == 0) {          // ---- the thread that issues tcgen05.mma
            const uint32_t a_addr = smem_u32(sm + HT_A), w_addr = smem_u32(sm + HT_W);
            const uint32_t idesc = umma_idesc_bf16(128, 64);
            uint32_t st = 0, ph = 0, ph_a = 0;
            // the six partial products, (A part, W part): hi hi, hi mid, mid hi, hi lo, lo hi, mid mid
            const int pa[6] = {0, 0, 1, 0, 2, 1}, pw[6] = {0, 1, 0, 2, 0, 1};
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(&bars[HB_AREADY], ph_a);
                ph_a ^= 1;
                tc_fence_after();
                for (int p = 0; p < 8; ++p) {
                    mbar_wait(&bars[HB_FULL0 + st], ph);
                    tc_fence_after();
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint64_t da = umma_desc_sw128(a_addr + pa[q] * 32768u);
                        const uint64_t db = umma_desc_sw128(w_addr + st * (uint32_t)REP3_PIECE + pw[q] * (uint32_t)REP3_IMG);
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks)
                            umma_bf16(tmem + p * 64, da + (uint64_t)(((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4),
                                      db + (uint64_t)(((ks >> 2) * 8192u + (ks & 3) * 32u) >> 4), idesc, (q > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(&bars[HB_EMPTY0 + st]);
                    tc_commit(&bars[HB_PDONE0 + p]);
                    if (++st == 2) st = 0, ph ^= 1;
                }
                tc_commit(&bars[HB_ADONE]);
            }
        }
    } else {
        // ---- compute warps: LayerNorm + operand split, then the epilogue of every piece
        const float4 g4 = *reinterpret_cast<const float4*>(gw + G_NORM + lane * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(gw + G_NORM + D + lane * 4);
        const float* s_hw = reinterpret_cast<const float*>(sm + HT_HW);
        const float* s_br = reinterpret_cast<const float*>(sm + HT_BR);
        float* s_yp = reinterpret_cast<float*>(sm + HT_YP);
        const int row = (warp & 3) * 32 + lane, half = warp >> 2;
        const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t ph_adone = 0, ph_p = 0;
        bool first = true;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long t0 = tile * 128;
            if (!first) {                                      // the previous tile's MMAs have read the A tiles
                mbar_wait(&bars[HB_ADONE], ph_adone);
                ph_adone ^= 1;
            }
            first = false;
            for (int r = warp; r < 128; r += 8) {
                const long long t = t0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t < tokens) v = *reinterpret_cast<const float4*>(X + t * D + lane * 4);
                const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
                const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
                const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
                const float rstd = 1.0f / sqrtf(var + 1e-5f);
                float z[4] = {dx * rstd * g4.x + b4.x, dy * rstd * g4.y + b4.y, dz * rstd * g4.z + b4.z, dw * rstd * g4.w + b4.w};
                if (t >= tokens) z[0] = z[1] = z[2] = z[3] = 0.f;
                // z = hi + mid + lo (bf16 each): three operand tiles
                uint2 part[3];
                float rem[4] = {z[0], z[1], z[2], z[3]};
#pragma unroll
                for (int s3 = 0; s3 < 3; ++s3) {
                    __nv_bfloat16 q[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        q[i] = __float2bfloat16_rn(rem[i]);
                        rem[i] -= __bfloat162float(q[i]);
                    }
                    part[s3].x = (uint32_t)__bfloat16_as_ushort(q[0]) | ((uint32_t)__bfloat16_as_ushort(q[1]) << 16);
                    part[s3].y = (uint32_t)__bfloat16_as_ushort(q[2]) | ((uint32_t)__bfloat16_as_ushort(q[3]) << 16);
                    *reinterpret_cast<uint2*>(sm + HT_A + s3 * 32768 + tile_off_bf16(r, lane * 4)) = part[s3];
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[HB_AREADY]);
            // ---- epilogue: this thread's 32 columns of every 64-column piece of its row
            float y0 = 0.f, y1 = 0.f, y2 = 0.f;
            const long long t = t0 + row;
            for (int p = 0; p < 8; ++p) {
                mbar_wait(&bars[HB_PDONE0 + p], ph_p);
                tc_fence_after();
                uint32_t acc[32];
                tmem_ld32(tbase + p * 64 + half * 32, acc);
                tmem_ld_wait();
                const int c0 = p * 64 + half * 32;
                float rv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    rv[i] = tanh_e2(__uint_as_float(acc[i]) + s_br[c0 + i]);
                    y0 = fmaf(rv[i], s_hw[c0 + i], y0);
                    y1 = fmaf(rv[i], s_hw[REP + c0 + i], y1);
                    y2 = fmaf(rv[i], s_hw[2 * REP + c0 + i], y2);
                }
                if (rep_out && t < tokens) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) stg256(rep_out + t * REP + c0 + i, rv + i);
                }
            }
            ph_p ^= 1;
            tc_fence_before();
            if (y) {
                float* yp = s_yp + (row * 2 + half) * 4;
                yp[0] = y0, yp[1] = y1, yp[2] = y2;
                asm volatile("bar.sync %0, 64;" ::"r"(2 + (warp & 3)) : "memory");   // the two warps that share these rows
                if (half == 0 && t < tokens) {
                    const float* o = s_yp + (row * 2 + 1) * 4;
                    y[t * 3 + 0] = y0 + o[0] + gw[G_HEADB + 0];
                    y[t * 3 + 1] = y1 + o[1] + gw[G_HEADB + 1];
                    y[t * 3 + 2] = y2 + o[2] + gw[G_HEADB + 2];
                }
                asm volatile("bar.sync %0, 64;" ::"r"(2 + (warp & 3)) : "memory");   // partials may be overwritten
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_head(const uint8_t* blob, const float* X, float* y, float* rep, long long tokens, cudaStream_t st, bool tensor_cores) {
    if (tokens <= 0) return KASF_OK;
    if (tensor_cores) {
        if ((((uintptr_t)X | (uintptr_t)rep) & 31) != 0) return KASF_EINVAL;
        cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_TOTAL);
        const long long nt = (tokens + 127) / 128;
        const int grid = (int)min(nt, (long long)sm_count());
        head_tc_kernel<<<grid, 320, HT_TOTAL, st>>>(blob, X, y, rep, tokens);
        return cuda_status();
    }
    cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM);
    const long long ntiles = (tokens + HT - 1) / HT;
    const int grid = (int)min(ntiles, (long long)sm_count() * 2);
    head_kernel<<<grid, 256, HEAD_SMEM, st>>>(reinterpret_cast<const float*>(blob), X, y, rep, tokens);
    return cuda_status();
}

}  // namespace kasf
