// kasf_head.cu -- K6 adaptive fusion and K7 output head.  Both fp32 on CUDA cores.
//
// K6 replaces the fusion of reference model/KASportsFormer.py:279-282:
//     alpha = softmax(Linear_384->3(cat(att, graph, bone)));  x = sum_i alpha_i * branch_i
//   HBM-bound: 3 x 512 B read + 512 B written per token; one warp per token, float4 per lane.
//
// K7 replaces reference model/KASportsFormer.py:339-345:
//     y = Linear_512->3(tanh(Linear_128->512(LayerNorm(x))))
//   Precision-critical: tf32/bf16 here breaks the 1e-2 mm bar, so it is
//   plain fp32 FMA with fp32 LayerNorm statistics and an accurate tanhf.
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

int launch_head(const uint8_t* blob, const float* X, float* y, float* rep, long long tokens, cudaStream_t st) {
    if (tokens <= 0) return KASF_OK;
    cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM);
    const long long ntiles = (tokens + HT - 1) / HT;
    const int grid = (int)min(ntiles, (long long)sm_count() * 2);
    head_kernel<<<grid, 256, HEAD_SMEM, st>>>(reinterpret_cast<const float*>(blob), X, y, rep, tokens);
    return cuda_status();
}

}  // namespace kasf
