// Micro-benchmark of the MLP GELU epilogue inner loop (bias add, 2*GELU via tanh, bf16 pack, swizzled STS.128),
// without tensor memory: what do 8 (or 16) warps on one SM achieve per 128x128 chunk, and how does the source-level
// pipelining depth change it?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_epi gelu_epi.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t tile_off_bf16(uint32_t r, uint32_t k) {
    return (k >> 6) * 16384u + r * 128u + ((((k & 63u) >> 3) ^ (r & 7u)) << 4) + ((k & 7u) << 1);
}
__device__ __forceinline__ float tanh_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
    return t;
}
__device__ __forceinline__ float gelu2(float v) {
    const float v2 = v * v;
    const float pl = fmaf(v2, fmaf(v2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
    return fmaf(v, tanh_fast(v * pl), v);
}

// ---- packed-half variants (what kasf_module.cu runs since r01p) ----
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ __half2 arg_h2(__half2 v) {
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
// 2*GELU(v) = (v + |v|) - G(min(|v|, 4)),  G(a) = a erfc(a / sqrt 2) as a degree-9 polynomial in t = a/2 - 1:
// FMA pipe only, no cancellation for negative v
__device__ __forceinline__ __half2 gelu2_poly_h2(__half2 v) {
    const __half2 a0 = __habs2(v);
    const __half2 a = __hmin2(a0, __float2half2_rn(4.0f));
    const __half2 t = __hfma2(a, __float2half2_rn(0.5f), __float2half2_rn(-1.0f));
    __half2 g = __float2half2_rn(5.194650522e-02f);
    g = __hfma2(g, t, __float2half2_rn(4.162358846e-02f));
    g = __hfma2(g, t, __float2half2_rn(-2.858623454e-01f));
    g = __hfma2(g, t, __float2half2_rn(3.757820403e-02f));
    g = __hfma2(g, t, __float2half2_rn(5.746606458e-01f));
    g = __hfma2(g, t, __float2half2_rn(-6.049317139e-01f));
    g = __hfma2(g, t, __float2half2_rn(3.410460008e-04f));
    g = __hfma2(g, t, __float2half2_rn(4.350995496e-01f));
    g = __hfma2(g, t, __float2half2_rn(-3.409467094e-01f));
    g = __hfma2(g, t, __float2half2_rn(9.094493380e-02f));
    return __hsub2(__hadd2(v, a0), g);
}

// VAR 0: as in kasf_module.cu (8-element groups as the compiler schedules them)
// VAR 1: explicit two-stage software pipeline over groups of G elements: tanh arguments of group g+1 are
//        computed while the MUFU results of group g are in flight
template <int VAR, int G, int COLS>
__global__ void __launch_bounds__(512, 1) k(const float* in, const float* bias, uint32_t* sink, long long* cyc, int iters) {
    extern __shared__ __align__(1024) uint8_t sm[];
    float* vb = reinterpret_cast<float*>(sm + 65536);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) vb[i] = bias[i];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2;   // part: column block of COLS columns
    float acc[COLS];
#pragma unroll
    for (int i = 0; i < COLS; ++i) acc[i] = in[(threadIdx.x * COLS + i) & 4095];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int c = it & 3, buf = it & 1;
        const float* b1 = vb + c * 128 + part * COLS;
        uint8_t* dst = sm + buf * 32768;
        if (VAR == 0) {
#pragma unroll
            for (int c8 = 0; c8 < COLS / 8; ++c8) {
                const float4 ba = *reinterpret_cast<const float4*>(b1 + c8 * 8);
                const float4 bb = *reinterpret_cast<const float4*>(b1 + c8 * 8 + 4);
                uint4 pk;
                pk.x = pack_bf16(gelu2(acc[c8 * 8 + 0] + ba.x), gelu2(acc[c8 * 8 + 1] + ba.y));
                pk.y = pack_bf16(gelu2(acc[c8 * 8 + 2] + ba.z), gelu2(acc[c8 * 8 + 3] + ba.w));
                pk.z = pack_bf16(gelu2(acc[c8 * 8 + 4] + bb.x), gelu2(acc[c8 * 8 + 5] + bb.y));
                pk.w = pack_bf16(gelu2(acc[c8 * 8 + 6] + bb.z), gelu2(acc[c8 * 8 + 7] + bb.w));
                *reinterpret_cast<uint4*>(dst + tile_off_bf16(row, part * COLS + c8 * 8)) = pk;
            }
        } else if (VAR == 2 || VAR == 3) {
            // packed half: groups of 8 columns = 4 pairs; VAR 3 sends pair 3 of every group through the FMA-pipe polynomial
            const uint4* b1h = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(vb) + c * 128 + part * COLS);
            constexpr int NM = VAR == 3 ? 3 : 4;       // pairs per group on the MUFU path
            __half2 v[2][4], w[2][4];
            auto stage1 = [&](int g, int s) {
                const uint4 bh = b1h[g];
                v[s][0] = __hadd2(u2h(pack_f16(acc[g * 8 + 0], acc[g * 8 + 1])), u2h(bh.x));
                v[s][1] = __hadd2(u2h(pack_f16(acc[g * 8 + 2], acc[g * 8 + 3])), u2h(bh.y));
                v[s][2] = __hadd2(u2h(pack_f16(acc[g * 8 + 4], acc[g * 8 + 5])), u2h(bh.z));
                v[s][3] = __hadd2(u2h(pack_f16(acc[g * 8 + 6], acc[g * 8 + 7])), u2h(bh.w));
#pragma unroll
                for (int i = 0; i < NM; ++i) w[s][i] = arg_h2(v[s][i]);
            };
            stage1(0, 0);
#pragma unroll
            for (int g = 0; g < COLS / 8; ++g) {
                const int s = g & 1;
                __half2 t[4], r[4];
#pragma unroll
                for (int i = 0; i < NM; ++i) t[i] = tanh_h2(w[s][i]);
                if (VAR == 3) r[3] = gelu2_poly_h2(v[s][3]);
                if (g + 1 < COLS / 8) stage1(g + 1, s ^ 1);
#pragma unroll
                for (int i = 0; i < NM; ++i) r[i] = __hfma2(v[s][i], t[i], v[s][i]);
                uint4 pk;
                pk.x = h2u(r[0]), pk.y = h2u(r[1]), pk.z = h2u(r[2]), pk.w = h2u(r[3]);
                *reinterpret_cast<uint4*>(dst + tile_off_bf16(row, part * COLS + g * 8)) = pk;
            }
        } else {
            float v[2][G], w[2][G];
            auto stage1 = [&](int g, int s) {
#pragma unroll
                for (int q = 0; q < G / 4; ++q) {
                    const float4 b = *reinterpret_cast<const float4*>(b1 + g * G + q * 4);
                    v[s][q * 4 + 0] = acc[g * G + q * 4 + 0] + b.x, v[s][q * 4 + 1] = acc[g * G + q * 4 + 1] + b.y;
                    v[s][q * 4 + 2] = acc[g * G + q * 4 + 2] + b.z, v[s][q * 4 + 3] = acc[g * G + q * 4 + 3] + b.w;
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const float v2 = v[s][i] * v[s][i];
                    w[s][i] = v[s][i] * fmaf(v2, fmaf(v2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
                }
            };
            stage1(0, 0);
#pragma unroll
            for (int g = 0; g < COLS / G; ++g) {
                const int s = g & 1;
                float t[G];
#pragma unroll
                for (int i = 0; i < G; ++i) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t[i]) : "f"(w[s][i]));
                if (g + 1 < COLS / G) stage1(g + 1, s ^ 1);
#pragma unroll
                for (int q = 0; q < G / 8; ++q) {
                    uint4 pk;
                    pk.x = pack_bf16(fmaf(v[s][q * 8 + 0], t[q * 8 + 0], v[s][q * 8 + 0]), fmaf(v[s][q * 8 + 1], t[q * 8 + 1], v[s][q * 8 + 1]));
                    pk.y = pack_bf16(fmaf(v[s][q * 8 + 2], t[q * 8 + 2], v[s][q * 8 + 2]), fmaf(v[s][q * 8 + 3], t[q * 8 + 3], v[s][q * 8 + 3]));
                    pk.z = pack_bf16(fmaf(v[s][q * 8 + 4], t[q * 8 + 4], v[s][q * 8 + 4]), fmaf(v[s][q * 8 + 5], t[q * 8 + 5], v[s][q * 8 + 5]));
                    pk.w = pack_bf16(fmaf(v[s][q * 8 + 6], t[q * 8 + 6], v[s][q * 8 + 6]), fmaf(v[s][q * 8 + 7], t[q * 8 + 7], v[s][q * 8 + 7]));
                    *reinterpret_cast<uint4*>(dst + tile_off_bf16(row, part * COLS + g * G + q * 8)) = pk;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < COLS; ++i) acc[i] += 1e-3f;
    }
    const long long t1 = clock64();
    __syncthreads();
    sink[threadIdx.x] = reinterpret_cast<uint32_t*>(sm)[threadIdx.x * 7];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int VAR, int G, int COLS>
void run(const char* name) {
    float *in, *bias;
    uint32_t* sink;
    long long* cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&bias, 512 * 4);
    cudaMalloc(&sink, 4096);
    cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 4);
    cudaMemset(bias, 0, 512 * 4);
    const int threads = 128 * (128 / COLS), iters = 400;
    cudaFuncSetAttribute(k<VAR, G, COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048);
    k<VAR, G, COLS><<<1, threads, 65536 + 2048>>>(in, bias, sink, cyc, iters);
    k<VAR, G, COLS><<<1, threads, 65536 + 2048>>>(in, bias, sink, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %4d threads: %7.0f cycles per 128x128 chunk  (%s)\n", name, threads, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<0, 8, 64>("as compiled (8 warps, 64 cols/thread)");
    run<1, 8, 64>("pipelined G=8 (8 warps)");
    run<1, 16, 64>("pipelined G=16 (8 warps)");
    run<1, 32, 64>("pipelined G=32 (8 warps)");
    run<2, 8, 64>("packed half, MUFU tanh (8 warps)");
    run<3, 8, 64>("packed half, 3 of 4 pairs MUFU + 1 poly (8 warps)");
    run<0, 8, 32>("as compiled (16 warps, 32 cols/thread)");
    run<1, 16, 32>("pipelined G=16 (16 warps)");
    return 0;
}
