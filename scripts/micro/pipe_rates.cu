// Micro-benchmark: issue rates of the instructions that bound the FormerModule epilogues (B200, sm_100a).
// Prints cycles per warp-instruction per SM sub-partition for MUFU.TANH / MUFU.EX2 / tanh.f16x2 / F2FP and
// for the whole GELU sequence, at 1, 2 and 4 warps per sub-partition (one CTA on one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

constexpr int ILP = 8, ITERS = 2048;

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    float v[ILP];
    float aux = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) {
                unsigned u = __float_as_uint(v[i]);
                asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
                v[i] = __uint_as_float(u);
            }
            if (OP == 3) {
                unsigned u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[i]), "f"(v[(i + 1) % ILP]));
                v[i] = __uint_as_float(u | 0x3f000000u);
            }
            if (OP == 4) {   // the GELU sequence of kasf_module.cu (bias add, poly, tanh, fma)
                const float x = v[i] + seed;
                const float x2 = x * x;
                const float pl = fmaf(x2, fmaf(x2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
                float t;
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * pl));
                v[i] = fmaf(x, t, x);
            }
            if (OP == 5) {   // GELU with a pure FMA-pipe rational/polynomial tanh replacement (no MUFU): degree-7 odd poly of clamp
                const float x = v[i] + seed;
                const float x2 = x * x;
                const float pl = fmaf(x2, fmaf(x2, -0.0003828259195935171f, 0.03722352208203997f), 0.7972238404651819f);
                float w = fminf(fmaxf(x * pl, -3.f), 3.f);
                const float w2 = w * w;
                float t = fmaf(w2, fmaf(w2, fmaf(w2, -0.0021f, 0.0331f), -0.2443f), 0.9837f) * w;
                v[i] = fmaf(x, t, x);
            }
            if (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 8) {   // HFMA2 alone
                unsigned u = __float_as_uint(v[i]);
                asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u) : "r"(0x3c003c00u));
                v[i] = __uint_as_float(u);
            }
            if (OP == 9) {   // MUFU.TANH + 4 FFMA (independent)
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
                float a = seed, b = seed + 1.f, c = seed + 2.f, d = seed + 3.f;
                asm volatile("fma.rn.f32 %0, %0, %0, %0;\n\tfma.rn.f32 %1, %1, %1, %1;\n\tfma.rn.f32 %2, %2, %2, %2;\n\tfma.rn.f32 %3, %3, %3, %3;"
                             : "+f"(a), "+f"(b), "+f"(c), "+f"(d));
                aux += a + b + c + d;
            }
            if (OP == 10) {  // MUFU.TANH + 4 HFMA2 (independent)
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
                unsigned a = 0x3c003c00u + i, b = a + 1, c = a + 2, d = a + 3;
                asm volatile("fma.rn.f16x2 %0, %0, %0, %0;\n\tfma.rn.f16x2 %1, %1, %1, %1;\n\tfma.rn.f16x2 %2, %2, %2, %2;\n\tfma.rn.f16x2 %3, %3, %3, %3;"
                             : "+r"(a), "+r"(b), "+r"(c), "+r"(d));
                aux += __uint_as_float(a ^ b ^ c ^ d);
            }
            if (OP == 11) {  // MUFU.TANH + 1 F2FP (independent)
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
                unsigned u;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(seed + i), "f"(seed));
                aux += __uint_as_float(u);
            }
            if (OP == 12) {  // F2FP f16x2 alone
                unsigned u;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[i]), "f"(seed));
                v[i] = __uint_as_float(u | 0x3c000000u);
            }
            if (OP == 13) {  // MUFU.TANH.F16 x2 (tanh.approx.f16x2) + 4 HFMA2
                unsigned u = __float_as_uint(v[i]);
                asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
                v[i] = __uint_as_float(u);
                unsigned a = 0x3c003c00u + i, b = a + 1, c = a + 2, d = a + 3;
                asm volatile("fma.rn.f16x2 %0, %0, %0, %0;\n\tfma.rn.f16x2 %1, %1, %1, %1;\n\tfma.rn.f16x2 %2, %2, %2, %2;\n\tfma.rn.f16x2 %3, %3, %3, %3;"
                             : "+r"(a), "+r"(b), "+r"(c), "+r"(d));
                aux += __uint_as_float(a ^ b ^ c ^ d);
            }
            if (OP == 7) v[i] = fmaf(v[i], seed, 0.5f);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    s += aux;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_elem) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * 4);
    cudaMalloc(&cyc, 8);
    printf("%-28s", name);
    for (int wps : {1, 2, 4, 8}) {
        const int threads = wps * 4 * 32;
        k<OP><<<1, threads>>>(out, cyc, 0.37f);
        k<OP><<<1, threads>>>(out, cyc, 0.37f);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // warp-level sequences issued per sub-partition: wps * ITERS * ILP
        printf("  %dw/smsp: %6.2f cyc/seq", wps, (double)c / ((double)wps * ITERS * ILP));
    }
    printf("   (%d instr/seq)\n", per_elem);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("MUFU.TANH f32", 1);
    run<1>("MUFU.EX2 f32", 1);
    run<2>("tanh.approx.f16x2", 1);
    run<3>("F2FP bf16x2 pack (+LOP)", 2);
    run<6>("MUFU.RCP", 1);
    run<7>("FFMA", 1);
    run<8>("HFMA2", 1);
    run<12>("F2FP f16x2", 1);
    run<9>("MUFU.TANH + 4 FFMA", 5);
    run<10>("MUFU.TANH + 4 HFMA2", 5);
    run<11>("MUFU.TANH + F2FP", 2);
    run<13>("tanh.f16x2 + 4 HFMA2", 7);
    run<4>("GELU seq (tanh MUFU)", 7);
    run<5>("GELU seq (FMA-pipe poly)", 13);
    return 0;
}
