// Micro-benchmark: throughput of the warp-level (legacy) tensor-core MMAs on B200 (sm_100a), per SM sub-partition:
// mma.sync m16n8k8 tf32 (the 3xTF32 similarity) and m16n8k16 bf16 (the attention core), ILP independent
// accumulator chains per warp, at 1/2/4 warps per sub-partition (one CTA on one SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rates mma_rates.cu && ./mma_rates
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 1024;

template <int OP, int ILP>
__global__ void k(float* out, long long* cyc, uint32_t seed) {
    float d[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b0 = seed ^ threadIdx.x, b1 = seed + 7;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 1024 * 4);
    cudaMalloc(&cyc, 8);
    printf("%-34s ILP %d:", name, ILP);
    for (int wps : {1, 2, 4}) {
        const int threads = wps * 4 * 32;
        k<OP, ILP><<<1, threads>>>(out, cyc, 0x3f800000u);
        k<OP, ILP><<<1, threads>>>(out, cyc, 0x3f800000u);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %dw/smsp: %6.2f cyc/mma/smsp", wps, (double)c / ((double)wps * ITERS * ILP));
    }
    printf("\n");
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0, 1>("mma.sync m16n8k8 tf32");
    run<0, 4>("mma.sync m16n8k8 tf32");
    run<0, 8>("mma.sync m16n8k8 tf32");
    run<1, 1>("mma.sync m16n8k16 bf16");
    run<1, 4>("mma.sync m16n8k16 bf16");
    run<1, 8>("mma.sync m16n8k16 bf16");
    return 0;
}
