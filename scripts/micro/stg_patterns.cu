// Micro-benchmark: cost of the output-epilogue store patterns on one SM / all SMs.
//   A: thread = row, STG.256 (one 32 B sector per lane, 32 different lines per instruction)   [kernel today]
//   B: 4 lanes = one 128 B line (after a 4x4 block transpose among the lanes), STG.256, 8 lines per instruction
//   C: SHFL throughput (cost of that transpose)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

template <int PAT>
__global__ void __launch_bounds__(256, 1) k(float* out, long long ntiles, int iters, long long* cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) % ntiles;
        float* base = out + tile * 128 * 128;
        if (PAT == 0) {
            const int row = (warp & 3) * 32 + lane, half = warp >> 2;
#pragma unroll
            for (int c = 0; c < 8; ++c) stg256(base + row * 128 + half * 64 + c * 8, v);
        } else if (PAT == 1) {
            // lanes 4q..4q+3 cover one 128-byte line of row q per instruction
            const int q = lane >> 2, s = lane & 3, half = warp >> 2;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int row = (warp & 3) * 32 + (c >> 1) * 8 + q;     // 4 instr pairs cover 32 rows x 2 lines
                stg256(base + row * 128 + half * 64 + (c & 1) * 32 + s * 8, v);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 48; ++c) v[c & 7] = __shfl_xor_sync(0xffffffffu, v[c & 7], 1 + (c & 1));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (v[0] == -1.f) out[0] = v[1] + v[2] + v[3] + v[4] + v[5] + v[6] + v[7];
}

template <int PAT>
void run(const char* name, float* out, long long ntiles, long long* cyc) {
    long long h[148];
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) k<PAT><<<grid, 256>>>(out, ntiles, 400, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-52s grid=%3d : %7.0f cycles per 64 KB tile\n", name, grid, (double)mx / 400);
    }
}

int main() {
    const long long ntiles = 16384;
    float* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)ntiles * 65536);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("A thread=row, 32 lines x 1 sector per STG.256", out, ntiles, cyc);
    run<1>("B 4 lanes = 1 line, 8 lines x 4 sectors per STG.256", out, ntiles, cyc);
    run<2>("C 48 SHFL per thread (8 warps)", out, ntiles, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
