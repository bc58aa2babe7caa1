// Micro-benchmark: how fast can every SM stream 32 KB weight chunks out of L2 with bulk copies (UBLKCP), as the
// FormerModule kernel's ring does?  All 148 CTAs walk the same chunk sequence (SAME=1) or disjoint ones (SAME=0),
// with R chunks in flight.  Also: a 64 KB row-tile gather by one warp of LDGSTS (16 B per lane, 512 B rows).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu && ./l2_stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int CHUNK = 32768;

__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* w, size_t nchunks_total, int iters, int R, int same, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 6 * CHUNK);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t base = same ? 0 : (size_t)blockIdx.x * 7;
        const long long t0 = clock64();
        for (int i = 0; i < iters + R; ++i) {
            if (i >= R) mbar_wait(&bars[(i - R) % R], ((i - R) / R) & 1);      // chunk i-R landed -> its slot is free again
            if (i < iters) {
                const size_t c = (base + i) % nchunks_total;
                mbar_expect(&bars[i % R], CHUNK);
                bulk_g2s(sm + (i % R) * CHUNK, w + c * CHUNK, CHUNK, &bars[i % R]);
            }
        }
        cyc[blockIdx.x] = clock64() - t0;
    }
}

__global__ void __launch_bounds__(128, 1) gather_kernel(const float* x, long long ntiles, int iters, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
    if (threadIdx.x == 0) {
        mbar_init(bar, 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) % ntiles;
            const float* g = x + tile * 128 * 128 + lane * 4;
            for (int r = 0; r < 128; ++r)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sm + r * 512 + ((lane ^ (r & 7)) << 4))), "l"(g + r * 128) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
            mbar_wait(bar, it & 1);
        }
        if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const size_t wbytes = 64ull << 20, nch = wbytes / CHUNK;
    uint8_t* w;
    float* x;
    long long* cyc;
    cudaMalloc(&w, wbytes);
    cudaMemset(w, 1, wbytes);
    const long long ntiles = 16384;                     // 1 GiB of rows: HBM-resident
    cudaMalloc(&x, (size_t)ntiles * 65536);
    cudaMemset(x, 0, (size_t)ntiles * 65536);
    cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * CHUNK + 64);
    cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    long long h[148];
    const int iters = 2000;
    for (int same = 1; same >= 0; --same)
        for (int R : {1, 2, 3, 4, 6}) {
            for (int rep = 0; rep < 2; ++rep) stream_kernel<<<148, 128, 6 * CHUNK + 64>>>(w, nch, iters, R, same, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("weights %s  R=%d : %7.0f cycles per 32 KB chunk  = %5.1f B/clk/SM  (%s)\n", same ? "same-seq" : "disjoint", R,
                   (double)mx / iters, (double)CHUNK * iters / mx, cudaGetErrorString(cudaGetLastError()));
        }
    for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) gather_kernel<<<grid, 128, 65536 + 64>>>(x, ntiles, 200, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("row gather (LDGSTS, 64 KB tile, HBM) grid=%3d : %7.0f cycles per tile = %5.1f B/clk/SM (%s)\n", grid, (double)mx / 200,
               65536.0 * 200 / mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
