// Micro-experiments for the two-tiles-in-flight FormerModule kernel (round 2):
//   1. tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form): layout check.  A [128 x 64] 16-bit is written
//      thread-per-row with tcgen05.st (column c of lane r = elements (r, 2c) | (r, 2c+1) << 16), B [128 x 64] sits in
//      shared memory as one 128-byte-swizzled sub-tile, D [128 x 128] fp32.  A decode pass with one-hot A prints
//      which logical k a (column, half) position feeds, should the straightforward guess be wrong.
//   2. N = 64 pieces out of a [128 x 128] weight chunk image: rows 64..127 of both K sub-tiles (two 8 KB blocks).
//   3. a 640-thread CTA (2 x 8 compute warps + service warpgroup) with setmaxnreg 128 / 96 / 32 and 227 KB of
//      dynamic shared memory launches and runs.
//   4. issue rates: SS vs TS MMAs (M128 N128 K16), N = 64 SS.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kasportsformer_b200/csrc -o ts_mma scripts/micro/ts_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "kasf_ptx.cuh"

using namespace kasf;
// bounded wait: a protocol error ends the kernel instead of hanging the box
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 200000000LL) return false;
    return true;
}
#define mbar_wait(b, p) wait_bounded(b, p)

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

// ---------------------------------------------------------------- 1 + 2
// mode 0: TS, A from TMEM (a: [128][64] fp32 -> bf16), B sub-tile [128 n][64 k]; D[128][128]
// mode 1: one-hot decode: TMEM A column `c0`, half `hf` = 1.0 for every lane, B[n][k] = k; prints D[r][0]
// mode 2: SS with N = 64 piece: A tile [128][128] in smem, B = rows 64..127 of a full chunk image; D[128][64]
__global__ void __launch_bounds__(128, 1) k_ts(const float* a, const float* b, float* d, int mode, int c0, int hf) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 98304);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 98304 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16);
    const int r = tid;
    if (mode == 0 || mode == 1) {
        uint32_t av[32];
        for (int c = 0; c < 32; ++c) {
            if (mode == 0) av[c] = pack_bf16(a[r * 64 + 2 * c], a[r * 64 + 2 * c + 1]);
            else av[c] = (c == c0) ? (hf ? 0x3f800000u : 0x00003f80u) : 0u;
        }
        tmem_st32(tb + 256, av);
        tmem_st_wait();
        // B sub-tile: row n = tid, 64 k
        for (int k = 0; k < 64; k += 2) {
            const float v0 = mode == 0 ? b[r * 64 + k] : (float)k, v1 = mode == 0 ? b[r * 64 + k + 1] : (float)(k + 1);
            *reinterpret_cast<uint32_t*>(sm + tile_off_bf16(r, k)) = pack_bf16(v0, v1);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, 128);
            const uint64_t db = umma_desc_sw128(smem_u32(sm));
            for (int ks = 0; ks < 4; ++ks) umma_ts(tmem, tmem + 256 + ks * 8, db + (uint64_t)((ks * 32) >> 4), idesc, ks > 0);
            tc_commit(bar);
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        for (int q = 0; q < 4; ++q) {
            uint32_t acc[32];
            tmem_ld32(tb + q * 32, acc);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) d[r * 128 + q * 32 + i] = __uint_as_float(acc[i]);
        }
    } else {
        // A tile [128][128] bf16 at sm+0, full chunk image [128 n][128 k] at sm+32768; use rows 64..127 of the chunk
        for (int k = 0; k < 128; k += 2) {
            *reinterpret_cast<uint32_t*>(sm + tile_off_bf16(r, k)) = pack_bf16(a[r * 128 + k], a[r * 128 + k + 1]);
            *reinterpret_cast<uint32_t*>(sm + 32768 + tile_off_bf16(r, k)) = pack_bf16(b[r * 128 + k], b[r * 128 + k + 1]);
        }
        // repack the piece as the ring slot would hold it: [sub-tile 0 rows 64..127 : 8 KB][sub-tile 1 rows 64..127 : 8 KB]
        __syncthreads();
        for (int i = tid; i < 8192 / 16; i += 128) {
            reinterpret_cast<uint4*>(sm + 65536)[i] = reinterpret_cast<const uint4*>(sm + 32768 + 64 * 128)[i];
            reinterpret_cast<uint4*>(sm + 65536 + 8192)[i] = reinterpret_cast<const uint4*>(sm + 32768 + 16384 + 64 * 128)[i];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t idesc = umma_idesc_bf16(128, 64);
            const uint64_t da = umma_desc_sw128(smem_u32(sm)), db = umma_desc_sw128(smem_u32(sm + 65536));
            for (int ks = 0; ks < 8; ++ks) {
                const uint64_t ka = (uint64_t)(((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4);
                const uint64_t kb = (uint64_t)(((ks >> 2) * 8192u + (ks & 3) * 32u) >> 4);
                umma_bf16(tmem, da + ka, db + kb, idesc, ks > 0);
            }
            tc_commit(bar);
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        for (int q = 0; q < 2; ++q) {
            uint32_t acc[32];
            tmem_ld32(tb + q * 32, acc);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) d[r * 64 + q * 32 + i] = __uint_as_float(acc[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- 3
__global__ void __launch_bounds__(640, 1) k_regs(int* flags) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int warp = threadIdx.x >> 5;
    // the CTA's pool is 640 x 96 registers: what the service warpgroup releases (128 x (96 - 32)) is exactly what the
    // eight warps of the first group take (256 x (128 - 96)); the second group keeps its 96
    if (warp >= 16) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (threadIdx.x == 512) flags[2] = 1;
    } else if (warp >= 8) {
        if (threadIdx.x == 256) flags[1] = 1;
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        if (threadIdx.x == 0) flags[0] = 1;
    }
    sm[threadIdx.x] = 1;
}

// ---------------------------------------------------------------- 4
// `n` MMAs of K = 16 back to back.  N: tile width, ts: A from TMEM, nd: accumulator tiles used round-robin,
// run: consecutive MMAs on the same accumulator before moving on
template <int N, int ts, int nd, int run>
__global__ void __launch_bounds__(128, 1) k_rate(int n, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 65536 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_f16(128, N);
        const uint64_t da = umma_desc_sw128(smem_u32(sm)), db = umma_desc_sw128(smem_u32(sm + 32768));
        const long long t0 = clock64();
        for (int it = 0; it < n; it += 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint64_t ko = (uint64_t)(((i & 3) * 32u) >> 4);          // inside sub-tile 0: N = 256 fits
                const uint32_t dcol = (uint32_t)(((i / run) % nd) * N);
                if (ts) umma_ts(tmem + dcol, tmem + 448 + (i & 7) * 8, db + ko, idesc, 1);
                else umma_bf16(tmem + dcol, da + ko, db + ko, idesc, 1);
            }
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
        cyc[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf16r(float v) {
    uint32_t u;
    memcpy(&u, &v, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    u &= 0xffff0000u;
    memcpy(&v, &u, 4);
    return v;
}

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaFuncSetAttribute(k_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 128);
    std::vector<float> a(128 * 128), b(128 * 128), d(128 * 128);
    srand(1);
    for (auto& v : a) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto& v : b) v = (rand() % 2001 - 1000) / 4000.0f;
    float *da, *db, *dd;
    cudaMalloc(&da, a.size() * 4), cudaMalloc(&db, b.size() * 4), cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    // ---- 1: TS
    cudaMemset(dd, 0, d.size() * 4);
    k_ts<<<1, 128, 98304 + 128>>>(da, db, dd, 0, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 128; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) s += (double)bf16r(a[r * 64 + k]) * bf16r(b[n * 64 + k]);
            maxerr = fmax(maxerr, fabs(s - d[r * 128 + n]));
        }
    printf("[1] TS-MMA (A in TMEM, packed pairs per column): %s, max |err| = %.3g  => %s\n", cudaGetErrorString(e), maxerr,
           maxerr < 1e-4 ? "LAYOUT OK" : "MISMATCH");
    if (maxerr >= 1e-4) {
        for (int c0 = 0; c0 < 10; ++c0)
            for (int hf = 0; hf < 2; ++hf) {
                k_ts<<<1, 128, 98304 + 128>>>(da, db, dd, 1, c0, hf);
                cudaDeviceSynchronize();
                cudaMemcpy(d.data(), dd, 128 * 4 * 128, cudaMemcpyDeviceToHost);
                printf("    decode: column %d half %d -> k = %.1f (row 0), %.1f (row 37), %.1f (row 100)\n", c0, hf, d[0], d[37 * 128], d[100 * 128]);
            }
    }
    // ---- 2: N = 64 piece
    k_ts<<<1, 128, 98304 + 128>>>(da, db, dd, 2, 0, 0);
    e = cudaDeviceSynchronize();
    cudaMemcpy(d.data(), dd, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    maxerr = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 128; ++k) s += (double)bf16r(a[r * 128 + k]) * bf16r(b[(64 + n) * 128 + k]);
            maxerr = fmax(maxerr, fabs(s - d[r * 64 + n]));
        }
    printf("[2] N=64 piece (rows 64..127 of both K sub-tiles): %s, max |err| = %.3g => %s\n", cudaGetErrorString(e), maxerr,
           maxerr < 1e-4 ? "OK" : "MISMATCH");
    // ---- 3: register hand-over with 20 warps
    int* flags;
    cudaMalloc(&flags, 16);
    cudaMemset(flags, 0, 16);
    cudaFuncSetAttribute(k_regs, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    k_regs<<<148, 640, 232448>>>(flags);
    e = cudaDeviceSynchronize();
    int hf[4] = {0, 0, 0, 0};
    cudaMemcpy(hf, flags, 12, cudaMemcpyDeviceToHost);
    printf("[3] 640 threads, setmaxnreg 128/96/32, 232448 B smem: %s, flags %d %d %d\n", cudaGetErrorString(e), hf[0], hf[1], hf[2]);
    // ---- 4: rates
    long long* cyc;
    cudaMalloc(&cyc, 8 * 148);
#define RATE(N, TS, ND, RUN)                                                                                        \
    for (int grid : {1, 148}) {                                                                                    \
        const int n = 4096;                                                                                        \
        cudaFuncSetAttribute(k_rate<N, TS, ND, RUN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 128);    \
        k_rate<N, TS, ND, RUN><<<grid, 128, 65536 + 128>>>(n, cyc);                                                \
        k_rate<N, TS, ND, RUN><<<grid, 128, 65536 + 128>>>(n, cyc);                                                \
        e = cudaDeviceSynchronize();                                                                               \
        long long h[148];                                                                                          \
        cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);                                                      \
        long long mx = 0;                                                                                          \
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;                                                 \
        printf("[4] grid %3d %s N=%3d, %d accumulators, runs of %d: %s, %.1f cycles per MMA (K=16)\n", grid,       \
               TS ? "TS" : "SS", N, ND, RUN, cudaGetErrorString(e), (double)mx / n);                               \
    }
    RATE(128, 0, 1, 1) RATE(128, 0, 2, 1) RATE(128, 0, 2, 8) RATE(64, 0, 1, 1) RATE(64, 0, 2, 1) RATE(64, 0, 4, 1)
    RATE(64, 0, 2, 8) RATE(64, 0, 4, 8) RATE(256, 0, 1, 1) RATE(32, 0, 1, 1) RATE(32, 0, 4, 1) RATE(128, 1, 1, 1)
    RATE(128, 1, 2, 1) RATE(128, 1, 2, 4) RATE(64, 1, 2, 1) RATE(64, 1, 1, 1)
    return 0;
}
