// Layout check for the tensor-core temporal GCN of the split path (long_gcn_tc_kernel):
//   A. similarity S = Z Z^T with ONE operand image: Z [256 rows x 128 cols] bf16 as two K-major, 128-byte-swizzled
//      column blocks of [256 x 64]; A = rows 0..127, B = all 256 rows (N = 256), K = 128 in eight steps.
//   B. aggregation O = P Z with the SAME image as an MN-major B operand (K = key row, N = column): descriptor
//      LBO = 32768 (next 64-column block), SBO = 1024 (next 8 keys), instruction descriptor bit 16 (B MN-major);
//      P [128 x 256] in tensor memory (".ts" form, 16-bit pairs, 8 columns per 16 keys).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kasportsformer_b200/csrc -o mn_major scripts/micro/mn_major.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "kasf_ptx.cuh"

using namespace kasf;
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 200000000LL) return false;
    return true;
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__host__ __device__ inline uint32_t z_off(uint32_t r, uint32_t c) {
    return (c >> 6) * 32768u + r * 128u + ((((c & 63u) >> 3) ^ (r & 7u)) << 4) + ((c & 7u) << 1);
}

__global__ void __launch_bounds__(128, 1) k_mn(const float* z, const float* pm, float* ds, float* dout, int lbo, int sbo) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 65536 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
    for (int i = tid; i < 256 * 128; i += 128) {
        const int r = i >> 7, c = i & 127;
        *reinterpret_cast<__nv_bfloat16*>(sm + z_off(r, c)) = __float2bfloat16(z[i]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t tb = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t zb = smem_u32(sm);
    // ---- A
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 256);
        for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * 32768u + (ks & 3) * 32u;
            umma_bf16(tmem, umma_desc_sw128(zb + off), umma_desc_sw128(zb + off), idesc, ks > 0);
        }
        tc_commit(bar);
    }
    if (!wait_bounded(bar, 0)) { if (tid == 0) printf("timeout A\n"); return; }
    tc_fence_after();
    for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(tb + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) ds[tid * 256 + c * 32 + i] = __uint_as_float(v[i]);
    }
    // ---- B: P -> TMEM columns 256..383
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) {
            const int k = (c * 32 + i) * 2;
            v[i] = pack_bf16(pm[tid * 256 + k], pm[tid * 256 + k + 1]);
        }
        tmem_st32(tb + 256 + c * 32, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, 128) | (1u << 16);
        for (int ks = 0; ks < 16; ++ks)
            umma_ts(tmem + 384, tmem + 256 + ks * 8, desc_mn_sw128(zb + ks * 2048u, lbo, sbo), idesc, ks > 0);
        tc_commit(bar);
    }
    if (!wait_bounded(bar, 1)) { if (tid == 0) printf("timeout B\n"); return; }
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tb + 384 + c * 32, v);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) dout[tid * 128 + c * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    setvbuf(stdout, nullptr, _IONBF, 0);
    std::vector<float> z(256 * 128), pm(128 * 256);
    srand(1);
    for (auto& v : z) v = bf((rand() % 2001 - 1000) / 500.0f);
    for (auto& v : pm) v = (rand() % 5 == 0) ? 1.0f : 0.0f;
    float *dz, *dp, *ds, *dout;
    cudaMalloc(&dz, z.size() * 4), cudaMalloc(&dp, pm.size() * 4), cudaMalloc(&ds, 128 * 256 * 4), cudaMalloc(&dout, 128 * 128 * 4);
    cudaMemcpy(dz, z.data(), z.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dp, pm.data(), pm.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 256);
    const int cand[3][2] = {{32768, 1024}, {1024, 32768}, {16, 1024}};
    for (int t = 0; t < 3; ++t) {
        cudaMemset(ds, 0, 128 * 256 * 4), cudaMemset(dout, 0, 128 * 128 * 4);
        k_mn<<<1, 128, 65536 + 256>>>(dz, dp, ds, dout, cand[t][0], cand[t][1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> s(128 * 256), o(128 * 128);
        cudaMemcpy(s.data(), ds, s.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        double es = 0, eo = 0;
        for (int r = 0; r < 128; ++r)
            for (int k = 0; k < 256; ++k) {
                double a = 0;
                for (int c = 0; c < 128; ++c) a += (double)z[r * 128 + c] * z[k * 128 + c];
                es = fmax(es, fabs(a - s[r * 256 + k]));
            }
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < 128; ++c) {
                double a = 0;
                for (int k = 0; k < 256; ++k) a += (double)pm[r * 256 + k] * z[k * 128 + c];
                eo = fmax(eo, fabs(a - o[r * 128 + c]));
            }
        printf("LBO %d SBO %d: similarity max err %.3g   aggregation (MN-major B, TS) max err %.3g\n", cand[t][0], cand[t][1], es, eo);
    }
    return 0;
}
