// kasf_internal.h -- declarations shared between the translation units of libkasf.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kasf_layout.h"
#include "kasf_ptx.cuh"
#include "kasf_tables.cuh"

// Temporal modules of sequences longer than this many frames take the split path (projection kernel, per-sequence
// mixer-core kernel, dense 128-row tail tiles) instead of the fused one-kernel path, which can only own whole
// sequences per 128-row tile.
#ifndef KASF_SPLIT_T
#define KASF_SPLIT_T 64
#endif

namespace kasf {

// CUDA launch status -> C-ABI code
inline int cuda_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? KASF_OK : -(1000 + (int)e);
}

// SMs of the current device (persistent kernels launch one CTA per SM)
inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        return 148;
    return n;
}

int pack_weights(const kasf_config* cfg, const float* image, void* packed, size_t cap, cudaStream_t st);

int launch_features(const uint8_t* blob, const float* x, float* bone, float* limb, float* X, float* XB, float* XL,
                    long long frames, cudaStream_t st);
// tensor_cores: head_tc_kernel (fp32-accurate bf16-triple MMAs); false: the fp32 FMA kernel (exact-precision path)
int launch_head(const uint8_t* blob, const float* X, float* y, float* rep, long long tokens, cudaStream_t st,
                bool tensor_cores = true);
int launch_fusion(const uint8_t* blob, int layer, const float* a, const float* g, const float* b, float* out,
                  long long tokens, cudaStream_t st);
// `scratch` (module_scratch_bytes(B, T) bytes, 256-byte aligned) is needed by temporal modules with T > KASF_SPLIT_T only
int launch_former_module(const uint8_t* blob, int layer, int kind, int mode, const float* in, const float* XL,
                         float* out, int B, int T, cudaStream_t st, unsigned long long* prof = nullptr,
                         void* scratch = nullptr, size_t scratch_bytes = 0, const void* limb_tiles = nullptr,
                         unsigned flags = 0);
// Pre-normalised limb rows as bf16 operand tiles in the tile order of `mode` (temporal, T > KASF_SPLIT_T: split-path row order): the
// optional `limb_tiles` argument of a bone module of the same mode.
size_t limb_tiles_bytes(int B, int T, int mode);
int launch_limb_tiles(const float* XL, void* tiles, int B, int T, int mode, cudaStream_t st);
size_t module_scratch_bytes(int B, int T);
int launch_metrics(int T, const float* pred, const float* pred_flip, const float* gt, const float* res,
                   const float* factor, const int32_t* action, int n_actions, double* sums, double* per_frame,
                   int B, cudaStream_t st);
// KASF_PRECISION_EXACT (kasf_exact.cu): the reference's fp32 arithmetic on CUDA cores, from the fp32 weight image
size_t exact_workspace_bytes(const kasf_config* cfg, int B);
int exact_forward(const kasf_config* cfg, const float* image, const uint8_t* blob, const float* x, float* y, float* rep,
                  int B, void* ws, size_t ws_bytes, cudaStream_t st);
double host_p_mpjpe(const double* p, const double* g);
int launch_flip(const float* in, float* out, long long frames, cudaStream_t st);
int launch_test_gemm(const float* a, const float* w, float* d, int M, int N, cudaStream_t st);

}  // namespace kasf
