// kasf_features.cu -- K1: kinematic anatomy features + the three input embeddings, one fused kernel.
//
// Replaces bone_decomposer (reference model/KASportsFormer.py:42-62), BoneRefusion.forward
// (model/modules/bone_refusion.py:61-70 + bone_MLP.py:16-27, 51 tiny n->16->1 MLPs on the RAW joints)
// and the joint / bone / limb embeddings + positional embeddings (model/KASportsFormer.py:325-330).
//
// HBM-bound by construction: reads 204 B per frame, writes 3 x 17 x 512 B per frame.  All fp32.
// Skeleton tables live in constant memory; limb-MLP weights are staged once per CTA in shared memory;
// each thread owns one of the 128 output channels, so every store instruction of a warp is one full
// 128-byte line.
#include "kasf_internal.h"

namespace kasf {

__constant__ int c_bone_child[16] = KASF_BONE_CHILD;
__constant__ int c_bone_parent[16] = KASF_BONE_PARENT;
__constant__ int c_limb_size[17] = KASF_LIMB_SIZE;
__constant__ int c_limb_member[68] = KASF_LIMB_MEMBER;

constexpr int FR = 8;             // frames per block iteration
constexpr int FEAT_THREADS = 128;

__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(FEAT_THREADS)
features_kernel(const uint8_t* __restrict__ blob, const float* __restrict__ x, float* __restrict__ bone_out,
                float* __restrict__ limb_out, float* __restrict__ X, float* __restrict__ XB,
                float* __restrict__ XL, long long frames) {
    __shared__ float s_limbw[17 * 3 * G_LIMB_STRIDE];
    __shared__ float s_in[FR][J][3];
    __shared__ float s_bone[FR][J][3];
    __shared__ float s_limb[FR][J][3];
    const float* gw = reinterpret_cast<const float*>(blob);
    const int tid = threadIdx.x;
    for (int i = tid; i < 17 * 3 * G_LIMB_STRIDE; i += FEAT_THREADS) s_limbw[i] = gw[G_LIMB + i];

    // embedding weights of this thread's channel: [e][in] + (bias) and positional rows
    float w[3][3], bias[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
#pragma unroll
        for (int in = 0; in < 3; ++in) w[e][in] = gw[G_EMB_W + (e * 3 + in) * D + tid];
        bias[e] = gw[G_EMB_B + e * D + tid];
    }
    const float* pos = gw + G_POS;

    const long long nblk = (frames + FR - 1) / FR;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long f0 = blk * FR;
        const int nf = (int)min((long long)FR, frames - f0);
        __syncthreads();   // previous iteration done with s_*
        for (int i = tid; i < nf * 51; i += FEAT_THREADS) (&s_in[0][0][0])[i] = x[f0 * 51 + i];
        __syncthreads();
        // ---- bones: one (frame, bone) per thread
        {
            const int f = tid >> 4, k = tid & 15;
            if (f < nf) {
                const int a = c_bone_child[k], b = c_bone_parent[k];
                const float dx = s_in[f][a][0] - s_in[f][b][0];
                const float dy = s_in[f][a][1] - s_in[f][b][1];
                float len = sqrtf(dx * dx + dy * dy);
                if (len == 0.f) len = 1.f;                    // KASportsFormer.py:51
                s_bone[f][k][0] = dx / len;
                s_bone[f][k][1] = dy / len;
                s_bone[f][k][2] = len;
            }
        }
        // ---- limb MLPs: (frame, group, channel) items
        for (int it = tid; it < nf * 51; it += FEAT_THREADS) {
            const int f = it / 51, g = (it % 51) / 3, ch = it % 3;
            const float* lw = s_limbw + (g * 3 + ch) * G_LIMB_STRIDE;
            const int n = c_limb_size[g];
            float in[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int jm = c_limb_member[g * 4 + m];
                in[m] = (m < n) ? s_in[f][jm < 0 ? 0 : jm][ch] : 0.f;
            }
            float acc = lw[96];
#pragma unroll
            for (int h = 0; h < LIMB_HID; ++h) {
                float pre = lw[64 + h];
#pragma unroll
                for (int m = 0; m < 4; ++m) pre = fmaf(in[m], lw[h * 4 + m], pre);   // padded weights are 0
                acc = fmaf(gelu_exact(pre), lw[80 + h], acc);
            }
            s_limb[f][g][ch] = acc;
        }
        __syncthreads();
        // ---- row 16 of the bone features = mean over the 16 bones (KASportsFormer.py:54-58)
        if (tid < nf * 3) {
            const int f = tid / 3, ch = tid % 3;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) s += s_bone[f][k][ch];
            s_bone[f][16][ch] = s * (1.0f / 16.0f);
        }
        __syncthreads();
        if (bone_out)
            for (int i = tid; i < nf * 51; i += FEAT_THREADS) bone_out[f0 * 51 + i] = (&s_bone[0][0][0])[i];
        if (limb_out)
            for (int i = tid; i < nf * 51; i += FEAT_THREADS) limb_out[f0 * 51 + i] = (&s_limb[0][0][0])[i];
        // ---- embeddings: thread = channel; (frame, joint) rows are 512-byte contiguous lines
        for (int f = 0; f < nf; ++f) {
            const size_t base = (size_t)(f0 + f) * J * D + tid;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float* a = s_in[f][j];
                const float* b = s_bone[f][j];
                const float* l = s_limb[f][j];
                const float vx = fmaf(a[2], w[0][2], fmaf(a[1], w[0][1], fmaf(a[0], w[0][0], bias[0])));
                const float vb = fmaf(b[2], w[1][2], fmaf(b[1], w[1][1], fmaf(b[0], w[1][0], bias[1])));
                const float vl = fmaf(l[2], w[2][2], fmaf(l[1], w[2][1], fmaf(l[0], w[2][0], bias[2])));
                X[base + j * D] = vx + pos[(0 * J + j) * D + tid];
                XB[base + j * D] = vb + pos[(1 * J + j) * D + tid];
                XL[base + j * D] = vl + pos[(2 * J + j) * D + tid];
            }
        }
    }
}

int launch_features(const uint8_t* blob, const float* x, float* bone, float* limb, float* X, float* XB, float* XL,
                    long long frames, cudaStream_t st) {
    if (frames <= 0) return KASF_OK;
    const long long nblk = (frames + FR - 1) / FR;
    const int grid = (int)min(nblk, (long long)sm_count() * 8);
    features_kernel<<<grid, FEAT_THREADS, 0, st>>>(blob, x, bone, limb, X, XB, XL, frames);
    return cuda_status();
}

}  // namespace kasf
