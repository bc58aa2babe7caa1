// kasf_long_gcn.cuh -- temporal GCN adjacency + aggregation of the split path on the tcgen05 tensor cores
// (included by kasf_module.cu; graph.py:99-134 of the reference: z = LN1(x), similarity z z^T, the four most similar
// frames of every frame, A_hat = D^-1/2 A D^-1/2, A_hat z).
//
// One CTA per (clip, joint) sequence of T <= 128 MT frames, MT in {1, 2} M-tiles of 128 rows, 128 MT threads
// (thread = row for everything that reads tensor memory, warp = row for the LayerNorm):
//
//   1. z = LN1(x) in fp32 (two-pass statistics), split into THREE bf16 pieces z = h + m + l (exact: 3 x 8 mantissa
//      bits), each piece one operand image [128 MT rows][128 columns] (two K-major, 128-byte-swizzled column blocks).
//   2. similarity S = z z^T with fp32 accuracy on the tensor cores: the six piece products mm, hl, lh, hm, mh, hh
//      (what is dropped, ml + lm + ll, is below 2^-23 relative), A = rows of M-tile mt, B = ALL rows of the same
//      images (N = 128 MT), accumulated in fp32 in tensor memory: 48 MMAs per M-tile instead of the 3xTF32 mma.sync
//      loop of the first version (which was bound by operand loads and splitting: 1.62 ms per launch at T = 243).
//   3. thread = row: 4th-largest value of the row (with multiplicity, torch.topk semantics) by sorting networks over
//      the tcgen05.ld chunks, adjacency bits (>= threshold), degree, d = degree^-1/2.
//   4. A_hat z = D^-1/2 A (D^-1/2 z) again on the tensor cores: the 0/1 adjacency row is exact in bf16 and goes to
//      tensor memory as the A operand (".ts" form, it replaces the similarity row in place), y = d_j z_j is split
//      into two bf16 pieces that overwrite h and m, and the SAME images serve as the MN-major B operand (K = frame,
//      N = column; LBO = column-block stride, SBO = 1024: scripts/micro/mn_major.cu).  The result is scaled by d_i
//      and rounded to bf16 (it is the A operand of the V GEMM of the tail kernel), so two pieces (2^-17) are enough.
//   5. row sums of A_hat in the order of the fused kernel (ascending frame index).
#pragma once

namespace lgt {

template <int MT>
struct Lay {
    static constexpr int ROWS = 128 * MT;
    static constexpr uint32_t CB = ROWS * 128u;           // one column block [ROWS][64] bf16
    static constexpr uint32_t PIECE = 2 * CB;
    static constexpr uint32_t ZH = 0, ZM = PIECE, ZL = 2 * PIECE;
    static constexpr uint32_t RSD = 3 * PIECE;            // f32 [ROWS]
    static constexpr uint32_t BARS = RSD + ROWS * 4;      // S ready [2], aggregation ready [2], tmem slot
    static constexpr uint32_t TOTAL = BARS + 64;
    static constexpr uint32_t TM_COLS = 256 * MT;         // per M-tile 256 columns: S [0, ROWS) -> P [0, ROWS/2); O [128, 256)
};

template <int MT>
__device__ __forceinline__ uint32_t z_off(uint32_t r, uint32_t c) {
    return (c >> 6) * Lay<MT>::CB + r * 128u + ((((c & 63u) >> 3) ^ (r & 7u)) << 4) + ((c & 7u) << 1);
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;               // next 64-column block (MN direction)
    d |= (uint64_t)(1024 >> 4) << 32;              // next 8 frames (K direction)
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }

template <int MT>
__global__ void __launch_bounds__(128 * MT, MT == 1 ? 2 : 1) long_gcn_tc_kernel(const ModParams p) {
    using L = Lay<MT>;
    constexpr int ROWS = L::ROWS, NT = 128 * MT, NW = 4 * MT;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::BARS + 32);
    float* rsd = reinterpret_cast<float*>(sm + L::RSD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = p.T;
    const long long seq = blockIdx.x, b = seq / J;
    const int j = (int)(seq % J);
    const float* vecg = reinterpret_cast<const float*>(p.mod);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, L::TM_COLS);
        tmem_relinquish();
    }
    // ---- 1. z = LN1(x), warp per row (lane = 4 columns), four rows in flight; pieces h | m | l
    {
        const float4 gam = __ldg(reinterpret_cast<const float4*>(vecg + V_N1W) + lane);
        const float4 bet = __ldg(reinterpret_cast<const float4*>(vecg + V_N1B) + lane);
        const float* xseq = p.in + ((b * T) * J + j) * D + lane * 4;
#pragma unroll 1
        for (int r0 = warp * 4; r0 < ROWS; r0 += NW * 4) {
            float4 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u;
                x[u] = r < T ? __ldg(reinterpret_cast<const float4*>(xseq + (long long)r * J * D)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + u;
                float s = (x[u].x + x[u].y) + (x[u].z + x[u].w);
#pragma unroll
                for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s * (1.0f / D);
                const float d0 = x[u].x - mean, d1 = x[u].y - mean, d2 = x[u].z - mean, d3 = x[u].w - mean;
                float q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
#pragma unroll
                for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                const float rstd = 1.0f / sqrtf(q * (1.0f / D) + 1e-5f);
                float z[4] = {fmaf(d0 * rstd, gam.x, bet.x), fmaf(d1 * rstd, gam.y, bet.y), fmaf(d2 * rstd, gam.z, bet.z),
                              fmaf(d3 * rstd, gam.w, bet.w)};
                uint2 pc[3];
                if (r < T) {
                    float res[4];
                    __nv_bfloat16 hb[4];
#pragma unroll
                    for (int pi = 0; pi < 3; ++pi) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            hb[i] = __float2bfloat16_rn(pi == 0 ? z[i] : res[i]);
                            res[i] = (pi == 0 ? z[i] : res[i]) - __bfloat162float(hb[i]);
                        }
                        pc[pi].x = (uint32_t)__bfloat16_as_ushort(hb[0]) | ((uint32_t)__bfloat16_as_ushort(hb[1]) << 16);
                        pc[pi].y = (uint32_t)__bfloat16_as_ushort(hb[2]) | ((uint32_t)__bfloat16_as_ushort(hb[3]) << 16);
                    }
                } else {
                    pc[0] = pc[1] = pc[2] = make_uint2(0u, 0u);
                }
                const uint32_t off = z_off<MT>(r, lane * 4);
                *reinterpret_cast<uint2*>(sm + L::ZH + off) = pc[0];
                *reinterpret_cast<uint2*>(sm + L::ZM + off) = pc[1];
                *reinterpret_cast<uint2*>(sm + L::ZL + off) = pc[2];
            }
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t zb = smem_u32(sm);
    // ---- 2. S = z z^T: six piece products, smallest first
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, ROWS);
        const uint32_t pa[6] = {L::ZM, L::ZH, L::ZL, L::ZH, L::ZM, L::ZH};
        const uint32_t pb[6] = {L::ZM, L::ZL, L::ZH, L::ZM, L::ZH, L::ZH};
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 1
            for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t koff = (ks >> 2) * L::CB + (ks & 3) * 32u;
                    umma_bf16(tmem + mt * 256, umma_desc_sw128(zb + pa[pr] + koff + mt * 16384u), umma_desc_sw128(zb + pb[pr] + koff),
                              idesc, (pr | ks) ? 1u : 0u);
                }
            }
            tc_commit(&bars[mt]);
        }
    }
    // ---- 3. thread = row: threshold, adjacency bits, degree
    const int mt = warp >> 2, row = mt * 128 + (warp & 3) * 32 + lane;
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + mt * 256;
    mbar_wait(&bars[mt], 0);
    tc_fence_after();
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
    for (int c = 0; c < ROWS / 32; ++c) {
        if (c * 32 >= T) break;                                    // (warp-uniform)
        uint32_t v[32];
        tmem_ld32(tb + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float c4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) c4[i] = (c * 32 + g * 4 + i < T) ? __uint_as_float(v[g * 4 + i]) : -INFINITY;
            sort4_desc(c4);
            merge_top4(best, c4);
        }
    }
    const float thr = best[3];
    uint32_t bits[ROWS / 32];
    int deg = 0;
#pragma unroll
    for (int c = 0; c < ROWS / 32; ++c) {
        bits[c] = 0;
        if (c * 32 < T) {
            uint32_t v[32];
            tmem_ld32(tb + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c * 32 + i < T && __uint_as_float(v[i]) >= thr) bits[c] |= 1u << i;
            deg += __popc(bits[c]);
        }
    }
    const float di = 1.0f / sqrtf((float)deg);
    if (row < T) rsd[row] = di;
    // the adjacency row (0 / 1, exact in bf16) replaces the similarity row: column w = frames (2w, 2w + 1)
#pragma unroll
    for (int c4 = 0; c4 < ROWS / 64; ++c4) {
        uint32_t pw[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t two = (bits[c4 * 2 + (i >> 4)] >> (2 * (i & 15))) & 3u;
            pw[i] = row < T ? ((two & 1u) * 0x3f80u) | ((two >> 1) * 0x3f800000u) : 0u;
        }
        tmem_st32(tb + c4 * 32, pw);
    }
    tmem_st_wait();
    if (MT == 2) mbar_wait(&bars[MT - 1], 0);     // every similarity MMA has read the images before they are rescaled
    tc_fence_before();
    __syncthreads();
    // ---- 4. y = d_j z_j -> two bf16 pieces over h | m (warp per row, lane = 4 columns); row sums of A_hat
#pragma unroll 1
    for (int r = warp; r < T; r += NW) {
        const uint32_t off = z_off<MT>(r, lane * 4);
        const uint2 h = *reinterpret_cast<const uint2*>(sm + L::ZH + off);
        const uint2 m = *reinterpret_cast<const uint2*>(sm + L::ZM + off);
        const uint2 l = *reinterpret_cast<const uint2*>(sm + L::ZL + off);
        const float dj = rsd[r];
        float y[4] = {((bf16_lo(h.x) + bf16_lo(m.x)) + bf16_lo(l.x)) * dj, ((bf16_hi(h.x) + bf16_hi(m.x)) + bf16_hi(l.x)) * dj,
                      ((bf16_lo(h.y) + bf16_lo(m.y)) + bf16_lo(l.y)) * dj, ((bf16_hi(h.y) + bf16_hi(m.y)) + bf16_hi(l.y)) * dj};
        __nv_bfloat16 hb[4], mb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            hb[i] = __float2bfloat16_rn(y[i]);
            mb[i] = __float2bfloat16_rn(y[i] - __bfloat162float(hb[i]));
        }
        uint2 oh, om;
        oh.x = (uint32_t)__bfloat16_as_ushort(hb[0]) | ((uint32_t)__bfloat16_as_ushort(hb[1]) << 16);
        oh.y = (uint32_t)__bfloat16_as_ushort(hb[2]) | ((uint32_t)__bfloat16_as_ushort(hb[3]) << 16);
        om.x = (uint32_t)__bfloat16_as_ushort(mb[0]) | ((uint32_t)__bfloat16_as_ushort(mb[1]) << 16);
        om.y = (uint32_t)__bfloat16_as_ushort(mb[2]) | ((uint32_t)__bfloat16_as_ushort(mb[3]) << 16);
        *reinterpret_cast<uint2*>(sm + L::ZH + off) = oh;
        *reinterpret_cast<uint2*>(sm + L::ZM + off) = om;
    }
    float rs = 0.f;
    if (row < T) {
#pragma unroll
        for (int c = 0; c < ROWS / 32; ++c) {
            uint32_t bw = bits[c];
            while (bw) {
                const int jr = 32 * c + __ffs(bw) - 1;
                bw &= bw - 1;
                rs += di * rsd[jr];
            }
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = umma_idesc_bf16(128, 128) | (1u << 16);       // B operand MN-major
#pragma unroll 1
        for (int m2 = 0; m2 < MT; ++m2) {
#pragma unroll 1
            for (int pc = 0; pc < 2; ++pc) {
                const uint32_t img = zb + (pc == 0 ? L::ZM : L::ZH);
#pragma unroll 4
                for (int ks = 0; ks < ROWS / 16; ++ks)
                    umma_ts(tmem + m2 * 256 + 128, tmem + m2 * 256 + ks * 8, desc_mn_sw128(img + ks * 2048u, L::CB), idesc,
                            (pc | ks) ? 1u : 0u);
            }
            tc_commit(&bars[2 + m2]);
        }
    }
    // ---- 5. d_i (A y) -> bf16 scratch rows; row sums
    mbar_wait(&bars[2 + mt], 0);
    tc_fence_after();
    const long long R = seq * T + row;
    __nv_bfloat16* dst = p.sq + R * D;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tb + 128 + c * 32, v);
        tmem_ld_wait();
        if (row < T) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 pk;
                pk.x = pack_bf16(di * __uint_as_float(v[q * 8 + 0]), di * __uint_as_float(v[q * 8 + 1]));
                pk.y = pack_bf16(di * __uint_as_float(v[q * 8 + 2]), di * __uint_as_float(v[q * 8 + 3]));
                pk.z = pack_bf16(di * __uint_as_float(v[q * 8 + 4]), di * __uint_as_float(v[q * 8 + 5]));
                pk.w = pack_bf16(di * __uint_as_float(v[q * 8 + 6]), di * __uint_as_float(v[q * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = pk;
            }
        }
    }
    if (row < T) p.srow[R] = rs;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, L::TM_COLS);
}

template <int MT>
static int launch_gcn_tc(const ModParams& p, int seqs, cudaStream_t st) {
    cudaFuncSetAttribute(long_gcn_tc_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<MT>::TOTAL);
    long_gcn_tc_kernel<MT><<<seqs, 128 * MT, Lay<MT>::TOTAL, st>>>(p);
    return cuda_status();
}

}  // namespace lgt
