// kasf_long_gcn.cuh -- temporal GCN adjacency + aggregation of the split path on the tcgen05 tensor cores
// (included by kasf_module.cu; graph.py:99-134 of the reference: z = LN1(x), similarity z z^T, the four most similar
// frames of every frame, A_hat = D^-1/2 A D^-1/2, A_hat z).
//
// One CTA per (clip, joint) sequence of T <= 128 MT frames, MT in {1, 2} M-tiles of 128 rows, 128 MT row threads
// (thread = row for everything that reads tensor memory, warp = row for the LayerNorm) + one MMA issuer warp:
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
    static constexpr uint32_t ADJ = RSD + ROWS * 4;       // u32 [ROWS / 32][ROWS]: adjacency bits, word-major (conflict-free)
    static constexpr uint32_t BARS = ADJ + ROWS * ROWS / 8;   // S ready [2], aggregation ready [2], images ready, rescaled ready; tmem slot
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
__global__ void __launch_bounds__(128 * MT + 32, MT == 1 ? 2 : 1) long_gcn_tc_kernel(const ModParams p) {
    using L = Lay<MT>;
    constexpr int ROWS = L::ROWS, NW = 4 * MT, RPW = ROWS / NW, NCH = ROWS / 32;
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L::BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + L::BARS + 56);
    float* rsd = reinterpret_cast<float*>(sm + L::RSD);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, T = p.T;
    const long long seq = blockIdx.x, b = seq / J;
    const int j = (int)(seq % J);
    const float* vecg = reinterpret_cast<const float*>(p.mod);
    // phase-cycle hook (kasf_former_module_profiled): thread 0's timeline, slots 0..6
    long long pt0 = p.prof ? clock64() : 0;
    auto mark = [&](int k) {
        if (p.prof && tid == 0) {
            const long long t1 = clock64();
            atomicAdd(p.prof + k, (unsigned long long)(t1 - pt0));
            pt0 = t1;
        }
    };
    // The issuer warp is driven by mbarriers only (images ready / rescaled images ready, one arrival per row warp);
    // the row warps meet each other at a named barrier of their own.
    auto row_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(128 * MT) : "memory"); };
    enum { B_S = 0, B_AGG = 2, B_IMG = 4, B_RESC = 5 };
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        mbar_init(&bars[B_IMG], NW);
        mbar_init(&bars[B_RESC], NW);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, L::TM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    mark(11);
    const uint32_t zb = smem_u32(sm);
    // =============================================================================== MMA issuer (one extra warp)
    // A thread that issues tcgen05.mma blocks once the MMA queue is full, i.e. for most of the run time of the 96
    // similarity MMAs; with the issue in a warp of its own the row warps of M-tile 0 take their thresholds while the
    // MMAs of M-tile 1 run.
    if (warp == NW) {
        if (lane == 0) {
            mbar_wait(&bars[B_IMG], 0);                    // operand images complete
            tc_fence_after();
            {
                const uint32_t idesc = umma_idesc_bf16(128, ROWS);
                const uint32_t pa[6] = {L::ZM, L::ZH, L::ZL, L::ZH, L::ZM, L::ZH};      // six piece products, smallest first
                const uint32_t pb[6] = {L::ZM, L::ZL, L::ZH, L::ZM, L::ZH, L::ZH};
#pragma unroll 1
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll 1
                    for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t koff = (ks >> 2) * L::CB + (ks & 3) * 32u;
                            umma_bf16(tmem + mt * 256, umma_desc_sw128(zb + pa[pr] + koff + mt * 16384u),
                                      umma_desc_sw128(zb + pb[pr] + koff), idesc, (pr | ks) ? 1u : 0u);
                        }
                    }
                    tc_commit(&bars[B_S + mt]);
                }
            }
            mbar_wait(&bars[B_RESC], 0);                   // adjacency rows in tensor memory, rescaled images complete
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
                tc_commit(&bars[B_AGG + m2]);
            }
        }
        return;
    }
    // ================================================================================================ row warps
    // ---- 1. z = LN1(x), warp per row (lane = 4 columns); every row of the warp is requested before the first is
    //         used (4 rows in flight per warp left this phase latency-bound: 32k of 75k cycles per sequence)
    {
        const float4 gam = __ldg(reinterpret_cast<const float4*>(vecg + V_N1W) + lane);
        const float4 bet = __ldg(reinterpret_cast<const float4*>(vecg + V_N1B) + lane);
        const float* xseq = p.in + ((b * T) * J + j) * D + lane * 4;
        // eight rows per round, their shuffle reductions interleaved by hand (the compiler keeps the shuffles of
        // different rows in program order, so row-at-a-time code runs one 500-cycle dependency chain after another);
        // the loads run two rounds ahead.  The round loop stays ROLLED: this kernel runs every instruction once per
        // CTA, and fully unrolled it spent as much time waiting for instruction fetches as for memory (ncu: stall
        // "no instruction" 1.7 per issue at 4096+ SASS instructions).
        constexpr int RB = 8;
        float4 xa[RB], xb[RB], xc[RB];
        auto load_round = [&](float4 (&x)[RB], int u0) {
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = warp + (u0 + u) * NW;
                x[u] = r < T ? __ldg(reinterpret_cast<const float4*>(xseq + (long long)r * J * D)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        load_round(xa, 0);
        load_round(xb, RB);
#pragma unroll 1
        for (int u0 = 0; u0 < RPW; u0 += RB) {
            load_round(xc, u0 + 2 * RB);                  // (rows >= T: zeros, no access)
            float s[RB], q[RB], mean[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) s[u] = (xa[u].x + xa[u].y) + (xa[u].z + xa[u].w);
#pragma unroll
            for (int o = 16; o; o >>= 1)
#pragma unroll
                for (int u = 0; u < RB; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                mean[u] = s[u] * (1.0f / D);
                const float d0 = xa[u].x - mean[u], d1 = xa[u].y - mean[u], d2 = xa[u].z - mean[u], d3 = xa[u].w - mean[u];
                q[u] = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1)
#pragma unroll
                for (int u = 0; u < RB; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], o);
            // branch-free tail (rsqrt + one Newton step instead of the IEEE sqrt / divide with their slow-path
            // branches, selects instead of `if (r < T)`): the eight rows' conversion chains interleave
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int r = warp + (u0 + u) * NW;
                const float a = q[u] * (1.0f / D) + 1e-5f;
                float rstd = rsqrtf(a);
                rstd = rstd * fmaf(-0.5f * a * rstd, rstd, 1.5f);
                const bool ok = r < T;
                const float z0 = ok ? fmaf((xa[u].x - mean[u]) * rstd, gam.x, bet.x) : 0.f, z1 = ok ? fmaf((xa[u].y - mean[u]) * rstd, gam.y, bet.y) : 0.f;
                const float z2 = ok ? fmaf((xa[u].z - mean[u]) * rstd, gam.z, bet.z) : 0.f, z3 = ok ? fmaf((xa[u].w - mean[u]) * rstd, gam.w, bet.w) : 0.f;
                // z = h + m + l exactly (packed conversions: F2FP converts two values per instruction)
                uint2 h, m, l;
                h.x = pack_bf16(z0, z1), h.y = pack_bf16(z2, z3);
                const float a0 = z0 - bf16_lo(h.x), a1 = z1 - bf16_hi(h.x), a2 = z2 - bf16_lo(h.y), a3 = z3 - bf16_hi(h.y);
                m.x = pack_bf16(a0, a1), m.y = pack_bf16(a2, a3);
                l.x = pack_bf16(a0 - bf16_lo(m.x), a1 - bf16_hi(m.x)), l.y = pack_bf16(a2 - bf16_lo(m.y), a3 - bf16_hi(m.y));
                const uint32_t off = z_off<MT>(r, lane * 4);
                *reinterpret_cast<uint2*>(sm + L::ZH + off) = h;
                *reinterpret_cast<uint2*>(sm + L::ZM + off) = m;
                *reinterpret_cast<uint2*>(sm + L::ZL + off) = l;
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) xa[u] = xb[u], xb[u] = xc[u];
        }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[B_IMG]);
    mark(0);
    // ---- 3. thread = row: threshold, adjacency bits, degree (tcgen05.ld of chunk c + 1 in flight under chunk c)
    const int mt = warp >> 2, row = mt * 128 + (warp & 3) * 32 + lane;
    const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + mt * 256;
    const int nch = (T + 31) >> 5;                         // chunks of 32 frames that hold valid columns
    mbar_wait(&bars[B_S + mt], 0);
    tc_fence_after();
    mark(1);
    // (rolled loops, two chunks per trip with the second tcgen05.ld in flight under the first chunk's network; the
    //  chunks that lie entirely inside the sequence need no column mask, the last, partial one is handled apart)
    uint32_t* adj = reinterpret_cast<uint32_t*>(sm + L::ADJ);
    const int nfull = T >> 5, tail = T & 31;
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    auto top4_chunk = [&](const uint32_t (&v)[32], int valid) {       // valid: 32, or the tail length (masked)
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float c4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) c4[i] = (g * 4 + i < valid) ? __uint_as_float(v[g * 4 + i]) : -INFINITY;
            sort4_desc(c4);
            merge_top4(best, c4);
        }
    };
    {
        uint32_t va[32], vb[32];
#pragma unroll 1
        for (int c = 0; c < nfull; c += 2) {
            tmem_ld32(tb + c * 32, va);
            if (c + 1 < nfull) tmem_ld32(tb + (c + 1) * 32, vb);
            tmem_ld_wait32(va);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float c4[4] = {__uint_as_float(va[g * 4]), __uint_as_float(va[g * 4 + 1]), __uint_as_float(va[g * 4 + 2]), __uint_as_float(va[g * 4 + 3])};
                sort4_desc(c4);
                merge_top4(best, c4);
            }
            if (c + 1 < nfull) {
                tmem_ld_wait32(vb);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    float c4[4] = {__uint_as_float(vb[g * 4]), __uint_as_float(vb[g * 4 + 1]), __uint_as_float(vb[g * 4 + 2]), __uint_as_float(vb[g * 4 + 3])};
                    sort4_desc(c4);
                    merge_top4(best, c4);
                }
            }
        }
        if (tail) {
            tmem_ld32(tb + nfull * 32, va);
            tmem_ld_wait32(va);
            top4_chunk(va, tail);
        }
    }
    const float thr = best[3];
    int deg = 0;
    {
        uint32_t va[32], vb[32];
        auto bits_of = [&](const uint32_t (&v)[32], int valid) {
            uint32_t bw = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < valid && __uint_as_float(v[i]) >= thr) bw |= 1u << i;
            return bw;
        };
#pragma unroll 1
        for (int c = 0; c < nfull; c += 2) {
            tmem_ld32(tb + c * 32, va);
            if (c + 1 < nfull) tmem_ld32(tb + (c + 1) * 32, vb);
            tmem_ld_wait32(va);
            uint32_t bw = bits_of(va, 32);
            deg += __popc(bw);
            adj[c * ROWS + row] = bw;
            if (c + 1 < nfull) {
                tmem_ld_wait32(vb);
                bw = bits_of(vb, 32);
                deg += __popc(bw);
                adj[(c + 1) * ROWS + row] = bw;
            }
        }
        if (tail) {
            tmem_ld32(tb + nfull * 32, va);
            tmem_ld_wait32(va);
            const uint32_t bw = bits_of(va, tail);
            deg += __popc(bw);
            adj[nfull * ROWS + row] = bw;
        }
#pragma unroll 1
        for (int c = nfull + (tail ? 1 : 0); c < NCH; ++c) adj[c * ROWS + row] = 0u;
    }
    const float di = 1.0f / sqrtf((float)deg);
    if (row < T) rsd[row] = di;
    mark(2);
    // the adjacency row (0 / 1, exact in bf16) replaces the similarity row: column w = frames (2w, 2w + 1)
#pragma unroll 1
    for (int c4 = 0; c4 < ROWS / 64; ++c4) {
        const uint32_t b0 = row < T ? adj[(2 * c4) * ROWS + row] : 0u, b1 = row < T ? adj[(2 * c4 + 1) * ROWS + row] : 0u;
        uint32_t pw[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const uint32_t two = ((i < 16 ? b0 : b1) >> (2 * (i & 15))) & 3u;
            pw[i] = ((two & 1u) * 0x3f80u) | ((two >> 1) * 0x3f800000u);
        }
        tmem_st32(tb + c4 * 32, pw);
    }
    tmem_st_wait();
    if (MT == 2) mbar_wait(&bars[B_S + MT - 1], 0);     // every similarity MMA has read the images before they are rescaled
    tc_fence_before();
    row_sync();                                        // every row's degree is known, every similarity row consumed
    tc_fence_after();
    mark(3);
    // ---- 4. y = d_j z_j -> two bf16 pieces over h | m (warp per row, lane = 4 columns, four rows per round);
    //         row sums of A_hat
#pragma unroll 1
    for (int r0 = warp; r0 < T; r0 += 4 * NW) {
        uint2 h[4], m[4], l[4];
        float dj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = min(r0 + u * NW, ROWS - 1);
            const uint32_t off = z_off<MT>(r, lane * 4);
            h[u] = *reinterpret_cast<const uint2*>(sm + L::ZH + off);
            m[u] = *reinterpret_cast<const uint2*>(sm + L::ZM + off);
            l[u] = *reinterpret_cast<const uint2*>(sm + L::ZL + off);
            dj[u] = rsd[r];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + u * NW;
            if (r < T) {
                const float y[4] = {((bf16_lo(h[u].x) + bf16_lo(m[u].x)) + bf16_lo(l[u].x)) * dj[u],
                                    ((bf16_hi(h[u].x) + bf16_hi(m[u].x)) + bf16_hi(l[u].x)) * dj[u],
                                    ((bf16_lo(h[u].y) + bf16_lo(m[u].y)) + bf16_lo(l[u].y)) * dj[u],
                                    ((bf16_hi(h[u].y) + bf16_hi(m[u].y)) + bf16_hi(l[u].y)) * dj[u]};
                uint2 oh, om;
                oh.x = pack_bf16(y[0], y[1]), oh.y = pack_bf16(y[2], y[3]);
                om.x = pack_bf16(y[0] - bf16_lo(oh.x), y[1] - bf16_hi(oh.x)), om.y = pack_bf16(y[2] - bf16_lo(oh.y), y[3] - bf16_hi(oh.y));
                const uint32_t off = z_off<MT>(r, lane * 4);
                *reinterpret_cast<uint2*>(sm + L::ZH + off) = oh;
                *reinterpret_cast<uint2*>(sm + L::ZM + off) = om;
            }
        }
    }
    float rs = 0.f;
    if (row < T) {
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            uint32_t bw = adj[c * ROWS + row];
            while (bw) {
                const int jr = 32 * c + __ffs(bw) - 1;
                bw &= bw - 1;
                rs += di * rsd[jr];
            }
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[B_RESC]);
    mark(4);
    // ---- 5. d_i (A y) -> bf16 scratch rows; row sums
    mbar_wait(&bars[B_AGG + mt], 0);
    tc_fence_after();
    mark(5);
    // (line-per-quad stores, quad_transpose8: the four lanes of a quad write one 128-byte line of one row per instruction)
    const long long R = seq * T + row;
    __nv_bfloat16* qbase = p.sq + (seq * T + (row & ~3)) * D;
#pragma unroll 1
    for (int c = 0; c < 4; c += 2) {
        uint32_t va[32], vb[32];
        tmem_ld32(tb + 128 + c * 32, va);
        tmem_ld32(tb + 128 + (c + 1) * 32, vb);
        tmem_ld_wait32(va);
        tmem_ld_wait32(vb);
        uint32_t w[4][8];                                  // 64 columns = 128 bytes = four 32-byte pieces
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = (q & 1) * 16 + 2 * i;
                w[q][i] = q < 2 ? pack_bf16(di * __uint_as_float(va[k]), di * __uint_as_float(va[k + 1]))
                                : pack_bf16(di * __uint_as_float(vb[k]), di * __uint_as_float(vb[k + 1]));
            }
        quad_transpose8(w, lane);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((row & ~3) + r < T) stg256u(qbase + r * D + c * 32 + (lane & 3) * 16, w[r]);
    }
    if (row < T) p.srow[R] = rs;
    mark(6);
    tc_fence_before();
    row_sync();
    if (warp == 0) tmem_dealloc(tmem, L::TM_COLS);
}

template <int MT>
static int launch_gcn_tc(const ModParams& p, int seqs, cudaStream_t st) {
    cudaFuncSetAttribute(long_gcn_tc_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Lay<MT>::TOTAL);
    long_gcn_tc_kernel<MT><<<seqs, 128 * MT + 32, Lay<MT>::TOTAL, st>>>(p);
    return cuda_status();
}

}  // namespace lgt
