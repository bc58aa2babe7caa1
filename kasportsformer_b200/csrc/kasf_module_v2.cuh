// kasf_module_v2.cuh -- the fused FormerModule kernel with TWO TILES IN FLIGHT per SM (included by kasf_module.cu).
//
// The one-tile kernel (former_module_kernel) runs the phases of a tile back to back: while its eight compute warps
// work on CUDA cores (LayerNorm, attention core / adjacency, GELU, epilogues) the tensor pipe idles, and while the
// projections run the warps wait (profiles/r01z_phase_cycles.json: 6-7k of 28-35k cycles per tile are MMA waits, issue
// slots 43 % used with two warps per scheduler).  This kernel splits a module at the point where the mixer's output
// has been added to the residual rows and pipelines the two halves over consecutive tiles:
//
//   mixer group (8 warps, 128 registers)   tile i+1: row gather, LN1, Q|K|V drains, attention core / adjacency +
//                                          aggregation, x1 = x + ls1 * mixer -> the output rows in global memory (L2)
//   MLP group   (8 warps,  96 registers)   tile i  : LN2(x1), eight GELU epilogues (64 hidden columns each), output
//                                          epilogue out = x1 + ls2 * (fc2 + b2)
//   service warpgroup (the CTA's first four warps): a weight producer and an MMA issuer PER GROUP.  The two streams
//                                          (mixer projections of tile i+1; fc1 / fc2 pieces of tile i) never wait
//                                          for each other; each has its own ring of 16 KB weight pieces.
//
// so that one tile's MMAs run under the other tile's CUDA-core phases and four warps per scheduler hide each other's
// latencies.  Every MMA has N = 64 .. 128 and its accumulators are double-buffered, so that a group drains one
// accumulator while the tensor cores fill the other (a hand-over per accumulator costs ~1-2k cycles next to sixteen busy
// warps: measured, scripts/trace_v2.py).  What had to shrink to make two tiles fit (measured: scripts/micro/ts_mma.cu):
//   tensor memory (512 columns): mixer accumulators 2 x 64 (the N-halves of K, V, Q, projection in turn) | fc1
//       accumulators 2 x 64 | the GELU output as the fc2 A OPERAND IN TENSOR MEMORY, 2 x 32 columns of packed fp16 pairs
//       (tcgen05.mma with A from TMEM: no shared-memory tile, no swizzled stores, the MMA reads only B from shared
//       memory) | fc2 accumulator 128.  No residual rows live in tensor memory: the mixer epilogue re-reads x and
//       hands x1 to the MLP group through the output rows themselves (L2 hits).
//   shared memory (227 KB): A1 32 | K|V / staged rows / z 64 | A2 (LN2 output) 32 | rings 2 + 3 x 16 | vectors | small arrays.
//       Weight pieces are 16 KB: an N-half [64 n x 128 k] of a chunk (two contiguous 8 KB blocks of the packed image)
//       for the mixer projections and fc1, a K-half [128 n x 64 k] for fc2 -- the blob is unchanged.
//   registers (640 threads x 96 at launch): the service warpgroup gives 64 of its 96 to the mixer group.
//
// Same arithmetic as the one-tile kernel (bf16 operands, fp16 hidden tile, fp32 everything else); the stage tests
// compare both against the same oracle.  Used for spatial modules and temporal modules with T <= 32 when the bone
// modules are fed from pre-normalised limb tiles (the path kasf_forward takes); everything else stays on the one-tile kernel.
namespace v2 {

constexpr int THREADS = 640;
// Warp roles.  The service warps are the FIRST warpgroup of the CTA (setmaxnreg works on aligned groups of four warps).
constexpr int W_PRODUCER_M = 0, W_MMA_M = 1, W_PRODUCER_P = 2, W_MMA_P = 3;
constexpr int W_G0 = 4, W_G1 = 12;      // first warp of the mixer group / of the MLP group (8 warps each)
constexpr int NSLOT = 5, MSLOTS = 2;    // ring slots: 0..MSLOTS-1 mixer projections, the rest fc1 / fc2 pieces
constexpr uint32_t SLOT = 16384;
constexpr uint32_t SM_A1 = 0, SM_KVZ = 32768, SM_A2 = 98304, SM_RING = 131072;
constexpr uint32_t SM_VEC = SM_RING + NSLOT * SLOT;
constexpr uint32_t SM_PART0 = SM_VEC + (uint32_t)MOD_VEC_BYTES;   // float2 [128][2]  mixer group
constexpr uint32_t SM_PART1 = SM_PART0 + 2048;                    // float2 [128][2]  MLP group
constexpr uint32_t SM_ADJ = SM_PART1 + 2048;                      // u32 [128][4]
constexpr uint32_t SM_ROWSUM = SM_ADJ + 2048;                     // f32 [128]
constexpr uint32_t SM_RSD = SM_ROWSUM + 512;                      // f32 [128]
constexpr uint32_t SM_BARS = SM_RSD + 512;
constexpr uint32_t SM_TOTAL = SM_BARS + 512;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
static_assert(SM_A1 == SM_A0 && SM_KVZ == SM_KV && SM_KVZ == SM_Z && SM_KVZ == SM_STAGE, "the shared helpers address A1 / K|V by the one-tile kernel's names");

struct LayG0 {
    static constexpr uint32_t PART = SM_PART0, ADJ = SM_ADJ, ROWSUM = SM_ROWSUM, RSD = SM_RSD;
    static constexpr int PAIR_BAR = 2;
};
struct LayG1 {
    static constexpr uint32_t PART = SM_PART1, ADJ = SM_ADJ, ROWSUM = SM_ROWSUM, RSD = SM_RSD;
    static constexpr int PAIR_BAR = 6;
};

// tensor memory columns: ACC0 | ACC1 mixer accumulators (64 each), H0 | H1 fc1 accumulators (64 each), HS0 | HS1 GELU
// pieces (fp16 pairs, 32 each), OUT fc2 accumulator (128)
constexpr uint32_t TM_ACC = 0, TM_H = 128, TM_HS = 256, TM_OUT = 384;

// mbarriers.  Rings: FULL (bulk-copy bytes) / EMPTY (tcgen05.commit).
//   mixer group -> issuer:  A1READY (A1 written: LN1 | attention output | A_hat z), ACCFREE0/1 (accumulator drained)
//   issuer -> mixer group:  ACCFULL0/1
//   mixer group -> MLP group: X1READY (x1 rows stored);  MLP group -> mixer group: X1TAKEN (the MLP group has started that
//                           tile: bounds the mixer group's lead, a waiter may lag one phase behind a barrier at most)
//   MLP group -> issuer:    A2READY (LN2 written), HSREADY0/1 (GELU piece in tensor memory, its fc1 accumulator drained)
//   issuer -> MLP group:    HFULL0/1, HSFREE0/1 (fc2 has read the piece), OUTFULL
//   ROWS: the row gather of a tile has landed;  bone: LIMBFULL (limb operand tile landed in A1), A1FREE (projection done)
enum { BB_FULL0 = 0, BB_EMPTY0 = NSLOT, BB_A1READY = 2 * NSLOT, BB_ACCFREE0, BB_ACCFREE1, BB_ACCFULL0, BB_ACCFULL1, BB_X1READY,
       BB_X1TAKEN, BB_A2READY, BB_HFREE0, BB_HFREE1, BB_HFULL0, BB_HFULL1, BB_HSREADY0, BB_HSREADY1, BB_HSFREE0, BB_HSFREE1,
       BB_OUTFULL, BB_ROWS, BB_LIMBFULL, BB_A1FREE, BB_COUNT };
static_assert(BB_COUNT <= 32 && BB_COUNT * 8 + 8 <= 512, "barrier block / one-register phase bits");

// The 64-bit shared-memory descriptor of a 128-byte-swizzled K-major operand differs between operands only in its low
// word (the start address in 16-byte units, LBO in the upper half); the high word (SBO = 1024 B, version 1, SWIZZLE_128B)
// is a constant that the MMA wrappers below splice in, so the issuers keep 32-bit values only.
// The issuer WARPS run their loops with all 32 lanes on warp-uniform values and hand each group of MMAs and its commits
// to ONE lane chosen by elect.sync: ptxas then keeps descriptors and addresses in uniform registers and emits
// back-to-back UTCHMMA (2-3 instructions per MMA).  Issued from inside an `if (lane == 0)` branch every MMA cost ~10
// instructions -- R2UR moves and an ELECT "waterfall" loop, because ptxas cannot know that one lane is active.
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
// A operand in tensor memory (16-bit pairs, lane = row, 8 columns per K = 16)
__device__ __forceinline__ void umma_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// 256-bit global accesses that bypass L1: x1 travels between the two groups through L2 (and a temporal module may work
// in place: the rows a tile re-reads must never come out of a stale L1 line)
__device__ __forceinline__ void ldcg256(const float* p, float* v) {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p)
                 : "memory");
}

// chunk index (kasf_layout.h) of mixer projection m
template <int KIND>
__device__ __forceinline__ int mixer_chunk(int m) {
    if (KIND == KASF_KIND_GRAPH) return m;                  // U, V
    return m == 0 ? 1 : (m == 1 ? 2 : (m == 2 ? 0 : 3));    // K, V, Q, proj
}

// Waiter side of the mbarriers of a compute thread (one register of phase bits, like Waiter), with the hardware
// suspending the thread between polls: here a waiting group shares its schedulers with a working one, and a hot
// polling loop would take issue slots from it.
struct WaiterS {
    uint32_t base;
    uint32_t phases;
    __device__ __forceinline__ void wait(int idx) {
        const uint32_t addr = base + idx * 8, parity = (phases >> idx) & 1u;
        uint32_t ok;
        do {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(addr), "r"(parity), "r"(0x989680u)
                : "memory");
        } while (!ok);
        phases ^= 1u << idx;
    }
};

// PROF builds: event trace of CTA 0, tiles 2..4: one region of TRACE_REGION events per role (0 mixer group thread 0,
// 1 MLP group thread 0, 2 / 3 the issuers, 4 / 5 the producers), plain stores (an atomic counter would put an L2
// round trip into every event): prof[24 + role] = events, prof[32 + 2 i] = tag | tile << 16, prof[33 + 2 i] = clock64.
// scripts/trace_v2.py prints the merged timeline.
constexpr int TRACE_REGION = 600;
template <bool PROF>
struct Tracer {
    int role, n;
    bool on;
    __device__ __forceinline__ void ev(const ModParams& p, int tag, int k) {
        if (PROF && on && k >= 2 && k <= 4 && n < TRACE_REGION) {
            const int i = role * TRACE_REGION + n;
            p.prof[32 + 2 * i] = (unsigned long long)(tag | (k << 16));
            p.prof[33 + 2 * i] = (unsigned long long)clock64();
            ++n;
        }
    }
    __device__ __forceinline__ void done(const ModParams& p) {
        if (PROF && on) p.prof[24 + role] = (unsigned long long)n;
    }
};

template <int KIND, int MODE, bool PROF = false>
__global__ void __launch_bounds__(THREADS, 1) former_module_v2_kernel(const ModParams p) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BARS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + SM_BARS + BB_COUNT * 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* vec = reinterpret_cast<const float*>(sm + SM_VEC);
    const uint8_t* chunks = p.mod + MOD_VEC_BYTES;

    if (tid == 0) {
        if ((smem_u32(sm) & 1023u) != 0) __trap();
        for (int i = 0; i < BB_COUNT; ++i) {
            const bool by_warps = i == BB_A1READY || i == BB_ACCFREE0 || i == BB_ACCFREE1 || i == BB_X1READY || i == BB_X1TAKEN ||
                                  i == BB_A2READY || i == BB_HFREE0 || i == BB_HFREE1 || i == BB_HSREADY0 || i == BB_HSREADY1;
            mbar_init(&bars[i], by_warps ? CW : (i == BB_ROWS ? CW * 32 * 2 : 1));
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    for (int i = tid; i < V_FLOATS / 4; i += THREADS)
        reinterpret_cast<float4*>(sm + SM_VEC)[i] = reinterpret_cast<const float4*>(p.mod)[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_local = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // >= 1: grid <= ntiles
    constexpr int NM = KIND == KASF_KIND_GRAPH ? 2 : 4;      // mixer projections per tile, two N-halves each

    if (warp < W_G0) {
        // ===================== service warpgroup =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == W_PRODUCER_M && lane == 0) {
            // ---- mixer ring: the N-halves of the tile's projections, in order
            uint32_t slot = 0, ph = 0, ph_a1free = 0;
            Tracer<PROF> tr{4, 0, blockIdx.x == 0};
#pragma unroll 1
            for (int k = 0; k < n_local; ++k) {
                const int tile = (int)blockIdx.x + k * (int)gridDim.x;
#pragma unroll 1
                for (int m = 0; m < NM; ++m) {
                    const uint8_t* c = chunks + (size_t)mixer_chunk<KIND>(m) * CHUNK_BYTES;
#pragma unroll 1
                    for (int h = 0; h < 2; ++h) {
                        // rows 64 h .. + 63 of both K sub-tiles (two 8 KB blocks)
                        mbar_wait_suspend(&bars[BB_EMPTY0 + slot], ph ^ 1);
                        mbar_arrive_expect_tx(&bars[BB_FULL0 + slot], SLOT);
                        uint8_t* dst = sm + SM_RING + slot * SLOT;
                        bulk_g2s(dst, c + h * 8192, 8192, &bars[BB_FULL0 + slot]);
                        bulk_g2s(dst + 8192, c + 16384 + h * 8192, 8192, &bars[BB_FULL0 + slot]);
                        tr.ev(p, 300 + m * 2 + h, k);
                        if (++slot == MSLOTS) slot = 0, ph ^= 1;
                        if (KIND == KASF_KIND_BONE && m == 0 && h == 1) {
                            // the tile's normalised limb rows: one 32 KB bulk copy into A1 once the previous tile's
                            // projection has read it (before V's pieces: those wait for K to be consumed, and K for this copy)
                            if (k > 0) {
                                mbar_wait_suspend(&bars[BB_A1FREE], ph_a1free);
                                ph_a1free ^= 1;
                            }
                            mbar_arrive_expect_tx(&bars[BB_LIMBFULL], TILE_BYTES);
                            bulk_g2s(sm + SM_A1, p.xlt + (size_t)tile * TILE_BYTES, TILE_BYTES, &bars[BB_LIMBFULL]);
                        }
                    }
                }
            }
            tr.done(p);
        } else if (warp == W_PRODUCER_P && lane == 0) {
            // ---- MLP ring: fc1(0), fc1(1), then fc2(q), fc1(q+2) for q = 0..7 (fc1 while q + 2 < 8)
            uint32_t slot = MSLOTS, ph = 0;
            Tracer<PROF> tr{5, 0, blockIdx.x == 0};
#pragma unroll 1
            for (int k = 0; k < n_local; ++k) {
#pragma unroll 1
                for (int j = 0; j < 18; ++j) {
                    const int q = j < 2 ? j : (j - 2) >> 1;
                    const bool fc1 = j < 2 || ((j - 2) & 1);
                    const int piece = j < 2 ? j : (fc1 ? q + 2 : q);
                    if (fc1 && piece >= 8) continue;
                    mbar_wait_suspend(&bars[BB_EMPTY0 + slot], ph ^ 1);
                    mbar_arrive_expect_tx(&bars[BB_FULL0 + slot], SLOT);
                    uint8_t* dst = sm + SM_RING + slot * SLOT;
                    if (fc1) {
                        // an N-half of a W1 chunk: rows 64 (piece & 1) .. + 63 of both K sub-tiles (two 8 KB blocks)
                        const uint8_t* c = chunks + (size_t)(4 + (piece >> 1)) * CHUNK_BYTES + (piece & 1) * 8192;
                        bulk_g2s(dst, c, 8192, &bars[BB_FULL0 + slot]);
                        bulk_g2s(dst + 8192, c + 16384, 8192, &bars[BB_FULL0 + slot]);
                    } else {
                        bulk_g2s(dst, chunks + (size_t)(8 + (piece >> 1)) * CHUNK_BYTES + (piece & 1) * SLOT, SLOT, &bars[BB_FULL0 + slot]);
                    }
                    tr.ev(p, 320 + j, k);
                    if (++slot == NSLOT) slot = MSLOTS, ph ^= 1;
                }
            }
            tr.done(p);
        } else if (warp == W_MMA_M) {
            // ---- issuer of the mixer projections (tile the mixer group works on): N-half h of projection m goes to
            //      accumulator h; it waits for that accumulator to be drained (ACCFREE h) and, where the A operand
            //      changes, for A1
            const uint32_t a1_addr = smem_u32(sm + SM_A1), ring_addr = smem_u32(sm + SM_RING);
            uint32_t phs = (1u << BB_ACCFREE0) | (1u << BB_ACCFREE1);   // both accumulators start out free
            auto wait = [&](int idx) {           // lane 0 polls, the warp reconverges: the issue code stays warp-uniform
                if (lane == 0) mbar_wait(&bars[idx], (phs >> idx) & 1u);
                __syncwarp();
                phs ^= 1u << idx;
            };
            uint32_t mslot = 0, mph = 0;
            Tracer<PROF> tr{2, 0, blockIdx.x == 0 && lane == 0};
#pragma unroll 1
            for (int k = 0; k < n_local; ++k) {
#pragma unroll 1
                for (int m = 0; m < NM; ++m) {
                    bool new_a, accumulate = false;
                    if (KIND == KASF_KIND_ATTENTION) new_a = m == 0 || m == 3;
                    else if (KIND == KASF_KIND_BONE) new_a = m != 1;
                    else new_a = true, accumulate = m == 1;
                    if (new_a) wait((KIND == KASF_KIND_BONE && m == 0) ? BB_LIMBFULL : BB_A1READY);
                    const uint32_t idesc = umma_idesc_bf16(128, 64);
                    const uint32_t a_lo = desc_lo(a1_addr);
#pragma unroll 1
                    for (uint32_t h = 0; h < 2; ++h) {
                        if (!accumulate) wait(BB_ACCFREE0 + h);
                        if (lane == 0) mbar_wait(&bars[BB_FULL0 + mslot], mph);
                        __syncwarp();
                        tc_fence_after();
                        tr.ev(p, 100 + m * 2 + h, k);
                        const uint32_t b_lo = desc_lo(ring_addr + mslot * SLOT);
                        if (elect_one()) {
#pragma unroll
                            for (uint32_t ks = 0; ks < 8; ++ks)
                                umma_lo(tmem + TM_ACC + h * 64, a_lo + (((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4),
                                        b_lo + (((ks >> 2) * 8192u + (ks & 3) * 32u) >> 4), idesc, (accumulate || ks > 0) ? 1u : 0u);
                            tc_commit(&bars[BB_EMPTY0 + mslot]);
                            tc_commit(&bars[BB_ACCFULL0 + h]);
                            if (KIND == KASF_KIND_BONE && m == 3 && h == 1) tc_commit(&bars[BB_A1FREE]);
                        }
                        __syncwarp();
                        if (++mslot == MSLOTS) mslot = 0, mph ^= 1;
                    }
                }
            }
            tr.done(p);
        } else if (warp == W_MMA_P) {
            // ---- issuer of the fc1 / fc2 pieces (tile the MLP group works on).  fc1 runs two pieces ahead of the GELU
            //      epilogues over the two accumulators; ONE trigger per piece (HSREADY q: the GELU piece is in tensor
            //      memory and its fc1 accumulator drained) releases fc2(q) and fc1(q+2) together -- a barrier check
            //      costs this warp ~100 cycles even when it has completed, sixteen of them per tile were its bottleneck
            const uint32_t a2_addr = smem_u32(sm + SM_A2), ring_addr = smem_u32(sm + SM_RING);
            uint32_t phs = 0;
            auto wait = [&](int idx) {
                if (lane == 0) mbar_wait(&bars[idx], (phs >> idx) & 1u);
                __syncwarp();
                phs ^= 1u << idx;
            };
            uint32_t pslot = MSLOTS, pph = 0;
            Tracer<PROF> tr{3, 0, blockIdx.x == 0 && lane == 0};
            auto slot_ready = [&]() -> uint32_t {
                if (lane == 0) mbar_wait(&bars[BB_FULL0 + pslot], pph);
                __syncwarp();
                tc_fence_after();
                return desc_lo(ring_addr + pslot * SLOT);
            };
            auto slot_next = [&]() {
                if (++pslot == NSLOT) pslot = MSLOTS, pph ^= 1;
            };
            auto fc1 = [&](int piece) {          // [128 x 128] x [64 x 128]^T -> H[piece & 1]
                const uint32_t b_lo = slot_ready();
                const uint32_t idesc = umma_idesc_bf16(128, 64);
                const uint32_t a_lo = desc_lo(a2_addr);
                if (elect_one()) {
#pragma unroll
                    for (uint32_t ks = 0; ks < 8; ++ks)
                        umma_lo(tmem + TM_H + (piece & 1) * 64, a_lo + (((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4),
                                b_lo + (((ks >> 2) * 8192u + (ks & 3) * 32u) >> 4), idesc, ks > 0 ? 1u : 0u);
                    tc_commit(&bars[BB_HFULL0 + (piece & 1)]);
                    tc_commit(&bars[BB_EMPTY0 + pslot]);
                }
                __syncwarp();
                slot_next();
            };
#pragma unroll 1
            for (int k = 0; k < n_local; ++k) {
                // (the accumulators are free: the MLP group drained pieces 6, 7 of the previous tile before it stored them,
                //  and those stores were this thread's last two triggers)
                wait(BB_A2READY);
                tr.ev(p, 160, k);
                fc1(0);
                fc1(1);
                tr.ev(p, 140, k);
#pragma unroll 1
                for (int q = 0; q < 8; ++q) {
                    wait(BB_HSREADY0 + (q & 1));
                    tr.ev(p, 161 + q, k);
                    {   // fc2 piece q: OUT (+)= GELU piece [128 x 64] (tensor memory, fp16) x [128 x 64]^T
                        const uint32_t b_lo = slot_ready();
                        const uint32_t idesc = KASF_HALF_GELU ? umma_idesc_f16(128, 128) : umma_idesc_bf16(128, 128);
                        const uint32_t hs = tmem + TM_HS + (q & 1) * 32;
                        if (elect_one()) {
#pragma unroll
                            for (uint32_t kk = 0; kk < 4; ++kk)
                                umma_ts_lo(tmem + TM_OUT, hs + kk * 8u, b_lo + kk * 2u, idesc, (q > 0 || kk > 0) ? 1u : 0u);
                            tc_commit(&bars[BB_HSFREE0 + (q & 1)]);
                            if (q == 7) tc_commit(&bars[BB_OUTFULL]);
                            tc_commit(&bars[BB_EMPTY0 + pslot]);
                        }
                        __syncwarp();
                        slot_next();
                    }
                    if (q + 2 < 8) fc1(q + 2);
                    tr.ev(p, 141 + q, k);
                }
            }
            tr.done(p);
        }
        __syncwarp();
    } else if (warp < W_G1) {
        // ===================== mixer group =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
        const int gw = warp - W_G0;        // warp of the group, 0..7 (gw & 3 == warp & 3: the TMEM lane quarter)
        EpiMap e;
        e.row = (gw & 3) * 32 + lane;
        e.half = gw >> 2;
        e.warp = gw;
        e.tbase = tmem + ((uint32_t)((gw & 3) * 32) << 16);
        WaiterS wt{smem_u32(bars), 0u};
        Tracer<PROF> tr{0, 0, blockIdx.x == 0 && gw == 0 && lane == 0};
        const bool mark = gw == 0 && lane == 0;
        long long pt0 = PROF ? clock64() : 0;
#define PMARK2(kk)                                                    \
    do {                                                              \
        if (PROF && mark) {                                           \
            const long long pt1 = clock64();                          \
            atomicAdd(p.prof + (kk), (unsigned long long)(pt1 - pt0)); \
            pt0 = pt1;                                                \
        }                                                             \
    } while (0)

        gather_rows<MODE>(p, sm, blockIdx.x, p.in, &bars[BB_ROWS], gw, lane, 0);
        gather_rows<MODE>(p, sm, blockIdx.x, p.in, &bars[BB_ROWS], gw, lane, 1);
#pragma unroll 1
        for (int k = 0; k < n_local; ++k) {
            const int tile = (int)blockIdx.x + k * (int)gridDim.x;
            const int nrows = tile_rows<MODE>(p, tile);
            const int gsize = MODE == KASF_MODE_SPATIAL ? J : p.T;
            const bool row_ok = e.row < nrows;
            const long long tok = row_ok ? row_token<MODE>(p, tile, e.row) : -1;
            float mean, rstd;
            float rs = 0.f;

            // N-half `nh` of Q (qkv 0), K (1) or V (2): accumulator nh, this thread's 32 columns -> bf16 in shared memory
            auto drain = [&](int qkv, int nh) {
                uint32_t acc[32];
                tmem_ld32(e.tbase + TM_ACC + nh * 64 + e.half * 32, acc);
                tmem_ld_wait();
                const int col0 = nh * 64 + e.half * 32;
                if (qkv == 0) {                        // query bias W_q beta_1 (LN1's affine lives in the weights)
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 bq = *reinterpret_cast<const float4*>(vec + V_BQ + col0 + c4 * 4);
                        acc[c4 * 4 + 0] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 0]) + bq.x);
                        acc[c4 * 4 + 1] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 1]) + bq.y);
                        acc[c4 * 4 + 2] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 2]) + bq.z);
                        acc[c4 * 4 + 3] = __float_as_uint(__uint_as_float(acc[c4 * 4 + 3]) + bq.w);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 pk;
                    pk.x = pack_bf16(__uint_as_float(acc[c * 8 + 0]), __uint_as_float(acc[c * 8 + 1]));
                    pk.y = pack_bf16(__uint_as_float(acc[c * 8 + 2]), __uint_as_float(acc[c * 8 + 3]));
                    pk.z = pack_bf16(__uint_as_float(acc[c * 8 + 4]), __uint_as_float(acc[c * 8 + 5]));
                    pk.w = pack_bf16(__uint_as_float(acc[c * 8 + 6]), __uint_as_float(acc[c * 8 + 7]));
                    if (qkv == 0) {
                        *reinterpret_cast<uint4*>(sm + SM_A1 + tile_off_bf16(e.row, col0 + c * 8)) = pk;
                    } else {
                        const uint32_t chunk = (qkv - 1) * 16 + (col0 >> 3) + c;
                        *reinterpret_cast<uint4*>(sm + SM_KVZ + f32_off(e.row, chunk)) = pk;
                    }
                }
            };
            // K then V: the halves alternate between the two accumulators, a drain runs under the next half's MMAs
            auto drain_kv = [&]() {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    wt.wait(BB_ACCFULL0 + (i & 1));
                    tc_fence_after();
                    drain(1 + (i >> 1), i & 1);
                    warp_arrive(&bars[BB_ACCFREE0 + (i & 1)], lane);
                }
            };

            {
                float xv[64];
                wt.wait(BB_ROWS);
                PMARK2(0);
                tr.ev(p, 0, k);
                read_staged(sm, e, xv, row_ok);
                if (KIND == KASF_KIND_BONE) csync();   // every staged row is in registers: the K|V drains may overwrite them
                ln_stats<LayG0>(sm, e, xv, mean, rstd);
                if (KIND == KASF_KIND_ATTENTION) {
                    ln_write<false, false>(sm, SM_A1, e, xv, mean, rstd, nullptr, nullptr, row_ok);
                    warp_arrive(&bars[BB_A1READY], lane);
                    PMARK2(1);
                    tr.ev(p, 1, k);
                    // the first half of K complete => every warp has arrived on A1READY, i.e. has read its staged rows
                    drain_kv();
                } else if (KIND == KASF_KIND_BONE) {
                    PMARK2(1);
                    tr.ev(p, 1, k);
                    drain_kv();                        // K, V of the limb tile; afterwards A1 is free
                    ln_write<false, false>(sm, SM_A1, e, xv, mean, rstd, nullptr, nullptr, row_ok);
                    warp_arrive(&bars[BB_A1READY], lane);
                } else {
                    csync();                           // z (fp32) replaces the staged rows
                    ln_write<true, true>(sm, SM_A1, e, xv, mean, rstd, vec + V_N1W, vec + V_N1B, row_ok);
                    warp_arrive(&bars[BB_A1READY], lane);
                    PMARK2(1);
                    tr.ev(p, 1, k);
                }
            }
            if (KIND != KASF_KIND_GRAPH) {
                // Q replaces LN1 in A1: both halves must be complete before the first one is drained
                wt.wait(BB_ACCFULL0);
                wt.wait(BB_ACCFULL1);
                tc_fence_after();
                drain(0, 0);
                drain(0, 1);
                {   // both accumulators are free for the projection
                    fence_proxy_async();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&bars[BB_ACCFREE0]);
                        mbar_arrive(&bars[BB_ACCFREE1]);
                    }
                }
                csync();                               // every warp reads the K|V rows of the others
                PMARK2(2);
                tr.ev(p, 2, k);
                attention_core<MODE, 0>(sm, gw, lane, gsize, nrows);
                warp_arrive(&bars[BB_A1READY], lane);
                csync();                               // K|V are dead: the next tile's rows may land there
            } else {
                csync();                               // z of the whole tile is in shared memory
                if (MODE == KASF_MODE_TEMPORAL) {
                    similarity_topk<0, LayG0>(sm, gw, lane, p.T, nrows);
                    csync();
                }
                PMARK2(2);
                tr.ev(p, 2, k);
                // ---- aggregation  agg_i = sum_j A_ij / sqrt(d_i d_j) * z_j : this thread's 64 columns of its row
                float ag[64];
#pragma unroll
                for (int i = 0; i < 64; ++i) ag[i] = 0.f;
                if (row_ok) {
                    if (MODE == KASF_MODE_SPATIAL) {
                        const int j = e.row % J, base = e.row - j;
                        const float di = c_rsd[c_deg[j]];
#pragma unroll 1
                        for (int n = 0; n < 4; ++n) {
                            const int nb = c_nbr[j * 4 + n];
                            if (nb < 0) break;
                            const float cf = di * c_rsd[c_deg[nb]];
                            rs += cf;
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const float4 z = *reinterpret_cast<const float4*>(sm + SM_KVZ + f32_off(base + nb, e.half * 16 + c));
                                ag[c * 4] = fmaf(cf, z.x, ag[c * 4]), ag[c * 4 + 1] = fmaf(cf, z.y, ag[c * 4 + 1]);
                                ag[c * 4 + 2] = fmaf(cf, z.z, ag[c * 4 + 2]), ag[c * 4 + 3] = fmaf(cf, z.w, ag[c * 4 + 3]);
                            }
                        }
                    } else {
                        const uint32_t* adj = reinterpret_cast<const uint32_t*>(sm + SM_ADJ);
                        const float* rsd = reinterpret_cast<const float*>(sm + SM_RSD);
                        const int gr0 = (e.row / p.T) * p.T;
                        const float di = rsd[e.row];
                        unsigned bits = adj[e.row * 4];          // T <= 32: one word
#pragma unroll 1
                        while (bits) {
                            const int jb = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const int jr = gr0 + jb;
                            const float cf = di * rsd[jr];
                            rs += cf;
#pragma unroll
                            for (int c = 0; c < 16; ++c) {
                                const float4 z = *reinterpret_cast<const float4*>(sm + SM_KVZ + f32_off(jr, e.half * 16 + c));
                                ag[c * 4] = fmaf(cf, z.x, ag[c * 4]), ag[c * 4 + 1] = fmaf(cf, z.y, ag[c * 4 + 1]);
                                ag[c * 4 + 2] = fmaf(cf, z.z, ag[c * 4 + 2]), ag[c * 4 + 3] = fmaf(cf, z.w, ag[c * 4 + 3]);
                            }
                        }
                    }
                }
                wt.wait(BB_ACCFULL0);                  // U z done (both halves): A1 may be overwritten
                wt.wait(BB_ACCFULL1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint4 pk;
                    pk.x = pack_bf16(ag[c * 8 + 0], ag[c * 8 + 1]), pk.y = pack_bf16(ag[c * 8 + 2], ag[c * 8 + 3]);
                    pk.z = pack_bf16(ag[c * 8 + 4], ag[c * 8 + 5]), pk.w = pack_bf16(ag[c * 8 + 6], ag[c * 8 + 7]);
                    *reinterpret_cast<uint4*>(sm + SM_A1 + tile_off_bf16(e.row, e.half * 64 + c * 8)) = pk;
                }
                warp_arrive(&bars[BB_A1READY], lane);
                csync();                               // z is dead (the epilogue recomputes it from x): next rows may land
            }
            PMARK2(3);
            tr.ev(p, 3, k);
            if (k + 1 < n_local) {
                gather_rows<MODE>(p, sm, tile + (int)gridDim.x, p.in, &bars[BB_ROWS], gw, lane, 0);
                gather_rows<MODE>(p, sm, tile + (int)gridDim.x, p.in, &bars[BB_ROWS], gw, lane, 1);
            }
            // ---- x1 = x + ls1 * mixer -> the output rows (L2).  This thread's columns: 32 of each accumulator half.
            //      x is re-read (L2), requested before the waits.
            {
                float xg[64];
                const long long rowoff = (tok >= 0 ? tok : 0) * D;
                if (row_ok) {
#pragma unroll
                    for (int b = 0; b < 2; ++b)
#pragma unroll
                        for (int c = 0; c < 4; ++c) ldcg256(p.in + rowoff + b * 64 + e.half * 32 + c * 8, xg + b * 32 + c * 8);
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) xg[i] = 0.f;
                }
                float bn_s = 1.f, bn_t = 0.f;
                if (KIND == KASF_KIND_GRAPH) {
                    const int node = MODE == KASF_MODE_SPATIAL ? e.row % J : e.row % p.T;
                    bn_s = vec[V_BNS + node];
                    bn_t = vec[V_BNT + node];
                }
                const float nm = -mean * rstd;
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    wt.wait(BB_ACCFULL0 + b);          // projection | += (A_hat z) V^T, N-half b
                    tc_fence_after();
                    if (b == 0) {
                        PMARK2(4);
                        tr.ev(p, 4, k);
                    }
                    uint32_t acc[32];
                    tmem_ld32(e.tbase + TM_ACC + b * 64 + e.half * 32, acc);
                    tmem_ld_wait();
                    warp_arrive(&bars[BB_ACCFREE0 + b], lane);
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const int col = b * 64 + e.half * 32 + c4 * 4;
                        const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS1 + col);
                        const float4 bm = *reinterpret_cast<const float4*>(vec + V_BMIX + col);
                        float m0 = __uint_as_float(acc[c4 * 4 + 0]) + bm.x, m1 = __uint_as_float(acc[c4 * 4 + 1]) + bm.y,
                              m2 = __uint_as_float(acc[c4 * 4 + 2]) + bm.z, m3 = __uint_as_float(acc[c4 * 4 + 3]) + bm.w;
                        const float x0 = xg[b * 32 + c4 * 4 + 0], x1 = xg[b * 32 + c4 * 4 + 1], x2 = xg[b * 32 + c4 * 4 + 2],
                                    x3 = xg[b * 32 + c4 * 4 + 3];
                        if (KIND == KASF_KIND_GRAPH) {
                            // mix = relu(z + BN_node(acc + bU + rowsum * bV)), z = LN1(x) recomputed exactly as ln_write does
                            const float4 bv = *reinterpret_cast<const float4*>(vec + V_BV + col);
                            const float4 g = *reinterpret_cast<const float4*>(vec + V_N1W + col);
                            const float4 be = *reinterpret_cast<const float4*>(vec + V_N1B + col);
                            const float z0 = row_ok ? fmaf(fmaf(x0, rstd, nm), g.x, be.x) : 0.f, z1 = row_ok ? fmaf(fmaf(x1, rstd, nm), g.y, be.y) : 0.f,
                                        z2 = row_ok ? fmaf(fmaf(x2, rstd, nm), g.z, be.z) : 0.f, z3 = row_ok ? fmaf(fmaf(x3, rstd, nm), g.w, be.w) : 0.f;
                            m0 = fmaxf(z0 + ((m0 + rs * bv.x) * bn_s + bn_t), 0.f);
                            m1 = fmaxf(z1 + ((m1 + rs * bv.y) * bn_s + bn_t), 0.f);
                            m2 = fmaxf(z2 + ((m2 + rs * bv.z) * bn_s + bn_t), 0.f);
                            m3 = fmaxf(z3 + ((m3 + rs * bv.w) * bn_s + bn_t), 0.f);
                        }
                        xg[b * 32 + c4 * 4 + 0] = fmaf(ls.x, m0, x0);
                        xg[b * 32 + c4 * 4 + 1] = fmaf(ls.y, m1, x1);
                        xg[b * 32 + c4 * 4 + 2] = fmaf(ls.z, m2, x2);
                        xg[b * 32 + c4 * 4 + 3] = fmaf(ls.w, m3, x3);
                    }
                    if (row_ok) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) stg256(p.out + rowoff + b * 64 + e.half * 32 + c * 8, xg + b * 32 + c * 8);
                    }
                }
                // the MLP group must have started the previous tile before X1READY completes another phase
                if (k > 0) wt.wait(BB_X1TAKEN);
                __threadfence_block();
                warp_arrive(&bars[BB_X1READY], lane);
            }
            PMARK2(5);
            tr.ev(p, 5, k);
        }
        tr.done(p);
    } else {
        // ===================== MLP group (keeps the 96 registers of the launch) =====================
        const int g = warp - W_G1;
        EpiMap e;
        e.row = (g & 3) * 32 + lane;
        e.half = g >> 2;
        e.warp = g;
        e.tbase = tmem + ((uint32_t)((g & 3) * 32) << 16);
        WaiterS wt{smem_u32(bars), (1u << BB_HSFREE0) | (1u << BB_HSFREE1)};   // the two GELU buffers start out free
        Tracer<PROF> tr{1, 0, blockIdx.x == 0 && g == 0 && lane == 0};
        const bool mark = g == 0 && lane == 0;
        long long pt0 = PROF ? clock64() : 0;
#pragma unroll 1
        for (int k = 0; k < n_local; ++k) {
            const int tile = (int)blockIdx.x + k * (int)gridDim.x;
            const int nrows = tile_rows<MODE>(p, tile);
            const bool row_ok = e.row < nrows;
            const long long tok = row_ok ? row_token<MODE>(p, tile, e.row) : -1;
            float* orow = p.out + (tok >= 0 ? tok : 0) * D + e.half * 64;
            // ---- LN2(x1) -> A2
            wt.wait(BB_X1READY);
            warp_arrive(&bars[BB_X1TAKEN], lane);
            PMARK2(8);
            tr.ev(p, 230, k);
            {
                float xv[64];
                if (row_ok) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) ldcg256(orow + c * 8, xv + c * 8);
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) xv[i] = 0.f;
                }
                float mean, rstd;
                ln_stats<LayG1>(sm, e, xv, mean, rstd);
                // (A2 is free: every fc1 piece of the previous tile completed before this thread saw its last HFULL)
                ln_write<false, false>(sm, SM_A2, e, xv, mean, rstd, nullptr, nullptr, row_ok);
            }
            warp_arrive(&bars[BB_A2READY], lane);
            PMARK2(9);
            tr.ev(p, 231, k);
            // ---- eight GELU epilogues: fc1 accumulator (64 hidden columns) -> 2*GELU -> packed fp16 pairs in tensor memory
#pragma unroll 1
            for (int q = 0; q < 8; ++q) {
                wt.wait(BB_HFULL0 + (q & 1));
                tc_fence_after();
                PMARK2(10);
                tr.ev(p, 200 + q, k);
                uint32_t acc[32];
                tmem_ld32(e.tbase + TM_H + (q & 1) * 64 + e.half * 32, acc);
                tmem_ld_wait();
                tr.ev(p, 210 + q, k);
                uint32_t hs[16];
#if KASF_HALF_GELU
                const uint4* b1h = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(vec + V_B1H) + q * 64 + e.half * 32);
                __half2 v[2][4], w[2][4];
                auto stage1 = [&](int gq, int s2) {
                    const uint4 bh = b1h[gq];
                    const uint32_t* a8 = &acc[gq * 8];
                    v[s2][0] = __hadd2(u2h(pack_f16(__uint_as_float(a8[0]), __uint_as_float(a8[1]))), u2h(bh.x));
                    v[s2][1] = __hadd2(u2h(pack_f16(__uint_as_float(a8[2]), __uint_as_float(a8[3]))), u2h(bh.y));
                    v[s2][2] = __hadd2(u2h(pack_f16(__uint_as_float(a8[4]), __uint_as_float(a8[5]))), u2h(bh.z));
                    v[s2][3] = __hadd2(u2h(pack_f16(__uint_as_float(a8[6]), __uint_as_float(a8[7]))), u2h(bh.w));
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[s2][i] = gelu2_arg_h2(v[s2][i]);
                };
                stage1(0, 0);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const int s2 = gq & 1;
                    __half2 t[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) t[i] = tanh_h2(w[s2][i]);
                    if (gq + 1 < 4) stage1(gq + 1, s2 ^ 1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) hs[gq * 4 + i] = h2u(__hfma2(v[s2][i], t[i], v[s2][i]));
                }
#else
                const float* b1 = vec + V_B1 + q * 64 + e.half * 32;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float v0 = __uint_as_float(acc[2 * i]) + b1[2 * i], v1 = __uint_as_float(acc[2 * i + 1]) + b1[2 * i + 1];
                    hs[i] = pack_bf16(fmaf(v0, gelu2_tanh(gelu2_arg(v0)), v0), fmaf(v1, gelu2_tanh(gelu2_arg(v1)), v1));
                }
#endif
                PMARK2(11);
                wt.wait(BB_HSFREE0 + (q & 1));         // fc2 of piece q-2 has read this buffer
                tc_fence_after();
                PMARK2(12);
                tmem_st16(e.tbase + TM_HS + (q & 1) * 32 + e.half * 16, hs);
                tmem_st_wait();
                warp_arrive(&bars[BB_HSREADY0 + (q & 1)], lane);
                tr.ev(p, 220 + q, k);
            }
            // ---- out = x1 + ls2 * (acc + b2): x1 re-read (L2), 256-bit stores of this thread's 64 columns
            float xr[64];
            if (row_ok) {
#pragma unroll
                for (int c = 0; c < 8; ++c) ldcg256(orow + c * 8, xr + c * 8);
            }
            wt.wait(BB_OUTFULL);
            tc_fence_after();
            PMARK2(13);
            {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    uint32_t acc[32];
                    tmem_ld32(e.tbase + TM_OUT + e.half * 64 + b * 32, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        const int col = e.half * 64 + b * 32 + c8 * 8;
                        float o[8];
#pragma unroll
                        for (int h4 = 0; h4 < 2; ++h4) {
                            const float4 ls = *reinterpret_cast<const float4*>(vec + V_LS2 + col + h4 * 4);
                            const float4 b2 = *reinterpret_cast<const float4*>(vec + V_B2 + col + h4 * 4);
                            const int i = c8 * 8 + h4 * 4;
                            o[h4 * 4 + 0] = fmaf(ls.x, __uint_as_float(acc[i + 0]) + b2.x, xr[b * 32 + i + 0]);
                            o[h4 * 4 + 1] = fmaf(ls.y, __uint_as_float(acc[i + 1]) + b2.y, xr[b * 32 + i + 1]);
                            o[h4 * 4 + 2] = fmaf(ls.z, __uint_as_float(acc[i + 2]) + b2.z, xr[b * 32 + i + 2]);
                            o[h4 * 4 + 3] = fmaf(ls.w, __uint_as_float(acc[i + 3]) + b2.w, xr[b * 32 + i + 3]);
                        }
                        if (tok >= 0) stg256(orow + b * 32 + c8 * 8, o);
                    }
                }
            }
            // (the fc2 accumulator is drained: the next tile's first fc2 piece follows this thread's next HSREADY)
            tc_fence_before();
            PMARK2(14);
            tr.ev(p, 232, k);
        }
        tr.done(p);
#undef PMARK2
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int KIND, int MODE>
static int launch_v2(const ModParams& p, cudaStream_t st, int sms) {
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    if (p.prof) {
        cudaFuncSetAttribute(former_module_v2_kernel<KIND, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
        former_module_v2_kernel<KIND, MODE, true><<<grid, THREADS, SM_TOTAL, st>>>(p);
        return cuda_status();
    }
    cudaFuncSetAttribute(former_module_v2_kernel<KIND, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    former_module_v2_kernel<KIND, MODE><<<grid, THREADS, SM_TOTAL, st>>>(p);
    return cuda_status();
}

}  // namespace v2
