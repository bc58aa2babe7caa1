// kasf_ptx.cuh -- thin inline-PTX wrappers for sm_100a: mbarrier, bulk async copy (TMA engine),
// tcgen05 (tensor memory alloc / MMA / commit / load) and the UMMA descriptors used by libkasf.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace kasf {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking test (try_wait may suspend the thread for a while; a poller of several barriers must not)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// same, but the thread is suspended by the hardware between polls (producer side: long waits, no issue slots)
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "KASF_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra KASF_DONE_%=;\n\t"
        "bra KASF_WAIT_%=;\n\t"
        "KASF_DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (UMMA / bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- bulk async copy (UBLKCP)
// global -> shared, completion signalled on an mbarrier by transaction bytes. 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok));
    return ok != 0;
}

// ---------------------------------------------------------------- 256-bit global accesses (LDG/STG.E.256)
// one full 32-byte sector per thread and instruction; addresses must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void stg256u(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// 4 x 4 transpose of 8-word pieces among the four lanes of a quad (two xor-shuffle stages): before, o[c] is piece c of
// the row this lane owns; after, o[r] is piece (lane & 3) of the row owned by lane (quad base + r).  The quad can then
// store ONE contiguous 128-byte line of row r per instruction (8 lines x 4 sectors per warp instruction) instead of 32
// single sectors of 32 different lines: 2.0k instead of 3.9k cycles per 64 KB on one SM (scripts/micro/stg_patterns.cu).
__device__ __forceinline__ void quad_transpose8(uint32_t (&o)[4][8], int lane) {
    const int s4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t send = (s4 & 2) ? o[i][k] : o[i + 2][k];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
            if (s4 & 2) o[i][k] = recv; else o[i + 2][k] = recv;
        }
#pragma unroll
    for (int i = 0; i < 4; i += 2)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t send = (s4 & 1) ? o[i][k] : o[i + 1][k];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
            if (s4 & 1) o[i][k] = recv; else o[i + 1][k] = recv;
        }
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, 16-bit float inputs (bf16 or fp16, see the instruction descriptor), fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same with the A operand in tensor memory (16-bit pairs, lane = row, 8 columns per K = 16): layout verified
// against the shared-memory form in scripts/micro/ts_mma.cu
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32b, N consecutive columns per thread (thread i of the warp <-> TMEM lane base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// wait for the outstanding tcgen05.ld of this thread; the destination registers are in/out operands so that the
// compiler cannot schedule a use of them above the wait (needed when independent work sits between ld and wait)
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// registers -> TMEM, 32 consecutive columns of this thread's lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- cp.async (LDGSTS), 16 bytes per thread
// src_bytes = 0 zero-fills the destination (padding rows)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes)
                 : "memory");
}
// the mbarrier receives one arrival (counted in its init count) once all prior cp.async of this thread landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B (64 bf16) wide,
// 8-row groups are 1024 B apart (SBO), version field = 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address  [0,14)
    d |= (uint64_t)1 << 16;                        // LBO (ignored for swizzled K-major) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // SBO = 1024 B   [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// Instruction descriptor: kind::f16, A=B=fp16 (format code 0), D=fp32, both K-major, shape M x N (K = 16).
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// Instruction descriptor: kind::f16, A=B=bf16, D=fp32, both K-major, shape M x N (K = 16).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4)            // D format: f32
           | (1u << 7)          // A format: bf16
           | (1u << 10)         // B format: bf16
           | ((N >> 3) << 17)   // N / 8
           | ((M >> 4) << 24);  // M / 16
}

// ---------------------------------------------------------------- operand tile layout
// A [128 rows x 128 k] bf16 operand tile = two sub-tiles of [128 rows x 64 k]; inside a sub-tile
// row r occupies bytes [128 r, 128 r + 128) and its eight 16-byte chunks are XOR-swizzled with
// (r & 7).  Byte offset of element (r, k):
__host__ __device__ constexpr uint32_t tile_off_bf16(uint32_t r, uint32_t k) {
    return (k >> 6) * 16384u + r * 128u + ((((k & 63u) >> 3) ^ (r & 7u)) << 4) + ((k & 7u) << 1);
}
constexpr uint32_t TILE_BYTES = 32768;   // one [128 x 128] bf16 operand tile (also one weight chunk)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 -> packed fp16x2 (round to nearest, saturating to the largest finite half instead of infinity)
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Issue the 8 K-steps (K = 128) of D[128 x N] (+)= A_tile * B_tile^T.  One thread.
__device__ __forceinline__ void umma_tile_k128(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t N,
                                               bool accumulate_first, bool fp16_operands = false) {
    const uint32_t idesc = fp16_operands ? umma_idesc_f16(128, N) : umma_idesc_bf16(128, N);
    // the start-address field (16-byte units) is the only part of the descriptor that moves with the K-step
    const uint64_t da = umma_desc_sw128(a_smem), db = umma_desc_sw128(b_smem);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
        // sub-tile (ks / 4), 32 bytes (16 bf16) per K-step inside the 128-byte swizzled row
        const uint64_t koff = (uint64_t)(((ks >> 2) * 16384u + (ks & 3) * 32u) >> 4);
        umma_bf16(tmem_d, da + koff, db + koff, idesc, (accumulate_first || ks > 0) ? 1u : 0u);
    }
}

}  // namespace kasf
