// tc_ptx.cuh - shared-memory operand layouts and the PTX wrappers (mbarrier, bulk copy, tcgen05, TMEM) shared by the
// tensor-core kernels (k_mlp_tc.cu, k_wave_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phn {

// ------------------------------------------------------------------------------------------------
// layout helpers (shared by host-side image builders and device writers)
// ------------------------------------------------------------------------------------------------
constexpr int TC_M = 128;          // frames per tile
constexpr int TC_NC = 128;         // hidden units per chunk
constexpr int TC_KB = 64;          // fp16 elements per 128-byte swizzle row
constexpr int TC_BLK = 128 * 128;  // bytes of one [128 rows x 64 fp16] block

// byte offset of element (row r < rows, column cc < 64) inside a K-major SW128 block
__host__ __device__ __forceinline__ uint32_t sw128_off(int r, int cc)
{
    return (uint32_t)r * 128u + ((((uint32_t)cc >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)cc & 7u) * 2u;
}

// The merger's activation image is M-major ("MN-major" A operand): its writers are the band nets' epilogues, whose
// threads own one tile ROW each (TMEM lane = row), so a warp store of one COLUMN must land on consecutive bytes.
// Canonical UMMA MN-major SWIZZLE_128B layout, per tile: [k / 8][row / 64][k % 8][128 B = 64 rows], 16-byte chunk
// index XOR k % 8; a 64-column block is 16 KB like the K-major one, one k-step of 16 is 4 KB.
//   descriptor: LBO = 1024 B (next 64 rows), SBO = 2048 B (next 8 columns)
__host__ __device__ __forceinline__ uint32_t mn128_off(int r, int k)
{
    return ((uint32_t)k >> 3) * 2048u + ((uint32_t)r >> 6) * 1024u + ((uint32_t)k & 7u) * 128u +
           (((((uint32_t)r & 63u) >> 3) ^ ((uint32_t)k & 7u)) << 4) + ((uint32_t)r & 7u) * 2u;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Warp-uniform wait: the loop condition is a vote, so the compiler sees uniform control flow around it and may keep
// warp-uniform values (descriptors, ring positions) in uniform registers across the wait.
__device__ __forceinline__ void mbar_wait_u(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!__all_sync(0xffffffffu, ok));
}
__device__ __forceinline__ uint32_t mbar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    // test_wait, not try_wait: try_wait may suspend the thread for a hardware time-out when the phase is still open
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_u(uint32_t bar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T   (both operands K-major)
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: 128 lanes x K/2 32-bit columns holding fp16 pairs; B K-major)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Issue-slot-lean forms for the steady-state path: descriptors travel as their low words (start address field +
// LBO) plus one shared high word, the accumulate flag is a compile-time constant.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
template <int ACC>
__device__ __forceinline__ void umma_ss_lo(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(kDescHi), "r"(idesc), "n"(ACC) : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(kDescHi), "r"(idesc), "n"(ACC) : "memory");
}
// ---- CTA-pair forms (cta_group::2, M = 256: 128 rows per CTA, every B tile split across the two CTAs)
template <int ACC>
__device__ __forceinline__ void umma2_ss_lo(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(kDescHi), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma2_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t blo, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "r"(blo), "r"(kDescHi), "r"(idesc), "n"(ACC) : "memory");
}
// completion of this thread's MMAs -> the mbarrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2_u(uint32_t bar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar_saddr), "h"((uint16_t)3) : "memory");
}
// arrive on the mbarrier at this shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)   // (a barrier other CTAs arrive on; same wait as CUTLASS' ClusterBarrier)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor: start address, LBO (unused for swizzled
// K-major) = 1, SBO = 1024 B between 8-row groups, descriptor version 1 (sm_100), layout type 2.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// M-major ("MN-major") SWIZZLE_128B A operand (mn128_off): LBO = 1024 B, SBO = 2048 B
__device__ __forceinline__ uint64_t make_mn128_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 16;
    d |= (uint64_t)(2048u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: fp16 x fp16 -> fp32, B K-major, A K-major or M-major (bit 15), M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn = false, int m = TC_M)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
#define PHN_TMEM_LD(NAME, SHAPE, N, OUTS, ...)                                                        \
    __device__ __forceinline__ void NAME(uint32_t taddr, uint32_t *v)                                  \
    {                                                                                                 \
        asm volatile("tcgen05.ld.sync.aligned.32x32b." SHAPE ".b32 {" OUTS "}, [%" #N "];" : __VA_ARGS__ : "r"(taddr)); \
    }
#define R4(b) "=r"(v[b]), "=r"(v[b + 1]), "=r"(v[b + 2]), "=r"(v[b + 3])
PHN_TMEM_LD(tmem_ld4, "x4", 4, "%0, %1, %2, %3", R4(0))
PHN_TMEM_LD(tmem_ld8, "x8", 8, "%0, %1, %2, %3, %4, %5, %6, %7", R4(0), R4(4))
PHN_TMEM_LD(tmem_ld16, "x16", 16, "%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15", R4(0), R4(4), R4(8), R4(12))
PHN_TMEM_LD(tmem_ld32, "x32", 32,
            "%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31",
            R4(0), R4(4), R4(8), R4(12), R4(16), R4(20), R4(24), R4(28))
#undef R4
#undef PHN_TMEM_LD
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c)
{
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// One lane of a converged warp (the same one every time: the lowest); tcgen05.commit tracks the MMAs
// of the thread that executes it, so the issuer's MMAs and commits must come from one elected lane.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

}  // namespace phn
