// k_mlp_tc.cu — K-mlp, tensor-core mode: one fused kernel per net,
//     P = fsoftmax(b2 + fsig(b1 + X W1^T) W2^T)
// on the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM),
// operands streamed into shared memory by the TMA engine (cp.async.bulk + mbarrier), the hidden
// layer never leaving the SM (TMEM accumulator -> registers -> TMEM operand).  Replaces NeuralNet::Forward (nn.cpp:872-950), fexp_sigmoid /
// fexp_softmax_v (fexp.h:33-78) and Traps::CalcInputFeaturesForMerger (traps.cpp:435-461) in
// reduced precision; the measured deviation from the exact mode is stated in DESIGN.md.
//
// Tiling: a persistent CTA (one per SM) owns 128-frame tiles (UMMA M = 128, cta_group::1) and walks the hidden
// layer in chunks of 128 units; over the CTA's linear chunk sequence g:
//     G1(g): D1[g&1] (TMEM, 128 cols)  = X[128 x K1] . W1[c]^T    K1/16 SS MMAs 128x128x16 (A, B in shared memory);
//                                        b1 rides along as two extra K columns (fp16 hi + lo) against X columns = 1
//     E1(g): H = fp16(fsig(D1))        -> TMEM (64 cols, fp16 pairs)             16 epilogue warps
//     G2(g): D2 (TMEM, N2P cols)      += H[128 x 128] . W2[:, c]^T  8 TS MMAs 128xN2Px16 (A = H from TMEM)
//     E2   : softmax over D2 + b2 once per tile, then posteriors + ln(posteriors) (merger) or
//            ln + merger input normalisation -> fp16 merger X image (band nets)
// Warp roles: warps 0..15 = epilogue (warp&3 = TMEM lane quarter, warp>>2 = column quarter), warp 16 = TMA
// producer, warp 17 = MMA issuer + TMEM allocator.  What shaped the schedule (measured with tools/umma_bench.cu
// and tools/tc_timeline.py, numbers in DESIGN.md):
//   * an SS MMA costs ~38 + N/2 clk, a TS MMA ~9 + N/2; interleaved SS/TS streams from ONE thread run at ~N/2,
//     two issuing threads serialise -> one issuer, burst(g) = G2(g) interleaved block-wise with G1(g+2),
//     software pipelined across chunks and tiles; E1(g+1) runs on the epilogue warps meanwhile;
//   * the issuer is instruction-fetch bound (it shares an SMSP and its L0 I-cache with four epilogue warps):
//     the steady-state burst is a lean single-instance path with chunk-level "full" barriers (3 waits per burst);
//   * the producer is a non-blocking two-cursor state machine (W1 + X stream, W2 stream), rings sized so that
//     the next chunk of both streams is resident while the current one is multiplied.
//
// All operands live in global memory as ready-made shared-memory images: 16 KB blocks of
// [128 rows x 64 fp16] in the canonical K-major SWIZZLE_128B layout (16-byte chunk index XOR
// row%8), so one 1-D bulk copy lands a block exactly as the UMMA descriptor expects it.  The
// weight images are built once (mlp_tc_prepare); K-stc and the band nets' E2 write activations
// straight into that layout.
#include "internal.h"
#include "device_math.cuh"

#include <cfloat>
#include <cstdlib>

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void umma_ss_lo(uint32_t tmem_d, uint32_t alo, uint32_t blo, uint32_t idesc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(alo), "r"(blo), "r"(kDescHi), "r"(idesc), "n"(ACC) : "memory");
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
// instruction descriptor: fp16 x fp16 -> fp32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n)
{
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
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

// The Quicknet bit-trick exponential (fexp.h:14-21) without a float->int conversion (F2I shares the
// quarter-rate XU pipe with MUFU.RCP, and the epilogue must keep pace with the tensor pipe):
//   D(y) = float whose exponent field is floor(t) and whose mantissa is frac(t),  t = y/ln2 + Ct,
//   Ct = 127 - 60801/2^20.  The caller forms u = sat(t / TMAX) with one FFMA.SAT (the saturation is
//   the clamp: t < 0 gives D = 0, t > TMAX gives D = 2^(TMAX-127)); a second FFMA forms
//   r = 2^23 + u * TMAX * 2^14: in [2^23, 2^24) the float's mantissa field IS round(t*2^14), so
//   bits(D) = bits(r) << 9.  t keeps 14 fractional bits (2^-14 relative in D; the reference keeps 20).
constexpr double kLn2 = 0.69314718055994530942;
constexpr double kCt = 127.0 - 60801.0 / 1048576.0;
constexpr float kSigTmax = 158.0f;               // sigmoid: D clamped to 2^31 (1/(1+D) is 0 in fp16 long before)
constexpr float kSmxTmax = 128.0f;               // softmax: y <= 0, so t <= Ct < 128
__device__ __forceinline__ float fexp_from_u(float u, float tmax)
{
    const float r = fmaf(u, tmax * 16384.0f, 8388608.0f);
    return __uint_as_float(__float_as_uint(r) << 9);
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
    const uint8_t *x_img;     // [tiles][KB1][16 KB]  activations, SW128 blocks
    const uint8_t *w1_img;    // [NCH][KB1][16 KB]
    const uint8_t *w2_img;    // [NCH][2][N2P*128 B]
    const float *b2;          // [N2P]      output bias; -FLT_MAX in the padding columns
    int n_tiles, KB1, NCH, S1, S2;
    int nks_last;             // k-steps (of 16) actually needed in the last k-block of layer 1
    int64_t nf;               // frames in this launch
    int nout;
    // outputs
    float *post; int ldpost;                          // merger: posteriors [nf][ldpost]
    float *logp;                                      // merger: ln(posteriors) [nf][ldpost] for the decoder, or nullptr
    uint8_t *xm_img; int xm_kb1; int xm_col0;         // band nets: merger input image, first column (multiple of 8)
    long long *dbg;                                   // optional timeline of CTA 0's second tile (tools/tc_timeline.py)
    int xm_bias;                                      // band 1: its columns nout, nout+1 are the merger's constant-1 bias inputs
    const float *mmean, *mdev;                        // merger input normalisation, indexed by image column
};

// debug timeline: dbg[(c * 16 + event)] = clock64() for chunk c of CTA 0's second tile
#define TC_DBG(ev, c)                                                                       \
    do {                                                                                    \
        if (a.dbg && blockIdx.x == 0 && tile == (int)gridDim.x) a.dbg[(c) * 16 + (ev)] = clock64(); \
    } while (0)
constexpr int TC_MAXS1 = 12;                        // upper bound on W1 ring stages (barrier array size)
constexpr int TC_EPI_WARPS = 16;                     // warps 0..15: epilogue; 16: TMA producer; 17: MMA issuer
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
// The SM's warp schedulers favour the highest warp id among eligible warps, so the two warps whose
// instruction streams gate everything else (TMA producer, MMA issuer) get the highest ids.
__device__ __forceinline__ void quarter_bar_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

template <int N2P>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_tc(TcArgs a)
{
    constexpr int W2_BLK = N2P * 128;       // bytes of one [N2P rows x 64 fp16] block
    constexpr int NQ = N2P / 4;             // D2 columns per epilogue column quarter (32, 36, 40 or 48)
    constexpr int NR = NQ - 32;             // columns beyond the first 32-column TMEM load (0, 4, 8 or 16)
    static_assert(NR == 0 || NR == 4 || NR == 8 || NR == 16, "unsupported output width");
    extern __shared__ uint8_t smem_raw[];
    // carve-up (all block bases 1024-byte aligned: the swizzle pattern is a function of the address)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sX = smem;                                        // KB1 x 16 KB
    uint8_t *sW1 = sX + (size_t)a.KB1 * TC_BLK;                // S1 x 16 KB ring
    uint8_t *sW2 = sW1 + (size_t)a.S1 * TC_BLK;                // S2 x W2_BLK ring
    float *s_b2 = reinterpret_cast<float *>(sW2 + (size_t)a.S2 * W2_BLK);   // [N2P]
    float *s_mm = s_b2 + N2P;                                  // [N2P] merger input mean  (band nets)
    float *s_md = s_mm + N2P;                                  // [N2P] merger input 1/std (band nets)
    float *s_red = s_md + N2P;                                 // [2][4][128] row max / row sum exchange
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_red + 8 * 128);
    uint64_t *x_full = bars;                 // [8]
    uint64_t *x_empty = bars + 8;            // [1]
    uint64_t *w1_full = bars + 9;            // [TC_MAXS1]
    uint64_t *w1_empty = w1_full + TC_MAXS1; // [TC_MAXS1]
    uint64_t *w2_empty = w1_empty + TC_MAXS1; // [4]
    uint64_t *w1c_full = w2_empty + 4;       // [4] chunk-level: all KB1 k-blocks of layer-1 chunk n have landed (n & 3)
    uint64_t *w2c_full = w1c_full + 4;       // [2] chunk-level: both k-blocks of layer-2 chunk n have landed (n & 1)
    uint64_t *d1_full = w2c_full + 2;        // [2]
    uint64_t *h_full = d1_full + 2, *h_empty = h_full + 1;
    uint64_t *d2_full = h_empty + 1, *d2_empty = d2_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d2_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARP_TMA = TC_EPI_WARPS, WARP_MMA = TC_EPI_WARPS + 1, EPI0 = 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&x_full[i], 1);
        for (int i = 0; i < TC_MAXS1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&w2_empty[i], 1); mbar_init(&w1c_full[i], a.KB1); }
        for (int i = 0; i < 2; ++i) mbar_init(&w2c_full[i], 2);
        mbar_init(x_empty, 1);
        for (int i = 0; i < 2; ++i) mbar_init(&d1_full[i], 1);
        mbar_init(h_full, TC_EPI_WARPS); mbar_init(h_empty, 1);
        mbar_init(d2_full, 1); mbar_init(d2_empty, TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {  // TMEM: 512 columns = D1 double buffer 2 x 128 | D2 up to 192 | H 64 (128 fp16 per lane)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < N2P; i += blockDim.x) {
        s_b2[i] = a.b2[i];
        s_mm[i] = a.mmean ? a.mmean[a.xm_col0 + i] : 0.0f;
        s_md[i] = a.mdev ? a.mdev[a.xm_col0 + i] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tD1[2] = {tmem, tmem + 128u};
    const uint32_t tD2 = tmem + 256u;
    const uint32_t tH = tmem + 448u;

    if (warp == WARP_TMA) {
        // ===================================================================== TMA producer
        // The whole warp walks the loops (warp-uniform control flow); one elected lane issues.  Loads follow the
        // issuer's consumption order over the CTA's linear chunk sequence g = tile_iter * NCH + c:
        //   X(tile 0), W1(0), W1(1), then per g:  [X(next tile) if chunk g+2 opens it]  W1(g+2)  W2(g)
        uint32_t ph_x_empty = 0, w1_stage = 0, ph_w1 = 0, w2_stage = 0, ph_w2 = 0;
        const int my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int G = my_tiles * a.NCH;
        // Normal case (the W1 ring holds a whole chunk): every k-block load of a chunk reports to the chunk's own
        // "full" barrier, so the issuer waits once per chunk; ring stages are still handed back one by one.
        // Two independent cursors (layer-1 weights + X, layer-2 weights), each advanced whenever its next ring stage
        // is free: a blocking in-order producer would sit on the shallow W2 ring and never prefetch W1 ahead.
        const bool chunk_bar = a.S1 >= a.KB1;
        uint32_t n1 = 0, n2 = 0;          // chunks fully issued per stream
        int kb1 = 0, kb2 = 0;             // next k-block inside the current chunk
        int c1 = 0, c2 = 0, tile1 = blockIdx.x;
        bool x_done = false, first_x = true;
        while (n1 < (uint32_t)G || n2 < (uint32_t)G) {
            if (n1 < (uint32_t)G) {
                if (c1 == 0 && kb1 == 0 && !x_done) {   // the chunk opens a tile: its X first
                    if (first_x || mbar_try(x_empty, ph_x_empty)) {
                        if (!first_x) ph_x_empty ^= 1;
                        first_x = false;
                        if (elect_one()) {
                            for (int kb = 0; kb < a.KB1; ++kb) {
                                mbar_expect_tx(&x_full[kb], TC_BLK);
                                tma_load_1d(sX + (size_t)kb * TC_BLK, a.x_img + ((size_t)tile1 * a.KB1 + kb) * TC_BLK, TC_BLK, &x_full[kb]);
                            }
                        }
                        __syncwarp();
                        x_done = true;
                    }
                } else if (mbar_try(&w1_empty[w1_stage], ph_w1 ^ 1)) {
                    if (elect_one()) {
                        uint64_t *fb = chunk_bar ? &w1c_full[n1 & 3] : &w1_full[w1_stage];
                        mbar_expect_tx(fb, TC_BLK);
                        tma_load_1d(sW1 + (size_t)w1_stage * TC_BLK, a.w1_img + ((size_t)c1 * a.KB1 + kb1) * TC_BLK, TC_BLK, fb);
                    }
                    __syncwarp();
                    if (++w1_stage == (uint32_t)a.S1) { w1_stage = 0; ph_w1 ^= 1; }
                    if (++kb1 == a.KB1) {
                        kb1 = 0; ++n1; x_done = false;
                        if (++c1 == a.NCH) { c1 = 0; tile1 += gridDim.x; }
                    }
                }
            }
            if (n2 < (uint32_t)G && mbar_try(&w2_empty[w2_stage], ph_w2 ^ 1)) {
                if (elect_one()) {
                    mbar_expect_tx(&w2c_full[n2 & 1], W2_BLK);
                    tma_load_1d(sW2 + (size_t)w2_stage * W2_BLK, a.w2_img + ((size_t)c2 * 2 + kb2) * W2_BLK, W2_BLK, &w2c_full[n2 & 1]);
                }
                __syncwarp();
                if (++w2_stage == (uint32_t)a.S2) { w2_stage = 0; ph_w2 ^= 1; }
                if (++kb2 == 2) {
                    kb2 = 0; ++n2;
                    if (++c2 == a.NCH) c2 = 0;
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ===================================================================== MMA issuer
        // One thread issues every MMA of the CTA (two issuing threads serialise in the tensor pipe; measured with
        // tools/umma_bench.cu).  Layer 1 reads both operands from shared memory (SS, ~38 + N/2 clk each when run
        // alone), layer 2 takes H from TMEM (TS, ~9 + N/2); interleaved one to one the pair runs at the nominal
        // N/2 per MMA because the SS operand fetch overlaps the TS math.  So the issue schedule over the CTA's
        // linear chunk sequence g is software pipelined:   burst(g) = G2(g) interleaved with G1(g+2),
        // issued as soon as E1(g) has published H(g) - which also means D1[g & 1] has been read and is free for
        // G1(g+2).  E1(g+1) (whose D1 came from the previous burst) runs on the epilogue warps meanwhile.
        const uint32_t idesc1 = make_idesc(TC_NC), idesc2 = make_idesc(N2P);
        uint32_t ph_x_full = 0, w1_stage = 0, ph_w1 = 0, w2_stage = 0, ph_h_full = 0, ph_d2_empty = 0;
        const uint64_t dX = make_sw128_desc(smem_u32(sX)), dW1 = make_sw128_desc(smem_u32(sW1));
        const uint64_t dW2 = make_sw128_desc(smem_u32(sW2));
        const int my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int G = my_tiles * a.NCH;
        // A chunk's layer-1 k-blocks are issued in groups of GK blocks: the whole chunk when the W1 ring can hold
        // it (S1 >= KB1, the normal case), else half a ring at a time (wide merger nets next to a wide X tile).
        const int GK = a.S1 >= a.KB1 ? a.KB1 : (a.S1 / 2 > 0 ? a.S1 / 2 : 1);
        // burst: layer 2 of chunk g2 (g2 < 0: none) interleaved one to one with layer 1 of chunk g1 (g1 < 0: none)
        // (c2 / c1: the chunks' positions inside their tiles; first2: no D2 has been handed to the epilogue yet)
        const bool chunk_bar = a.S1 >= a.KB1;
        uint32_t n1 = 0, n2 = 0;   // running chunk counters of the two weight streams (chunk-level "full" barriers)
        int c2 = 0, c1 = 0, d1buf = 0;
        bool first2 = true;
        int tile = blockIdx.x;   // (only for the debug timeline)
        auto burst = [&](bool do_g2, bool do_g1) {
            const int g1 = do_g1 ? d1buf : -1;   // only the D1 buffer index of the layer-1 chunk matters below
            uint32_t w2b = w2_stage + 1;
            if (w2b == (uint32_t)a.S2) w2b = 0;
            const bool opens = do_g1 && c1 == 0;
            const bool wait_d2 = do_g2 && c2 == 0 && !first2;
            bool g2_pending = do_g2;
            for (int kb0 = 0; kb0 < (g1 >= 0 ? a.KB1 : 1); kb0 += GK) {
                const int nkb = g1 >= 0 ? (a.KB1 - kb0 < GK ? a.KB1 - kb0 : GK) : 0;
                // operands of this group: weights first (loaded long ago in the steady state), H last (the freshest)
                if (nkb > 0) {
                    if (chunk_bar) {
                        mbar_wait(&w1c_full[n1 & 3], (n1 >> 2) & 1);
                    } else {
                        uint32_t st = w1_stage, ph = ph_w1;
                        for (int k = 0; k < nkb; ++k) {
                            mbar_wait(&w1_full[st], ph);
                            if (++st == (uint32_t)a.S1) { st = 0; ph ^= 1; }
                        }
                    }
                    if (opens)
                        for (int k = 0; k < nkb; ++k) mbar_wait(&x_full[kb0 + k], ph_x_full);
                }
                if (g2_pending) {
                    mbar_wait(&w2c_full[n2 & 1], (n2 >> 1) & 1);
                    if (wait_d2) mbar_wait(d2_empty, ph_d2_empty);
                    mbar_wait(h_full, ph_h_full);
                }
                tc_fence_after();
                if (elect_one()) {
                    // interleave in blocks of four k-steps (one k-block): SS block of layer 1, TS block of layer 2, ...
                    const int nblk = g2_pending ? (nkb > 2 ? nkb : 2) : nkb;
                    uint32_t st = w1_stage;
                    for (int k = 0; k < nblk; ++k) {
                        if (k < nkb) {
                            const int kb = kb0 + k;
                            const int nks = kb == a.KB1 - 1 ? a.nks_last : 4;
                            const uint64_t ad = dX + (uint64_t)((kb * TC_BLK) >> 4), bd = dW1 + (uint64_t)((st * TC_BLK) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (ks < nks) umma_f16_ss(tD1[g1 & 1], ad + 2 * ks, bd + 2 * ks, idesc1, (kb | ks) ? 1u : 0u);
                            tc_commit(&w1_empty[st]);
                            if (++st == (uint32_t)a.S1) st = 0;
                        }
                        if (g2_pending && k < 2) {   // A = H[:, 64 k + 16 ks .. +15] = 8 TMEM columns per k-step
                            const uint32_t w2s = k ? w2b : w2_stage;
                            const uint64_t bd = dW2 + (uint64_t)((w2s * W2_BLK) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_f16_ts(tD2, tH + (uint32_t)(k * 4 + ks) * 8u, bd + 2 * ks, idesc2, (c2 | k | ks) ? 1u : 0u);
                            tc_commit(&w2_empty[w2s]);
                        }
                    }
                    if (g2_pending) {
                        tc_commit(h_empty);
                        if (c2 == a.NCH - 1) tc_commit(d2_full);
                    }
                    if (g1 >= 0 && kb0 + nkb == a.KB1) {
                        tc_commit(&d1_full[g1 & 1]);
                        if (c1 == a.NCH - 1) tc_commit(x_empty);
                    }
                }
                __syncwarp();
                w1_stage += nkb;
                if (w1_stage >= (uint32_t)a.S1) { w1_stage -= a.S1; ph_w1 ^= 1; }
                if (g2_pending) {
                    ph_h_full ^= 1;
                    if (wait_d2) ph_d2_empty ^= 1;
                    w2_stage += 2;
                    if (w2_stage >= (uint32_t)a.S2) w2_stage -= a.S2;
                    g2_pending = false;
                }
            }
            if (opens) ph_x_full ^= 1;
            if (do_g1) { ++n1; d1buf ^= 1; if (++c1 == a.NCH) c1 = 0; }
            if (do_g2) { ++n2; first2 = false; if (++c2 == a.NCH) c2 = 0; }
        };
        // Steady state (both chunks inside their tiles, W1 chunk resident): the same burst with nothing but the
        // instructions it needs - this warp shares an SMSP (and its 6 KB L0 instruction cache) with four epilogue
        // warps, and every instruction-fetch miss of the issuer is a bubble in the tensor pipe.
        const uint32_t xlo = (uint32_t)dX, w1lo = (uint32_t)dW1, w2lo = (uint32_t)dW2;
        auto lean = [&]() {
            if (lane == 0) TC_DBG(3, c2);
            mbar_wait(&w1c_full[n1 & 3], (n1 >> 2) & 1);
            mbar_wait(&w2c_full[n2 & 1], (n2 >> 1) & 1);
            mbar_wait(h_full, ph_h_full);
            if (lane == 0) TC_DBG(5, c2);   // operands there
            tc_fence_after();
            uint32_t w2b = w2_stage + 1;
            if (w2b == (uint32_t)a.S2) w2b = 0;
            if (elect_one()) {
                const uint32_t td1 = tD1[d1buf];
                uint32_t st = w1_stage, alo = xlo;
                for (int k = 0; k < a.KB1; ++k) {
                    const uint32_t blo = w1lo + st * (TC_BLK >> 4);
                    if (k == 0) umma_ss_lo<0>(td1, alo, blo, idesc1); else umma_ss_lo<1>(td1, alo, blo, idesc1);
                    if (k < a.KB1 - 1 || a.nks_last > 1) umma_ss_lo<1>(td1, alo + 2, blo + 2, idesc1);
                    if (k < a.KB1 - 1 || a.nks_last > 2) umma_ss_lo<1>(td1, alo + 4, blo + 4, idesc1);
                    if (k < a.KB1 - 1 || a.nks_last > 3) umma_ss_lo<1>(td1, alo + 6, blo + 6, idesc1);
                    tc_commit(&w1_empty[st]);
                    if (++st == (uint32_t)a.S1) st = 0;
                    alo += TC_BLK >> 4;
                    if (k < 2) {   // A = H[:, 64 k + 16 ks .. +15] = 8 TMEM columns per k-step
                        const uint32_t w2s = k ? w2b : w2_stage;
                        const uint32_t b2lo = w2lo + w2s * (W2_BLK >> 4);
                        const uint32_t ta = tH + (uint32_t)k * 32u;
                        umma_ts_lo<1>(tD2, ta, b2lo, idesc2);
                        umma_ts_lo<1>(tD2, ta + 8u, b2lo + 2, idesc2);
                        umma_ts_lo<1>(tD2, ta + 16u, b2lo + 4, idesc2);
                        umma_ts_lo<1>(tD2, ta + 24u, b2lo + 6, idesc2);
                        tc_commit(&w2_empty[w2s]);
                    }
                }
                tc_commit(h_empty);
                if (c2 == a.NCH - 1) tc_commit(d2_full);
                tc_commit(&d1_full[d1buf]);
                if (c1 == a.NCH - 1) tc_commit(x_empty);
                TC_DBG(13, c2);                 // all MMAs and commits issued
            }
            __syncwarp();
            w1_stage += a.KB1;
            if (w1_stage >= (uint32_t)a.S1) { w1_stage -= a.S1; ph_w1 ^= 1; }
            ph_h_full ^= 1;
            w2_stage += 2;
            if (w2_stage >= (uint32_t)a.S2) w2_stage -= a.S2;
            ++n1; d1buf ^= 1; if (++c1 == a.NCH) c1 = 0;
            ++n2; if (++c2 == a.NCH) c2 = 0;
        };
        const bool lean_ok = chunk_bar && a.KB1 >= 2;
        int pro = G > 1 ? 2 : G;   // prologue: layer 1 of the first two chunks on its own
        for (int g = 0; g < G;) {
            bool t_g2[2] = {false, false}, t_g1[2] = {true, true};
            int nt = 1;
            const int c = c2;
            if (pro > 0) {
                --pro;
            } else {
                const bool has_g1 = g + 2 < G;
                if (lane == 0) TC_DBG(2, c);   // burst starts
                if (has_g1 && c1 != 0 && c2 != 0 && lean_ok) {
                    lean();
                    ++g;
                    if (lane == 0) TC_DBG(4, c);   // burst issued
                    if (c2 == 0) tile += gridDim.x;
                    continue;
                }
                if (has_g1 && c1 == 0) {
                    // G1(g+2) opens a tile whose X may still be in flight: do not hold layer 2 of this chunk back for it
                    nt = 2; t_g2[0] = true; t_g1[0] = false;
                } else {
                    t_g2[0] = true; t_g1[0] = has_g1;
                }
                ++g;
            }
            for (int t = 0; t < nt; ++t) burst(t_g2[t], t_g1[t]);   // (one call site: one copy of the general path)
            if (c2 == 0 && t_g2[0]) tile += gridDim.x;
        }
    } else {
        // ===================================================================== epilogue warps
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int cq = (warp - EPI0) >> 2;       // column quarter 0..3
        const int row = q * 32 + lane;           // tile row == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t ph_d1_full0 = 0, ph_d1_full1 = 0, ph_h_empty = 0, ph_d2_full = 0;
        bool first_h = true;
        int g = 0;   // the CTA's linear chunk index: D1 buffer = g & 1 (NCH may be odd)
        // H[row][cq*32 .. +31] (fp16 pairs) = TMEM lane `row`, columns cq*16 .. +15 of the H operand
        // u = sat(t / TMAX), t = -(x + b1)/ln2 + Ct; b1 is already inside x (two extra K columns of layer 1)
        const float sigA = (float)(-1.0 / (kLn2 * (double)kSigTmax)), sigB = (float)(kCt / (double)kSigTmax);
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            for (int c = 0; c < a.NCH; ++c, ++g) {
                const int b = g & 1;
                if (threadIdx.x == EPI0 * 32) TC_DBG(8, c);    // e1: waiting for D1
                if (b == 0) { mbar_wait(&d1_full[0], ph_d1_full0); ph_d1_full0 ^= 1; }
                else        { mbar_wait(&d1_full[1], ph_d1_full1); ph_d1_full1 ^= 1; }
                tc_fence_after();
                if (threadIdx.x == EPI0 * 32) TC_DBG(9, c);    // e1: D1 seen
                // (no "D1 free" signal: the issuer reuses D1[b] only after this chunk's H has been published)
                // fsig(x) = 1 / (1 + D(-x - b1)); two reciprocals share one MUFU: 1/a = a' / (a a'), 1/a' = a / (a a')
                uint32_t hp[16];
                {
                    uint32_t acc[32];
                    tmem_ld32((b ? tD1[1] : tD1[0]) + lane_addr + cq * 32, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float a0 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 0]), sigA, sigB), kSigTmax);
                        const float a1 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 1]), sigA, sigB), kSigTmax);
                        const float a2 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 2]), sigA, sigB), kSigTmax);
                        const float a3 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 3]), sigA, sigB), kSigTmax);
                        const float r01 = rcp_approx(a0 * a1), r23 = rcp_approx(a2 * a3);
                        hp[g4 * 2] = pack_half2(r01 * a1, r01 * a0);
                        hp[g4 * 2 + 1] = pack_half2(r23 * a3, r23 * a2);
                    }
                }
                if (threadIdx.x == EPI0 * 32) TC_DBG(10, c);   // e1: math done
                if (!first_h) { mbar_wait(h_empty, ph_h_empty); ph_h_empty ^= 1; }
                first_h = false;
                if (threadIdx.x == EPI0 * 32) TC_DBG(11, c);   // e1: H buffer free
                tc_fence_after();
                tmem_st16(tH + lane_addr + cq * 16, hp);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_full);
                if (threadIdx.x == EPI0 * 32) TC_DBG(12, c);   // e1: H published
            }
            // ------------------------------------------------------------- E2: softmax + outputs
            // column quarter cq owns D2 columns [cq*NQ, (cq+1)*NQ); the 4 warps of a row quarter
            // exchange row max / row sum through shared memory and their own named barrier
            const int n0 = cq * NQ;
            if (threadIdx.x == EPI0 * 32) TC_DBG(13, 0);       // e2: waiting for D2
            mbar_wait(d2_full, ph_d2_full); ph_d2_full ^= 1;
            tc_fence_after();
            if (threadIdx.x == EPI0 * 32) TC_DBG(14, 0);       // e2: D2 seen
            float o[NQ];
            {
                uint32_t raw[NQ];
                tmem_ld32(tD2 + lane_addr + n0, raw);
                if (NR == 4) tmem_ld4(tD2 + lane_addr + n0 + 32, raw + 32);
                if (NR == 8) tmem_ld8(tD2 + lane_addr + n0 + 32, raw + 32);
                if (NR == 16) tmem_ld16(tD2 + lane_addr + n0 + 32, raw + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(d2_empty);
#pragma unroll
                for (int j = 0; j < NQ / 4; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(s_b2 + n0 + 4 * j);   // -FLT_MAX in padding columns
                    o[4 * j + 0] = __uint_as_float(raw[4 * j + 0]) + b4.x;
                    o[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b4.y;
                    o[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b4.z;
                    o[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b4.w;
                }
            }
            float mx = o[0];
#pragma unroll
            for (int i = 1; i < NQ; ++i) mx = fmaxf(mx, o[i]);
            s_red[cq * 128 + row] = mx;
            quarter_bar_sync(q);
            mx = fmaxf(fmaxf(s_red[row], s_red[128 + row]), fmaxf(s_red[256 + row], s_red[384 + row]));
            float sum = 0.0f;
            const float smxA = (float)(1.0 / (kLn2 * (double)kSmxTmax)), smxB = (float)(kCt / (double)kSmxTmax);
#pragma unroll
            for (int i = 0; i < NQ; ++i) {     // fexp_softmax_v (fexp.h:49-78): e = D(o - max)
                const float e = fexp_from_u(fma_sat(o[i] - mx, smxA, smxB), kSmxTmax);
                o[i] = e;
                sum += e;
            }
            s_red[512 + cq * 128 + row] = sum;
            quarter_bar_sync(q);
            sum = (s_red[512 + row] + s_red[512 + 128 + row]) + (s_red[512 + 256 + row] + s_red[512 + 384 + row]);
            const float sc = 1.0f / sum;
            const int64_t f = (int64_t)tile * TC_M + row;
            if (f < a.nf) {
                if (a.post) {
                    float *dst = a.post + f * a.ldpost + n0;
#pragma unroll
                    for (int j = 0; j < NQ / 4; ++j)
                        if (n0 + 4 * j < a.ldpost)
                            *reinterpret_cast<float4 *>(dst + 4 * j) = make_float4(o[4 * j] * sc, o[4 * j + 1] * sc, o[4 * j + 2] * sc, o[4 * j + 3] * sc);
                    if (a.logp) {   // decoder soft function (srec.cpp:1088-1097), fused: ln p for the token passing kernel
                        float *ldst = a.logp + f * a.ldpost + n0;
#pragma unroll
                        for (int j = 0; j < NQ / 4; ++j)
                            if (n0 + 4 * j < a.ldpost)
                                *reinterpret_cast<float4 *>(ldst + 4 * j) = make_float4(__logf(o[4 * j] * sc), __logf(o[4 * j + 1] * sc),
                                                                                       __logf(o[4 * j + 2] * sc), __logf(o[4 * j + 3] * sc));
                    }
                } else {
                    // merger input: sLn(p), merger input normalisation, fp16, 8-byte pieces of the merger's X image
                    const int nlim = (a.nout + 7) & ~7;   // this net's share of the image: nout rounded up to 8 columns
#pragma unroll
                    for (int j = 0; j < NQ / 4; ++j) {
                        const int n = n0 + 4 * j;
                        if (n < nlim) {
                            const float4 m4 = *reinterpret_cast<const float4 *>(s_mm + n), d4 = *reinterpret_cast<const float4 *>(s_md + n);
                            const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, md[4] = {d4.x, d4.y, d4.z, d4.w};
                            float xn[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float p = o[4 * j + i] * sc;
                                const float v = p > 0.0f ? __logf(p) : 0.0f;
                                xn[i] = n + i < a.nout ? (v - mm[i]) * md[i] : ((a.xm_bias && n + i < a.nout + 2) ? 1.0f : 0.0f);
                            }
                            const int cm = a.xm_col0 + n;
                            uint8_t *blk = a.xm_img + ((size_t)tile * a.xm_kb1 + (cm >> 6)) * TC_BLK;
                            *reinterpret_cast<uint2 *>(blk + (size_t)row * 128 + ((((cm >> 3) & 7) ^ (row & 7)) << 4) + ((cm & 4) << 1)) =
                                make_uint2(pack_half2(xn[0], xn[1]), pack_half2(xn[2], xn[3]));
                        }
                    }
                }
            }
            if (threadIdx.x == EPI0 * 32) TC_DBG(15, 0);       // e2 done
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// weight images (built once per context)
// ------------------------------------------------------------------------------------------------
struct TcNetImages {
    uint8_t *w1_img = nullptr, *w2_img = nullptr;
    float *b2 = nullptr;
    float *mean_img = nullptr, *dev_img = nullptr;   // merger only: input normalisation in image column order
    int KB1 = 0, NCH = 0, N2P = 0, nks_last = 4, kin = 0;
};

// Image column k of the merger's layer 1 <- network input: the band-1 half starts at `split8`
// (= band outputs rounded up to 8) so that both band nets write whole 16-byte chunks.
__host__ __device__ __forceinline__ int img_col_to_input(int k, int split, int split8)
{
    if (split <= 0) return k;                 // band nets: identity
    if (k < split) return k;
    if (k < split8) return -1;
    return k - split8 + split;
}

// Layer-1 weight image.  Image columns kin, kin+1 carry the bias: fp16(b1) and fp16(b1 - fp16(b1)); the
// activations hold 1.0 in both, so the tensor core adds b1 with ~2^-22 relative error.
__global__ void k_build_w1_img(const float *__restrict__ w1, const float *__restrict__ b1, int nin, int nhid, int nin4,
                               uint8_t *img, int KB1, int NCH, int split, int split8, int kin)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)NCH * 128 * KB1 * 64;
    if (idx >= total) return;
    const int k = (int)(idx % (KB1 * 64));
    const int n = (int)(idx / (KB1 * 64));
    const int ki = img_col_to_input(k, split, split8);
    float v = (n < nhid && k < kin && ki >= 0 && ki < nin) ? w1[(int64_t)n * nin4 + ki] : 0.0f;
    if (n < nhid && k == kin) v = b1[n];
    if (n < nhid && k == kin + 1) v = b1[n] - __half2float(__float2half_rn(b1[n]));
    const int c = n >> 7, r = n & 127, kb = k >> 6, cc = k & 63;
    *reinterpret_cast<__half *>(img + ((size_t)c * KB1 + kb) * TC_BLK + sw128_off(r, cc)) = __float2half_rn(v);
}

// Constant-1 bias inputs of an activation image (columns kin, kin+1 of every row of every tile).
__global__ void k_fill_bias_cols(uint8_t *img, int64_t rows, int KB1, int kin)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * 2) return;
    const int64_t row = idx >> 1;
    const int k = kin + (int)(idx & 1);
    *reinterpret_cast<__half *>(img + ((size_t)(row >> 7) * KB1 + (k >> 6)) * TC_BLK + sw128_off((int)(row & 127), k & 63)) = __float2half_rn(1.0f);
}

__global__ void k_build_w2_img(const float *__restrict__ w2, int nhid, int nout, int nhid4, uint8_t *img, int N2P, int NCH)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)N2P * NCH * 128;
    if (idx >= total) return;
    const int k = (int)(idx % (NCH * 128));
    const int n = (int)(idx / (NCH * 128));
    const float v = (n < nout && k < nhid) ? w2[(int64_t)n * nhid4 + k] : 0.0f;
    const int c = k >> 7, kb = (k >> 6) & 1, cc = k & 63;
    *reinterpret_cast<__half *>(img + ((size_t)c * 2 + kb) * (N2P * 128) + sw128_off(n, cc)) = __float2half_rn(v);
}

__global__ void k_build_mnorm(const float *__restrict__ mean, const float *__restrict__ dev, int nin, float *mean_img,
                              float *dev_img, int ncols, int split, int split8)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncols) return;
    const int ki = img_col_to_input(k, split, split8);
    mean_img[k] = (ki >= 0 && ki < nin) ? mean[ki] : 0.0f;
    dev_img[k] = (ki >= 0 && ki < nin) ? dev[ki] : 0.0f;
}

__global__ void k_build_bias(const float *__restrict__ b2, int nout, float *b2p, int N2P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N2P) b2p[i] = i < nout ? b2[i] : -FLT_MAX;   // padding columns drop out of the softmax (D(-huge) = 0)
}

struct TcState {
    TcNetImages net[3];
    bool ready = false;
};

int mlp_tc_prepare(phn_ctx *c)
{
    if (!c->tc) c->tc = new TcState();
    TcState &st = *static_cast<TcState *>(c->tc);
    if (st.ready) return PHN_OK;
    for (int i = 0; i < 3; ++i) {
        DevNet &n = c->net[i];
        TcNetImages &im = st.net[i];
        const int split = i == 2 ? c->net[0].nout : 0, split8 = (split + 7) / 8 * 8;
        im.kin = i == 2 ? split8 + split : n.nin;   // image columns that carry data; two bias columns follow
        im.KB1 = (im.kin + 2 + 63) / 64;
        im.NCH = (n.nhid + 127) / 128;
        im.N2P = (n.nout + 15) / 16 * 16;
        const int rem = im.kin + 2 - (im.KB1 - 1) * 64;
        im.nks_last = (rem + 15) / 16;
        if (im.KB1 > 8) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: more than 512 network inputs\n");
        if (im.N2P != 128 && im.N2P != 144 && im.N2P != 160 && im.N2P != 192)
            return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: %d network outputs not instantiated\n", n.nout);
        const size_t w1b = (size_t)im.NCH * im.KB1 * TC_BLK, w2b = (size_t)im.NCH * 2 * im.N2P * 128;
        PHN_CUDA(c, cudaMalloc((void **)&im.w1_img, w1b));
        PHN_CUDA(c, cudaMalloc((void **)&im.w2_img, w2b));
        PHN_CUDA(c, cudaMalloc((void **)&im.b2, sizeof(float) * im.N2P));
        const int64_t t1 = (int64_t)im.NCH * 128 * im.KB1 * 64, t2 = (int64_t)im.N2P * im.NCH * 128;
        k_build_w1_img<<<(unsigned)((t1 + 255) / 256), 256, 0, c->stream>>>(n.w1, n.b1, n.nin, n.nhid, n.nin4, im.w1_img, im.KB1, im.NCH, split, split8, im.kin);
        if (i == 2) {
            // + 16: a band net reads its N2P (= outputs rounded up to 16) columns starting at its first image column
            const int ncol = im.KB1 * 64 + 16;
            PHN_CUDA(c, cudaMalloc((void **)&im.mean_img, sizeof(float) * ncol));
            PHN_CUDA(c, cudaMalloc((void **)&im.dev_img, sizeof(float) * ncol));
            k_build_mnorm<<<(ncol + 127) / 128, 128, 0, c->stream>>>(n.mean, n.dev, n.nin, im.mean_img, im.dev_img, ncol, split, split8);
        }
        k_build_w2_img<<<(unsigned)((t2 + 255) / 256), 256, 0, c->stream>>>(n.w2, n.nhid, n.nout, n.nhid4, im.w2_img, im.N2P, im.NCH);
        k_build_bias<<<(im.N2P + 255) / 256, 256, 0, c->stream>>>(n.b2, n.nout, im.b2, im.N2P);
        PHN_CUDA(c, cudaGetLastError());
        n.w1h = reinterpret_cast<__half *>(im.w1_img);  // owned by the context from here on (freed in phn_destroy)
        n.w2h = reinterpret_cast<__half *>(im.w2_img);
        n.nhidP = im.NCH * 128;
        n.noutP = im.N2P;
    }
    st.ready = true;
    return PHN_OK;
}

void mlp_tc_release(phn_ctx *c)
{
    if (!c->tc) return;
    TcState *st = static_cast<TcState *>(c->tc);
    for (auto &im : st->net) {
        if (im.b2) cudaFree(im.b2);
        if (im.mean_img) cudaFree(im.mean_img);
        if (im.dev_img) cudaFree(im.dev_img);
    }
    delete st;
    c->tc = nullptr;
}

template <int N2P>
static int launch_one(phn_ctx *c, const TcArgs &a, size_t smem_bytes, int grid)
{
    PHN_CUDA(c, cudaFuncSetAttribute(k_mlp_tc<N2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    k_mlp_tc<N2P><<<grid, TC_THREADS, smem_bytes, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

static int run_net_tc(phn_ctx *c, int which, const uint8_t *x_img, int64_t nf, int64_t f0)
{
    TcState &st = *static_cast<TcState *>(c->tc);
    const TcNetImages &im = st.net[which];
    const DevNet &n = c->net[which];
    TcArgs a{};
    a.x_img = x_img; a.w1_img = im.w1_img; a.w2_img = im.w2_img; a.b2 = im.b2;
    a.n_tiles = (int)((nf + TC_M - 1) / TC_M);
    a.dbg = (which == c->tc_dbg_net) ? (long long *)c->tc_dbg : nullptr;
    a.KB1 = im.KB1; a.NCH = im.NCH; a.nks_last = im.nks_last; a.nf = nf; a.nout = n.nout;
    if (which < 2) {
        a.post = nullptr;
        a.xm_img = (uint8_t *)c->d_xmh.p; a.xm_kb1 = st.net[2].KB1; a.xm_col0 = which * ((n.nout + 7) / 8 * 8);
        a.mmean = st.net[2].mean_img; a.mdev = st.net[2].dev_img;
        a.xm_bias = which == 1;
    } else {
        a.post = (float *)c->d_post.p + f0 * c->ldp; a.ldpost = c->ldp;
        a.logp = c->fuse_logp ? (float *)c->d_logp.p + f0 * c->ldp : nullptr;
    }
    // shared memory plan: X (KB1 blocks) + H (2 blocks) + constants + barriers are fixed; the rest is split
    // between the W2 ring (S2 k-blocks of N2P x 64) and the W1 ring (S1 blocks of 128 x 64)
    const size_t w2_blk = (size_t)im.N2P * 128;
    const size_t fixed = (size_t)im.KB1 * TC_BLK + sizeof(float) * (3 * (size_t)im.N2P + 8 * 128) + 64 * 8 + 1024;
    const size_t max_smem = 232448;
    // Ring plan.  Both weight streams want two chunks resident (the one being multiplied and the one in flight):
    // W2 ring 4 stages, W1 ring 2 KB1 stages.  When that does not fit (wide merger next to its wide X tile), W2
    // falls back to 3 and then 2 stages and W1 takes whatever is left (at least 2 stages; the issuer then walks a
    // chunk in groups, see k_mlp_tc).
    int S2 = 4;
    if (fixed + S2 * w2_blk + 2 * (size_t)im.KB1 * TC_BLK > max_smem) S2 = 3;
    if (fixed + S2 * w2_blk + (size_t)(im.KB1 + 1) * TC_BLK > max_smem) S2 = 2;
    if (fixed + S2 * w2_blk + 2 * (size_t)TC_BLK > max_smem) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: shared memory plan does not fit\n");
    int S1 = (int)((max_smem - fixed - S2 * w2_blk) / TC_BLK);
    if (S1 > TC_MAXS1) S1 = TC_MAXS1;
    a.S1 = S1; a.S2 = S2;
    const size_t smem_bytes = fixed + (size_t)S1 * TC_BLK + (size_t)S2 * w2_blk;
    int grid = a.n_tiles < c->num_sms ? a.n_tiles : c->num_sms;
    if (const char *e = getenv("PHNREC_TC_GRID")) {  // debugging aid: force several tiles per CTA
        const int g = atoi(e);
        if (g > 0 && g < grid) grid = g;
    }
    int rc;
    switch (im.N2P) {
        case 128: rc = launch_one<128>(c, a, smem_bytes, grid); break;
        case 144: rc = launch_one<144>(c, a, smem_bytes, grid); break;
        case 160: rc = launch_one<160>(c, a, smem_bytes, grid); break;
        default: rc = launch_one<192>(c, a, smem_bytes, grid); break;
    }
    c->k_launches[PHN_K_MLP] += 1;
    return rc;
}

// The merger's activation image is written by the band nets' epilogues; its constant-1 bias columns that lie
// outside what those epilogues write are filled once per (re)allocation.
int mlp_tc_fill_merger_bias(phn_ctx *c, int64_t rows)
{
    TcState &st = *static_cast<TcState *>(c->tc);
    if (rows <= 0) return PHN_OK;
    k_fill_bias_cols<<<(unsigned)((rows * 2 + 255) / 256), 256, 0, c->stream>>>((uint8_t *)c->d_xmh.p, rows, st.net[2].KB1, st.net[2].kin);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

int launch_mlp_tc(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    int rc;
    if ((rc = run_net_tc(c, 0, (const uint8_t *)c->d_x0h.p, nf, f0))) return rc;
    if ((rc = run_net_tc(c, 1, (const uint8_t *)c->d_x1h.p, nf, f0))) return rc;
    return run_net_tc(c, 2, (const uint8_t *)c->d_xmh.p, nf, f0);
}

}  // namespace phn
