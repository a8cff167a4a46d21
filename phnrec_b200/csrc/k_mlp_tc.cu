// k_mlp_tc.cu — K-mlp, tensor-core mode: one fused kernel per net,
//     P = fsoftmax(b2 + fsig(b1 + X W1^T) W2^T)
// on the 5th-generation tensor cores (tcgen05.mma, fp16 operands, fp32 accumulators in TMEM),
// operands streamed into shared memory by the TMA engine (cp.async.bulk + mbarrier), the hidden
// layer never leaving the SM.  Replaces NeuralNet::Forward (nn.cpp:872-950), fexp_sigmoid /
// fexp_softmax_v (fexp.h:33-78) and Traps::CalcInputFeaturesForMerger (traps.cpp:435-461) in
// reduced precision; the measured deviation from the exact mode is stated in DESIGN.md.
//
// Tiling: a CTA owns 128 frames (UMMA M = 128, cta_group::1) and walks the hidden layer in
// chunks of 128 units:
//     G1(c): D1[c&1] (TMEM, 128 cols)  = X[128 x K1] . W1[c]^T          K1/16 MMAs of 128x128x16
//     E1(c): H = fp16(fsig(D1 + b1))  -> shared memory (K-major, 128B swizzle)     8 epilogue warps
//     G2(c): D2 (TMEM, N2P cols)     += H[128 x 128] . W2[:, c]^T       8 MMAs of 128xN2Px16
//     E2   : softmax over D2 + b2, then posteriors (merger) or ln + merger input norm (band nets)
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..9 = epilogue
// (two warps per TMEM lane quarter, each taking half of the columns).
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
    const float *sig_k;       // [NCH*128]  per hidden unit: (Ct - b1/ln2) / kSigTmax
    const float *b2;          // [N2P]      output bias; -FLT_MAX in the padding columns
    int n_tiles, KB1, NCH, S1, S2;
    int nks_last;             // k-steps (of 16) actually needed in the last k-block of layer 1
    int64_t nf;               // frames in this launch
    int nout;
    // outputs
    float *post; int ldpost;                          // merger: posteriors [nf][ldpost]
    float *logp;                                      // merger: ln(posteriors) [nf][ldpost] for the decoder, or nullptr
    uint8_t *xm_img; int xm_kb1; int xm_col0;         // band nets: merger input image, first column (multiple of 8)
    const float *mmean, *mdev;                        // merger input normalisation, indexed by image column
};

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
    uint8_t *sH = sW2 + (size_t)a.S2 * W2_BLK;                 // 2 x 16 KB
    float *s_sigk = reinterpret_cast<float *>(sH + 2 * TC_BLK);  // [NCH*128]
    float *s_b2 = s_sigk + a.NCH * TC_NC;                      // [N2P]
    float *s_mm = s_b2 + N2P;                                  // [N2P] merger input mean  (band nets)
    float *s_md = s_mm + N2P;                                  // [N2P] merger input 1/std (band nets)
    float *s_red = s_md + N2P;                                 // [2][4][128] row max / row sum exchange
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_red + 8 * 128);
    uint64_t *x_full = bars;                 // [8]
    uint64_t *x_empty = bars + 8;            // [1]
    uint64_t *w1_full = bars + 9;            // [8]
    uint64_t *w1_empty = bars + 17;          // [8]
    uint64_t *w2_full = bars + 25;           // [4]
    uint64_t *w2_empty = bars + 29;          // [4]
    uint64_t *d1_full = bars + 33;           // [2]
    uint64_t *d1_empty = bars + 35;          // [2]
    uint64_t *h_full = bars + 37, *h_empty = bars + 38;
    uint64_t *d2_full = bars + 39, *d2_empty = bars + 40;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 41);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARP_TMA = TC_EPI_WARPS, WARP_MMA = TC_EPI_WARPS + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(&x_full[i], 1); mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1); }
        mbar_init(x_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], TC_EPI_WARPS); }
        mbar_init(h_full, TC_EPI_WARPS); mbar_init(h_empty, 1);
        mbar_init(d2_full, 1); mbar_init(d2_empty, TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {  // TMEM: 512 columns (D1 double buffer 2 x 128, D2 up to 192)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < a.NCH * TC_NC; i += blockDim.x) s_sigk[i] = a.sig_k[i];
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

    if (warp == WARP_TMA) {
        // ===================================================================== TMA producer
        // The whole warp walks the loops (warp-uniform control flow); one elected lane issues.
        uint32_t ph_x_empty = 0, w1_stage = 0, ph_w1 = 0, w2_stage = 0, ph_w2 = 0;
        bool first_tile = true;
        auto load_w1 = [&](int c) {
            for (int kb = 0; kb < a.KB1; ++kb) {
                mbar_wait(&w1_empty[w1_stage], ph_w1 ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&w1_full[w1_stage], TC_BLK);
                    tma_load_1d(sW1 + (size_t)w1_stage * TC_BLK, a.w1_img + ((size_t)c * a.KB1 + kb) * TC_BLK, TC_BLK, &w1_full[w1_stage]);
                }
                __syncwarp();
                if (++w1_stage == (uint32_t)a.S1) { w1_stage = 0; ph_w1 ^= 1; }
            }
        };
        auto load_w2 = [&](int c) {
            for (int kb = 0; kb < 2; ++kb) {
                mbar_wait(&w2_empty[w2_stage], ph_w2 ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&w2_full[w2_stage], W2_BLK);
                    tma_load_1d(sW2 + (size_t)w2_stage * W2_BLK, a.w2_img + ((size_t)c * 2 + kb) * W2_BLK, W2_BLK, &w2_full[w2_stage]);
                }
                __syncwarp();
                if (++w2_stage == (uint32_t)a.S2) { w2_stage = 0; ph_w2 ^= 1; }
            }
        };
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            // Later tiles: the first weight chunk is prefetched while the previous tile still owns X -
            // but only when it fits the ring entirely (KB1 <= S1); otherwise the issuer, which waits
            // for X before it frees a ring stage, and this warp would wait for each other.
            const bool prefetch_w1 = !first_tile && a.KB1 <= a.S1;
            if (prefetch_w1) load_w1(0);
            if (!first_tile) { mbar_wait(x_empty, ph_x_empty); ph_x_empty ^= 1; }
            if (elect_one()) {
                for (int kb = 0; kb < a.KB1; ++kb) {
                    mbar_expect_tx(&x_full[kb], TC_BLK);
                    tma_load_1d(sX + (size_t)kb * TC_BLK, a.x_img + ((size_t)tile * a.KB1 + kb) * TC_BLK, TC_BLK, &x_full[kb]);
                }
            }
            __syncwarp();
            if (!prefetch_w1) load_w1(0);
            first_tile = false;
            for (int c = 0; c < a.NCH; ++c) {   // same order as the issuer consumes: G1(c+1) before G2(c)
                if (c + 1 < a.NCH) load_w1(c + 1);
                load_w2(c);
            }
        }
    } else if (warp == WARP_MMA) {
        // ===================================================================== MMA issuer
        // Warp-uniform loops; the MMAs and commits of one k-block are issued by the elected lane.
        const uint32_t idesc1 = make_idesc(TC_NC), idesc2 = make_idesc(N2P);
        uint32_t ph_x_full = 0, w1_stage = 0, ph_w1 = 0, w2_stage = 0, ph_w2 = 0, ph_h_full = 0, ph_d2_empty = 0;
        uint32_t ph_d1_empty0 = 0, ph_d1_empty1 = 0, n_d1_use0 = 0, n_d1_use1 = 0;
        bool first_d2 = true;
        // descriptors differ only in the 14-bit start-address field: base + (byte offset >> 4)
        const uint64_t dX = make_sw128_desc(smem_u32(sX)), dW1 = make_sw128_desc(smem_u32(sW1));
        const uint64_t dW2 = make_sw128_desc(smem_u32(sW2)), dH = make_sw128_desc(smem_u32(sH));
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            auto g1 = [&](int c) {
                const int b = c & 1;
                if (b == 0) {
                    if (n_d1_use0 > 0) { mbar_wait(&d1_empty[0], ph_d1_empty0); ph_d1_empty0 ^= 1; }
                    ++n_d1_use0;
                } else {
                    if (n_d1_use1 > 0) { mbar_wait(&d1_empty[1], ph_d1_empty1); ph_d1_empty1 ^= 1; }
                    ++n_d1_use1;
                }
                const uint32_t td = b ? tD1[1] : tD1[0];
                for (int kb = 0; kb < a.KB1; ++kb) {
                    if (c == 0) mbar_wait(&x_full[kb], ph_x_full);
                    mbar_wait(&w1_full[w1_stage], ph_w1);
                    tc_fence_after();
                    if (elect_one()) {
                        const int nks = kb == a.KB1 - 1 ? a.nks_last : 4;
                        const uint64_t ad = dX + (uint64_t)((kb * TC_BLK) >> 4), bd = dW1 + (uint64_t)((w1_stage * TC_BLK) >> 4);
                        for (int ks = 0; ks < nks; ++ks) umma_f16_ss(td, ad + 2 * ks, bd + 2 * ks, idesc1, (kb | ks) ? 1u : 0u);
                        tc_commit(&w1_empty[w1_stage]);
                        if (kb == a.KB1 - 1) {
                            tc_commit(&d1_full[b]);
                            if (c == a.NCH - 1) tc_commit(x_empty);
                        }
                    }
                    __syncwarp();
                    if (++w1_stage == (uint32_t)a.S1) { w1_stage = 0; ph_w1 ^= 1; }
                }
            };
            auto g2 = [&](int c) {
                if (c == 0 && !first_d2) { mbar_wait(d2_empty, ph_d2_empty); ph_d2_empty ^= 1; }
                first_d2 = false;
                mbar_wait(h_full, ph_h_full); ph_h_full ^= 1;
                for (int kb = 0; kb < 2; ++kb) {
                    mbar_wait(&w2_full[w2_stage], ph_w2);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = dH + (uint64_t)((kb * TC_BLK) >> 4), bd = dW2 + (uint64_t)((w2_stage * W2_BLK) >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma_f16_ss(tD2, ad + 2 * ks, bd + 2 * ks, idesc2, (c | kb | ks) ? 1u : 0u);
                        tc_commit(&w2_empty[w2_stage]);
                        if (kb == 1) {
                            tc_commit(h_empty);
                            if (c == a.NCH - 1) tc_commit(d2_full);
                        }
                    }
                    __syncwarp();
                    if (++w2_stage == (uint32_t)a.S2) { w2_stage = 0; ph_w2 ^= 1; }
                }
            };
            g1(0);
            for (int c = 0; c < a.NCH; ++c) {
                if (c + 1 < a.NCH) g1(c + 1);
                g2(c);
            }
            ph_x_full ^= 1;
        }
    } else {
        // ===================================================================== epilogue warps
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int cq = warp >> 2;                // column quarter 0..3
        const int row = q * 32 + lane;           // tile row == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t ph_d1_full0 = 0, ph_d1_full1 = 0, ph_h_empty = 0, ph_d2_full = 0;
        bool first_h = true;
        // H[row][cq*32 .. +31] lives in k-block cq>>1 of the H operand, 16-byte chunks (cq&1)*4 .. +3
        uint8_t *hrow = sH + (size_t)(cq >> 1) * TC_BLK + (size_t)row * 128;
        const int hchunk0 = (cq & 1) * 4;
        const float sigA = (float)(-1.0 / (kLn2 * (double)kSigTmax));
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
            for (int c = 0; c < a.NCH; ++c) {
                const int b = c & 1;
                if (b == 0) { mbar_wait(&d1_full[0], ph_d1_full0); ph_d1_full0 ^= 1; }
                else        { mbar_wait(&d1_full[1], ph_d1_full1); ph_d1_full1 ^= 1; }
                tc_fence_after();
                uint32_t acc[32];
                tmem_ld32((b ? tD1[1] : tD1[0]) + lane_addr + cq * 32, acc);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d1_empty[b]);
                // fsig(x) = 1 / (1 + D(-x - b1)); two reciprocals share one MUFU: 1/a = a' / (a a'), 1/a' = a / (a a')
                uint32_t hp[16];
                const float4 *kp = reinterpret_cast<const float4 *>(s_sigk + c * TC_NC + cq * 32);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 k4 = kp[g];
                    const float a0 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g * 4 + 0]), sigA, k4.x), kSigTmax);
                    const float a1 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g * 4 + 1]), sigA, k4.y), kSigTmax);
                    const float a2 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g * 4 + 2]), sigA, k4.z), kSigTmax);
                    const float a3 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g * 4 + 3]), sigA, k4.w), kSigTmax);
                    const float r01 = rcp_approx(a0 * a1), r23 = rcp_approx(a2 * a3);
                    hp[g * 2] = pack_half2(r01 * a1, r01 * a0);
                    hp[g * 2 + 1] = pack_half2(r23 * a3, r23 * a2);
                }
                if (!first_h) { mbar_wait(h_empty, ph_h_empty); ph_h_empty ^= 1; }
                first_h = false;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                    const uint4 v = make_uint4(hp[ch * 4], hp[ch * 4 + 1], hp[ch * 4 + 2], hp[ch * 4 + 3]);
                    *reinterpret_cast<uint4 *>(hrow + (((hchunk0 + ch) ^ (row & 7)) << 4)) = v;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_full);
            }
            // ------------------------------------------------------------- E2: softmax + outputs
            // column quarter cq owns D2 columns [cq*NQ, (cq+1)*NQ); the 4 warps of a row quarter
            // exchange row max / row sum through shared memory and their own named barrier
            const int n0 = cq * NQ;
            mbar_wait(d2_full, ph_d2_full); ph_d2_full ^= 1;
            tc_fence_after();
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
                                xn[i] = n + i < a.nout ? (v - mm[i]) * md[i] : 0.0f;
                            }
                            const int cm = a.xm_col0 + n;
                            uint8_t *blk = a.xm_img + ((size_t)tile * a.xm_kb1 + (cm >> 6)) * TC_BLK;
                            *reinterpret_cast<uint2 *>(blk + (size_t)row * 128 + ((((cm >> 3) & 7) ^ (row & 7)) << 4) + ((cm & 4) << 1)) =
                                make_uint2(pack_half2(xn[0], xn[1]), pack_half2(xn[2], xn[3]));
                        }
                    }
                }
            }
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
    float *sig_k = nullptr, *b2 = nullptr;
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

__global__ void k_build_w1_img(const float *__restrict__ w1, int nin, int nhid, int nin4, uint8_t *img, int KB1, int NCH,
                               int split, int split8)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)NCH * 128 * KB1 * 64;
    if (idx >= total) return;
    const int k = (int)(idx % (KB1 * 64));
    const int n = (int)(idx / (KB1 * 64));
    const int ki = img_col_to_input(k, split, split8);
    const float v = (n < nhid && ki >= 0 && ki < nin) ? w1[(int64_t)n * nin4 + ki] : 0.0f;
    const int c = n >> 7, r = n & 127, kb = k >> 6, cc = k & 63;
    *reinterpret_cast<__half *>(img + ((size_t)c * KB1 + kb) * TC_BLK + sw128_off(r, cc)) = __float2half_rn(v);
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

__global__ void k_build_bias(const float *__restrict__ b1, int nhid, float *sig_k, int nhidP, const float *__restrict__ b2,
                             int nout, float *b2p, int N2P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // u = sat(t / TMAX),  t = -(x + b1)/ln2 + Ct : the per-unit addend of the epilogue's FFMA.SAT
    if (i < nhidP) sig_k[i] = (float)((kCt - (double)(i < nhid ? b1[i] : 0.0f) / kLn2) / (double)kSigTmax);
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
        im.kin = i == 2 ? split8 + split : n.nin;   // image columns that carry data
        im.KB1 = (im.kin + 63) / 64;
        im.NCH = (n.nhid + 127) / 128;
        im.N2P = (n.nout + 15) / 16 * 16;
        const int rem = im.kin - (im.KB1 - 1) * 64;
        im.nks_last = (rem + 15) / 16;
        if (im.KB1 > 8) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: more than 512 network inputs\n");
        if (im.N2P != 128 && im.N2P != 144 && im.N2P != 160 && im.N2P != 192)
            return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: %d network outputs not instantiated\n", n.nout);
        const size_t w1b = (size_t)im.NCH * im.KB1 * TC_BLK, w2b = (size_t)im.NCH * 2 * im.N2P * 128;
        PHN_CUDA(c, cudaMalloc((void **)&im.w1_img, w1b));
        PHN_CUDA(c, cudaMalloc((void **)&im.w2_img, w2b));
        PHN_CUDA(c, cudaMalloc((void **)&im.sig_k, sizeof(float) * im.NCH * 128));
        PHN_CUDA(c, cudaMalloc((void **)&im.b2, sizeof(float) * im.N2P));
        const int64_t t1 = (int64_t)im.NCH * 128 * im.KB1 * 64, t2 = (int64_t)im.N2P * im.NCH * 128;
        k_build_w1_img<<<(unsigned)((t1 + 255) / 256), 256, 0, c->stream>>>(n.w1, n.nin, n.nhid, n.nin4, im.w1_img, im.KB1, im.NCH, split, split8);
        if (i == 2) {
            // + 16: a band net reads its N2P (= outputs rounded up to 16) columns starting at its first image column
            const int ncol = im.KB1 * 64 + 16;
            PHN_CUDA(c, cudaMalloc((void **)&im.mean_img, sizeof(float) * ncol));
            PHN_CUDA(c, cudaMalloc((void **)&im.dev_img, sizeof(float) * ncol));
            k_build_mnorm<<<(ncol + 127) / 128, 128, 0, c->stream>>>(n.mean, n.dev, n.nin, im.mean_img, im.dev_img, ncol, split, split8);
        }
        k_build_w2_img<<<(unsigned)((t2 + 255) / 256), 256, 0, c->stream>>>(n.w2, n.nhid, n.nout, n.nhid4, im.w2_img, im.N2P, im.NCH);
        const int nb = im.NCH * 128 > im.N2P ? im.NCH * 128 : im.N2P;
        k_build_bias<<<(nb + 255) / 256, 256, 0, c->stream>>>(n.b1, n.nhid, im.sig_k, im.NCH * 128, n.b2, n.nout, im.b2, im.N2P);
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
        if (im.sig_k) cudaFree(im.sig_k);
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
    a.x_img = x_img; a.w1_img = im.w1_img; a.w2_img = im.w2_img; a.sig_k = im.sig_k; a.b2 = im.b2;
    a.n_tiles = (int)((nf + TC_M - 1) / TC_M);
    a.KB1 = im.KB1; a.NCH = im.NCH; a.nks_last = im.nks_last; a.nf = nf; a.nout = n.nout;
    if (which < 2) {
        a.post = nullptr;
        a.xm_img = (uint8_t *)c->d_xmh.p; a.xm_kb1 = st.net[2].KB1; a.xm_col0 = which * ((n.nout + 7) / 8 * 8);
        a.mmean = st.net[2].mean_img; a.mdev = st.net[2].dev_img;
    } else {
        a.post = (float *)c->d_post.p + f0 * c->ldp; a.ldpost = c->ldp;
        a.logp = c->fuse_logp ? (float *)c->d_logp.p + f0 * c->ldp : nullptr;
    }
    // shared memory plan: X (KB1 blocks) + H (2 blocks) + constants + barriers are fixed; the rest is split
    // between the W2 ring (S2 k-blocks of N2P x 64) and the W1 ring (S1 blocks of 128 x 64)
    const size_t w2_blk = (size_t)im.N2P * 128;
    const size_t fixed = (size_t)im.KB1 * TC_BLK + 2 * TC_BLK + sizeof(float) * ((size_t)im.NCH * TC_NC + 3 * im.N2P + 8 * 128) + 48 * 8 + 1024;
    const size_t max_smem = 232448;
    int S1 = 0, S2 = 4;
    for (; S2 >= 2; --S2) {
        if (fixed + S2 * w2_blk + 2 * (size_t)TC_BLK > max_smem) continue;
        S1 = (int)((max_smem - fixed - S2 * w2_blk) / TC_BLK);
        if (S1 >= (S2 == 2 ? 2 : 4)) break;
    }
    if (S2 < 2 || S1 < 2) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: shared memory plan does not fit\n");
    if (S1 > 8) S1 = 8;
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

int launch_mlp_tc(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    int rc;
    if ((rc = run_net_tc(c, 0, (const uint8_t *)c->d_x0h.p, nf, f0))) return rc;
    if ((rc = run_net_tc(c, 1, (const uint8_t *)c->d_x1h.p, nf, f0))) return rc;
    return run_net_tc(c, 2, (const uint8_t *)c->d_xmh.p, nf, f0);
}

}  // namespace phn
