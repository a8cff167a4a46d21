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
#include "tc_ptx.cuh"

#include <cfloat>
#include <type_traits>
#include <cstdlib>

namespace phn {

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one issue slot for two lanes of fp32 work).  The epilogues are bound by
// the SM's issue slots (a sigmoid per hidden unit and frame: ~6.5 instructions each in scalar form), so every pair of
// scalar FFMA / FADD / FMUL that can travel as one packed instruction is a direct gain.
#ifndef PHN_TC_PACKED
#define PHN_TC_PACKED 1
#endif
#ifndef PHN_TC_E1_LD32       // (kernel development switches, A/B'd on the GPU; the defaults are what measured best)
#define PHN_TC_E1_LD32 0
#endif
#ifndef PHN_TC_D1_EARLY   // 1: d1_empty right after the second TMEM load of E1 instead of after its arithmetic (measured: 3.12 against 3.10 ms, not kept)
#define PHN_TC_D1_EARLY 0
#endif
#ifndef PHN_TC_HRING   // 1: H lives in a ring of 16-column quarter slots behind D2 (see "H ring" in the kernel) instead of one 64-column buffer
#define PHN_TC_HRING 1
#endif
#ifndef PHN_TC_RCP4
#define PHN_TC_RCP4 0
#endif
#ifndef PHN_TC_PSLEEP
#define PHN_TC_PSLEEP 0
#endif
// PHN_TC_SHF=1 (kernel development): bits << 9 on the integer ALU (SHF) instead of the compiler's IMAD.SHL, which shares the FMA
// pipe with the epilogue's FFMA / FFMA2 / FMUL2 work (ncu: math-pipe throttle on exactly these instructions).  Measured and
// not kept: 3.150 against 3.100 ms per step for the three nets, three A/B pairs on one box.
#ifndef PHN_TC_SHF
#define PHN_TC_SHF 0
#endif
__device__ __forceinline__ uint32_t shl9(uint32_t v)
{
#if PHN_TC_SHF
    uint32_t r;
    asm("shf.l.clamp.b32 %0, %1, %2, 9;" : "=r"(r) : "r"(0u), "r"(v));
    return r;
#else
    return v << 9;
#endif
}
struct f2 { float x, y; };
__device__ __forceinline__ uint64_t pk2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 up2(uint64_t v)
{
    f2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c)
{
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return up2(d);
}
__device__ __forceinline__ f2 fadd2(f2 a, f2 b)
{
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return up2(d);
}
__device__ __forceinline__ f2 fmul2(f2 a, f2 b)
{
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return up2(d);
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
// softmax: y = o - max <= 0.  The numerators are scaled by 2^24 (it cancels against the row sum) and t is kept in
// [1, 152]: e' = 2^24 D(y) stays a NORMAL float down to y/ln2 = -150, where the reference's own float D(y) runs out of
// denormals and becomes 0 (fexp.h:14-21 cast to float).  Without the shift every y < -88 gave e = 0, i.e. sLn's guard
// ln := 0 (dspc.h:155-160) instead of ln p ~ -90: one frame in ~10^5 of the merger's input was then far off.
// u = sat((t - 1) / (TMAX - 1)); r = 2^23 + 2^14 (1 + u (TMAX - 1)); bits(e') = bits(r) << 9.  u == 0 <=> e' == 2^-126: "zero".
constexpr float kSmxTmax = 152.0f;
constexpr double kSmxOff = 24.0;
constexpr float kSmxZero = 1.17549435e-38f;      // 2^-126: the clamp value, treated as the reference's exact 0
__device__ __forceinline__ float fexp_from_u(float u, float tmax)
{
    const float r = fmaf(u, tmax * 16384.0f, 8388608.0f);
    return __uint_as_float(__float_as_uint(r) << 9);
}

// 16 hidden pre-activations (b1 already inside, fp32 bit patterns from the accumulator) -> 16 fp16 activations as 8 pairs:
// fsig(x) = 1 / (1 + D(-x)) with the conversion-free bit-trick exponential above; elements (0, 2) and (1, 3) of every four
// travel as packed fp32 pairs, and two reciprocals share one MUFU: 1/a0 = a1 / (a0 a1), 1/a1 = a0 / (a0 a1).
__device__ __forceinline__ void sigmoid16(const uint32_t *acc, uint32_t *hp, float sigA, float sigB)
{
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const f2 uA = {fma_sat(__uint_as_float(acc[g4 * 4 + 0]), sigA, sigB), fma_sat(__uint_as_float(acc[g4 * 4 + 2]), sigA, sigB)};
        const f2 uB = {fma_sat(__uint_as_float(acc[g4 * 4 + 1]), sigA, sigB), fma_sat(__uint_as_float(acc[g4 * 4 + 3]), sigA, sigB)};
        const f2 kk = {kSigTmax * 16384.0f, kSigTmax * 16384.0f}, mg = {8388608.0f, 8388608.0f}, one = {1.0f, 1.0f};
        const f2 rA = ffma2(uA, kk, mg), rB = ffma2(uB, kk, mg);
        const f2 dA = {__uint_as_float(shl9(__float_as_uint(rA.x))), __uint_as_float(shl9(__float_as_uint(rA.y)))};
        const f2 dB = {__uint_as_float(shl9(__float_as_uint(rB.x))), __uint_as_float(shl9(__float_as_uint(rB.y)))};
        const f2 aA = fadd2(dA, one), aB = fadd2(dB, one);
        const f2 pr = fmul2(aA, aB);
        const f2 rc = {rcp_approx(pr.x), rcp_approx(pr.y)};
        const f2 hA = fmul2(rc, aB), hB = fmul2(rc, aA);          // (h0, h2), (h1, h3)
        hp[g4 * 2] = pack_half2(hA.x, hB.x);
        hp[g4 * 2 + 1] = pack_half2(hA.y, hB.y);
    }
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
    int XR;                   // slots of the X ring: KB1 (the tile's blocks, reloaded in place) or KB1 + 1 (one spare slot: the
                              // next tile's first block is already there when the last layer-1 chunk of this tile ends)
    int nks_last;             // k-steps (of 16) actually needed in the last k-block of layer 1
    int64_t nf;               // frames in this launch
    int nout;
    // outputs
    float *post; int ldpost;                          // merger: posteriors [nf][ldpost]
    float *logp;                                      // merger: ln(posteriors) for the decoder, TILED [tile][ldpost][128 rows], or nullptr
    int64_t logp_tile0;                               //         global tile index of this launch's first tile
    int x_mn;                                         // the activation image is M-major (mn128_off): the merger's
    uint8_t *xm_img; int xm_kb1; int xm_col0;         // band nets: merger input image, first column (multiple of 8)
    long long *dbg;                                   // optional timeline of CTA 0's second tile (tools/tc_timeline.py)
    int xm_bias;                                      // band 1: its columns nout, nout+1 are the merger's constant-1 bias inputs
    const float *mmean, *mdev;                        // merger input normalisation, indexed by image column
};

// debug timeline: dbg[(c * 16 + event)] = clock64() for chunk c of CTA 0's second tile
#define TC_DBG(ev, c)                                                                       \
    do {                                                                                    \
        if (DBG && a.dbg && blockIdx.x == 0 && tile == tile0 + tstep) a.dbg[(c) * 16 + (ev)] = clock64(); \
    } while (0)
// E2 stage events of that tile go to row 15: 0 A begins (waits for D2), 1 D2 seen, 2 A done, 3 B begins, 4 B done, 5 C begins, 6 C done
#define TC_DBG2(ev)                                                                                                   \
    do {                                                                                                              \
        if (DBG && a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && e2_tile == tile0 + tstep) a.dbg[15 * 16 + (ev)] = clock64(); \
    } while (0)
constexpr int TC_MAXS1 = 24;                        // upper bound on W1 ring stages (barrier array size)
constexpr int TC_MAXS2 = 8;                         // ... W2 ring stages (4 chunks)
constexpr int TC_EPI_WARPS = 16;                     // warps 0..15: epilogue; 16: TMA producer; 17, 18: MMA issuers (layer 1, layer 2)
constexpr int TC_THREADS = (TC_EPI_WARPS + 3) * 32;
// The SM's warp schedulers favour the highest warp id among eligible warps, so the two warps whose
// instruction streams gate everything else (TMA producer, MMA issuer) get the highest ids.
__device__ __forceinline__ void quarter_bar_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

__device__ __forceinline__ float lg2_approx(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Band nets, E2 stage C: the thread's NQ soft-max numerators e (row `row`, image columns c0 .. c0+NQ-1, c0 % 8 == SH)
// -> normalised merger inputs, fp16, into the merger's M-major image.  x = lg2(e / sum) * A + B with the per-column
// constants of the kernel prologue; e == 0 (exponent underflow, padding columns) takes sLn's guard: ln := 0 -> x = B.
// All address arithmetic is per thread and loop invariant: 8 pointers (one per column % 8), immediates for the rest.
template <int NQ, int SH>
__device__ __forceinline__ void band_out(const float (&o)[NQ], float lg2_sc, const float *sA, const float *sB, uint8_t *img,
                                         int row, int c0, int ncols)
{
    uint8_t *ptr[8];
    const uint32_t rc = ((uint32_t)row & 63u) >> 3;
    uint8_t *rb = img + (size_t)((c0 - SH) >> 3) * 2048 + ((uint32_t)row >> 6) * 1024u + ((uint32_t)row & 7u) * 2u;
#pragma unroll
    for (int t = 0; t < 8; ++t) ptr[t] = rb + t * 128 + ((rc ^ (uint32_t)t) << 4);
#pragma unroll
    for (int j = 0; j < NQ / 4; ++j) {
        if (4 * j < ncols) {   // (the image share of a net ends on a multiple of 8 columns, c0 is a multiple of 4)
            const float4 a4 = *reinterpret_cast<const float4 *>(sA + 4 * j), b4 = *reinterpret_cast<const float4 *>(sB + 4 * j);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cc = SH + 4 * j + i;
                const float e = o[4 * j + i];
                const float x = e > kSmxZero ? fmaf(lg2_approx(e), av[i], fmaf(lg2_sc, av[i], bv[i])) : bv[i];
                *reinterpret_cast<__half *>(ptr[cc & 7] + (cc >> 3) * 2048) = __float2half_rn(x);
            }
        }
    }
}

// XMN: the activation image is M-major (the merger's), else K-major.
// PAIR: the kernel runs as clusters of two CTAs driving cta_group::2 MMAs (M = 256): each CTA owns one 128-frame tile
// and keeps only HALF of every weight block in its shared memory (W1: 64 of the chunk's 128 hidden rows, W2: N2P/2
// output rows) - half the TMA weight stream and 6 KB instead of 8 KB of shared-memory operand reads per layer-1 MMA,
// which is what bounds the single-CTA kernel.  CTA 0 of the pair issues every MMA; its completion signals are
// multicast to both CTAs; CTA 1's issuer warps only relay "my operands have landed" to CTA 0, and its epilogue warps
// arrive on CTA 0's barriers.  Everything else (producer, epilogues) is per CTA and unchanged.
// DBG: the instantiation with the clock64() timeline hooks (tools/tc_timeline.py); the product kernels carry none.
template <int N2P, bool XMN, bool PAIR, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1) k_mlp_tc(TcArgs a)
{
    constexpr int W2_BLK = N2P * 128;       // bytes of one [N2P rows x 64 fp16] block of the weight image
    constexpr int W1_ST = PAIR ? TC_BLK / 2 : TC_BLK;      // bytes of one W1 ring stage in this CTA's shared memory
    constexpr int W2_ST = PAIR ? W2_BLK / 2 : W2_BLK;      // ... W2 ring stage
    const uint32_t rank = PAIR ? (blockIdx.x & 1u) : 0u;   // CTA rank in the pair (cluster dims (2,1,1))
    // tile sequence of this CTA: tile0, tile0 + tstep, ... (n_my of them; a pair's two CTAs run in lock step)
    const int n_units = PAIR ? (a.n_tiles + 1) / 2 : a.n_tiles, unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int ustep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_my = unit0 < n_units ? (n_units - 1 - unit0) / ustep + 1 : 0;
    const int tile0 = PAIR ? 2 * unit0 + (int)rank : unit0, tstep = PAIR ? 2 * ustep : ustep;
    constexpr int NQ = N2P / 4;             // D2 columns per epilogue column quarter (32, 36, 40 or 48)
    constexpr int NR = NQ - 32;             // columns beyond the first 32-column TMEM load (0, 4, 8 or 16)
    static_assert(NR == 0 || NR == 4 || NR == 8 || NR == 16, "unsupported output width");
    extern __shared__ uint8_t smem_raw[];
    // carve-up (all block bases 1024-byte aligned: the swizzle pattern is a function of the address)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sX = smem;                                        // XR x 16 KB: block k of the CTA's tile i sits in slot (i KB1 + k) mod XR
    uint8_t *sW1 = sX + (size_t)a.XR * TC_BLK;                 // S1 x 16 KB ring
    uint8_t *sW2 = sW1 + (size_t)a.S1 * W1_ST;                 // S2 x W2_ST ring
    float *s_b2 = reinterpret_cast<float *>(sW2 + (size_t)a.S2 * W2_ST);    // [N2P]
    float *s_mm = s_b2 + N2P;                                  // [N2P] merger input mean  (band nets)
    float *s_md = s_mm + N2P;                                  // [N2P] merger input 1/std (band nets)
    float *s_red = s_md + N2P;                                 // [2][4][128] row max / row sum exchange
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_red + 8 * 128);
    uint64_t *x_full = bars;                 // [8]
    uint64_t *x_empty = bars + 8;            // [8] per k-block: the tile's last layer-1 chunk has consumed it
    uint64_t *w1_full = bars + 16;           // [TC_MAXS1]
    uint64_t *w1_empty = w1_full + TC_MAXS1; // [TC_MAXS1]
    uint64_t *w2_empty = w1_empty + TC_MAXS1; // [TC_MAXS2]
    uint64_t *w1c_full = w2_empty + TC_MAXS2;       // [4] chunk-level: all KB1 k-blocks of layer-1 chunk n have landed (n & 3)
    uint64_t *w2c_full = w1c_full + 4;       // [4] chunk-level: both k-blocks of layer-2 chunk n have landed (n & 3)
    uint64_t *d1_full = w2c_full + 4;        // [2]
    uint64_t *h_full = d1_full + 2, *h_empty = h_full + 1;
    uint64_t *d2_full = h_empty + 1, *d2_empty = d2_full + 1;
    uint64_t *d1_empty = d2_empty + 1;       // [2] E1 has read the accumulator buffer
    uint64_t *pw1_full = d1_empty + 2;       // [4] PAIR, CTA 0: the peer's layer-1 operands of chunk n have landed (n & 3)
    uint64_t *pw2_full = pw1_full + 4;       // [4] PAIR, CTA 0: the peer's layer-2 weights of chunk n have landed (n & 3)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pw2_full + 4);
    volatile int *s_par = reinterpret_cast<volatile int *>(bars + 100);   // [8] loop bounds for the issuer warps (see there)
    // H ring: the hidden activations of a chunk are four QUARTERS of 16 TMEM columns (32 hidden units as fp16 pairs; quarter j is
    // written by the epilogue warps of column quarter j and read by layer-2 k-steps 2j, 2j + 1).  Quarter j of chunk g lives in
    // slot (4 g + j) mod HR of the HR = (256 - N2P) / 16 slots behind D2 (7 for 144 outputs, 6 for 160, 8 for 128, 4 for 192),
    // and every slot has its own full / free barrier pair: an epilogue warp stores its quarter as soon as THAT slot's previous
    // reader - two layer-2 MMAs of an earlier chunk - has completed, instead of waiting for the whole of layer 2 of the
    // previous chunk (one 64-column buffer).  With 7 slots only the warps of quarter 3 ever wait, and only for the first two
    // MMAs of the previous chunk.
    constexpr int HR = (256 - N2P) / 16;
    static_assert(HR >= 4 && HR <= 8, "H ring: 4..8 quarter slots");
    uint64_t *hq_full = bars + 104;          // [8] the four (PAIR: eight) warps of a quarter have stored it
    uint64_t *hq_free = bars + 112;          // [8] the two MMAs that read the slot have completed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARP_TMA = TC_EPI_WARPS, WARP_MMA1 = TC_EPI_WARPS + 1, WARP_MMA2 = TC_EPI_WARPS + 2, EPI0 = 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&x_full[i], 1);
        for (int i = 0; i < TC_MAXS1; ++i) { mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1); }
        for (int i = 0; i < TC_MAXS2; ++i) mbar_init(&w2_empty[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&w1c_full[i], a.KB1); mbar_init(&w2c_full[i], 2); }
        for (int i = 0; i < 8; ++i) mbar_init(&x_empty[i], 1);
        constexpr uint32_t EPI_ARRIVALS = PAIR ? 2 * TC_EPI_WARPS : TC_EPI_WARPS;   // PAIR: both CTAs' epilogue warps arrive on CTA 0
        for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], EPI_ARRIVALS); }
        mbar_init(h_full, EPI_ARRIVALS); mbar_init(h_empty, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&hq_full[i], EPI_ARRIVALS / 4); mbar_init(&hq_free[i], 1); }
        mbar_init(d2_full, 1); mbar_init(d2_empty, EPI_ARRIVALS);
        for (int i = 0; i < 4; ++i) mbar_init(&pw1_full[i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&pw2_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_par[0] = a.KB1; s_par[1] = a.XR; s_par[2] = a.S1; s_par[3] = a.NCH; s_par[4] = a.nks_last; s_par[5] = a.S2;
    }
    if (warp == WARP_MMA1) {  // TMEM: 512 columns = D1 double buffer 2 x 128 | D2 up to 192 | H 64 (128 fp16 per lane)
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < N2P; i += blockDim.x) {
        s_b2[i] = a.b2[i];
        // band nets: merger input x = (sLn(p) - mean) * dev = lg2(p) * (ln2 * dev) - mean * dev; beyond the net's outputs
        // the image holds the two constant-1 bias inputs (band 1) and zeros
        const bool live = a.mmean && i < a.nout;
        s_mm[i] = live ? (float)kLn2 * a.mdev[a.xm_col0 + i] : 0.0f;
        s_md[i] = live ? -a.mmean[a.xm_col0 + i] * a.mdev[a.xm_col0 + i] : ((a.xm_bias && i < a.nout + 2) ? 1.0f : 0.0f);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tD1[2] = {tmem, tmem + 128u};
    const uint32_t tD2 = tmem + 256u;
    const uint32_t tH = PHN_TC_HRING ? tmem + 256u + (uint32_t)N2P : tmem + 448u;   // H ring base / the single H buffer

    if (warp == WARP_TMA) {
        // ===================================================================== TMA producer
        // The whole warp walks the loops (warp-uniform control flow); one elected lane issues.  Loads follow the
        // issuer's consumption order over the CTA's linear chunk sequence g = tile_iter * NCH + c:
        //   X(tile 0), W1(0), W1(1), then per g:  [X(next tile) if chunk g+2 opens it]  W1(g+2)  W2(g)
        uint32_t w1_stage = 0, ph_w1 = 0, w2_stage = 0, ph_w2 = 0;
        const int my_tiles = n_my;
        const int G = my_tiles * a.NCH;
        // Normal case (the W1 ring holds a whole chunk): every k-block load of a chunk reports to the chunk's own
        // "full" barrier, so the issuer waits once per chunk; ring stages are still handed back one by one.
        // Three independent cursors (X tiles, layer-1 weights, layer-2 weights), each advanced whenever its next
        // buffer is free: a blocking in-order producer would sit on the shallow W2 ring and never prefetch W1 ahead.
        // X k-blocks are handed back one by one while the tile's last layer-1 chunk is being multiplied, so the next
        // tile's X (already in L2, see the prefetch) streams in behind it and is there when its first chunk is issued.
        const bool chunk_bar = a.S1 >= a.KB1;
        uint32_t n1 = 0, n2 = 0;          // chunks fully issued per stream
        int kb1 = 0, kb2 = 0;             // next k-block inside the current chunk
        int c1 = 0, c2 = 0;
        int xt = 0, xk = 0, xtile = tile0;   // X cursor: tile iteration, k-block, global tile
        int xs = 0; uint32_t xwrap = 0;      //           ring slot of that block and how often the ring has wrapped (= uses of the slot so far)
        while (n1 < (uint32_t)G || n2 < (uint32_t)G || xt < my_tiles) {
#if PHN_TC_PSLEEP
            bool progress = false;
#endif
            if (xt < my_tiles && (xwrap == 0 || mbar_try(&x_empty[xs], (xwrap - 1) & 1u))) {
#if PHN_TC_PSLEEP
                progress = true;
#endif
                if (elect_one()) {
                    // (PAIR: the odd CTA of the last pair may own a tile past the end; it multiplies the last real tile again
                    // and writes nothing)
                    const int xsrc = xtile < a.n_tiles ? xtile : a.n_tiles - 1;
                    mbar_expect_tx(&x_full[xs], TC_BLK);
                    tma_load_1d(sX + (size_t)xs * TC_BLK, a.x_img + ((size_t)xsrc * a.KB1 + xk) * TC_BLK, TC_BLK, &x_full[xs]);
                    // the CTA's next tile goes to L2 now: when its turn comes (one tile time from here) the
                    // load that sits between two tiles' MMAs is an L2 hit
                    if (xk == 0 && xtile + tstep < a.n_tiles)
                        l2_prefetch(a.x_img + (size_t)(xtile + tstep) * a.KB1 * TC_BLK, (uint32_t)a.KB1 * TC_BLK);
                }
                __syncwarp();
                if (++xk == a.KB1) { xk = 0; ++xt; xtile += tstep; }
                if (++xs == a.XR) { xs = 0; ++xwrap; }
            }
            if (n1 < (uint32_t)G && mbar_try(&w1_empty[w1_stage], ph_w1 ^ 1)) {
#if PHN_TC_PSLEEP
                progress = true;
#endif
                if (elect_one()) {
                    uint64_t *fb = chunk_bar ? &w1c_full[n1 & 3] : &w1_full[w1_stage];
                    // (PAIR: this CTA's half of the block = 64 of the chunk's 128 hidden rows, contiguous in the K-major image)
                    mbar_expect_tx(fb, W1_ST);
                    tma_load_1d(sW1 + (size_t)w1_stage * W1_ST, a.w1_img + ((size_t)c1 * a.KB1 + kb1) * TC_BLK + (size_t)rank * W1_ST, W1_ST, fb);
                }
                __syncwarp();
                if (++w1_stage == (uint32_t)a.S1) { w1_stage = 0; ph_w1 ^= 1; }
                if (++kb1 == a.KB1) {
                    kb1 = 0; ++n1;
                    if (++c1 == a.NCH) c1 = 0;
                }
            }
            if (n2 < (uint32_t)G && mbar_try(&w2_empty[w2_stage], ph_w2 ^ 1)) {
#if PHN_TC_PSLEEP
                progress = true;
#endif
                if (elect_one()) {
                    mbar_expect_tx(&w2c_full[n2 & 3], W2_ST);
                    tma_load_1d(sW2 + (size_t)w2_stage * W2_ST, a.w2_img + ((size_t)c2 * 2 + kb2) * W2_BLK + (size_t)rank * W2_ST, W2_ST, &w2c_full[n2 & 3]);
                }
                __syncwarp();
                if (++w2_stage == (uint32_t)a.S2) { w2_stage = 0; ph_w2 ^= 1; }
                if (++kb2 == 2) {
                    kb2 = 0; ++n2;
                    if (++c2 == a.NCH) c2 = 0;
                }
            }
#if PHN_TC_PSLEEP
            // nothing to load right now: this warp has the highest scheduling priority of its SM sub-partition, and a polling
            // loop there starves the four epilogue warps it shares the issue port with
            if (!progress) __nanosleep(PHN_TC_PSLEEP);
#endif
        }
    } else if (warp == WARP_MMA1) {
        // ===================================================================== MMA issuer, layer 1
        // Two issuer warps, one per layer: an issuing warp is instruction-bound (descriptor arithmetic lives in
        // ordinary registers and every tcgen05.mma needs them moved to uniform registers: ~12 SASS instructions per
        // MMA from ONE warp, at single-warp issue latency - measured ~115 clk per MMA against the tensor pipe's
        // ~64-72), and the two layers' streams are independent: different accumulators, different operand rings,
        // their own barriers.  The tensor pipe interleaves them (SS operand fetch overlaps the TS math).
        // This warp: for every chunk g of the CTA's linear sequence, D1[g & 1] = X . W1[c]^T (K1/16 SS MMAs), as
        // soon as the weights have landed and E1(g - 2) has read the accumulator buffer.
        constexpr uint32_t idesc1 = make_idesc(TC_NC, XMN, PAIR ? 2 * TC_M : TC_M);
        constexpr uint32_t akst = XMN ? (4096u >> 4) : 2u;     // A descriptor advance per k-step of 16 columns
        const uint64_t dX = XMN ? make_mn128_desc(smem_u32(sX)) : make_sw128_desc(smem_u32(sX));
        const uint64_t dW1 = make_sw128_desc(smem_u32(sW1));
        const uint32_t xlo = (uint32_t)dX, xhi = (uint32_t)(dX >> 32), w1lo = (uint32_t)dW1;
        const uint32_t bar_w1e = smem_u32(w1_empty), bar_xe = smem_u32(x_empty), bar_d1f = smem_u32(d1_full);
        // (loop bounds read once through volatile shared memory, which pins them in registers: kernel parameters are
        // re-read from the constant bank inside the issuing loop, and every such load is a dependent stall between two
        // MMAs of the warp that feeds the tensor pipe)
        const int p_kb1 = s_par[0], p_xr = s_par[1], p_s1 = s_par[2], p_nch = s_par[3], p_nksl = s_par[4];
        const int my_tiles = n_my;
        const int G = my_tiles * p_nch;
        const bool chunk_bar = PAIR || p_s1 >= p_kb1;   // the W1 ring holds a whole chunk: one "full" barrier per chunk (PAIR: always, the host plan guarantees it)
        uint32_t st = 0, ph_w1 = 0;
        int xs0 = 0; uint32_t xw0 = 0;          // ring slot of the current tile's block 0, wrap count of the ring at that block
        int c1 = 0, tile = tile0;               // (tile: only for the debug timeline)
        const bool leader = elect_one();
        if (PAIR && rank != 0) {
            // CTA 1 of a pair issues nothing: this warp tells CTA 0 when the layer-1 operands of chunk g (its half of the
            // W1 chunk, and its X tile when the chunk opens one) have landed in THIS CTA's shared memory
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                mbar_wait(&w1c_full[g & 3], (uint32_t)(g >> 2) & 1u);
                if (c1 == 0) {
                    int s = xs0; uint32_t w = xw0;
                    for (int k = 0; k < p_kb1; ++k) {
                        mbar_wait(&x_full[s], w & 1u);
                        if (++s == p_xr) { s = 0; ++w; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&pw1_full[g & 3], 0);
                if (++c1 == p_nch) {
                    c1 = 0;
                    xs0 += p_kb1;
                    if (xs0 >= p_xr) { xs0 -= p_xr; ++xw0; }
                }
            }
        } else
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const bool opens = c1 == 0, closes = c1 == p_nch - 1;
            const uint32_t td1 = (g & 1) ? tmem + 128u : tmem;
            if (lane == 0) TC_DBG(2, c1);   // layer 1 of chunk c1: begins
            // (the weights are normally there long before the accumulator buffer is handed back: the barrier that really gates
            // the chunk is waited for last, so nothing stands between its completion and the first MMA)
            if (chunk_bar) mbar_wait(&w1c_full[g & 3], (uint32_t)(g >> 2) & 1u);
            if (PAIR) mbar_wait_cluster(&pw1_full[g & 3], (uint32_t)(g >> 2) & 1u);
            if (g >= 2) {
                if (PAIR) mbar_wait_cluster(&d1_empty[g & 1], (uint32_t)((g >> 1) - 1) & 1u);
                else mbar_wait(&d1_empty[g & 1], (uint32_t)((g >> 1) - 1) & 1u);
            }
            if (lane == 0) TC_DBG(3, c1);   // accumulator + weights there
            tc_fence_after();
            int xs = xs0; uint32_t xw = xw0;
            uint32_t alo = xlo + (uint32_t)xs0 * (TC_BLK >> 4);
            // The k-block loop exists in four instances (chunk opens / closes a tile or not, compile-time): the steady-state
            // one - ten chunks out of twelve - contains no barrier wait and no conditional commit.
            auto kloop = [&](auto opens_c, auto closes_c) {
            constexpr bool opens = decltype(opens_c)::value, closes = decltype(closes_c)::value;
#pragma unroll 1
            for (int k = 0; k < p_kb1; ++k) {   // (kept rolled: a small loop body stays in the SMSP's instruction cache)
                if (!chunk_bar) mbar_wait(&w1_full[st], ph_w1);
                if (opens) mbar_wait(&x_full[xs], xw & 1u);
                if (!chunk_bar || opens) tc_fence_after();
                const uint32_t blo = w1lo + st * (W1_ST >> 4);
                const uint32_t bw1 = bar_w1e + st * 8u, bxe = bar_xe + (uint32_t)xs * 8u;
                if (leader) {
                    if (PAIR) {
                        if (k == 0) umma2_ss_lo<0>(td1, alo, xhi, blo, idesc1); else umma2_ss_lo<1>(td1, alo, xhi, blo, idesc1);
                        if (k < p_kb1 - 1 || p_nksl == 4) {
                            umma2_ss_lo<1>(td1, alo + akst, xhi, blo + 2, idesc1);
                            umma2_ss_lo<1>(td1, alo + 2 * akst, xhi, blo + 4, idesc1);
                            umma2_ss_lo<1>(td1, alo + 3 * akst, xhi, blo + 6, idesc1);
                        } else {
                            if (p_nksl > 1) umma2_ss_lo<1>(td1, alo + akst, xhi, blo + 2, idesc1);
                            if (p_nksl > 2) umma2_ss_lo<1>(td1, alo + 2 * akst, xhi, blo + 4, idesc1);
                        }
                        tc_commit2_u(bw1);
                        if (closes) tc_commit2_u(bxe);
                    } else {
                        if (k == 0) umma_ss_lo<0>(td1, alo, xhi, blo, idesc1); else umma_ss_lo<1>(td1, alo, xhi, blo, idesc1);
                        if (k < p_kb1 - 1 || p_nksl == 4) {
                            umma_ss_lo<1>(td1, alo + akst, xhi, blo + 2, idesc1);
                            umma_ss_lo<1>(td1, alo + 2 * akst, xhi, blo + 4, idesc1);
                            umma_ss_lo<1>(td1, alo + 3 * akst, xhi, blo + 6, idesc1);
                        } else {
                            if (p_nksl > 1) umma_ss_lo<1>(td1, alo + akst, xhi, blo + 2, idesc1);
                            if (p_nksl > 2) umma_ss_lo<1>(td1, alo + 2 * akst, xhi, blo + 4, idesc1);
                        }
                        tc_commit_u(bw1);
                        if (closes) tc_commit_u(bxe);
                    }
                }
                __syncwarp();
                if (++st == (uint32_t)p_s1) { st = 0; ph_w1 ^= 1; }
                alo += TC_BLK >> 4;
                if (++xs == p_xr) { xs = 0; ++xw; alo = xlo; }
            }
            };
            using T_ = std::true_type; using F_ = std::false_type;
            if (opens) { if (closes) kloop(T_{}, T_{}); else kloop(T_{}, F_{}); }
            else       { if (closes) kloop(F_{}, T_{}); else kloop(F_{}, F_{}); }
            if (leader) { if (PAIR) tc_commit2_u(bar_d1f + (uint32_t)(g & 1) * 8u); else tc_commit_u(bar_d1f + (uint32_t)(g & 1) * 8u); }
            __syncwarp();
            if (lane == 0) TC_DBG(4, c1);   // issued
            if (++c1 == p_nch) {
                c1 = 0; tile += tstep;
                xs0 += p_kb1;
                if (xs0 >= p_xr) { xs0 -= p_xr; ++xw0; }
            }
        }
    } else if (warp == WARP_MMA2) {
        // ===================================================================== MMA issuer, layer 2
        // D2 (+)= H(g) . W2[:, c]^T: 8 TS MMAs per chunk (A = H from TMEM, 8 columns per k-step), as soon as E1(g) has
        // published H and the chunk's two W2 k-blocks have landed; H goes back to the epilogue warps right behind them.
        constexpr uint32_t idesc2 = make_idesc(N2P, false, PAIR ? 2 * TC_M : TC_M);
        const uint32_t w2lo = (uint32_t)make_sw128_desc(smem_u32(sW2));
        const uint32_t bar_w2e = smem_u32(w2_empty), bar_he = smem_u32(h_empty), bar_d2f = smem_u32(d2_full);
        const int p_s2 = s_par[5], p_nch = s_par[3];
        const int my_tiles = n_my;
        const int G = my_tiles * p_nch;
        uint32_t w2s = 0, ph_d2_empty = 0;
        int c2 = 0, tile = tile0;
        const bool leader = elect_one();
        if (PAIR && rank != 0) {
            // CTA 1 of a pair: relay "my half of the chunk's W2 blocks has landed" to CTA 0
#pragma unroll 1
            for (int g = 0; g < G; ++g) {
                mbar_wait(&w2c_full[g & 3], (uint32_t)(g >> 2) & 1u);
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&pw2_full[g & 3], 0);
            }
        } else
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            if (lane == 0) TC_DBG(5, c2);   // layer 2 of chunk c2: begins
            mbar_wait(&w2c_full[g & 3], (uint32_t)(g >> 2) & 1u);
            if (PAIR) mbar_wait_cluster(&pw2_full[g & 3], (uint32_t)(g >> 2) & 1u);
            if (c2 == 0 && g > 0) { if (PAIR) mbar_wait_cluster(d2_empty, ph_d2_empty); else mbar_wait(d2_empty, ph_d2_empty); ph_d2_empty ^= 1; }
#if PHN_TC_HRING
            const uint32_t bar_hqf = smem_u32(hq_free);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {   // quarter j: k-steps 2 j, 2 j + 1 of the chunk (k-block j >> 1 of the W2 chunk)
                const int lin = 4 * g + j, slot = lin % HR;
                const uint32_t use = (uint32_t)(lin / HR);
                if (PAIR) mbar_wait_cluster(&hq_full[slot], use & 1u); else mbar_wait(&hq_full[slot], use & 1u);
                if (lane == 0 && j == 0) TC_DBG(6, c2);   // first quarter of H + weights there
                tc_fence_after();
                const uint32_t b2lo = w2lo + w2s * (W2_ST >> 4) + (uint32_t)(j & 1) * 4u;
                const uint32_t ta = tH + (uint32_t)slot * 16u;
                const uint32_t acc0 = (j | c2) ? 1u : 0u;
                if (leader) {
                    if (PAIR) {
                        umma2_f16_ts(tD2, ta, ((uint64_t)kDescHi << 32) | b2lo, idesc2, acc0);
                        umma2_ts_lo<1>(tD2, ta + 8u, b2lo + 2, idesc2);
                        tc_commit2_u(bar_hqf + (uint32_t)slot * 8u);
                        if (j & 1) tc_commit2_u(bar_w2e + w2s * 8u);
                    } else {
                        umma_f16_ts(tD2, ta, ((uint64_t)kDescHi << 32) | b2lo, idesc2, acc0);
                        umma_ts_lo<1>(tD2, ta + 8u, b2lo + 2, idesc2);
                        tc_commit_u(bar_hqf + (uint32_t)slot * 8u);
                        if (j & 1) tc_commit_u(bar_w2e + w2s * 8u);
                    }
                }
                __syncwarp();
                if ((j & 1) && ++w2s == (uint32_t)p_s2) w2s = 0;
            }
            if (leader && c2 == p_nch - 1) { if (PAIR) tc_commit2_u(bar_d2f); else tc_commit_u(bar_d2f); }
            (void)bar_he;
#else
            if (PAIR) mbar_wait_cluster(h_full, (uint32_t)g & 1u); else mbar_wait(h_full, (uint32_t)g & 1u);
            if (lane == 0) TC_DBG(6, c2);   // H + weights there
            tc_fence_after();
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const uint32_t b2lo = w2lo + w2s * (W2_ST >> 4);
                const uint32_t ta = tH + (uint32_t)k * 32u;
                const uint32_t bw2 = bar_w2e + w2s * 8u;
                const uint32_t acc0 = (k | c2) ? 1u : 0u;
                if (leader) {
                    if (PAIR) {
                        umma2_f16_ts(tD2, ta, ((uint64_t)kDescHi << 32) | b2lo, idesc2, acc0);
                        umma2_ts_lo<1>(tD2, ta + 8u, b2lo + 2, idesc2);
                        umma2_ts_lo<1>(tD2, ta + 16u, b2lo + 4, idesc2);
                        umma2_ts_lo<1>(tD2, ta + 24u, b2lo + 6, idesc2);
                        tc_commit2_u(bw2);
                    } else {
                        umma_f16_ts(tD2, ta, ((uint64_t)kDescHi << 32) | b2lo, idesc2, acc0);
                        umma_ts_lo<1>(tD2, ta + 8u, b2lo + 2, idesc2);
                        umma_ts_lo<1>(tD2, ta + 16u, b2lo + 4, idesc2);
                        umma_ts_lo<1>(tD2, ta + 24u, b2lo + 6, idesc2);
                        tc_commit_u(bw2);
                    }
                }
                __syncwarp();
                if (++w2s == (uint32_t)p_s2) w2s = 0;
            }
            if (leader) {
                if (PAIR) { tc_commit2_u(bar_he); if (c2 == p_nch - 1) tc_commit2_u(bar_d2f); }   // H goes back to the epilogue warps
                else      { tc_commit_u(bar_he);  if (c2 == p_nch - 1) tc_commit_u(bar_d2f); }
            }
#endif
            __syncwarp();
            if (lane == 0) TC_DBG(7, c2);   // issued
            if (++c2 == p_nch) { c2 = 0; tile += tstep; }
        }
    } else {
        // ===================================================================== epilogue warps
        // E1 runs once per chunk; E2 of a tile is cut into three stages that ride behind the E1s of the NEXT tile's
        // first chunks (A: drain D2 into registers and hand D2 back, row max; B: exponentials, row sum; C: outputs),
        // so the tensor pipe never waits for a soft-max.  After the CTA's last chunk the stages run back to back.
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int cq = (warp - EPI0) >> 2;       // column quarter 0..3
        const int row = q * 32 + lane;           // tile row == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        uint32_t ph_d1_full0 = 0, ph_d1_full1 = 0, ph_h_empty = 0, ph_d2_full = 0;
        bool first_h = true;
        // H[row][cq*32 .. +31] (fp16 pairs) = TMEM lane `row`, columns cq*16 .. +15 of the H operand
        // u = sat(t / TMAX), t = -(x + b1)/ln2 + Ct; b1 is already inside x (two extra K columns of layer 1)
        const float sigA = (float)(-1.0 / (kLn2 * (double)kSigTmax)), sigB = (float)(kCt / (double)kSigTmax);
        const float smxA = (float)(1.0 / (kLn2 * ((double)kSmxTmax - 1.0))), smxB = (float)((kCt + kSmxOff - 1.0) / ((double)kSmxTmax - 1.0));
        const float smxK = (kSmxTmax - 1.0f) * 16384.0f, smxM = 8388608.0f + 16384.0f;
        const int my_tiles = n_my;
        const int G = my_tiles * a.NCH;
        // the consumers of these three signals are the issuer warps: this CTA's, or (PAIR) CTA 0's for both CTAs
        auto signal = [&](uint64_t *bar) { if (PAIR) mbar_arrive_cluster(bar, 0); else mbar_arrive(bar); };
        const int n0 = cq * NQ;                  // column quarter cq owns D2 columns [cq*NQ, (cq+1)*NQ)
        float o[NQ];
        float mx = 0.0f;
        int e2_stage = 3, e2_tile = 0;           // 3: no soft-max pending; 0..2: next stage of tile e2_tile
        int c = 0, tile = tile0;
        for (int g = 0; g <= G; ++g) {           // g == G: drain iteration (no E1)
            const bool drain = g == G;
            if (!drain) {
                const int b = g & 1;             // D1 buffer = linear chunk index & 1 (NCH may be odd)
                if (threadIdx.x == EPI0 * 32) TC_DBG(8, c);    // e1: waiting for D1
                if (b == 0) { mbar_wait(&d1_full[0], ph_d1_full0); ph_d1_full0 ^= 1; }
                else        { mbar_wait(&d1_full[1], ph_d1_full1); ph_d1_full1 ^= 1; }
                tc_fence_after();
                if (threadIdx.x == EPI0 * 32) TC_DBG(9, c);    // e1: D1 seen
                // fsig(x) = 1 / (1 + D(-x - b1)); two reciprocals share one MUFU: 1/a = a' / (a a'), 1/a' = a / (a a')
                uint32_t hp[16];
#if PHN_TC_E1_LD32
                // both halves of the warp's 32 accumulator columns in one TMEM load; the accumulator buffer goes back to the
                // layer-1 issuer before the arithmetic starts
                uint32_t acc32[32];
                tmem_ld32((b ? tD1[1] : tD1[0]) + lane_addr + cq * 32, acc32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) signal(&d1_empty[b]);
#endif
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#if PHN_TC_E1_LD32
                    const uint32_t *acc = acc32 + 16 * half;
#else
                    uint32_t acc[16];
                    tmem_ld16((b ? tD1[1] : tD1[0]) + lane_addr + cq * 32 + half * 16, acc);
                    tmem_ld_wait();
#if PHN_TC_D1_EARLY
                    if (half == 1) {   // the accumulator buffer is in registers: the layer-1 issuer may overwrite it while the second
                        tc_fence_before();   // half's sigmoids are still being computed (half a chunk of slack for the next D1)
                        __syncwarp();
                        if (lane == 0) signal(&d1_empty[b]);
                    }
#endif
#endif
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
#if PHN_TC_PACKED
                        // four sigmoids: elements (0, 2) and (1, 3) travel as packed pairs; 1/a0 = a1 / (a0 a1), 1/a1 = a0 / (a0 a1)
                        const f2 uA = {fma_sat(__uint_as_float(acc[g4 * 4 + 0]), sigA, sigB), fma_sat(__uint_as_float(acc[g4 * 4 + 2]), sigA, sigB)};
                        const f2 uB = {fma_sat(__uint_as_float(acc[g4 * 4 + 1]), sigA, sigB), fma_sat(__uint_as_float(acc[g4 * 4 + 3]), sigA, sigB)};
                        const f2 kk = {kSigTmax * 16384.0f, kSigTmax * 16384.0f}, mg = {8388608.0f, 8388608.0f}, one = {1.0f, 1.0f};
                        const f2 rA = ffma2(uA, kk, mg), rB = ffma2(uB, kk, mg);
                        const f2 dA = {__uint_as_float(shl9(__float_as_uint(rA.x))), __uint_as_float(shl9(__float_as_uint(rA.y)))};
                        const f2 dB = {__uint_as_float(shl9(__float_as_uint(rB.x))), __uint_as_float(shl9(__float_as_uint(rB.y)))};
                        const f2 aA = fadd2(dA, one), aB = fadd2(dB, one);
                        const f2 pr = fmul2(aA, aB);
#if PHN_TC_RCP4
                        // one reciprocal for all four: 1 / (a0 a1) = (a2 a3) / (a0 a1 a2 a3)   (a <= 1 + 2^31: the product stays finite)
                        const float r4 = rcp_approx(pr.x * pr.y);
                        const f2 rc = fmul2(f2{r4, r4}, f2{pr.y, pr.x});
#else
                        const f2 rc = {rcp_approx(pr.x), rcp_approx(pr.y)};
#endif
                        const f2 hA = fmul2(rc, aB), hB = fmul2(rc, aA);          // (h0, h2), (h1, h3)
                        hp[half * 8 + g4 * 2] = pack_half2(hA.x, hB.x);
                        hp[half * 8 + g4 * 2 + 1] = pack_half2(hA.y, hB.y);
#else
                        const float a0 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 0]), sigA, sigB), kSigTmax);
                        const float a1 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 1]), sigA, sigB), kSigTmax);
                        const float a2 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 2]), sigA, sigB), kSigTmax);
                        const float a3 = 1.0f + fexp_from_u(fma_sat(__uint_as_float(acc[g4 * 4 + 3]), sigA, sigB), kSigTmax);
                        const float r01 = rcp_approx(a0 * a1), r23 = rcp_approx(a2 * a3);
                        hp[half * 8 + g4 * 2] = pack_half2(r01 * a1, r01 * a0);
                        hp[half * 8 + g4 * 2 + 1] = pack_half2(r23 * a3, r23 * a2);
#endif
                    }
                }
#if !PHN_TC_E1_LD32 && !PHN_TC_D1_EARLY
                tc_fence_before();
                __syncwarp();
                if (lane == 0) signal(&d1_empty[b]);      // (the layer-1 issuer may overwrite this accumulator buffer)
#endif
                if (threadIdx.x == EPI0 * 32) TC_DBG(10, c);   // e1: math done
#if PHN_TC_HRING
                {
                    const int lin = 4 * g + cq, slot = lin % HR;
                    const int use = lin / HR;
                    if (use > 0) mbar_wait(&hq_free[slot], (uint32_t)(use - 1) & 1u);   // (the slot's previous readers have completed)
                    if (threadIdx.x == EPI0 * 32) TC_DBG(11, c);   // e1: H slot free
                    tc_fence_after();
                    tmem_st16(tH + lane_addr + (uint32_t)slot * 16u, hp);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) signal(&hq_full[slot]);
                    (void)first_h; (void)ph_h_empty;
                }
#else
                if (!first_h) { mbar_wait(h_empty, ph_h_empty); ph_h_empty ^= 1; }
                first_h = false;
                if (threadIdx.x == EPI0 * 32) TC_DBG(11, c);   // e1: H buffer free
                tc_fence_after();
                tmem_st16(tH + lane_addr + cq * 16, hp);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) signal(h_full);
#endif
                if (threadIdx.x == EPI0 * 32) TC_DBG(12, c);   // e1: H published
            }
            // ------------------------------------------------------------- E2 stages: soft-max + outputs of tile e2_tile
            // one stage per chunk; everything that is left when a tile ends or in the drain iteration
            int budget = (drain || c == a.NCH - 1) ? 3 : 1;
            while (e2_stage < 3 && budget-- > 0) {
                if (e2_stage == 0) {
                    // A: D2 -> registers (+ b2), D2 handed back to the issuer, row max of this warp's columns
                    TC_DBG2(0);
                    mbar_wait(d2_full, ph_d2_full); ph_d2_full ^= 1;
                    tc_fence_after();
                    TC_DBG2(1);
                    uint32_t raw[NQ];
                    tmem_ld32(tD2 + lane_addr + n0, raw);
                    if (NR == 4) tmem_ld4(tD2 + lane_addr + n0 + 32, raw + 32);
                    if (NR == 8) tmem_ld8(tD2 + lane_addr + n0 + 32, raw + 32);
                    if (NR == 16) tmem_ld16(tD2 + lane_addr + n0 + 32, raw + 32);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) signal(d2_empty);
#pragma unroll
                    for (int j = 0; j < NQ / 4; ++j) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(s_b2 + n0 + 4 * j);   // -FLT_MAX in padding columns
#if PHN_TC_PACKED
                        const f2 s01 = fadd2(f2{__uint_as_float(raw[4 * j + 0]), __uint_as_float(raw[4 * j + 1])}, f2{b4.x, b4.y});
                        const f2 s23 = fadd2(f2{__uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])}, f2{b4.z, b4.w});
                        o[4 * j + 0] = s01.x; o[4 * j + 1] = s01.y; o[4 * j + 2] = s23.x; o[4 * j + 3] = s23.y;
#else
                        o[4 * j + 0] = __uint_as_float(raw[4 * j + 0]) + b4.x;
                        o[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + b4.y;
                        o[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + b4.z;
                        o[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + b4.w;
#endif
                    }
                    mx = o[0];
#pragma unroll
                    for (int i = 1; i < NQ; ++i) mx = fmaxf(mx, o[i]);
                    s_red[cq * 128 + row] = mx;
                    TC_DBG2(2);
                } else if (e2_stage == 1) {
                    // B: row max across the 4 column-quarter warps of this lane quarter, e = D(o - max) (fexp.h:49-78), row sum
                    TC_DBG2(3);
                    quarter_bar_sync(q);
                    mx = fmaxf(fmaxf(s_red[row], s_red[128 + row]), fmaxf(s_red[256 + row], s_red[384 + row]));
#if PHN_TC_PACKED
                    f2 sum2 = {0.0f, 0.0f};
                    const f2 nmx = {-mx, -mx}, kk = {smxK, smxK}, mg = {smxM, smxM};
#pragma unroll
                    for (int i = 0; i < NQ; i += 2) {
                        const f2 d = fadd2(f2{o[i], o[i + 1]}, nmx);
                        const f2 u = {fma_sat(d.x, smxA, smxB), fma_sat(d.y, smxA, smxB)};
                        const f2 r = ffma2(u, kk, mg);
                        const f2 e = {__uint_as_float(shl9(__float_as_uint(r.x))), __uint_as_float(shl9(__float_as_uint(r.y)))};
                        o[i] = e.x; o[i + 1] = e.y;
                        sum2 = fadd2(sum2, e);
                    }
                    const float sum = sum2.x + sum2.y;
#else
                    float sum = 0.0f;
#pragma unroll
                    for (int i = 0; i < NQ; ++i) {
                        const float e = __uint_as_float(__float_as_uint(fmaf(fma_sat(o[i] - mx, smxA, smxB), smxK, smxM)) << 9);
                        o[i] = e;
                        sum += e;
                    }
#endif
                    s_red[512 + cq * 128 + row] = sum;
                    TC_DBG2(4);
                } else {
                    // C: normalise, outputs
                    TC_DBG2(5);
                    quarter_bar_sync(q);
                    const float sum = (s_red[512 + row] + s_red[512 + 128 + row]) + (s_red[512 + 256 + row] + s_red[512 + 384 + row]);
                    const int64_t f = (int64_t)e2_tile * TC_M + row;
                    if (f < a.nf) {
                        if (!a.xm_img) {
                            if (a.post) {
                                const float sc = 1.0f / sum;
                                float *dst = a.post + f * a.ldpost + n0;
#pragma unroll
                                for (int j = 0; j < NQ / 4; ++j)
                                    if (n0 + 4 * j < a.ldpost)
                                        *reinterpret_cast<float4 *>(dst + 4 * j) = make_float4(o[4 * j] * sc, o[4 * j + 1] * sc, o[4 * j + 2] * sc, o[4 * j + 3] * sc);
                            }
                            if (a.logp) {   // decoder soft function (srec.cpp:1088-1097), fused: ln p for the token passing kernel,
                                            // column-major inside the tile so that a warp's store of one column is one 128-byte line
                                float *ldst = a.logp + ((a.logp_tile0 + e2_tile) * a.ldpost + n0) * TC_M + row;
                                const float ln_sc = -__logf(sum);
#pragma unroll
                                for (int i = 0; i < NQ; ++i)
                                    if (n0 + i < a.ldpost) ldst[i * TC_M] = fmaf(lg2_approx(o[i]), (float)kLn2, ln_sc);
                            }
                        } else {
                            // merger input: sLn(p), merger input normalisation, fp16, into the merger's M-major X image: a warp's
                            // store of one column is 64 consecutive bytes
                            const int nlim = (a.nout + 7) & ~7;   // this net's share of the image: nout rounded up to 8 columns
                            uint8_t *img = a.xm_img + (size_t)e2_tile * a.xm_kb1 * TC_BLK;
                            const float lg2_sc = -lg2_approx(sum);
                            const int c0 = a.xm_col0 + n0;
                            if (c0 & 4) band_out<NQ, 4>(o, lg2_sc, s_mm + n0, s_md + n0, img, row, c0, nlim - n0);
                            else        band_out<NQ, 0>(o, lg2_sc, s_mm + n0, s_md + n0, img, row, c0, nlim - n0);
                        }
                    }
                    TC_DBG2(6);
                }
                ++e2_stage;
            }
            if (!drain && ++c == a.NCH) {        // the tile's last H is on its way: its soft-max becomes pending
                c = 0;
                e2_stage = 0; e2_tile = tile;
                tile += tstep;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // the peer is done with this CTA's barriers, shared memory and tensor memory
    if (warp == WARP_MMA1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
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
    // (only the merger's image is filled this way, and the merger's image is M-major)
    *reinterpret_cast<__half *>(img + (size_t)(row >> 7) * KB1 * TC_BLK + mn128_off((int)(row & 127), k)) = __float2half_rn(1.0f);
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

template <int N2P, bool XMN, bool PAIR, bool DBG>
static int launch_inst_k(phn_ctx *c, const TcArgs &a, size_t smem_bytes, int grid)
{
    PHN_CUDA(c, cudaFuncSetAttribute(k_mlp_tc<N2P, XMN, PAIR, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PHN_CUDA(c, cudaLaunchKernelEx(&cfg, k_mlp_tc<N2P, XMN, PAIR, DBG>, a));
    return PHN_OK;
}

template <int N2P, bool XMN, bool PAIR>
static int launch_inst(phn_ctx *c, const TcArgs &a, size_t smem_bytes, int grid)
{
    // (the timeline instantiation only when tools/tc_timeline.py asked for one on this net)
    if (a.dbg) return launch_inst_k<N2P, XMN, PAIR, true>(c, a, smem_bytes, grid);
    return launch_inst_k<N2P, XMN, PAIR, false>(c, a, smem_bytes, grid);
}

template <int N2P>
static int launch_one(phn_ctx *c, const TcArgs &a, size_t smem_bytes, int grid, bool pair)
{
    int rc;
    if (pair) rc = a.x_mn ? launch_inst<N2P, true, true>(c, a, smem_bytes, grid) : launch_inst<N2P, false, true>(c, a, smem_bytes, grid);
    else      rc = a.x_mn ? launch_inst<N2P, true, false>(c, a, smem_bytes, grid) : launch_inst<N2P, false, false>(c, a, smem_bytes, grid);
    if (rc) return rc;
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
        // audio -> labels path: the decoder only needs ln p, so the linear posteriors are not written at all
        a.post = c->fuse_logp ? nullptr : (float *)c->d_post.p + f0 * c->ldp; a.ldpost = c->ldp;
        a.logp = c->fuse_logp ? (float *)c->d_logp.p : nullptr;
        a.logp_tile0 = f0 / TC_M;
        a.x_mn = 1;
    }
    // shared memory plan: X (KB1 blocks) + H (2 blocks) + constants + barriers are fixed; the rest is split
    // between the W2 ring (S2 k-blocks of N2P x 64) and the W1 ring (S1 blocks of 128 x 64)
    // CTA pairs (cta_group::2): each CTA keeps half of every weight block.  PHNREC_TC_PAIR=0 selects the single-CTA kernel.
    bool pair = a.n_tiles >= 2;
    if (const char *e = getenv("PHNREC_TC_PAIR")) pair = pair && atoi(e) != 0;
    const size_t w1_blk = pair ? TC_BLK / 2 : TC_BLK;
    const size_t w2_blk = (size_t)im.N2P * 128 / (pair ? 2 : 1);
    // X ring: a spare slot where it is cheap (narrow nets: 16 KB out of >= 3 chunks of weight ring), none for the merger,
    // whose ring is short already.  PHNREC_TC_XSPARE=0/1 overrides (kernel development).
    int xspare = pair && im.KB1 <= 3;
    if (const char *e = getenv("PHNREC_TC_XSPARE")) xspare = atoi(e) != 0 && im.KB1 < 8;
    a.XR = im.KB1 + (xspare ? 1 : 0);
    const size_t fixed = (size_t)a.XR * TC_BLK + sizeof(float) * (3 * (size_t)im.N2P + 8 * 128) + 128 * 8 + 1024;
    const size_t max_smem = 232448;
    // Ring plan.  Both weight streams want two chunks resident (the one being multiplied and the one in flight):
    // W2 ring 4 stages, W1 ring 2 KB1 stages.  When that does not fit (wide merger next to its wide X tile), W2
    // falls back to 3 and then 2 stages and W1 takes whatever is left (at least 2 stages; the issuer then walks a
    // chunk in groups, see k_mlp_tc).
    // Pairs have the room for deeper rings: a weight block arrives ~3000 clk after its request when every SM streams
    // the whole net out of L2, more than one chunk time - so up to four chunks are kept in flight.
    int S2 = pair ? TC_MAXS2 : 4;
    if (pair && fixed + S2 * w2_blk + 3 * (size_t)im.KB1 * w1_blk > max_smem) S2 = 6;
    if (pair && fixed + S2 * w2_blk + 2 * (size_t)im.KB1 * w1_blk > max_smem) S2 = 4;
    if (fixed + S2 * w2_blk + 2 * (size_t)im.KB1 * w1_blk > max_smem) S2 = 3;
    if (fixed + S2 * w2_blk + (size_t)(im.KB1 + 1) * w1_blk > max_smem) S2 = 2;
    if (fixed + S2 * w2_blk + 2 * (size_t)w1_blk > max_smem) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: shared memory plan does not fit\n");
    int S1 = (int)((max_smem - fixed - S2 * w2_blk) / w1_blk);
    if (S1 > TC_MAXS1) S1 = TC_MAXS1;
    if (S1 > 4 * im.KB1) S1 = 4 * im.KB1;   // (four chunk-level "full" barriers)
    if (pair && S1 < im.KB1) return fail(c, PHN_ERR_UNSUPPORTED, "tensor-core mode: shared memory plan does not fit (pair)\n");
    a.S1 = S1; a.S2 = S2;
    const size_t smem_bytes = fixed + (size_t)S1 * w1_blk + (size_t)S2 * w2_blk;
    int grid = a.n_tiles < c->num_sms ? a.n_tiles : c->num_sms;
    if (pair) {
        const int npairs = (a.n_tiles + 1) / 2, maxp = c->num_sms / 2;
        grid = 2 * (npairs < maxp ? npairs : maxp);
    }
    if (const char *e = getenv("PHNREC_TC_GRID")) {  // debugging aid: force several tiles per CTA
        const int g = atoi(e);
        if (g > 0 && g < grid) grid = g;
        if (pair) grid = grid < 2 ? 2 : (grid & ~1);
    }
    int rc;
    switch (im.N2P) {
        case 128: rc = launch_one<128>(c, a, smem_bytes, grid, pair); break;
        case 144: rc = launch_one<144>(c, a, smem_bytes, grid, pair); break;
        case 160: rc = launch_one<160>(c, a, smem_bytes, grid, pair); break;
        default: rc = launch_one<192>(c, a, smem_bytes, grid, pair); break;
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
