// k_wave_tc16.cu — K-wave for the tensor-core pipeline, 16 kHz systems (16-bit linear input, 25 ms window = 400 samples, 512-point
// transform, 256 bins): the windowed DFT of every frame as a GEMM on tcgen05, like k_wave_tc.cu for the 8 kHz systems.
//
// Replaces ConvertWaveformFormat (srec.cpp:709-791), MelBanks::ProcessFrame (melbanks.cpp:111-204), cFour1 / _mbApply
// (dspc.cpp:24-78, 236-269), cPower / sLn (dspc.h:141-160) and FrameBasedNormalization (srec.cpp:1594-1620).
//
// The [400 x 512] matrix (hi + lo halves: 819 KB) does not fit beside the operands, so the product is cut four ways and the machinery
// of k_wave_tc.cu runs over the pieces: two PASSES (bins 0..127, 128..255: N = 256 each, one accumulator each - the epilogue of
// pass 0 runs under the MMAs of pass 1), two K-HALVES (samples 0..199, 200..399 of the window: the A stage of the 8 kHz kernel,
// 208 columns) that accumulate into the pass's accumulator, and for each of them the two PARTS of a 16-bit sample (fp16(sample)
// against W_hi and W_lo, the rounding error against W_hi).  Per tile: four PRODUCTIONS of an A stage - (khalf 0, hi), (khalf 1, hi),
// (khalf 0, lo), (khalf 1, lo); stage = khalf (two stages) - each serving two PIECES (pass 0, pass 1) = 8 pieces, 156 MMAs.
// The matrix of a piece's (pass, khalf) - 7 blocks of 16 KB per CTA - comes from L2 by bulk copies into a buffer of three units
// (W_hi blocks, W_lo blocks, tail block) with a full / free barrier pair each: a unit is refilled for its next piece while the
// MMAs of the other units run (lo pieces need W_hi only and alternate between the two block units).
// Producers (8 warps, 16 rows each): lane = 16-byte chunk of a row, the words of 8 rows requested before the first is decoded (rows
// are 20 chunks apart: little to share between them); per tile a loop without bounds / lengths when all of a warp's rows are whole
// windows inside the buffer.  Epilogue: the filterbank walk of k_wave_tc.cu over 256 bins, its state carried across the two
// passes of a tile in registers; weights in shared memory, shift / store masks as kernel parameters.
// History on 998 000 frames: 2.72 ms (eight productions per tile) -> 1.91 (a K-half's stages serve both passes) -> 1.30 (8 producer
// warps) -> 1.18 (fast producer loop) -> 1.10 ms (matrix units refilled under the MMAs); register FFT: 2.28 ms.
#include "internal.h"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace phn {

constexpr int W16_KH = 200;               // samples of the window per K-half
constexpr int W16_KC = 26;                // 16-byte chunks (8 samples) of a K-half fed to the tensor cores (208 columns)
constexpr int W16_NBIN = 256;             // bins of the 512-point transform
constexpr int W16_NCH = W16_NBIN / 16;    // chunks of 16 bins
constexpr int W16_PROD = 8, W16_EPI = 8;
constexpr int W16_RPW = 128 / W16_PROD;     // rows of a CTA's half tile per producer warp
constexpr int W16_THREADS = (W16_EPI + 1 + W16_PROD) * 32;
constexpr int W16_BLK = 16384;
constexpr int W16_BBLK = 7;
constexpr int W16_HCH = 10;               // chunks one epilogue half may walk
constexpr size_t W16_SMEM = (size_t)(W16_BBLK + 6 + 1) * W16_BLK + 160 + 2 * W16_HCH * 128;   // the base must be 1024-byte aligned (checked)

struct Wave16Half {
    uint32_t shift[W16_NBIN / 32], emit[W16_NBIN / 32];
    int c_begin, c_end, cur0, flush_lo, flush_hi, z_begin, z_end;
};

struct Wave16Args {
    const uint8_t *audio, *audio_end;
    const int64_t *byte_off, *frame_off;
    int n_utt;
    int64_t f_begin, f_end;
    int vs, step, nbanks;
    float frame_shift, frame_floor;
    const uint8_t *w_img;                 // [4 pieces = 2 pass + khalf][2 ranks][7][16 KB]
    const float4 *wtab;                   // [2 halves][W16_HCH][8]: wlo (4 x float4) then whi of the chunks a half walks
    float *mel;
    Wave16Half h[2];
};

__device__ __noinline__ uint32_t w16_word_tail(const uint8_t *p, const uint8_t *end)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        if (p + i < end) r |= (uint32_t)p[i] << (8 * i);
    return r;
}
__device__ __forceinline__ void w16_st_if(float *p, float v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((uint32_t)pred) : "memory");
}
__device__ __forceinline__ float w16_lg2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ int w16_find_utt(const int64_t *off, int n, int64_t f)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(W16_THREADS, 1) k_wave_tc16(const __grid_constant__ Wave16Args a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    if (smem_u32(smem) & 1023u) __trap();                 // (the swizzle pattern is a function of the address)
    uint8_t *sB = smem;
    uint8_t *sA = sB + (size_t)W16_BBLK * W16_BLK;        // 2 stages x 3 blocks
    uint8_t *sT = sA + (size_t)6 * W16_BLK;               // the stages' shared tail block
    uint64_t *bars = reinterpret_cast<uint64_t *>(sT + W16_BLK);
    uint64_t *a_full = bars, *a_empty = bars + 2, *d_full = bars + 4, *d_empty = bars + 6;
    uint64_t *m_full = bars + 8, *m_pfull = bars + 11, *m_free = bars + 14;   // matrix units X0, X1, X2 (see the MMA warp)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 17);
    float4 *s_w = reinterpret_cast<float4 *>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;
    const int64_t nf = a.f_end - a.f_begin;
    const int64_t n_units = (nf + 255) / 256, ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
    const int64_t unit0 = cl * n_units / ncl;
    const int n_my = (int)((cl + 1) * n_units / ncl - unit0);
    constexpr int WARP_MMA = W16_EPI, PROD0 = W16_EPI + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 2 * W16_PROD); mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 2 * W16_EPI);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(&m_full[i], 1); mbar_init(&m_pfull[i], 1); mbar_init(&m_free[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * W16_HCH * 8; i += blockDim.x) s_w[i] = a.wtab[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == WARP_MMA) {
        // ===================================================================== matrix pieces + MMA issue
        constexpr uint32_t idesc = make_idesc(256, false, 256);
        const uint64_t dA = make_sw128_desc(smem_u32(sA)), dB = make_sw128_desc(smem_u32(sB)), dT = make_sw128_desc(smem_u32(sT));
        const uint32_t alo0 = (uint32_t)dA, blo0 = (uint32_t)dB, tlo0 = (uint32_t)dT, hi = (uint32_t)(dA >> 32);
        const uint32_t bar_ae = smem_u32(a_empty), bar_df = smem_u32(d_full), bar_fr = smem_u32(m_free);
        const bool leader = elect_one();
        // A tile has four PRODUCTIONS of an A stage - (khalf 0, hi), (khalf 1, hi), (khalf 0, lo), (khalf 1, lo); hi = fp16(sample),
        // lo = its rounding error; stage = khalf - and each serves two PIECES (pass 0, pass 1): a piece multiplies the stage by the
        // matrix of (pass, khalf) into its pass's accumulator (hi against W_hi and W_lo, lo against W_hi).
        // The matrix buffer is three UNITS with a full / free barrier pair each - X0 = blocks 0..2, X1 = blocks 3..5, X2 = the tail
        // block - and a unit is reloaded as soon as the MMAs that read it have completed, while those of the other units run:
        //   hi piece: X0 = W_hi, X1 = W_lo, X2 = both tails (12 + 12 + 2 MMAs); lo pieces need W_hi only and alternate between
        //   X0 and X1 for it (12 + 1 MMAs).  Uses of a unit per tile: X0 in pieces 0 1 2 3 4 6, X1 in 0 1 2 3 5 7, X2 in all eight
        //   (even counts: the barrier parity of a use depends on the piece's place in the tile only).
        const int n_piece = 8 * n_my;
        auto use_parity = [](int X, int q) -> uint32_t { return (uint32_t)((X == 2 || q < 4 ? q : q >> 1) & 1); };
        auto load_unit = [&](int X, int pc) {   // the contents piece pc needs in unit X (the caller has waited for the unit to be free)
            if (!leader) return;
            const int q = pc & 7, pass = q & 1, kh = (q >> 1) & 1;
            const bool low = q >= 4;
            const uint8_t *src = a.w_img + ((size_t)((pass * 2 + kh) * 2 + (int)rank) * W16_BBLK) * W16_BLK;
            if (X == 2) {
                mbar_expect_tx(&m_full[2], W16_BLK);
                tma_load_1d(sB + (size_t)6 * W16_BLK, src + (size_t)6 * W16_BLK, W16_BLK, &m_full[2]);
            } else {
                const int sb = (X == 1 && !low) ? 3 : 0;                 // W_lo for a hi piece's X1, W_hi otherwise
                mbar_expect_tx(&m_full[X], 3 * W16_BLK);
                for (int b = 0; b < 3; ++b) tma_load_1d(sB + (size_t)(3 * X + b) * W16_BLK, src + (size_t)(sb + b) * W16_BLK, W16_BLK, &m_full[X]);
            }
        };
        auto next_use = [](int X, int pc) {     // the next piece that reads unit X after piece pc
            const int q = pc & 7, base = pc - q;
            if (X == 2 || q < 3) return pc + 1;
            if (q == 3) return base + (X == 0 ? 4 : 5);
            return q >= 6 ? base + 8 : pc + 2;
        };
        if (n_piece > 0) { load_unit(0, 0); load_unit(1, 0); load_unit(2, 0); }
        int prevX = -1, prev_pc = 0;
#pragma unroll 1
        for (int pc = 0; pc < n_piece; ++pc) {
            const int q = pc & 7, tile = pc >> 3, pi = q >> 1, pass = q & 1, kh = pi & 1;
            const bool low = pi >> 1;
            const int n_units = low ? 2 : 3;
#pragma unroll 1
            for (int k = 0; k < n_units; ++k) {
                const int X = low ? (k == 0 ? (q & 1) : 2) : k;
                const uint32_t par = use_parity(X, q);
                mbar_wait(&m_full[X], par);
                if (rank != 0) {
                    if (lane == 0) mbar_arrive_cluster(&m_pfull[X], 0);
                } else {
                    mbar_wait_cluster(&m_pfull[X], par);
                    if (k == 0) {
                        if (pass == 0) mbar_wait_cluster(&a_full[kh], low ? 1u : 0u);   // (two productions per stage and tile: parity = part)
                        if (pi == 0 && tile >= 1) mbar_wait_cluster(&d_empty[pass], ((uint32_t)tile & 1u) ^ 1u);
                    }
                    tc_fence_after();
                    if (leader) {
                        const uint32_t td = tmem + 256u * (uint32_t)pass;
                        const uint32_t alo = alo0 + (uint32_t)kh * (3u * (W16_BLK >> 4));   // A stage = K-half
                        if (X < 2) {
                            const uint32_t blo = blo0 + (uint32_t)X * (3u * (W16_BLK >> 4));
#pragma unroll
                            for (int b = 0; b < 3; ++b)
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t o = (uint32_t)b * (W16_BLK >> 4) + 2u * (uint32_t)ks;
                                    if (k == 0 && b == 0 && ks == 0 && pi == 0) umma2_ss_lo<0>(td, alo + o, hi, blo + o, idesc);
                                    else umma2_ss_lo<1>(td, alo + o, hi, blo + o, idesc);
                                }
                        } else {
                            umma2_ss_lo<1>(td, tlo0 + 2u * (uint32_t)kh, hi, blo0 + 6u * (W16_BLK >> 4), idesc);
                            if (!low) umma2_ss_lo<1>(td, tlo0 + 2u * (uint32_t)kh, hi, blo0 + 6u * (W16_BLK >> 4) + 2u, idesc);
                        }
                        tc_commit2_u(bar_fr + 8u * (uint32_t)X);
                        if (X == 2) {                                    // the piece's last unit
                            if (pass == 1) tc_commit2_u(bar_ae + 8u * (uint32_t)kh);
                            if (pi == 3) tc_commit2_u(bar_df + 8u * (uint32_t)pass);
                        }
                    }
                }
                __syncwarp();
                // the unit issued before this one: once its MMAs have completed (those just issued are running), its next contents
                if (prevX >= 0) {
                    const int nx = next_use(prevX, prev_pc);
                    if (nx < n_piece) {
                        mbar_wait(&m_free[prevX], use_parity(prevX, prev_pc & 7));
                        load_unit(prevX, nx);
                    }
                }
                prevX = X; prev_pc = pc;
                __syncwarp();
            }
        }
    } else if (warp >= PROD0) {
        // ===================================================================== producers: audio -> A stage
        const int pw = warp - PROD0;
        int u = 0;
        int64_t fo_cur = 0, fo_next = -1, b0 = 0, len = 0;
        int64_t soff = 0, soff0 = 0;
        uint32_t soff32 = 0;
        bool fast = false;
        int lim = -1;
        auto st_shared = [](uint32_t addr, const uint4 &v) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        };
#pragma unroll 1
        for (int na = 0; na < 4 * n_my; ++na) {                          // productions: (khalf 0, hi), (khalf 1, hi), (khalf 0, lo), (khalf 1, lo)
            const int tile = na >> 2, pi = na & 3, kh = pi & 1;
            const bool lowp = pi >> 1;
            if (pi == 0) {   // this lane's row of the tile: byte offset of its first sample, samples of its window inside the signal
                const int64_t g = a.f_begin + ((unit0 + tile) * 256 + (int64_t)rank * 128) + pw * W16_RPW + (lane & (W16_RPW - 1));
                soff = 0; lim = -1;
                if (g < a.f_end) {
                    if (g >= fo_next) {
                        if (fo_next >= 0 && u + 2 <= a.n_utt && g < a.frame_off[u + 2]) ++u; else u = w16_find_utt(a.frame_off, a.n_utt, g);
                        fo_cur = a.frame_off[u]; fo_next = a.frame_off[u + 1];
                        b0 = a.byte_off[u]; len = (a.byte_off[u + 1] - b0) / 2;
                    }
                    const int64_t s0 = (g - fo_cur) * a.step, left = len - s0;
                    soff = b0 + 2 * s0;
                    lim = left < a.vs ? (left < 0 ? 0 : (int)left) : a.vs;
                }
                // the common case - every row of this warp a whole 400-sample window, all reads inside the buffer: a loop without
                // per-row bounds, lengths and 64-bit shuffles (offsets from the warp's first row)
                soff0 = __shfl_sync(0xffffffffu, soff, 0);
                fast = __all_sync(0xffffffffu, lim == 2 * W16_KH && soff >= soff0 && soff - soff0 < (int64_t)1 << 30 &&
                                                   a.audio + soff + 2 * (2 * W16_KH) + 4 <= a.audio_end);
                soff32 = (uint32_t)(soff - soff0);
            }
            bool waited = na < 2;                                        // (a stage's first production has nothing to wait for)
            if (fast) {
                const uint8_t *wsrc = a.audio + soff0 + 2 * (W16_KH * kh + 8 * lane);
                const bool ld = lane < W16_KH / 8;                       // chunks 0..24 hold samples, chunk 25 is the K padding
#pragma unroll 1
                for (int gi = 0; gi < W16_RPW / 8; ++gi) {
                    uint32_t t[8][5];
                    uint32_t shr[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const uint8_t *src = wsrc + __shfl_sync(0xffffffffu, soff32, 8 * gi + rr);
                        const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
                        shr[rr] = ((uint32_t)reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
#pragma unroll
                        for (int k = 0; k < 5; ++k) t[rr][k] = ld ? __ldg(p + k) : 0u;
                    }
                    if (!waited) { mbar_wait(&a_empty[kh], lowp ? 0u : 1u); waited = true; }
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        uint4 oh;
                        uint32_t *ph_ = &oh.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t v = __funnelshift_r(t[rr][k], t[rr][k + 1], shr[rr]);
                            const uint32_t xh = __byte_perm(v, 0x64646464u, 0x4341u) ^ 0x00800080u, xl = __byte_perm(v, 0x64646464u, 0x4240u);
                            const __half2 th = __hmul2(__hadd2(*reinterpret_cast<const __half2 *>(&xh), __float2half2_rn(-1152.0f)), __float2half2_rn(256.0f));
                            const __half2 lf = __hadd2(*reinterpret_cast<const __half2 *>(&xl), __float2half2_rn(-1024.0f));
                            const __half2 hh = __hadd2(th, lf);
                            const __half2 r = lowp ? __hsub2(lf, __hsub2(hh, th)) : hh;
                            ph_[k] = ld ? *reinterpret_cast<const uint32_t *>(&r) : 0u;
                        }
                        const int r = 8 * gi + rr, c = lane;
                        if (lane < W16_KC) {
                            const uint32_t rowoff = (uint32_t)(pw * W16_RPW + r) * 128u;
                            if (c < 24) st_shared(smem_u32(sA) + (uint32_t)(3 * kh + (c >> 3)) * W16_BLK + rowoff + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4), oh);
                            else st_shared(smem_u32(sT) + rowoff + (((2u * (uint32_t)kh + (uint32_t)(c & 7)) ^ ((uint32_t)r & 7u)) << 4), oh);
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int gi = 0; gi < W16_RPW / 8; ++gi) {
                uint32_t t[8][5];
                uint32_t shr[8];
                int nn[8];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int64_t so = __shfl_sync(0xffffffffu, soff, 8 * gi + rr);
                    const int lm = __shfl_sync(0xffffffffu, lim, 8 * gi + rr);
                    int inside = lm - W16_KH * kh;                       // samples of this K-half inside the signal
                    inside = inside < 0 ? 0 : (inside > W16_KH ? W16_KH : inside);
                    nn[rr] = lane < W16_KC ? inside - 8 * lane : 0;
#pragma unroll
                    for (int k = 0; k < 5; ++k) t[rr][k] = 0u;
                    const uint8_t *src = a.audio + so + 2 * (W16_KH * kh + 8 * lane);
                    const uint8_t *p = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
                    shr[rr] = ((uint32_t)reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
                    if (nn[rr] > 0) {
                        if (p + 20 <= a.audio_end) {
#pragma unroll
                            for (int k = 0; k < 5; ++k) t[rr][k] = __ldg(reinterpret_cast<const unsigned int *>(p + 4 * k));
                        } else {
#pragma unroll
                            for (int k = 0; k < 5; ++k) t[rr][k] = w16_word_tail(p + 4 * k, a.audio_end);
                        }
                    }
                }
                if (!waited) { mbar_wait(&a_empty[kh], lowp ? 0u : 1u); waited = true; }   // (the stage's previous production: parity = its part)   // (the words are on their way meanwhile)
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    uint4 oh;
                    uint32_t *ph_ = &oh.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // two little-endian samples: 256 h and l as halves (0x6400 | byte is the half 1024 + byte), their sum rounded once
                        const uint32_t v = __funnelshift_r(t[rr][k], t[rr][k + 1], shr[rr]);
                        const uint32_t xh = __byte_perm(v, 0x64646464u, 0x4341u) ^ 0x00800080u, xl = __byte_perm(v, 0x64646464u, 0x4240u);
                        const __half2 th = __hmul2(__hadd2(*reinterpret_cast<const __half2 *>(&xh), __float2half2_rn(-1152.0f)), __float2half2_rn(256.0f));
                        const __half2 lf = __hadd2(*reinterpret_cast<const __half2 *>(&xl), __float2half2_rn(-1024.0f));
                        const __half2 hh = __hadd2(th, lf);
                        const __half2 r = lowp ? __hsub2(lf, __hsub2(hh, th)) : hh;   // fp16(sample), or what the rounding dropped (Fast2Sum: exact)
                        ph_[k] = *reinterpret_cast<const uint32_t *>(&r);
                    }
                    const int n = nn[rr];                                // zeros beyond the signal and in rows past the end
                    if (n < 8) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (2 * i >= n) ph_[i] = 0u;
                            else if (2 * i + 1 >= n) ph_[i] &= 0x0000FFFFu;
                        }
                    }
                    const int r = 8 * gi + rr, c = lane;
                    if (lane < W16_KC) {
                        const uint32_t rowoff = (uint32_t)(pw * W16_RPW + r) * 128u;
                        if (c < 24) st_shared(smem_u32(sA) + (uint32_t)(3 * kh + (c >> 3)) * W16_BLK + rowoff + ((((uint32_t)c & 7u) ^ ((uint32_t)r & 7u)) << 4), oh);
                        else st_shared(smem_u32(sT) + rowoff + (((2u * (uint32_t)kh + (uint32_t)(c & 7)) ^ ((uint32_t)r & 7u)) << 4), oh);
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&a_full[kh], 0);
        }
    } else {
        // ===================================================================== epilogue: |Z|^2 -> filterbank -> ln -> mel
        const int qd = warp & 3, half = warp >> 2;
        const Wave16Half &tb = a.h[half];
        const float4 *sw = s_w + half * W16_HCH * 8;
        const int nb = a.nbanks, c_begin = tb.c_begin, c_end = tb.c_end;
        const float fl_eff = a.frame_floor != -9999.9f ? a.frame_floor : -INFINITY;
        auto finish = [&](float acc) {
            float o = acc > 0.0f ? w16_lg2(acc) * 0.69314718055994530942f : 0.0f;
            return fmaxf(o + a.frame_shift, fl_eff);
        };
#pragma unroll 1
        for (int tile = 0; tile < n_my; ++tile) {
            const int64_t g = a.f_begin + ((unit0 + tile) * 256 + (int64_t)rank * 128 + qd * 32 + lane);
            const bool live = g < a.f_end;
            float *dst = a.mel + (live ? g : a.f_begin) * nb;
            float acc_lo = 0.0f, acc_hi = 0.0f;
            float *p = dst + tb.cur0 - 1;
            uint32_t v0[32];
            auto bins16 = [&](const uint32_t *v, int c) {
                const uint32_t sm = tb.shift[c >> 1] >> (16 * (c & 1)), em = live ? tb.emit[c >> 1] >> (16 * (c & 1)) : 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 l4 = sw[8 * (c - c_begin) + q], h4 = sw[8 * (c - c_begin) + 4 + q];
                    const float lw[4] = {l4.x, l4.y, l4.z, l4.w}, hw[4] = {h4.x, h4.y, h4.z, h4.w};
                    if (((sm >> (4 * q)) & 0xFu) == 0u) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = 4 * q + j;
                            const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                            const float pwr = fmaf(im, im, re * re);
                            acc_lo = fmaf(pwr, lw[j], acc_lo);
                            acc_hi = fmaf(pwr, hw[j], acc_hi);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = 4 * q + j;
                            const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                            const float pwr = fmaf(im, im, re * re);
                            const bool sh = (sm >> i) & 1u, st = (em >> i) & 1u;
                            w16_st_if(p, finish(acc_lo), st);
                            acc_lo = sh ? acc_hi : acc_lo;
                            acc_hi = sh ? 0.0f : acc_hi;
                            p += sh ? 1 : 0;
                            acc_lo = fmaf(pwr, lw[j], acc_lo);
                            acc_hi = fmaf(pwr, hw[j], acc_hi);
                        }
                    }
                }
            };
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                mbar_wait(&d_full[pass], (uint32_t)tile & 1u);
                tc_fence_after();
                const uint32_t tbase = tmem + 256u * (uint32_t)pass + ((uint32_t)(qd * 32) << 16);
                const int cb = c_begin > 8 * pass ? c_begin : 8 * pass, ce = c_end < 8 * pass + 8 ? c_end : 8 * pass + 8;
                bool released = false;
#pragma unroll 1
                for (int c = cb; c < ce; ++c) {
                    tmem_ld32(tbase + 32u * (uint32_t)(c - 8 * pass), v0);
                    tmem_ld_wait();
                    if (c + 1 >= ce) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&d_empty[pass], 0);
                        released = true;
                    }
                    bins16(v0, c);
                }
                if (!released) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&d_empty[pass], 0);
                }
            }
            if (live) {
                if (tb.flush_lo) p[0] = finish(acc_lo);
                if (tb.flush_hi) p[1] = finish(acc_hi);
                for (int b = tb.z_begin; b < tb.z_end; ++b) dst[b] = finish(0.0f);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Wave16State {
    uint8_t *w_img = nullptr;
    float4 *wtab = nullptr;
    Wave16Half h[2];
    bool ok = false;
};

static bool wave_tc16_fits(const phn_ctx *c)
{
    return c->mt.logN == 9 && c->vs <= 2 * W16_KH && c->step == 160 && c->nbanks >= 1 && c->nbanks <= 64 && !c->plp && !c->z_mean && c->preem == 0.0f;
}

int wave_tc16_prepare(phn_ctx *c)
{
    if (c->wave_tc16) return PHN_OK;
    Wave16State *st = new Wave16State();
    c->wave_tc16 = st;
    if (!wave_tc16_fits(c)) return PHN_OK;
    const MelTables &mt = c->mt;
    const int nb = mt.nbanks;
    // ---- filterbank tables (the construction of k_wave_tc.cu, over 256 bins)
    std::vector<int> sh(W16_NBIN, 0), cur_at(W16_NBIN + 1, 0);
    {
        int prev = 0;
        for (int k = 0; k < W16_NBIN; ++k) {
            cur_at[k] = prev;
            const int sgm = mt.banks[k];
            if (sgm >= 0) {
                if (sgm < prev || sgm - prev > 1) return PHN_OK;   // banks narrower than a bin: the FFT kernels serve this model
                sh[k] = sgm - prev;
                prev = sgm;
            }
        }
        cur_at[W16_NBIN] = prev;
    }
    auto bin_lo_w = [&](int k) { const int sgm = mt.banks[k]; return sgm >= 1 && sgm <= nb ? mt.coeffs[k] : 0.0f; };
    auto bin_hi_w = [&](int k) { const int sgm = mt.banks[k]; return sgm >= 0 && sgm < nb ? 1.0f - mt.coeffs[k] : 0.0f; };
    struct HalfTab { Wave16Half h; std::vector<float> wlo, whi; };
    auto build_half = [&](HalfTab &t, int bb, int be) {
        memset(&t.h, 0, sizeof(t.h));
        t.wlo.assign(W16_NBIN, 0.0f); t.whi.assign(W16_NBIN, 0.0f);
        int kmin = W16_NBIN, kmax = -1;
        for (int k = 0; k < W16_NBIN; ++k) {
            const int sgm = mt.banks[k];
            if (sgm < 0) continue;
            const bool lo_in = sgm - 1 >= bb && sgm - 1 < be, hi_in = sgm >= bb && sgm < be;
            t.wlo[k] = lo_in ? bin_lo_w(k) : 0.0f;
            t.whi[k] = hi_in ? bin_hi_w(k) : 0.0f;
            if (lo_in || hi_in) { kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
        }
        if (kmax < 0) { t.h.z_begin = bb; t.h.z_end = be; return; }
        t.h.c_begin = kmin / 16; t.h.c_end = kmax / 16 + 1;
        t.h.cur0 = cur_at[16 * t.h.c_begin];
        int cur = t.h.cur0;
        std::vector<char> stored(nb + 2, 0);
        for (int k = 16 * t.h.c_begin; k < 16 * t.h.c_end; ++k) {
            if (!sh[k]) continue;
            t.h.shift[k >> 5] |= 1u << (k & 31);
            const int b = cur - 1;
            if (b >= bb && b < be) { t.h.emit[k >> 5] |= 1u << (k & 31); stored[b] = 1; }
            ++cur;
        }
        t.h.flush_lo = cur - 1 >= bb && cur - 1 < be;
        t.h.flush_hi = cur >= bb && cur < be;
        if (t.h.flush_lo) stored[cur - 1] = 1;
        if (t.h.flush_hi) stored[cur] = 1;
        int zb = bb, ze = bb;
        for (int b = bb; b < be; ++b)
            if (!stored[b]) { if (ze == zb) zb = b; ze = b + 1; }
        t.h.z_begin = zb; t.h.z_end = ze;
    };
    auto check_half = [&](const HalfTab &t, int bb, int be) {   // replay of the device walk against the filterbank's definition
        std::vector<double> P(W16_NBIN), want(nb, 0.0), got(nb, -1.0);
        for (int k = 0; k < W16_NBIN; ++k) P[k] = 1.0 + 0.37 * k + (k % 7) * 0.11;
        for (int k = 0; k < W16_NBIN; ++k) {
            const int sgm = mt.banks[k];
            if (sgm < 0) continue;
            if (sgm >= 1 && sgm <= nb) want[sgm - 1] += (double)bin_lo_w(k) * P[k];
            if (sgm < nb) want[sgm] += (double)bin_hi_w(k) * P[k];
        }
        double lo = 0.0, hi = 0.0;
        int p = t.h.cur0 - 1;
        for (int k = 16 * t.h.c_begin; k < 16 * t.h.c_end; ++k) {
            const bool s1 = (t.h.shift[k >> 5] >> (k & 31)) & 1u, e1 = (t.h.emit[k >> 5] >> (k & 31)) & 1u;
            if (e1) { if (p < 0 || p >= nb) return false; got[p] = lo; }
            if (s1) { lo = hi; hi = 0.0; ++p; }
            lo += (double)t.wlo[k] * P[k];
            hi += (double)t.whi[k] * P[k];
        }
        if (t.h.flush_lo) { if (p < 0 || p >= nb) return false; got[p] = lo; }
        if (t.h.flush_hi) { if (p + 1 < 0 || p + 1 >= nb) return false; got[p + 1] = hi; }
        for (int b = t.h.z_begin; b < t.h.z_end; ++b) { if (got[b] >= 0.0) return false; got[b] = 0.0; }
        for (int b = bb; b < be; ++b)
            if (got[b] < 0.0 || fabs(got[b] - want[b]) > 1e-9 * (1.0 + fabs(want[b]))) return false;
        return true;
    };
    int best = -1, best_cost = 1 << 30;
    HalfTab t0, t1;
    for (int split = 0; split <= nb; ++split) {
        build_half(t0, 0, split);
        build_half(t1, split, nb);
        if (!check_half(t0, 0, split) || !check_half(t1, split, nb)) continue;
        const int cost = std::max(t0.h.c_end - t0.h.c_begin, t1.h.c_end - t1.h.c_begin);
        if (cost > W16_HCH) continue;
        if (cost < best_cost) { best_cost = cost; best = split; }
    }
    if (best < 0) return PHN_OK;
    build_half(t0, 0, best);
    build_half(t1, best, nb);
    st->h[0] = t0.h; st->h[1] = t1.h;
    std::vector<float4> wtab((size_t)2 * W16_HCH * 8, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int hf = 0; hf < 2; ++hf) {
        const HalfTab &t = hf ? t1 : t0;
        for (int cl = 0; cl < W16_HCH; ++cl) {
            const int cidx = t.h.c_begin + cl;
            if (cidx >= W16_NCH) break;
            for (int q = 0; q < 4; ++q) {
                const int k = 16 * cidx + 4 * q;
                wtab[(size_t)(hf * W16_HCH + cl) * 8 + q] = make_float4(t.wlo[k], t.wlo[k + 1], t.wlo[k + 2], t.wlo[k + 3]);
                wtab[(size_t)(hf * W16_HCH + cl) * 8 + 4 + q] = make_float4(t.whi[k], t.whi[k + 1], t.whi[k + 2], t.whi[k + 3]);
            }
        }
    }
    // ---- the windowed DFT matrix, cut into (pass, khalf) pieces; per piece and CTA rank the shared-memory image of k_wave_tc.cu
    std::vector<uint8_t> img((size_t)4 * 2 * W16_BBLK * W16_BLK, 0);
    const double w0 = 6.283185307179586476925286766559 / 512.0;
    for (int piece = 0; piece < 4; ++piece) {
        const int pass = piece >> 1, kh = piece & 1;
        for (int n = 0; n < 256; ++n) {
            const int rk = n >> 7, nl = n & 127, j = 128 * pass + (n >> 1);
            uint8_t *ri = img.data() + ((size_t)(piece * 2 + rk) * W16_BBLK) * W16_BLK;
            for (int kk = 0; kk < 208; ++kk) {
                const int k = W16_KH * kh + kk;
                double w = 0.0;
                if (kk < W16_KH && k < mt.vs) {
                    const double ang = w0 * (double)((j * k) & 511);
                    w = (double)mt.hamming[k] * ((n & 1) ? -sin(ang) : cos(ang));
                }
                const __half whi = __float2half_rn((float)w);
                const __half wlo = __float2half_rn((float)(w - (double)__half2float(whi)));
                for (int part = 0; part < 2; ++part) {
                    const size_t off = kk < 192 ? (size_t)(3 * part + kk / 64) * W16_BLK + sw128_off(nl, kk % 64)
                                                : (size_t)6 * W16_BLK + sw128_off(nl, 16 * part + (kk - 192));
                    *reinterpret_cast<__half *>(ri + off) = part ? wlo : whi;
                }
            }
        }
    }
    PHN_CUDA(c, cudaMalloc((void **)&st->w_img, img.size()));
    PHN_CUDA(c, cudaMemcpy(st->w_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    PHN_CUDA(c, cudaMalloc((void **)&st->wtab, wtab.size() * sizeof(float4)));
    PHN_CUDA(c, cudaMemcpy(st->wtab, wtab.data(), wtab.size() * sizeof(float4), cudaMemcpyHostToDevice));
    st->ok = true;
    return PHN_OK;
}

void wave_tc16_release(phn_ctx *c)
{
    if (!c->wave_tc16) return;
    Wave16State *st = static_cast<Wave16State *>(c->wave_tc16);
    if (st->w_img) cudaFree(st->w_img);
    if (st->wtab) cudaFree(st->wtab);
    delete st;
    c->wave_tc16 = nullptr;
}

bool wave_tc16_applies(phn_ctx *c)
{
    static const bool off = getenv("PHNREC_WAVE_TC") && atoi(getenv("PHNREC_WAVE_TC")) == 0;
    if (off || !c->wave_tc16) return false;
    const Wave16State *st = static_cast<const Wave16State *>(c->wave_tc16);
    return st->ok && c->fmt == PHN_WAVE_LIN16 && c->dc_shift == 0.0f && c->scale == 1.0f && wave_tc16_fits(c);
}

int launch_wave_tc16(phn_ctx *c, const void *d_audio, int64_t f_begin, int64_t f_end)
{
    const Wave16State *st = static_cast<const Wave16State *>(c->wave_tc16);
    Wave16Args a;
    a.audio = (const uint8_t *)d_audio; a.audio_end = a.audio + c->total_bytes;
    a.byte_off = (const int64_t *)c->d_byte_off.p; a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt; a.f_begin = f_begin; a.f_end = f_end;
    a.vs = c->vs; a.step = c->step; a.nbanks = c->nbanks;
    a.frame_shift = c->frame_shift; a.frame_floor = c->frame_floor;
    a.w_img = st->w_img; a.wtab = st->wtab; a.mel = (float *)c->d_mel.p;
    a.h[0] = st->h[0]; a.h[1] = st->h[1];
    PHN_CUDA(c, cudaFuncSetAttribute(k_wave_tc16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W16_SMEM));
    const int64_t units = (f_end - f_begin + 255) / 256;
    const int64_t maxp = c->num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (units < maxp ? units : maxp))); cfg.blockDim = dim3(W16_THREADS);
    cfg.dynamicSmemBytes = W16_SMEM; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PHN_CUDA(c, cudaLaunchKernelEx(&cfg, k_wave_tc16, a));
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

}  // namespace phn
