// k_wave_tc.cu — K-wave for the tensor-core pipeline: the windowed DFT of every frame as ONE GEMM on tcgen05.
//
// Replaces, like k_wave.cu, ConvertWaveformFormat (srec.cpp:709-791), MelBanks::ProcessFrame (melbanks.cpp:111-204),
// cFour1 / _mbApply (dspc.cpp:24-78, 236-269), cPower / sLn (dspc.h:141-160) and FrameBasedNormalization
// (srec.cpp:1594-1620) - for the configuration the benchmark is quoted on (8 kHz A-law, 25 ms window = 200 samples,
// 256-point transform, plain front end) and the same models fed 16-bit linear samples (LIN16, see the kernel).  Everything
// else keeps the register-FFT kernels of k_wave.cu.
//
// Why a GEMM: the FFT kernel spends ~17 000 thread-instructions per frame on butterflies and their two shared-memory
// transposes and is bound by the SM's issue slots (0.76 ms for 998 000 frames).  The same spectrum is
//     Z[f][n] = sum_k x[f][k] . W[k][n],   W[k][2j] = ham[k] cos(2 pi j k / 256),  W[k][2j+1] = -ham[k] sin(2 pi j k / 256)
// i.e. [frames x 208] . [208 x 256]: 106 kFLOP per frame, 0.1 TFLOP per batch - a tenth of a millisecond of tensor time.
// Precision: an A-law sample is an integer of at most 6 significant bits times a power of two - EXACT in fp16; the
// matrix goes in as two fp16 terms (W = hi + lo, residual below 2^-22 of the window's scale), the products are exact
// and the sums are fp32 in TMEM: the spectrum is as good as an fp32 DFT (better conditioned than the fp32 FFT it
// replaces, whose twiddle products round once per stage).
//
// One cluster of two CTAs (cta_group::2, M = 256 frames, N = 256 = 128 bins x {re, im}) walks frame tiles:
//   producers (8 warps per CTA)  audio bytes -> fp16 -> the A tile in shared memory, canonical K-major SWIZZLE_128B
//                                (lane = one 16-byte chunk = 8 samples of a frame: one coalesced 208-byte read per row,
//                                two samples per instruction in half2 arithmetic); two stages
//   MMA warp (CTA 0)             26 MMAs of K = 16 per tile: (3 x 4 + 1) k-steps against W_hi, then against W_lo
//   epilogue (4 warps per CTA)   thread = frame: TMEM -> |Z|^2 -> the triangular filterbank as a running pair of
//                                accumulators (Banks[] is non-decreasing in the bin, dspc.cpp:236-269) -> guarded ln ->
//                                frame normalisation -> mel; the accumulator is double-buffered, so the epilogue of tile
//                                i runs under the MMAs of tile i + 1
// The matrix (2 x 7 blocks of 16 KB per CTA: each CTA holds the 128 output columns it contributes) is loaded once.
// The 8-sample tail of the window (k = 192..207) would waste 3/4 of a 64-column block in every operand: the two A
// stages share one tail block (stage s in k-step s of it), and W_hi / W_lo share one (k-steps 0 / 1).
#include "internal.h"
#include "device_math.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace phn {

constexpr int WT_K = 208;                 // window columns fed to the tensor cores (13 k-steps of 16; the window is <= 208)
constexpr int WT_NBIN = 128;              // bins 0..127 of the 256-point transform
// producer warps per CTA: 8 (16 rows of the tile each) for A-law - 0.245 -> 0.217 ms on 998 000 frames: a producer warp is one
// dependent instruction stream - 4 (32 rows each) for lin16, whose two parts per tile then fit 128 registers (8 warps: 0.49 against 0.42 ms)
template <bool LIN16> constexpr int WT_PROD = LIN16 ? 4 : 8;
constexpr int WT_EPI = 8;                 // epilogue warps per CTA (warp % 4 = TMEM lane quarter, warp / 4 = half of the bins)
template <bool LIN16> constexpr int WT_THREADS = (WT_EPI + 1 + WT_PROD<LIN16>) * 32;
constexpr int WT_BLK = 16384;             // [128 rows x 64 fp16], SWIZZLE_128B
constexpr int WT_BBLK = 7;                // matrix blocks per CTA: 3 hi, 3 lo, 1 shared tail
constexpr int WT_HCH = 7;                  // chunks of 16 bins one epilogue half may walk (its weights sit in shared memory)
constexpr size_t WT_SMEM = (size_t)(WT_BBLK + 2 * 3 + 1) * WT_BLK + 128 + 2 * WT_HCH * 128 + 1024;   // + barriers + filterbank weights + alignment slack

// The filterbank as one half of the epilogue sees it (kernel parameters: constant bank).  Banks[] (dspc.cpp:236-269) does not
// decrease with the bin; a half owns a contiguous range of banks and walks the 16-bin chunks that hold their bins.
struct WaveTcHalf {
    float4 wlo[WT_NBIN / 4];              // bin k -> c[k] for bank Banks[k] - 1   (0 when that bank is not this half's)
    float4 whi[WT_NBIN / 4];              // bin k -> 1 - c[k] for bank Banks[k]
    uint32_t shift[WT_NBIN / 32];         // bit k: Banks[] grows (by one) at bin k, relative to the last bin inside the filterbank
    uint32_t emit[WT_NBIN / 32];          // bit k: the bank completed by that step is this half's
    int c_begin, c_end;                   // chunks of 16 bins
    int cur0;                             // Banks[] in effect at the first bin of chunk c_begin
    int flush_lo, flush_hi;               // banks cur - 1 / cur pending after the last chunk are this half's
    int b_begin, b_end;                   // the half's banks that the walk stores (host-side check only)
    int z_begin, z_end;                   // banks no bin belongs to (value: ln of silence)
};
struct WaveTcTab { WaveTcHalf h[2]; };

struct WaveTcArgs {
    const uint8_t *audio, *audio_end;     // the batch's audio and one past its last byte
    const int64_t *byte_off, *frame_off;
    int n_utt;
    int64_t f_begin, f_end;               // frames of this launch
    int vs, step, nbanks;
    long long *tl;                        // kernel development (PHNREC_WTC_DBG & 8): clock64() timeline of CTA 0, [role][tile][event]
    int dbg;                              // kernel development (PHNREC_WTC_DBG): 1 producers store zeros, 2 epilogue skips the filterbank, 4 no MMAs
    float frame_shift, frame_floor;
    const uint8_t *w_img;                 // [2 ranks][WT_BBLK][16 KB]
    float *mel;
    WaveTcTab tab;
};

// two A-law bytes (bits 0..7 and 16..23 of x) -> half2 bits of 8 * ALawTableD5 (alaw.cpp:14-48, the value alaw8_float of
// k_wave.cu produces).  With t = byte ^ 0xD5, segment s = t[6:4], mantissa m = t[3:0]: for s >= 1 the value
// 8 (2m + 33) 2^(s-1) is the half with exponent field 22 + s and mantissa (2m + 1) << 5, i.e. 0x5820 + (t[6:0] << 6);
// that pattern read for s = 0 is f = 8m + 132 where 16m + 8 = 2f - 256 is wanted, and 2f - 256 < f exactly when s = 0.
__device__ __forceinline__ uint32_t alaw2_half2(uint32_t x)
{
    const uint32_t t = x ^ 0x00D500D5u;
    const uint32_t fb = ((t & 0x007F007Fu) << 6) + 0x58205820u;
    const __half2 f = *reinterpret_cast<const __half2 *>(&fb);
    const __half2 g = __hfma2(f, __float2half2_rn(2.0f), __float2half2_rn(-256.0f));
    const __half2 r = __hmin2(f, g);
    return *reinterpret_cast<const uint32_t *>(&r) | ((t & 0x00800080u) << 8);
}

// one word at p when fewer than 4 bytes are left in the batch's audio buffer (its very last chunks)
__device__ __noinline__ uint32_t load_word_tail(const uint8_t *p, const uint8_t *end)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        if (p + i < end) r |= (uint32_t)p[i] << (8 * i);
    return r;
}

// predicated store: the value is computed whether or not it is stored (no branch around it)
__device__ __forceinline__ void st_global_if(float *p, float v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((uint32_t)pred) : "memory");
}

__device__ __forceinline__ float lg2_approx_wt(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ int find_utt(const int64_t *off, int n, int64_t f)
{
    int lo = 0, hi = n;  // off[lo] <= f < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

// DBG: the instantiation with the development switches (a.dbg) and the clock64() timeline; the product kernel carries neither.
#define WT_TL(role, ev) do { if (DBG && a.tl && blockIdx.x == 0 && lane == 0 && it < 32) a.tl[((role) * 32 + it) * 8 + (ev)] = clock64(); } while (0)

// LIN16: 16-bit linear samples.  A sample is hi + lo with hi = fp16(sample) and lo the rounding error (an integer of at most 4 bits),
// both exact halves, and the tile goes through the tensor cores as two parts that share the accumulator: part 0 = hi against W_hi and
// W_lo, part 1 = lo against W_hi (lo W_lo is below 2^-22 of the sample's own magnitude).  A part takes the place of a tile in the two-stage A ring (iteration = 2 tile + part, stage =
// part), so shared memory, barriers and phases are those of the A-law kernel; only the accumulator hand-over is per tile.
template <bool DBG, bool LIN16>
__global__ void __launch_bounds__(WT_THREADS<LIN16>, 1) k_wave_tc(const __grid_constant__ WaveTcArgs a)
{
    const int dbg = DBG ? a.dbg : 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *sB = smem;                                   // WT_BBLK blocks
    uint8_t *sA = sB + (size_t)WT_BBLK * WT_BLK;          // 2 stages x 3 blocks
    uint8_t *sT = sA + (size_t)6 * WT_BLK;                // the stages' shared tail block
    uint64_t *bars = reinterpret_cast<uint64_t *>(sT + WT_BLK);
    uint64_t *a_full = bars;          // [2] CTA 0: both CTAs' producers have written stage s
    uint64_t *a_empty = bars + 2;     // [2] the tile's MMAs have read stage s (multicast commit)
    uint64_t *d_full = bars + 4;      // [2] accumulator s is complete (multicast commit)
    uint64_t *d_empty = bars + 6;     // [2] CTA 0: both CTAs' epilogues have read accumulator s
    uint64_t *b_full = bars + 8;      // this CTA's half of the matrix has landed
    uint64_t *pb_full = bars + 9;     // CTA 0: the peer's has
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);
    float4 *s_w = reinterpret_cast<float4 *>(bars + 16);   // [2 halves][WT_HCH chunks][8]: wlo (4 x float4) then whi of the chunks a half walks

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;
    const int64_t nf = a.f_end - a.f_begin;
    // a cluster walks a contiguous range of 256-frame units: consecutive tiles mostly stay inside one utterance
    const int64_t n_units = (nf + 255) / 256, ncl = gridDim.x >> 1, cl = blockIdx.x >> 1;
    const int64_t unit0 = cl * n_units / ncl;
    const int n_my = (int)((cl + 1) * n_units / ncl - unit0);
    const int n_it = LIN16 ? 2 * n_my : n_my;   // iterations of the A ring
    constexpr int WARP_MMA = WT_EPI, PROD0 = WT_EPI + 1;

    if (threadIdx.x == 0) {
        if (DBG && a.tl && blockIdx.x == 0) a.tl[4 * 32 * 8] = clock64();
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 2 * WT_PROD<LIN16>); mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 2 * WT_EPI);
        }
        mbar_init(b_full, 1); mbar_init(pb_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * WT_HCH * 8; i += blockDim.x) {
        const int hf = i / (WT_HCH * 8), cl = (i / 8) % WT_HCH, q = i & 7, c = a.tab.h[hf].c_begin + cl;
        s_w[i] = c < WT_NBIN / 16 ? (q < 4 ? a.tab.h[hf].wlo[4 * c + q] : a.tab.h[hf].whi[4 * c + q - 4]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == WARP_MMA) {
        // ===================================================================== matrix load + MMA issue
        if (elect_one()) {
            mbar_expect_tx(b_full, WT_BBLK * WT_BLK);
            for (int b = 0; b < WT_BBLK; ++b)
                tma_load_1d(sB + (size_t)b * WT_BLK, a.w_img + ((size_t)rank * WT_BBLK + b) * WT_BLK, WT_BLK, b_full);
        }
        __syncwarp();
        mbar_wait(b_full, 0);
        if (rank != 0) {
            if (lane == 0) mbar_arrive_cluster(pb_full, 0);
        } else {
            mbar_wait_cluster(pb_full, 0);
            constexpr uint32_t idesc = make_idesc(256, false, 256);
            const uint64_t dA = make_sw128_desc(smem_u32(sA)), dB = make_sw128_desc(smem_u32(sB)), dT = make_sw128_desc(smem_u32(sT));
            const uint32_t alo0 = (uint32_t)dA, blo0 = (uint32_t)dB, tlo0 = (uint32_t)dT, hi = (uint32_t)(dA >> 32);
            const uint32_t bar_ae = smem_u32(a_empty), bar_df = smem_u32(d_full);
            const bool leader = elect_one();
#pragma unroll 1
            for (int it = 0; it < n_it; ++it) {
                const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
                const int tile = LIN16 ? it >> 1 : it;
                const bool low = LIN16 && (it & 1);                       // the low bytes' part: accumulates on top of the high bytes'
                const uint32_t ds = LIN16 ? (uint32_t)tile & 1u : s, dph = LIN16 ? ((uint32_t)tile >> 1) & 1u : ph;
                WT_TL(0, 0);
                mbar_wait_cluster(&a_full[s], ph);
                WT_TL(0, 1);
                if (!low && tile >= 2) mbar_wait_cluster(&d_empty[ds], dph ^ 1u);
                WT_TL(0, 2);
                tc_fence_after();
                if (leader) {
                    const uint32_t td = tmem + 256u * ds;
                    if (!(dbg & 4)) {
                    const uint32_t alo = alo0 + s * (3u * (WT_BLK >> 4));
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        if (low && part == 1) break;
                        const uint32_t blo = blo0 + (uint32_t)part * (3u * (WT_BLK >> 4));
#pragma unroll
                        for (int b = 0; b < 3; ++b)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t o = (uint32_t)b * (WT_BLK >> 4) + 2u * (uint32_t)ks;
                                if (part == 0 && b == 0 && ks == 0 && !low) umma2_ss_lo<0>(td, alo + o, hi, blo + o, idesc);
                                else umma2_ss_lo<1>(td, alo + o, hi, blo + o, idesc);
                            }
                        // the window's tail: stage s of the A tail block against part `part` of the matrix tail block
                        umma2_ss_lo<1>(td, tlo0 + 2u * s, hi, blo0 + 6u * (WT_BLK >> 4) + 2u * (uint32_t)part, idesc);
                    }
                    }
                    tc_commit2_u(bar_ae + 8u * s);
                    if (!LIN16 || low) tc_commit2_u(bar_df + 8u * ds);
                }
                __syncwarp();
                WT_TL(0, 3);
            }
        }
    } else if (warp >= PROD0) {
        // ===================================================================== producers: audio -> A tile
        // A warp owns 32 consecutive rows of the tile (lane = row for the bookkeeping), as four groups of 8.  Row r of a
        // group starts 10 chunks (80 samples) after row r - 1, so chunk c of row r is chunk c - 10 of row r + 1: a group of
        // 8 consecutive frames of one utterance has 7 * 10 + 26 = 96 DIFFERENT chunks = three per lane, each decoded once
        // and stored to the (up to three) rows that contain it.  The words of a group of tile i + 1 are requested as soon as the
        // same group of tile i has been decoded (its registers are free): a whole tile time passes before they are needed.  Groups that straddle two
        // utterances, touch the end of the batch's audio or hold an incomplete window go row by row (one in a hundred).
        constexpr int RPW = 128 / WT_PROD<LIN16>, NG = RPW / 8;
        static_assert(RPW == 16 || RPW == 32, "rows of a warp on its lanes");
        const int pw = warp - PROD0;
        const int64_t audio_len = a.audio_end - a.audio;
        const uint32_t abase = (uint32_t)reinterpret_cast<uintptr_t>(a.audio);
        // this lane's utterance (kept across tiles: the lane's next row is 256 frames further)
        int u = 0;
        int64_t fo_cur = 0, fo_next = -1, b0 = 0, len = 0;
        // where this lane's three chunks of a group go: chunk q = lane + 32 i = 10 r + c belongs to row r0 = q / 10 as
        // column c0 = q % 10, to row r0 - 1 as c0 + 10 and to row r0 - 2 as c0 + 20 (SWIZZLE_128B: 16-byte column XOR
        // row % 8; the window's tail, columns 24 and 25, lives in the stages' shared block at column 2 s + c - 24).
        // Shared-memory addresses for stage 0, group 0; 0 = no such row.
        uint32_t st_addr[3][3];
        bool st_tail[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int q = lane + 32 * i, r0 = (q * 205) >> 11, c0 = q - 10 * r0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int r = r0 - j, c = c0 + 10 * j;
                const bool ok = r >= 0 && r < 8 && c < WT_K / 8;
                st_tail[i][j] = c >= 24;
                const uint32_t blk = c < 24 ? smem_u32(sA) + (uint32_t)(c >> 3) * WT_BLK : smem_u32(sT);
                st_addr[i][j] = ok ? blk + (uint32_t)(pw * RPW + r) * 128u + ((((uint32_t)c & 7u) ^ (uint32_t)r) << 4) : 0u;
            }
        }
        auto st_shared = [](uint32_t addr, const uint4 &v) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        };
        constexpr int BPS = LIN16 ? 2 : 1;                   // bytes per sample
        constexpr int CB = 8 * BPS;                          // bytes per chunk (8 samples)
        constexpr int NW = CB / 4 + 1;                       // aligned words that cover a chunk at any byte alignment
        // chunk of 8 samples -> 8 halves.  A-law: the G.711 expansion.  lin16 (two little-endian samples per word): the sample's
        // fp16 rounding, or (`lowp`) what the rounding dropped.
        auto decode_chunk = [&](const uint32_t (&wv)[NW], uint32_t sh, bool lowp) {
            uint32_t v[NW - 1];
#pragma unroll
            for (int k = 0; k < NW - 1; ++k) v[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
            uint4 out;
            if constexpr (!LIN16) {
                out.x = alaw2_half2(__byte_perm(v[0], 0u, 0x4140u));
                out.y = alaw2_half2(__byte_perm(v[0], 0u, 0x4342u));
                out.z = alaw2_half2(__byte_perm(v[1], 0u, 0x4140u));
                out.w = alaw2_half2(__byte_perm(v[1], 0u, 0x4342u));
            } else {
                uint32_t *o = &out.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // t = 256 h and l as halves (0x6400 | byte is the half 1024 + byte), then the sample as hi + lo with
                    // hi = fp16(256 h + l) and lo = the rounding error, both exact (Fast2Sum: |256 h| >= l or h = 0)
                    const uint32_t xh = __byte_perm(v[k], 0x64646464u, 0x4341u) ^ 0x00800080u, xl = __byte_perm(v[k], 0x64646464u, 0x4240u);
                    const __half2 t = __hmul2(__hadd2(*reinterpret_cast<const __half2 *>(&xh), __float2half2_rn(-1152.0f)), __float2half2_rn(256.0f));
                    const __half2 lf = __hadd2(*reinterpret_cast<const __half2 *>(&xl), __float2half2_rn(-1024.0f));
                    const __half2 hi = __hadd2(t, lf);
                    const __half2 r = lowp ? __hsub2(lf, __hsub2(hi, t)) : hi;
                    o[k] = *reinterpret_cast<const uint32_t *>(&r);
                }
            }
            return out;
        };
        auto load_chunk = [&](const uint8_t *p, uint32_t (&wv)[NW]) {   // p: 4-byte aligned
            if (p + 4 * NW <= a.audio_end) {
#pragma unroll
                for (int k = 0; k < NW; ++k) wv[k] = __ldg(reinterpret_cast<const unsigned int *>(p + 4 * k));
            } else {                                         // (the last bytes of the batch)
#pragma unroll
                for (int k = 0; k < NW; ++k) wv[k] = load_word_tail(p + 4 * k, a.audio_end);
            }
        };
        int64_t soff = 0;
        int lim = -1;
        uint32_t seq = 0;
        // row bookkeeping of tile `t` -> (bsoff, blim, bseq): byte offset of the row's first sample, samples of its window inside the signal
        int64_t bsoff = 0;
        int blim = -1;
        uint32_t bseq = 0;
        auto book = [&](int t) {
            const int64_t g = a.f_begin + ((unit0 + t) * 256 + (int64_t)rank * 128) + pw * RPW + (lane & (RPW - 1));   // (RPW = 16: the upper lanes repeat the rows)
            bsoff = 0; blim = -1;                            // -1: a row past the end of the launch (zeros)
            if (g < a.f_end && !(dbg & 1)) {
                if (g >= fo_next) {
                    // the next utterance, or a search when the row jumped further
                    if (fo_next >= 0 && u + 2 <= a.n_utt && g < a.frame_off[u + 2]) ++u; else u = find_utt(a.frame_off, a.n_utt, g);
                    fo_cur = a.frame_off[u]; fo_next = a.frame_off[u + 1];
                    b0 = a.byte_off[u]; len = (a.byte_off[u + 1] - b0) / BPS;
                }
                const int64_t s0 = (g - fo_cur) * a.step, left = len - s0;
                bsoff = b0 + s0 * BPS;
                blim = left < a.vs ? (left < 0 ? 0 : (int)left) : a.vs;   // samples of the window inside the signal
            }
            const int64_t so0 = __shfl_sync(0xffffffffu, bsoff, lane & ~7);         // the group's first row
            const bool in_seq = blim == a.vs && bsoff == so0 + (int64_t)(lane & 7) * (80 * BPS) && so0 + CB * 96 + 4 <= audio_len;
            bseq = __ballot_sync(0xffffffffu, in_seq);
        };
        // a group that is not 8 whole windows of one utterance: row by row, lane = chunk; the eight rows' words are requested
        // before the first is decoded
        auto slow_group = [&](int gi, uint32_t s, bool lowp) {
            const uint32_t s_blk = s * (3u * WT_BLK);
            uint32_t t[8][NW];
            uint32_t shr[8];
            int nn[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int64_t so = __shfl_sync(0xffffffffu, soff, 8 * gi + rr);
                const int lm = __shfl_sync(0xffffffffu, lim, 8 * gi + rr);
                nn[rr] = lane < WT_K / 8 ? lm - 8 * lane : 0;   // samples of this chunk inside the signal
#pragma unroll
                for (int k = 0; k < NW; ++k) t[rr][k] = 0u;
                const uint8_t *src = a.audio + so + CB * lane;
                shr[rr] = ((uint32_t)reinterpret_cast<uintptr_t>(src) & 3u) * 8u;
                if (nn[rr] > 0) load_chunk(reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3), t[rr]);
            }
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                uint4 out = decode_chunk(t[rr], shr[rr], lowp);
                const int n = nn[rr];                        // zeros beyond the signal (melbanks.cpp:151-170) and in rows past the end
                uint32_t *o = &out.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (2 * i >= n) o[i] = 0u;
                    else if (2 * i + 1 >= n) o[i] &= 0x0000FFFFu;
                }
                const int r = 8 * gi + rr, c = lane;
                const uint32_t blk = c < 24 ? smem_u32(sA) + s_blk + (uint32_t)(c >> 3) * WT_BLK : smem_u32(sT);
                const uint32_t col = c < 24 ? (uint32_t)(c & 7) : 2u * s + (uint32_t)(c & 7);
                if (lane < WT_K / 8) st_shared(blk + (uint32_t)(pw * RPW + r) * 128u + ((col ^ ((uint32_t)r & 7u)) << 4), out);
            }
        };
        // the three chunks of this lane in a whole group: decoded once, stored to every row that holds them
        auto fast_group = [&](int gi, uint32_t s, const uint32_t (&wg)[3][NW], uint32_t sh, bool lowp) {
            const uint32_t s_blk = s * (3u * WT_BLK), s_tail = s * 32u;   // stage 1: three blocks further / tail column XOR 2
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint4 val = decode_chunk(wg[i], sh, lowp);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    if (st_addr[i][j])
                        st_shared((st_tail[i][j] ? st_addr[i][j] ^ s_tail : st_addr[i][j] + s_blk) + (uint32_t)gi * 1024u, val);
            }
        };
        // the words of group gi of the tile whose bookkeeping is in (so_, seq_)
        auto request = [&](int gi, int64_t so_, uint32_t seq_, uint32_t (&wg)[3][NW]) {
            const int64_t sg = __shfl_sync(0xffffffffu, so_, 8 * gi);
            if (((seq_ >> (8 * gi)) & 0xFFu) == 0xFFu) {
                const uint8_t *p = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(a.audio + sg + CB * lane) & ~(uintptr_t)3);
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int k = 0; k < NW; ++k) wg[i][k] = __ldg(reinterpret_cast<const unsigned int *>(p + 32 * CB * i + 4 * k));
            }
        };
        if constexpr (!LIN16) {
            // A-law: the words of a group of tile i + 1 are requested as soon as the same group of tile i has been decoded
            uint32_t w[NG][3][NW];
            if (n_my > 0) {
                book(0);
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) request(gi, bsoff, bseq, w[gi]);
            }
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
                if (pw == 0) WT_TL(1, 0);
                soff = bsoff; lim = blim; seq = bseq;        // this tile's rows; then the next tile's, whose words are requested
                const bool more = it + 1 < n_my;             // group by group as this tile's registers become free
                if (more) book(it + 1);
                if (it >= 2) mbar_wait(&a_empty[s], ph ^ 1u);
                if (pw == 0) WT_TL(1, 1);
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) {
                    const int64_t sg = __shfl_sync(0xffffffffu, soff, 8 * gi);
                    if (((seq >> (8 * gi)) & 0xFFu) == 0xFFu) fast_group(gi, s, w[gi], ((abase + (uint32_t)sg) & 3u) * 8u, false);
                    else slow_group(gi, s, false);
                    if (more) request(gi, bsoff, bseq, w[gi]);
                }
                if (pw == 0) WT_TL(1, 2);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&a_full[s], 0);
                if (pw == 0) WT_TL(1, 3);
            }
        } else {
            // lin16: two iterations per tile (high bytes into stage 0, low bytes into stage 1); a group's words are requested while
            // the group before it is decoded
            uint32_t w[2][3][NW];
#pragma unroll 1
            for (int it = 0; it < n_it; ++it) {
                const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
                const bool lowp = it & 1;
                if (pw == 0) WT_TL(1, 0);
                if (!lowp) { book(it >> 1); soff = bsoff; lim = blim; seq = bseq; }
                request(0, soff, seq, w[0]);
                if (it >= 2) mbar_wait(&a_empty[s], ph ^ 1u);
                if (pw == 0) WT_TL(1, 1);
#pragma unroll
                for (int gi = 0; gi < NG; ++gi) {
                    if (gi + 1 < NG) request(gi + 1, soff, seq, w[(gi + 1) & 1]);
                    const int64_t sg = __shfl_sync(0xffffffffu, soff, 8 * gi);
                    if (((seq >> (8 * gi)) & 0xFFu) == 0xFFu) fast_group(gi, s, w[gi & 1], ((abase + (uint32_t)sg) & 3u) * 8u, lowp);
                    else slow_group(gi, s, lowp);
                }
                if (pw == 0) WT_TL(1, 2);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&a_full[s], 0);
                if (pw == 0) WT_TL(1, 3);
            }
        }
    } else {
        // ===================================================================== epilogue: |Z|^2 -> filterbank -> ln -> mel
        // thread = frame (TMEM lane), two warps per lane quarter: each `half` owns a range of banks and walks the chunks of
        // 16 bins that hold them.  Bank `cur` collects (1 - c) P from the bins with Banks = cur (acc_hi) and bank cur - 1
        // collects c P from the same bins (acc_lo); where Banks moves on, bank cur - 1 is complete.  No branch in the walk (a
        // lone warp pays tens of clocks for each): the logarithm is taken at every bin and stored under a predicate.
        const int qd = warp & 3, half = warp >> 2;
        const WaveTcHalf &tb = a.tab.h[half];
        const float4 *sw = s_w + half * WT_HCH * 8;
        const int nb = a.nbanks, c_begin = tb.c_begin, c_end = tb.c_end;
        const float fl_eff = a.frame_floor != -9999.9f ? a.frame_floor : -INFINITY;   // srec.cpp:1594-1620 (floor switched off: -9999.9)
        // one finished bank: sLn (dspc.h:155-160: ln, digital silence -> 0) as lg2.approx * ln 2, then the frame normalisation
        auto finish = [&](float acc) {
            float o = acc > 0.0f ? lg2_approx_wt(acc) * 0.69314718055994530942f : 0.0f;
            return fmaxf(o + a.frame_shift, fl_eff);
        };
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
            const int64_t g = a.f_begin + ((unit0 + it) * 256 + (int64_t)rank * 128 + qd * 32 + lane);
            const bool live = g < a.f_end;
            float *dst = a.mel + (live ? g : a.f_begin) * nb;
            float acc_lo = 0.0f, acc_hi = 0.0f;
            float *p = dst + tb.cur0 - 1;                    // where bank cur - 1 goes
            uint32_t v0[32];
            auto bins16 = [&](const uint32_t *v, int c) {    // bins 16 c + (0..15)
                const uint32_t sm = tb.shift[c >> 1] >> (16 * (c & 1)), em = live ? tb.emit[c >> 1] >> (16 * (c & 1)) : 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 l4 = sw[8 * (c - c_begin) + q], h4 = sw[8 * (c - c_begin) + 4 + q];
                    const float lw[4] = {l4.x, l4.y, l4.z, l4.w}, hw[4] = {h4.x, h4.y, h4.z, h4.w};
                    if (((sm >> (4 * q)) & 0xFu) == 0u) {    // (uniform) Banks[] constant over these four bins: accumulate only
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = 4 * q + j;
                            const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                            const float pwr = fmaf(im, im, re * re);  // cPower (dspc.h:141-146)
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
                            st_global_if(p, finish(acc_lo), st);
                            acc_lo = sh ? acc_hi : acc_lo;
                            acc_hi = sh ? 0.0f : acc_hi;
                            p += sh ? 1 : 0;
                            acc_lo = fmaf(pwr, lw[j], acc_lo);
                            acc_hi = fmaf(pwr, hw[j], acc_hi);
                        }
                    }
                }
            };
            if (qd == 0) WT_TL(2 + half, 0);
            mbar_wait(&d_full[s], ph);
            if (qd == 0) WT_TL(2 + half, 1);
            tc_fence_after();
            const uint32_t tbase = tmem + 256u * s + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {          // (one chunk in registers at a time: 17 warps share the register file)
                tmem_ld32(tbase + 32u * (uint32_t)c, v0);
                tmem_ld_wait();
                if (qd == 0) WT_TL(2 + half, 2 + (c - c_begin));
                if (c + 1 >= c_end) {                        // the half's part of the accumulator is in registers: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&d_empty[s], 0);
                }
                if (!(dbg & 2)) bins16(v0, c);
            }
            if (c_begin >= c_end) {                          // (a half without banks)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&d_empty[s], 0);
            }
            if (live) {
                if (tb.flush_lo) p[0] = finish(acc_lo);
                if (tb.flush_hi) p[1] = finish(acc_hi);
                for (int b = tb.z_begin; b < tb.z_end; ++b) dst[b] = finish(0.0f);
            }
            if (qd == 0) WT_TL(2 + half, 7);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer is done with this CTA's barriers, shared memory and tensor memory
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct WaveTcState {
    uint8_t *w_img = nullptr;
    WaveTcTab tab;
    bool ok = false;      // the model's front end fits this kernel
};

static bool wave_tc_fits(const phn_ctx *c)
{
    return c->mt.logN == 8 && c->vs <= WT_K && c->step == 80 && c->nbanks >= 1 && c->nbanks <= 64 && !c->plp && !c->z_mean && c->preem == 0.0f;
}

// The windowed DFT matrix as the two CTAs' shared-memory images, and the filterbank tables.
int wave_tc_prepare(phn_ctx *c)
{
    if (c->wave_tc) return PHN_OK;
    WaveTcState *st = new WaveTcState();
    c->wave_tc = st;
    if (!wave_tc_fits(c)) return PHN_OK;
    const MelTables &mt = c->mt;
    std::vector<uint8_t> img((size_t)2 * WT_BBLK * WT_BLK, 0);
    const double w0 = 6.283185307179586476925286766559 / 256.0;
    for (int n = 0; n < 256; ++n) {
        const int rank = n >> 7, nl = n & 127, j = n >> 1;
        uint8_t *ri = img.data() + (size_t)rank * WT_BBLK * WT_BLK;
        for (int k = 0; k < WT_K; ++k) {
            double w = 0.0;
            if (k < mt.vs) {
                const double ang = w0 * (double)((j * k) & 255);
                w = (double)mt.hamming[k] * ((n & 1) ? -sin(ang) : cos(ang));
            }
            const __half hi = __float2half_rn((float)w);
            const __half lo = __float2half_rn((float)(w - (double)__half2float(hi)));
            for (int part = 0; part < 2; ++part) {
                const size_t off = k < 192 ? (size_t)(3 * part + k / 64) * WT_BLK + sw128_off(nl, k % 64)
                                           : (size_t)6 * WT_BLK + sw128_off(nl, 16 * part + (k - 192));
                *reinterpret_cast<__half *>(ri + off) = part ? lo : hi;
            }
        }
    }
    PHN_CUDA(c, cudaMalloc((void **)&st->w_img, img.size()));
    PHN_CUDA(c, cudaMemcpy(st->w_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    // Filterbank tables.  Banks[] must grow by single steps (no bank narrower than a bin), else the FFT kernels serve the model.
    memset(&st->tab, 0, sizeof(st->tab));
    const int nb = mt.nbanks;
    std::vector<int> sh(WT_NBIN, 0), cur_at(WT_NBIN + 1, 0);   // shift at bin k; Banks[] in effect before bin k
    {
        int prev = 0;
        for (int k = 0; k < WT_NBIN; ++k) {
            cur_at[k] = prev;
            const int sgm = mt.banks[k];
            if (sgm >= 0) {
                if (sgm < prev || sgm - prev > 1) return PHN_OK;
                sh[k] = sgm - prev;
                prev = sgm;
            }
        }
        cur_at[WT_NBIN] = prev;
    }
    auto bin_lo_w = [&](int k) { const int sgm = mt.banks[k]; return sgm >= 1 && sgm <= nb ? mt.coeffs[k] : 0.0f; };         // -> bank sgm - 1
    auto bin_hi_w = [&](int k) { const int sgm = mt.banks[k]; return sgm >= 0 && sgm < nb ? 1.0f - mt.coeffs[k] : 0.0f; };   // -> bank sgm
    // the half that owns banks [bb, be): tables, chunk range, pending banks; false if it cannot be expressed
    auto build_half = [&](WaveTcHalf &h, int bb, int be) {
        memset(&h, 0, sizeof(h));
        h.b_begin = bb; h.b_end = be;
        int kmin = WT_NBIN, kmax = -1;
        for (int k = 0; k < WT_NBIN; ++k) {
            const int sgm = mt.banks[k];
            if (sgm < 0) continue;
            const float wl = (sgm - 1 >= bb && sgm - 1 < be) ? bin_lo_w(k) : 0.0f, wh = (sgm >= bb && sgm < be) ? bin_hi_w(k) : 0.0f;
            (&h.wlo[k >> 2].x)[k & 3] = wl;
            (&h.whi[k >> 2].x)[k & 3] = wh;
            if (sgm - 1 >= bb && sgm - 1 < be) { kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
            if (sgm >= bb && sgm < be) { kmin = std::min(kmin, k); kmax = std::max(kmax, k); }
        }
        if (kmax < 0) { h.c_begin = h.c_end = 0; h.z_begin = bb; h.z_end = be; h.b_begin = h.b_end = 0; return; }
        h.c_begin = kmin / 16; h.c_end = kmax / 16 + 1;
        h.cur0 = cur_at[16 * h.c_begin];
        int cur = h.cur0;
        std::vector<char> stored(nb + 2, 0);
        for (int k = 16 * h.c_begin; k < 16 * h.c_end; ++k) {
            if (!sh[k]) continue;
            h.shift[k >> 5] |= 1u << (k & 31);
            const int b = cur - 1;
            if (b >= bb && b < be) { h.emit[k >> 5] |= 1u << (k & 31); stored[b] = 1; }
            ++cur;
        }
        h.flush_lo = cur - 1 >= bb && cur - 1 < be;
        h.flush_hi = cur >= bb && cur < be;
        if (h.flush_lo) stored[cur - 1] = 1;
        if (h.flush_hi) stored[cur] = 1;
        // banks of the range that no bin reaches (below the first / above the last filter): one contiguous run at either end
        int zb = bb, ze = bb;
        for (int b = bb; b < be; ++b)
            if (!stored[b]) { if (ze == zb) zb = b; ze = b + 1; }
        h.z_begin = zb; h.z_end = ze;
        int fb = be, fe = bb;
        for (int b = bb; b < be; ++b)
            if (stored[b]) { fb = std::min(fb, b); fe = std::max(fe, b + 1); }
        h.b_begin = fb; h.b_end = fe > fb ? fe : fb;
    };
    // replay of the device walk on a test spectrum against the filterbank's definition
    auto check_half = [&](const WaveTcHalf &h, int bb, int be) {
        std::vector<double> P(WT_NBIN), want(nb, 0.0), got(nb, -1.0);
        for (int k = 0; k < WT_NBIN; ++k) P[k] = 1.0 + 0.37 * k + (k % 7) * 0.11;
        for (int k = 0; k < WT_NBIN; ++k) {
            const int sgm = mt.banks[k];
            if (sgm < 0) continue;
            if (sgm >= 1 && sgm <= nb) want[sgm - 1] += (double)bin_lo_w(k) * P[k];
            if (sgm < nb) want[sgm] += (double)bin_hi_w(k) * P[k];
        }
        double lo = 0.0, hi = 0.0;
        int p = h.cur0 - 1;
        for (int k = 16 * h.c_begin; k < 16 * h.c_end; ++k) {
            const bool s1 = (h.shift[k >> 5] >> (k & 31)) & 1u, e1 = (h.emit[k >> 5] >> (k & 31)) & 1u;
            if (e1) { if (p < 0 || p >= nb) return false; got[p] = lo; }
            if (s1) { lo = hi; hi = 0.0; ++p; }
            lo += (double)(&h.wlo[k >> 2].x)[k & 3] * P[k];
            hi += (double)(&h.whi[k >> 2].x)[k & 3] * P[k];
        }
        if (h.flush_lo) { if (p < 0 || p >= nb) return false; got[p] = lo; }
        if (h.flush_hi) { if (p + 1 < 0 || p + 1 >= nb) return false; got[p + 1] = hi; }
        for (int b = h.z_begin; b < h.z_end; ++b) got[b] = 0.0;
        for (int b = bb; b < be; ++b)
            if (fabs(got[b] - want[b]) > 1e-9 * (1.0 + fabs(want[b]))) return false;
        for (int b = h.b_begin; b < h.b_end; ++b)
            if (got[b] < 0.0) return false;      // (a bank inside the finishing range that was never stored)
        for (int b = h.z_begin; b < h.z_end; ++b)
            if (b >= h.b_begin && b < h.b_end) return false;
        return true;
    };
    // split the banks where the two halves walk about the same number of chunks
    int best = -1, best_cost = 1 << 30;
    for (int split = 0; split <= nb; ++split) {
        WaveTcHalf h0, h1;
        build_half(h0, 0, split);
        build_half(h1, split, nb);
        if (!check_half(h0, 0, split) || !check_half(h1, split, nb)) continue;
        const int cost = std::max(h0.c_end - h0.c_begin, h1.c_end - h1.c_begin);
        if (cost > WT_HCH) continue;
        if (cost < best_cost) { best_cost = cost; best = split; }
    }
    if (best < 0) return PHN_OK;
    build_half(st->tab.h[0], 0, best);
    build_half(st->tab.h[1], best, nb);
    st->ok = true;
    return PHN_OK;
}

void wave_tc_release(phn_ctx *c)
{
    if (!c->wave_tc) return;
    WaveTcState *st = static_cast<WaveTcState *>(c->wave_tc);
    if (st->w_img) cudaFree(st->w_img);
    delete st;
    c->wave_tc = nullptr;
}

// the tensor-core front end serves this call (model fits, plain waveform scaling, not switched off)
bool wave_tc_applies(phn_ctx *c)
{
    static const bool off = getenv("PHNREC_WAVE_TC") && atoi(getenv("PHNREC_WAVE_TC")) == 0;
    if (off || !c->wave_tc) return false;
    const WaveTcState *st = static_cast<const WaveTcState *>(c->wave_tc);
    return st->ok && (c->fmt == PHN_WAVE_ALAW || c->fmt == PHN_WAVE_LIN16) && c->dc_shift == 0.0f && c->scale == 1.0f && wave_tc_fits(c);
}

int launch_wave_tc(phn_ctx *c, const void *d_audio, int64_t f_begin, int64_t f_end)
{
    const WaveTcState *st = static_cast<const WaveTcState *>(c->wave_tc);
    WaveTcArgs a;
    a.audio = (const uint8_t *)d_audio; a.audio_end = a.audio + c->total_bytes;
    a.byte_off = (const int64_t *)c->d_byte_off.p; a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt; a.f_begin = f_begin; a.f_end = f_end;
    a.vs = c->vs; a.step = c->step; a.nbanks = c->nbanks;
    a.frame_shift = c->frame_shift; a.frame_floor = c->frame_floor;
    a.w_img = st->w_img; a.mel = (float *)c->d_mel.p;
    a.tab = st->tab;
    static const int dbg = getenv("PHNREC_WTC_DBG") ? atoi(getenv("PHNREC_WTC_DBG")) : 0;
    a.dbg = dbg;
    a.tl = nullptr;
    static long long *d_tl = nullptr;
    if (dbg & 8) {
        if (!d_tl) PHN_CUDA(c, cudaMalloc((void **)&d_tl, sizeof(long long) * (4 * 32 * 8 + 8)));
        PHN_CUDA(c, cudaMemsetAsync(d_tl, 0, sizeof(long long) * (4 * 32 * 8 + 8), c->stream));
        a.tl = d_tl;
    }
    const bool lin16 = c->fmt == PHN_WAVE_LIN16;
    auto kern = dbg ? (lin16 ? k_wave_tc<true, true> : k_wave_tc<true, false>) : (lin16 ? k_wave_tc<false, true> : k_wave_tc<false, false>);
    // (a function attribute belongs to the device's context: set per launch, like k_mlp_tc - contexts on several GPUs share this code)
    PHN_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM));
    const int64_t units = (f_end - f_begin + 255) / 256;
    const int64_t maxp = c->num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (units < maxp ? units : maxp))); cfg.blockDim = dim3(lin16 ? WT_THREADS<true> : WT_THREADS<false>);
    cfg.dynamicSmemBytes = WT_SMEM; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PHN_CUDA(c, cudaLaunchKernelEx(&cfg, kern, a));
    PHN_CUDA(c, cudaGetLastError());
    if (a.tl) {   // print the timeline of tiles 8..15 (clocks relative to the first event of tile 8)
        std::vector<long long> h(4 * 32 * 8 + 8);
        PHN_CUDA(c, cudaStreamSynchronize(c->stream));
        PHN_CUDA(c, cudaMemcpy(h.data(), d_tl, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        const char *names[4] = {"mma", "prod", "epi_lo", "epi_hi"};
        const long long t0 = h[(0 * 32 + 8) * 8];
        fprintf(stderr, "wtc kernel start %lld; mma issue-done per tile:", h[4 * 32 * 8] - t0);
        for (int it = 0; it < 32; ++it) fprintf(stderr, " %lld", h[(0 * 32 + it) * 8 + 3] - t0);
        fprintf(stderr, "\n");
        for (int it = 8; it < 16; ++it)
            for (int r = 0; r < 4; ++r) {
                fprintf(stderr, "wtc tile %2d %-6s", it, names[r]);
                for (int e = 0; e < 8; ++e) fprintf(stderr, " %8lld", h[(r * 32 + it) * 8 + e] ? h[(r * 32 + it) * 8 + e] - t0 : -1);
                fprintf(stderr, "\n");
            }
    }
    return PHN_OK;
}

}  // namespace phn
