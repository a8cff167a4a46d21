// k_wave_tc.cu — K-wave for the tensor-core pipeline: the windowed DFT of every frame as ONE GEMM on tcgen05.
//
// Replaces, like k_wave.cu, ConvertWaveformFormat (srec.cpp:709-791), MelBanks::ProcessFrame (melbanks.cpp:111-204),
// cFour1 / _mbApply (dspc.cpp:24-78, 236-269), cPower / sLn (dspc.h:141-160) and FrameBasedNormalization
// (srec.cpp:1594-1620) - for the configuration the benchmark is quoted on (8 kHz A-law, 25 ms window = 200 samples,
// 256-point transform, plain front end).  Everything else keeps the register-FFT kernels of k_wave.cu.
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
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace phn {

constexpr int WT_K = 208;                 // window columns fed to the tensor cores (13 k-steps of 16; the window is <= 208)
constexpr int WT_NBIN = 128;              // bins 0..127 of the 256-point transform
constexpr int WT_PROD = 8;                // producer warps per CTA
constexpr int WT_EPI = 4;                 // epilogue warps per CTA (warp = TMEM lane quarter)
constexpr int WT_THREADS = (WT_EPI + 1 + WT_PROD) * 32;
constexpr int WT_BLK = 16384;             // [128 rows x 64 fp16], SWIZZLE_128B
constexpr int WT_BBLK = 7;                // matrix blocks per CTA: 3 hi, 3 lo, 1 shared tail
constexpr size_t WT_SMEM = (size_t)(WT_BBLK + 2 * 3 + 1) * WT_BLK + 128 + 1024;   // + barriers + alignment slack

struct WaveTcTab {                        // kernel parameters: read with compile-time offsets from the constant bank
    float4 wlo[WT_NBIN / 4];              // bin k -> c[k] for bank Banks[k] - 1   (0 when that bank does not exist)
    float4 whi[WT_NBIN / 4];              // bin k -> 1 - c[k] for bank Banks[k]
    uint32_t shift[WT_NBIN / 32];         // bit k % 32 of word k / 32: Banks[] grows (by one) at bin k, relative to the last bin inside the filterbank
};

struct WaveTcArgs {
    const uint8_t *audio, *audio_end;     // the batch's audio and one past its last byte
    const int64_t *byte_off, *frame_off;
    int n_utt;
    int64_t f_begin, f_end;               // frames of this launch
    int vs, step, nbanks;
    int dbg;                              // kernel development (PHNREC_WTC_DBG): 1 producers store zeros, 2 epilogue skips the filterbank, 4 no MMAs
    float frame_shift, frame_floor;
    const uint8_t *w_img;                 // [2 ranks][WT_BBLK][16 KB]
    float *mel;
    WaveTcTab tab;
};

__device__ __forceinline__ void mbar_arrive_cluster_rel(uint64_t *bar, uint32_t cta)
{
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

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

// three consecutive words at p when fewer than 12 bytes are left in the batch's audio buffer (its very last chunks)
__device__ __noinline__ uint3 load3_tail(const uint8_t *p, const uint8_t *end)
{
    uint32_t r[3] = {0u, 0u, 0u};
    for (int i = 0; i < 12; ++i)
        if (p + i < end) r[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
    return make_uint3(r[0], r[1], r[2]);
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

__global__ void __launch_bounds__(WT_THREADS, 1) k_wave_tc(const __grid_constant__ WaveTcArgs a)
{
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = blockIdx.x & 1u;
    const int64_t nf = a.f_end - a.f_begin;
    const int n_units = (int)((nf + 255) / 256), unit0 = (int)(blockIdx.x >> 1), ustep = (int)(gridDim.x >> 1);
    const int n_my = unit0 < n_units ? (n_units - 1 - unit0) / ustep + 1 : 0;
    constexpr int WARP_MMA = WT_EPI, PROD0 = WT_EPI + 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 2 * WT_PROD); mbar_init(&a_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 2 * WT_EPI);
        }
        mbar_init(b_full, 1); mbar_init(pb_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
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
            if (lane == 0) mbar_arrive_cluster_rel(pb_full, 0);
        } else {
            mbar_wait_acq_cluster(pb_full, 0);
            constexpr uint32_t idesc = make_idesc(256, false, 256);
            const uint64_t dA = make_sw128_desc(smem_u32(sA)), dB = make_sw128_desc(smem_u32(sB)), dT = make_sw128_desc(smem_u32(sT));
            const uint32_t alo0 = (uint32_t)dA, blo0 = (uint32_t)dB, tlo0 = (uint32_t)dT, hi = (uint32_t)(dA >> 32);
            const uint32_t bar_ae = smem_u32(a_empty), bar_df = smem_u32(d_full);
            const bool leader = elect_one();
#pragma unroll 1
            for (int it = 0; it < n_my; ++it) {
                const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
                mbar_wait_acq_cluster(&a_full[s], ph);
                if (it >= 2) mbar_wait_acq_cluster(&d_empty[s], ph ^ 1u);
                tc_fence_after();
                if (leader) {
                    const uint32_t td = tmem + 256u * s;
                    if (!(a.dbg & 4)) {
                    const uint32_t alo = alo0 + s * (3u * (WT_BLK >> 4));
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        const uint32_t blo = blo0 + (uint32_t)part * (3u * (WT_BLK >> 4));
#pragma unroll
                        for (int b = 0; b < 3; ++b)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t o = (uint32_t)b * (WT_BLK >> 4) + 2u * (uint32_t)ks;
                                if (part == 0 && b == 0 && ks == 0) umma2_ss_lo<0>(td, alo + o, hi, blo + o, idesc);
                                else umma2_ss_lo<1>(td, alo + o, hi, blo + o, idesc);
                            }
                        // the window's tail: stage s of the A tail block against part `part` of the matrix tail block
                        umma2_ss_lo<1>(td, tlo0 + 2u * s, hi, blo0 + 6u * (WT_BLK >> 4) + 2u * (uint32_t)part, idesc);
                    }
                    }
                    tc_commit2_u(bar_ae + 8u * s);
                    tc_commit2_u(bar_df + 8u * s);
                }
                __syncwarp();
            }
        }
    } else if (warp >= PROD0) {
        // ===================================================================== producers: audio -> A tile
        // A warp owns 16 consecutive rows of the tile.  Per tile: lanes 0..15 look up their row's utterance (source byte
        // offset, samples inside the signal), then the words of 8 rows are requested before the first is decoded - two
        // memory latencies per tile, not one per row - and the stage is only waited for when the stores begin.
        constexpr int RPW = 128 / WT_PROD, RG = 8;           // rows per warp and tile; rows per round of loads
        const int pw = warp - PROD0;
        const int cb = lane >> 3, cc = lane & 7;             // this lane's chunk: block, 16-byte column inside the block
        const bool active = lane < WT_K / 8;
        int u = 0;
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
            const int64_t g0 = a.f_begin + ((int64_t)(unit0 + it * ustep) * 256 + (int64_t)rank * 128) + pw * RPW;
            // row info on lanes 0..15
            int64_t soff = 0;
            int lim = -1;                                    // -1: a row past the end of the launch (zeros)
            {
                const int64_t g = g0 + (lane & (RPW - 1));
                if (g < a.f_end && !(a.dbg & 1)) {
                    if (!(a.frame_off[u] <= g && g < a.frame_off[u + 1])) u = find_utt(a.frame_off, a.n_utt, g);
                    const int64_t b0 = a.byte_off[u], len = a.byte_off[u + 1] - b0;
                    const int64_t s0 = (g - a.frame_off[u]) * a.step, left = len - s0;
                    soff = b0 + s0;
                    lim = left < a.vs ? (left < 0 ? 0 : (int)left) : a.vs;   // samples of the window inside the signal
                }
            }
            // where this lane's chunk of the warp's row 0 goes; a row is 128 bytes further, its chunk index XORed with row % 8
            uint8_t *base = (cb < 3 ? sA + ((size_t)s * 3 + cb) * WT_BLK : sT) + (size_t)pw * RPW * 128;
            const uint32_t col = cb < 3 ? (uint32_t)cc : 2u * s + (uint32_t)cc;
            const uint8_t *lane_src = a.audio + 8 * lane;
#pragma unroll 1
            for (int h = 0; h < RPW / RG; ++h) {
                uint32_t w[RG][3];
#pragma unroll
                for (int rr = 0; rr < RG; ++rr) {
                    const int64_t so = __shfl_sync(0xffffffffu, soff, h * RG + rr);
                    const int lm = __shfl_sync(0xffffffffu, lim, h * RG + rr);
                    w[rr][0] = w[rr][1] = w[rr][2] = 0u;
                    if (active && lm >= 0) {
                        const uint8_t *p = reinterpret_cast<const uint8_t *>(reinterpret_cast<uintptr_t>(lane_src + so) & ~(uintptr_t)3);
                        if (p + 12 <= a.audio_end) {
                            w[rr][0] = __ldg(reinterpret_cast<const unsigned int *>(p));
                            w[rr][1] = __ldg(reinterpret_cast<const unsigned int *>(p + 4));
                            w[rr][2] = __ldg(reinterpret_cast<const unsigned int *>(p + 8));
                        } else {
                            const uint3 t = load3_tail(p, a.audio_end);
                            w[rr][0] = t.x; w[rr][1] = t.y; w[rr][2] = t.z;
                        }
                    }
                }
                if (h == 0 && it >= 2) mbar_wait(&a_empty[s], ph ^ 1u);
#pragma unroll
                for (int rr = 0; rr < RG; ++rr) {
                    const int64_t so = __shfl_sync(0xffffffffu, soff, h * RG + rr);
                    const int lm = __shfl_sync(0xffffffffu, lim, h * RG + rr);
                    const uint32_t sh = (((uint32_t)reinterpret_cast<uintptr_t>(a.audio) + (uint32_t)so) & 3u) * 8u;   // (8 * lane keeps the alignment)
                    const uint32_t lo = __funnelshift_r(w[rr][0], w[rr][1], sh), hi = __funnelshift_r(w[rr][1], w[rr][2], sh);
                    uint4 out;
                    out.x = alaw2_half2(__byte_perm(lo, 0u, 0x4140u));
                    out.y = alaw2_half2(__byte_perm(lo, 0u, 0x4342u));
                    out.z = alaw2_half2(__byte_perm(hi, 0u, 0x4140u));
                    out.w = alaw2_half2(__byte_perm(hi, 0u, 0x4342u));
                    if (lm < a.vs) {   // an utterance shorter than one window: zeros beyond the signal (melbanks.cpp:151-170); a row past the end: zeros
                        const int n = lm - 8 * lane;         // samples of this chunk inside the signal
                        uint32_t *o = &out.x;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (2 * i >= n) o[i] = 0u;
                            else if (2 * i + 1 >= n) o[i] &= 0x0000FFFFu;
                        }
                    }
                    // (row % 8 = rr: RG = 8 and a group's first row is a multiple of 8)
                    if (active) *reinterpret_cast<uint4 *>(base + (size_t)(h * RG + rr) * 128 + ((col ^ (uint32_t)rr) << 4)) = out;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_rel(&a_full[s], 0);
        }
    } else {
        // ===================================================================== epilogue: |Z|^2 -> filterbank -> ln -> mel
        const int nb = a.nbanks;
        const float fl_eff = a.frame_floor != -9999.9f ? a.frame_floor : -INFINITY;   // srec.cpp:1594-1620 (floor switched off: -9999.9)
        // one finished bank: sLn (dspc.h:155-160: ln, digital silence -> 0) as lg2.approx * ln 2, then the frame normalisation
        auto finish = [&](float acc) {
            float o = acc > 0.0f ? lg2_approx_wt(acc) * 0.69314718055994530942f : 0.0f;
            return fmaxf(o + a.frame_shift, fl_eff);
        };
#pragma unroll 1
        for (int it = 0; it < n_my; ++it) {
            const uint32_t s = (uint32_t)it & 1u, ph = ((uint32_t)it >> 1) & 1u;
            const int64_t g = a.f_begin + ((int64_t)(unit0 + it * ustep) * 256 + (int64_t)rank * 128 + warp * 32 + lane);
            const bool live = g < a.f_end;
            float *dst = a.mel + (live ? g : a.f_begin) * nb;
            mbar_wait(&d_full[s], ph);
            tc_fence_after();
            const uint32_t tbase = tmem + 256u * s + ((uint32_t)(warp * 32) << 16);
            // Banks[] (dspc.cpp:236-269) does not decrease with the bin: bank `cur` collects (1 - c) P from the bins with
            // Banks = cur (acc_hi) and bank cur - 1 collects c P from the same bins (acc_lo); when Banks moves on - at most one
            // step per bin, wave_tc_prepare checks - bank cur - 1 is complete.  The tables are kernel parameters (constant bank).
            float acc_lo = 0.0f, acc_hi = 0.0f;
            int cur = 0;
            uint32_t v0[32], v1[32];
            auto bins16 = [&](const uint32_t *v, int c2, int half) {   // bins 32 c2 + 16 half + (0..15)
                const uint32_t mask = a.tab.shift[c2] >> (16 * half);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 l4 = a.tab.wlo[8 * c2 + 4 * half + q], h4 = a.tab.whi[8 * c2 + 4 * half + q];
                    const float lw[4] = {l4.x, l4.y, l4.z, l4.w}, hw[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = 4 * q + j;
                        const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                        const float pwr = fmaf(im, im, re * re);  // cPower (dspc.h:141-146)
                        if (mask & (1u << i)) {
                            if (live && cur >= 1 && cur <= nb) dst[cur - 1] = finish(acc_lo);
                            acc_lo = acc_hi; acc_hi = 0.0f; ++cur;
                        }
                        acc_lo = fmaf(pwr, lw[j], acc_lo);
                        acc_hi = fmaf(pwr, hw[j], acc_hi);
                    }
                }
            };
            tmem_ld32(tbase, v0);
#pragma unroll 1
            for (int c2 = 0; c2 < 4; ++c2) {                 // 32 bins per round: the code stays small (tables at run-time offsets)
                tmem_ld_wait();
                tmem_ld32(tbase + 64u * (uint32_t)c2 + 32u, v1);
                if (!(a.dbg & 2)) bins16(v0, c2, 0);
                tmem_ld_wait();
                if (c2 < 3) tmem_ld32(tbase + 64u * (uint32_t)c2 + 64u, v0);
                else {                                       // the accumulator is in registers: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_rel(&d_empty[s], 0);
                }
                if (!(a.dbg & 2)) bins16(v1, c2, 1);
            }
            if (live) {
                if (cur >= 1 && cur <= nb) dst[cur - 1] = finish(acc_lo);
                if (cur < nb) dst[cur] = finish(acc_hi);
                for (int b = cur + 1; b < nb; ++b) dst[b] = finish(0.0f);
            }
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
    return c->mt.logN == 8 && c->vs <= WT_K && c->nbanks >= 1 && c->nbanks <= 64 && !c->plp && !c->z_mean && c->preem == 0.0f;
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
    int prev = 0;
    memset(&st->tab, 0, sizeof(st->tab));
    for (int k = 0; k < WT_NBIN; ++k) {
        const int sgm = mt.banks[k];
        if (sgm < 0) continue;
        if (sgm < prev || sgm - prev > 1) return PHN_OK;     // banks narrower than a bin: the FFT kernels serve this model
        if (sgm > prev) st->tab.shift[k >> 5] |= 1u << (k & 31);
        prev = sgm;
        if (sgm >= 1 && sgm <= mt.nbanks) (&st->tab.wlo[k >> 2].x)[k & 3] = mt.coeffs[k];
        if (sgm < mt.nbanks) (&st->tab.whi[k >> 2].x)[k & 3] = 1.0f - mt.coeffs[k];
    }
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

// the tensor-core front end serves this call (model fits, A-law, plain waveform scaling, not switched off)
bool wave_tc_applies(phn_ctx *c)
{
    static const bool off = getenv("PHNREC_WAVE_TC") && atoi(getenv("PHNREC_WAVE_TC")) == 0;
    if (off || !c->wave_tc) return false;
    const WaveTcState *st = static_cast<const WaveTcState *>(c->wave_tc);
    return st->ok && c->fmt == PHN_WAVE_ALAW && c->dc_shift == 0.0f && c->scale == 1.0f && wave_tc_fits(c);
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
    static bool attr_set = false;
    if (!attr_set) {
        PHN_CUDA(c, cudaFuncSetAttribute(k_wave_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM));
        attr_set = true;
    }
    const int64_t units = (f_end - f_begin + 255) / 256;
    const int64_t maxp = c->num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * (units < maxp ? units : maxp))); cfg.blockDim = dim3(WT_THREADS);
    cfg.dynamicSmemBytes = WT_SMEM; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PHN_CUDA(c, cudaLaunchKernelEx(&cfg, k_wave_tc, a));
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

}  // namespace phn
