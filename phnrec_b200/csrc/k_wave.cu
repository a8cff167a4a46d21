// k_wave.cu — K-wave: waveform decode + framing + Hamming + FFT + power + mel filterbank + guarded ln.
//
// Replaces ConvertWaveformFormat (srec.cpp:709-791), MelBanks::GetFeatures/ProcessFrame
// (melbanks.cpp:111-204), cFour1 / _mbApply (dspc.cpp:24-78, 236-269), cPower / sLn (dspc.h:141-160)
// and FrameBasedNormalization (srec.cpp:1594-1620) for a ragged batch of utterances.
//
// One warp owns one frame at a time.  The N-point FFT is the reference's radix-2
// decimation-in-time, regrouped into register passes of three stages (radix 8; the last pass is
// radix N/64) with two shared-memory transposes in between; each lane fetches its own samples
// straight from the audio in bit-reversed order.  Butterflies are independent within a stage, so
// regrouping changes no rounding: in the EXACT instantiation (reference twiddles in double, built
// on the host by the same recurrence, and the reference's rounding points) mel values are
// bit-identical to the CPU implementation.  The !EXACT instantiation (tensor-core pipeline) uses
// fp32 FMAs.
#include "internal.h"
#include <cstdlib>
#include "device_math.cuh"

#include <algorithm>
#include <type_traits>

namespace phn {

struct WaveArgs {
    const uint8_t *audio;
    const int64_t *byte_off, *frame_off, *pair_off;
    int n_utt;
    int64_t total_frames;
    int64_t f_begin, f_end;   // frame range of this launch (a group of whole utterances)
    int64_t p_begin, p_end;   // the same range in frame pairs (k_wave_pair)
    int fmt, vs, step, N, logN, nbanks;
    float scale, dc_shift, frame_shift, frame_floor, preem;
    int z_mean;
    int raw_energy;           // 1: bank energies as they are (no logarithm, no frame normalisation): the input of K-plp
    int mel_len;              // k_wave_pair: longest bank range in bins (rows of the filterbank weight table)
    const float *hamming, *coeffs;
    const int *banks, *klo, *khi;
    const double2 *tw;
    float *mel;
};

constexpr int kWaveWarps = 8;

__device__ __forceinline__ int find_utt(const int64_t *off, int n, int64_t f)
{
    int lo = 0, hi = n;  // off[lo] <= f < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

// One radix-2 butterfly of cFour1 (dspc.cpp:55-76) on (lo, up) = (element i, element i+h).
// EXACT: the reference's arithmetic - twiddle in double, products and their sum/difference rounded
// in double, one rounding to float, then float add/sub.  Twiddle (1, 0) needs no arithmetic at all:
// fl32(fl64(1*x) - fl64(0*y)) == x (up to the sign of a zero, which |X|^2 cannot see).
// !EXACT: plain fp32 FMAs.
template <bool EXACT, typename TW>
__device__ __forceinline__ void bfly(float2 &lo, float2 &up, const TW w, const bool trivial)
{
    float tr, ti;
    if (trivial) {
        tr = up.x; ti = up.y;
    } else if (EXACT) {
        const double kr = (double)up.x, ki = (double)up.y;
        tr = __double2float_rn(__dsub_rn(__dmul_rn((double)w.x, kr), __dmul_rn((double)w.y, ki)));
        ti = __double2float_rn(__dadd_rn(__dmul_rn((double)w.x, ki), __dmul_rn((double)w.y, kr)));
    } else {
        const float wr = (float)w.x, wi = (float)w.y;
        tr = fmaf(wr, up.x, -wi * up.y);
        ti = fmaf(wr, up.y, wi * up.x);
    }
    up.x = __fsub_rn(lo.x, tr); up.y = __fsub_rn(lo.y, ti);
    lo.x = __fadd_rn(lo.x, tr); lo.y = __fadd_rn(lo.y, ti);
}

// LEVELS consecutive radix-2 stages on R = 2^LEVELS elements held in registers.
// v[r] is element e = base + S*r of the length-N array; stage half-sizes S, 2S, 4S.
// Twiddle of the pair (e, e+h) is tw[h-1 + (e mod h)], e mod h = base_mod + S*(r mod hh).
template <bool EXACT, int LEVELS, typename TW>
__device__ __forceinline__ void fft_pass(float2 *v, const TW *s_tw, int S, int base_mod)
{
    constexpr int R = 1 << LEVELS;
#pragma unroll
    for (int lv = 0; lv < LEVELS; ++lv) {
        const int hh = 1 << lv;
        const int h = S * hh;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r & hh) continue;
            const int m = base_mod + S * (r & (hh - 1));
            bfly<EXACT, TW>(v[r], v[r + hh], s_tw[h - 1 + m], m == 0);
        }
    }
}

__device__ __forceinline__ int pad_idx(int e) { return e + (e >> 5); }  // breaks the stride-8/64 bank patterns

template <bool EXACT, int LOGN>
__global__ void __launch_bounds__(kWaveWarps * 32) k_wave(WaveArgs a)
{
    constexpr int N = 1 << LOGN, N2 = N / 2;
    constexpr int OCT = N / 256;            // octets per lane in passes 1 and 2 (1 or 2)
    constexpr int L3 = LOGN - 6;            // levels of the last pass (2 or 3)
    constexpr int R3 = 1 << L3;             // its radix (4 or 8); N/R3 = 64 groups -> 2 per lane
    constexpr int WORK = N + N / 32 + N2 / 2;  // float2 units per warp: padded data[] + pw[N2]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using TW = typename std::conditional<EXACT, double2, float2>::type;   // fp32 pipeline: twiddles rounded to float once
    TW *s_tw = reinterpret_cast<TW *>(smem_raw);                    // [N-1] (+1 pad) in a double2-sized slot each
    float *s_ham = reinterpret_cast<float *>(smem_raw + sizeof(double2) * N);  // [N], zero beyond vs
    float *s_coef = s_ham + N;                                      // [N2]
    int *s_bank = reinterpret_cast<int *>(s_coef + N2);             // [N2]
    double *s_logtab = reinterpret_cast<double *>(s_bank + N2);     // [32] glibc logf table
    float2 *s_work = reinterpret_cast<float2 *>(s_logtab + 32);
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < N - 1; i += blockDim.x) {
        const double2 w = a.tw[i];
        if (EXACT) reinterpret_cast<double2 *>(s_tw)[i] = w;
        else reinterpret_cast<float2 *>(s_tw)[i] = make_float2((float)w.x, (float)w.y);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_ham[i] = i < a.vs ? a.hamming[i] : 0.0f;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) { s_coef[i] = a.coeffs[i]; s_bank[i] = a.banks[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 *data = s_work + (size_t)warp * WORK;
    float *pw = reinterpret_cast<float *>(data + N + N / 32);
    const int bps = a.fmt == PHN_WAVE_LIN16 ? 2 : 1;
    const bool plain = !a.z_mean && a.preem == 0.0f;

    for (int64_t f = a.f_begin + (int64_t)blockIdx.x * kWaveWarps + warp; f < a.f_end; f += (int64_t)gridDim.x * kWaveWarps) {
        const int u = find_utt(a.frame_off, a.n_utt, f);
        const int64_t t = f - a.frame_off[u];
        const int64_t b0 = a.byte_off[u];
        const int64_t len = (a.byte_off[u + 1] - b0) / bps;  // samples in this utterance
        const int64_t s0 = t * a.step;

        auto sample = [&](int i) -> float {  // decode (srec.cpp:742-743 / 768-769), dc shift, scale; 0 beyond the signal
            float x = 0.0f;
            if (i < a.vs && s0 + i < len) {
                if (a.fmt == PHN_WAVE_LIN16) {
                    const uint8_t *p = a.audio + b0 + 2 * (s0 + i);
                    x = (float)(short)((unsigned)p[0] | ((unsigned)p[1] << 8));
                } else {
                    x = __fmul_rn(8.0f, (float)alaw_d5(a.audio[b0 + s0 + i]));
                }
                if (a.dc_shift != 0.0f) x = __fadd_rn(x, a.dc_shift);
                if (a.scale != 1.0f) x = __fmul_rn(x, a.scale);
            }
            return x;
        };

        float *stage = reinterpret_cast<float *>(data);  // N floats; data[] is free until pass 1 stores
        if (!plain) {
            // optional sSubtractAverage / sPreemphasis (dspc.h:63-84), not used by the shipped systems
            for (int i = lane; i < N; i += 32) stage[i] = sample(i);
            __syncwarp();
            float avg = 0.0f;
            if (a.z_mean) {
                if (lane == 0) {
                    float s = 0.0f;
                    for (int i = 0; i < a.vs; ++i) s = __fadd_rn(s, stage[i]);
                    pw[0] = __fdiv_rn(s, (float)a.vs);
                }
                __syncwarp();
                avg = pw[0];
                __syncwarp();
            }
            float y[N / 32];
#pragma unroll
            for (int q = 0; q < N / 32; ++q) {
                const int i = lane + 32 * q;
                float x = 0.0f;
                if (i < a.vs) {
                    x = a.z_mean ? __fsub_rn(stage[i], avg) : stage[i];
                    if (a.preem != 0.0f) {
                        if (i == 0) x = __fmul_rn(x, __fsub_rn(1.0f, a.preem));
                        else {
                            const float xp = a.z_mean ? __fsub_rn(stage[i - 1], avg) : stage[i - 1];
                            x = __fsub_rn(x, __fmul_rn(a.preem, xp));
                        }
                    }
                }
                y[q] = x;
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < N / 32; ++q) stage[lane + 32 * q] = y[q];
            __syncwarp();
        }

        // ---- pass 1 (stages h = 1, 2, 4): elements 8o .. 8o+7 = windowed inputs at bit-reversed positions
        float2 v[OCT][8];
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = (int)(__brev((unsigned)(8 * o + r)) >> (32 - LOGN));
                const float x = plain ? sample(i) : stage[i];
                v[oc][r] = make_float2(__fmul_rn(x, s_ham[i]), 0.0f);
            }
        }
        if (!plain) __syncwarp();
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            fft_pass<EXACT, 3, TW>(v[oc], s_tw, 1, 0);
#pragma unroll
            for (int r = 0; r < 8; ++r) data[pad_idx(8 * o + r)] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 2 (h = 8, 16, 32): elements low3 + 8r + 64*high
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
#pragma unroll
            for (int r = 0; r < 8; ++r) v[oc][r] = data[pad_idx(low3 + 8 * r + 64 * high)];
        }
        __syncwarp();
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
            fft_pass<EXACT, 3, TW>(v[oc], s_tw, 8, low3);
#pragma unroll
            for (int r = 0; r < 8; ++r) data[pad_idx(low3 + 8 * r + 64 * high)] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 3 (h = 64 .. N/2): elements low6 + 64r, two groups per lane; then |X|^2 of bins < N/2
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const int low6 = lane + 32 * gq;
            float2 w3[R3];
#pragma unroll
            for (int r = 0; r < R3; ++r) w3[r] = data[pad_idx(low6 + 64 * r)];
            fft_pass<EXACT, L3, TW>(w3, s_tw, 64, low6);
#pragma unroll
            for (int r = 0; r < R3 / 2; ++r)   // cPower (dspc.h:141-146)
                pw[low6 + 64 * r] = __fadd_rn(__fmul_rn(w3[r].x, w3[r].x), __fmul_rn(w3[r].y, w3[r].y));
        }
        __syncwarp();

        // ---- mel filterbank (dspc.cpp:236-269): lane b accumulates bank b in ascending bin order
        if (lane < a.nbanks) {
            float acc = 0.0f;
            const int hi = a.khi[lane];
            for (int k = a.klo[lane]; k <= hi; ++k) {
                const float p = pw[k];
                const float v2 = __fmul_rn(s_coef[k], p);
                acc = __fadd_rn(acc, s_bank[k] == lane ? __fsub_rn(p, v2) : v2);
            }
            float o = acc;
            if (!a.raw_energy) {
                o = ln_guarded(acc, s_logtab);
                if (a.frame_shift != 0.0f) o = __fadd_rn(o, a.frame_shift);           // srec.cpp:1594-1620
                if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
            }
            a.mel[f * a.nbanks + lane] = o;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core pipeline (fp32 front end): TWO real frames (2i, 2i + 1 of one utterance) per complex FFT.  z = a + i b; by linearity and the
// conjugate symmetry of real signals  A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / 2i,  so
// |A[k]|^2 = ((Zr[k] + Zr[N-k])^2 + (Zi[k] - Zi[N-k])^2) / 4,  |B[k]|^2 = ((Zr[k] - Zr[N-k])^2 + (Zi[k] + Zi[N-k])^2) / 4.
// Half the butterflies per frame; the price is fp32 cross-talk between the two frames at the 1e-7 level of the
// louder one.  An all-zero frame (digital silence) is detected on load and keeps the reference's exact answer
// (sLn's guard: mel = 0), whatever its partner holds.  Only for the plain front end (no z_mean / pre-emphasis).
//
// The kernel is bound by the shared-memory pipe (ncu: LSU wavefronts 87 % of peak before this layout), so everything
// that is loop invariant per lane lives in registers and only the two FFT transposes, the power spectra and the
// filterbank go through shared memory:
//   * the twiddles of a lane's butterflies do not depend on the frame: 7 + 7 + 2 (R3 - 1) float2 registers, loaded once;
//   * A-law bytes are expanded arithmetically (a 256-entry table indexed by random bytes costs ~3 bank conflicts per load);
//   * the filterbank is a dense per-bank weight table w[j][bank] (j = bin - klo[bank]; (1 - c) or c, 0 beyond the bank's
//     range): one conflict-free weight load + one spectrum load + one FMA per bin, the same trip count on every lane.

// LEVELS radix-2 stages on registers, twiddles in registers: twr[hh - 1 + j] belongs to stage half-size hh (in units of
// the pass stride), pair index j.  UNIT0: the j == 0 twiddle of every stage is (1, 0) (first pass) - no arithmetic.
template <int LEVELS, bool UNIT0>
__device__ __forceinline__ void fft_pass_r(float2 *v, const float2 *twr)
{
    constexpr int R = 1 << LEVELS;
#pragma unroll
    for (int lv = 0; lv < LEVELS; ++lv) {
        const int hh = 1 << lv;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r & hh) continue;
            const int j = r & (hh - 1);
            float2 &lo = v[r], &up = v[r + hh];
            float tr, ti;
            if (UNIT0 && j == 0) {
                tr = up.x; ti = up.y;
            } else {
                const float2 w = twr[hh - 1 + j];
                tr = fmaf(w.x, up.x, -w.y * up.y);
                ti = fmaf(w.x, up.y, w.y * up.x);
            }
            up.x = lo.x - tr; up.y = lo.y - ti;
            lo.x = lo.x + tr; lo.y = lo.y + ti;
        }
    }
}

// A-law byte -> 8 * ALawTableD5 value as a float, built in the exponent/mantissa fields (alaw.cpp:14-48; same value as
// 8 * alaw_d5): with t = byte ^ 0xD5 (bit 7: negative), segment s = t[6:4] >= 1 and mantissa m = t[3:0] the G.711 value
// (2m + 33) << (s - 1) is the float with exponent 131 + s and mantissa (2m + 1) / 32, i.e. its bits are a constant PLUS
// t[6:0] << 19 (the constant's exponent bits overlap the segment's); segment 0 holds 2m + 1 = twice that pattern's value minus 32.
__device__ __forceinline__ float alaw8_float(unsigned byte)
{
    const unsigned t = byte ^ 0xD5u;
    float v = __uint_as_float(((t << 19) & 0x03F80000u) + 0x43040000u);   // 8 * (2m + 33) * 2^(s - 1)
    if ((t & 0x70u) == 0u) v = fmaf(v, 2.0f, -256.0f);                    // segment 0: 8 * (2m + 1)
    return __uint_as_float(__float_as_uint(v) | ((t << 24) & 0x80000000u));
}

// utterance of work unit p: off[u] <= p < off[u + 1]; the warp probes 32 pivots per round (two rounds for 1000 utterances)
__device__ __forceinline__ int find_utt_warp(const int64_t *off, int n, int64_t p, int lane)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int stepw = (hi - lo + 31) >> 5;
        const int idx = lo + lane * stepw;
        const bool le = idx < hi && off[idx] <= p;
        const int cnt = __popc(__ballot_sync(0xffffffffu, le));   // off[] is non-decreasing and off[lo] <= p: cnt >= 1
        lo += (cnt - 1) * stepw;
        hi = min(hi, lo + stepw);
    }
    return lo;
}

#ifndef PHN_PAIR_WARPS
#define PHN_PAIR_WARPS 4
#endif
#ifndef PHN_PAIR_MINB
#define PHN_PAIR_MINB 4
#endif
constexpr int kPairWarps = PHN_PAIR_WARPS;

// per-warp shared memory: FFT work array (N float2 + 1/8 padding) and the filterbank table T
//   exchange 1 (written 8 consecutive elements per lane, read with stride 8):  element e at e + (e >> 3)
//   exchange 2 (written with stride 8, read with stride 64):                   element e at e + 8 (e >> 6)
// both conflict-free for 64-bit accesses (each half-warp covers 16 distinct 8-byte banks) with compile-time offsets.
// T[j * 33 + col]: contribution of the j-th bin of a bank's range to filterbank column col (= bank, + 16 for the second
// frame when both frames fit one warp; else the second frame has its own table).  The lanes that hold the power
// spectrum scatter weight * power into T (each bin feeds at most two banks); lane col then adds up its column.
// Slots that no bin maps to are zeroed once and stay zero.
__host__ __device__ constexpr int pair_tsz(int mel_len) { return (33 * mel_len + 32 + 3) & ~3; }   // + the trash slots (33 mel_len, + 16 for frame B)

template <int LOGN, bool ALAW>
__global__ void __launch_bounds__(kPairWarps * 32, LOGN == 8 ? PHN_PAIR_MINB : (LOGN == 9 ? 2 : 1)) k_wave_pair(WaveArgs a)
{
    constexpr int N = 1 << LOGN;
    constexpr int OCT = N / 256;
    constexpr int L3 = LOGN - 6;
    constexpr int R3 = 1 << L3;
    constexpr int NB = R3;                   // power-spectrum bins per lane: k = lane + 32 gq + 64 r, gq < 2, r < R3/2
    constexpr int DATA = N + N / 8;          // float2 units
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool both = a.nbanks <= 16;        // bank b of frame A on lane b, of frame B on lane 16 + b; else one frame after the other
    const int TSZ = pair_tsz(a.mel_len);
    const int twarp = both ? TSZ : 2 * TSZ;  // floats of T per warp
    double *s_logtab = reinterpret_cast<double *>(smem_raw);            // [32] glibc logf table
    float2 *s_data = reinterpret_cast<float2 *>(s_logtab + 32);         // kPairWarps x DATA
    float *s_T = reinterpret_cast<float *>(s_data + (size_t)kPairWarps * DATA);   // kPairWarps x twarp
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < kPairWarps * twarp; i += blockDim.x) s_T[i] = 0.0f;
    __syncthreads();

    float2 *data = s_data + (size_t)warp * DATA;
    float *T = s_T + (size_t)warp * twarp;
    constexpr int bps = ALAW ? 1 : 2;

    // ---- per-lane constants
    // twiddles tw[h - 1 + (e mod h)] of this lane's butterflies (float, rounded once)
    auto twf = [&](int i) { const double2 w = a.tw[i]; return make_float2((float)w.x, (float)w.y); };
    float2 tw1[7], tw2[7], tw3[2][R3 - 1];
#pragma unroll
    for (int i = 0; i < 7; ++i) tw1[i] = twf(i);                                 // h = 1, 2, 4: e mod h = r mod h
#pragma unroll
    for (int lv = 0; lv < 3; ++lv)
#pragma unroll
        for (int j = 0; j < (1 << lv); ++j) tw2[(1 << lv) - 1 + j] = twf((8 << lv) - 1 + (lane & 7) + 8 * j);       // h = 8, 16, 32
#pragma unroll
    for (int gq = 0; gq < 2; ++gq)
#pragma unroll
        for (int lv = 0; lv < L3; ++lv)
#pragma unroll
            for (int j = 0; j < (1 << lv); ++j) tw3[gq][(1 << lv) - 1 + j] = twf((64 << lv) - 1 + lane + 32 * gq + 64 * j);   // h = 64 ..
    // Hamming window at this lane's (bit-reversed) sample positions, zero beyond the window
    float ham[OCT][8];
#pragma unroll
    for (int oc = 0; oc < OCT; ++oc) {
        const int ib = (int)(__brev((unsigned)(lane + 32 * oc)) >> (32 - (LOGN - 3)));
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = ib + ((((r & 1) << 2) | (r & 2) | ((r >> 2) & 1)) << (LOGN - 3));
            ham[oc][r] = i < a.vs ? a.hamming[i] : 0.0f;
        }
    }
    // filterbank (dspc.cpp:236-269): bin k with Banks[k] = s gives (1 - c) P to bank s and c P to bank s - 1
    int slot_r[NB], slot_f[NB];
    float wt_r[NB], wt_f[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        const int k = lane + 32 * (t / (R3 / 2)) + 64 * (t % (R3 / 2));
        const int sgm = a.banks[k];
        const float cf = a.coeffs[k];
        slot_r[t] = slot_f[t] = 33 * a.mel_len; wt_r[t] = wt_f[t] = 0.0f;   // bins outside the filterbank: weight 0 into the trash slot
        if (sgm >= 0 && sgm < a.nbanks) { slot_r[t] = (k - a.klo[sgm]) * 33 + sgm; wt_r[t] = 1.0f - cf; }
        if (sgm >= 1 && sgm <= a.nbanks) { slot_f[t] = (k - a.klo[sgm - 1]) * 33 + sgm - 1; wt_f[t] = cf; }
    }
    const int offB = both ? 16 : TSZ;                         // frame B's columns
    const int mbk = both ? (lane & 15) : lane;                // the bank this lane sums up
    const float dc = a.dc_shift, sc = a.scale;

    // work unit = frames (2i, 2i + 1) of ONE utterance (an odd last frame goes alone): which frames share an FFT does not
    // depend on what else is in the batch or on how the batch was cut into launches
    for (int64_t p = a.p_begin + (int64_t)blockIdx.x * kPairWarps + warp; p < a.p_end; p += (int64_t)gridDim.x * kPairWarps) {
        const int u = find_utt_warp(a.pair_off, a.n_utt, p, lane);
        const int64_t t0 = 2 * (p - a.pair_off[u]), T_u = a.frame_off[u + 1] - a.frame_off[u];
        const int64_t fA = a.frame_off[u] + t0, fB = fA + 1;
        const bool haveB = t0 + 1 < T_u;
        // per frame: pointer to its first sample and the number of samples the window may read (0 beyond the signal)
        const uint8_t *src[2];
        int lim[2];
        {
            const int64_t b0 = a.byte_off[u];
            const int64_t len = (a.byte_off[u + 1] - b0) / bps;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                const int64_t s0 = (t0 + (w && haveB ? 1 : 0)) * a.step;
                src[w] = a.audio + b0 + bps * s0;
                const int64_t left = len - s0;
                lim[w] = (w && !haveB) ? 0 : (int)(left < a.vs ? (left < 0 ? 0 : left) : a.vs);
            }
        }
        auto decode = [&](const uint8_t *q) -> float {  // srec.cpp:742-743 / 768-769, then dc shift and scale in the reference's order
            if (ALAW) return __fmul_rn(__fadd_rn(alaw8_float(q[0]), dc), sc);
            return __fmul_rn(__fadd_rn((float)(short)((unsigned)q[0] | ((unsigned)q[1] << 8)), dc), sc);
        };

        // ---- pass 1 (stages h = 1, 2, 4): elements 8o .. 8o+7 = windowed inputs at bit-reversed positions.
        // brev(8o + r) = brev(o) + (brev3(r) << (LOGN - 3)): one lane-dependent base, compile-time offsets.
        float2 v[OCT][8];
        bool nzA = false, nzB = false;
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int ib = (int)(__brev((unsigned)o) >> (32 - (LOGN - 3)));
            const uint8_t *qa = src[0] + bps * ib, *qb = src[1] + bps * ib;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int off = (((r & 1) << 2) | (r & 2) | ((r >> 2) & 1)) << (LOGN - 3);   // brev3(r) << (LOGN-3)
                const float xa = ib + off < lim[0] ? decode(qa + bps * off) : 0.0f;
                const float xb = ib + off < lim[1] ? decode(qb + bps * off) : 0.0f;
                nzA |= xa != 0.0f; nzB |= xb != 0.0f;
                v[oc][r] = make_float2(xa * ham[oc][r], xb * ham[oc][r]);
            }
        }
        const bool liveA = __any_sync(0xffffffffu, nzA), liveB = __any_sync(0xffffffffu, nzB);
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            fft_pass_r<3, true>(v[oc], tw1);
            float2 *d1 = data + 9 * o;                                 // exchange 1: (8o + r) + ((8o + r) >> 3)
#pragma unroll
            for (int r = 0; r < 8; ++r) d1[r] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 2 (h = 8, 16, 32): elements low3 + 8r + 64*high
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
            const float2 *d2 = data + low3 + 72 * high;                // exchange 1: e + (e >> 3), e >> 3 = r + 8 high
#pragma unroll
            for (int r = 0; r < 8; ++r) v[oc][r] = d2[9 * r];
        }
        __syncwarp();
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
            fft_pass_r<3, false>(v[oc], tw2);
            float2 *d2 = data + low3 + 72 * high;                      // exchange 2: e + 8 (e >> 6), e >> 6 = high
#pragma unroll
            for (int r = 0; r < 8; ++r) d2[8 * r] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 3 (h = 64 .. N/2): elements low6 + 64r, two groups per lane; all N bins are needed now
        float2 w3[2][R3];
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const float2 *d3 = data + lane + 32 * gq;                  // exchange 2: e >> 6 = r
#pragma unroll
            for (int r = 0; r < R3; ++r) w3[gq][r] = d3[72 * r];
        }
        __syncwarp();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            fft_pass_r<L3, false>(w3[gq], tw3[gq]);
            float2 *d3 = data + lane + 32 * gq;
#pragma unroll
            for (int r = R3 / 2; r < R3; ++r) d3[72 * r] = w3[gq][r];   // only the upper half is read back (bins N - k)
        }
        __syncwarp();
        // ---- separate the two spectra (bin k = low6 + 64r, r < R3/2, pairs with bin N - k) and scatter the weighted
        // powers into the filterbank table
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const int low6 = lane + 32 * gq;
#pragma unroll
            for (int r = 0; r < R3 / 2; ++r) {
                const int k = low6 + 64 * r, t = gq * (R3 / 2) + r;
                const float2 z = w3[gq][r];
                const int m = N - k;
                const float2 y = k == 0 ? z : data[m + 8 * (m >> 6)];   // k = 0 pairs with itself
                const float sr = z.x + y.x, di = z.y - y.y, dr = z.x - y.x, si = z.y + y.y;
                const float pa = 0.25f * fmaf(sr, sr, di * di);      // cPower (dspc.h:141-146) of frame A
                const float pb = 0.25f * fmaf(dr, dr, si * si);      //                          frame B
                T[slot_r[t]] = wt_r[t] * pa;
                T[slot_f[t]] = wt_f[t] * pa;
                T[slot_r[t] + offB] = wt_r[t] * pb;
                T[slot_f[t] + offB] = wt_f[t] * pb;
            }
        }
        __syncwarp();

        // ---- sum the columns: ln, shift, floor, store
#pragma unroll 1
        for (int pass = 0; pass < (both ? 1 : 2); ++pass) {
            const int w = both ? (lane >> 4) : pass;
            const float *col = T + lane + (both ? 0 : pass * TSZ);
            float acc0 = 0.0f, acc1 = 0.0f;
            int j = 0;
#pragma unroll 4
            for (; j + 1 < a.mel_len; j += 2) {
                acc0 += col[33 * j];
                acc1 += col[33 * j + 33];
            }
            if (j < a.mel_len) acc0 += col[33 * j];
            if (mbk < a.nbanks && (w == 0 || haveB)) {
                float acc = acc0 + acc1;
                if (!(w ? liveB : liveA)) acc = 0.0f;                // digital silence stays exactly silent
                float o = ln_guarded(acc, s_logtab);
                if (a.frame_shift != 0.0f) o += a.frame_shift;                        // srec.cpp:1594-1620
                if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
                a.mel[(w ? fB : fA) * a.nbanks + mbk] = o;
            }
        }
        __syncwarp();
    }
}

template <int LOGN, bool ALAW>
static int launch_wave_pair_k(phn_ctx *c, const WaveArgs &a, size_t smem)
{
    PHN_CUDA(c, cudaFuncSetAttribute(k_wave_pair<LOGN, ALAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    PHN_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave_pair<LOGN, ALAW>, kPairWarps * 32, smem));
    int64_t blocks = (a.p_end - a.p_begin + kPairWarps - 1) / kPairWarps;
    // Grid = a few times what is resident at once: the previous batch's decoder may hold part of the register file when this
    // kernel starts (it runs on its own stream), and a grid of exactly the resident count would leave the CTAs that did not
    // fit waiting with a fixed share of the work.  PHNREC_FRONT_OVERSUB overrides the factor (kernel development).
    static const int oversub = getenv("PHNREC_FRONT_OVERSUB") ? atoi(getenv("PHNREC_FRONT_OVERSUB")) : 4;
    static const int cta_cap = getenv("PHNREC_WAVE_CTAS") ? atoi(getenv("PHNREC_WAVE_CTAS")) : 0;   // (kernel development)
    (void)cta_cap;
    const int64_t cap = (int64_t)c->num_sms * (per_sm > 0 ? per_sm : 1) * (oversub > 0 ? oversub : 1);
    if (blocks > cap) blocks = cap;
    k_wave_pair<LOGN, ALAW><<<(unsigned)blocks, kPairWarps * 32, smem, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

template <int LOGN>
static int launch_wave_pair_t(phn_ctx *c, WaveArgs a)
{
    constexpr int N = 1 << LOGN;
    a.mel_len = 1;
    for (int b = 0; b < c->mt.nbanks; ++b) a.mel_len = std::max(a.mel_len, c->mt.bank_khi[b] - c->mt.bank_klo[b] + 1);
    const size_t twarp = (size_t)pair_tsz(a.mel_len) * (c->mt.nbanks <= 16 ? 1 : 2);
    const size_t smem = sizeof(double) * 32 + (size_t)kPairWarps * (sizeof(float2) * (N + N / 8) + sizeof(float) * twarp);
    return a.fmt == PHN_WAVE_ALAW ? launch_wave_pair_k<LOGN, true>(c, a, smem) : launch_wave_pair_k<LOGN, false>(c, a, smem);
}

template <bool EXACT, int LOGN>
static int launch_wave_t(phn_ctx *c, const WaveArgs &a)
{
    constexpr int N = 1 << LOGN, N2 = N / 2;
    const size_t smem = sizeof(double2) * N + sizeof(float) * (N + N2) + sizeof(int) * N2 + sizeof(double) * 32 +
                        sizeof(float2) * (size_t)kWaveWarps * (N + N / 32 + N2 / 2);
    PHN_CUDA(c, cudaFuncSetAttribute(k_wave<EXACT, LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool pair = !EXACT && !a.z_mean && a.preem == 0.0f;
    if (pair) return launch_wave_pair_t<LOGN>(c, a);
    int64_t blocks = (a.f_end - a.f_begin + kWaveWarps - 1) / kWaveWarps;
    int per_sm = 1;   // persistent grid: exactly the blocks that are resident at once (a larger grid runs its tail at low occupancy)
    PHN_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wave<EXACT, LOGN>, kWaveWarps * 32, smem));
    const int64_t cap = (int64_t)c->num_sms * (per_sm > 0 ? per_sm : 1);
    if (blocks > cap) blocks = cap;
    k_wave<EXACT, LOGN><<<(unsigned)blocks, kWaveWarps * 32, smem, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

int launch_wave(phn_ctx *c, const void *d_audio, int u0, int u1)
{
    if (u1 < 0) u1 = c->n_utt;
    if (u1 <= u0) return PHN_OK;
    const int64_t f_begin = c->h_frame_off[u0], f_end = c->h_frame_off[u1];
    if (f_end <= f_begin) return PHN_OK;
    WaveArgs a;
    a.audio = (const uint8_t *)d_audio;
    a.byte_off = (const int64_t *)c->d_byte_off.p;
    a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt;
    a.total_frames = c->total_frames;
    a.f_begin = f_begin; a.f_end = f_end;
    a.pair_off = (const int64_t *)c->d_pair_off.p;
    a.p_begin = c->h_pair_off[u0]; a.p_end = c->h_pair_off[u1];
    a.fmt = c->fmt; a.vs = c->vs; a.step = c->step; a.N = c->mt.N; a.logN = c->mt.logN; a.nbanks = c->nbanks;
    a.scale = c->scale; a.dc_shift = c->dc_shift; a.frame_shift = c->frame_shift; a.frame_floor = c->frame_floor;
    a.preem = c->preem; a.z_mean = c->z_mean;
    a.raw_energy = c->plp;
    a.hamming = c->tab.hamming; a.coeffs = c->tab.coeffs; a.banks = c->tab.banks;
    a.klo = c->tab.bank_klo; a.khi = c->tab.bank_khi; a.tw = c->tab.tw;
    a.mel = (float *)c->d_mel.p;
    // The exact instantiation always serves the stage-wise API (phn_mel: what `-t par` saves must be
    // the reference's bits); the fused tensor-core pipeline takes the fp32 one.
    const bool exact = c->mlp_mode != PHN_MLP_TC_F16 || c->force_exact_wave;
    int rc;
    if (!exact && !c->wave_tc && (rc = wave_tc_prepare(c))) return rc;   // (first call in this mode: the DFT matrix image)
    if (!exact && !c->wave_tc16 && (rc = wave_tc16_prepare(c))) return rc;
    if (!exact && wave_tc16_applies(c)) {   // ... of the 16 kHz systems (k_wave_tc16.cu)
        if ((rc = launch_wave_tc16(c, d_audio, f_begin, f_end))) return rc;
        c->k_launches[PHN_K_WAVE] += 1;
        return PHN_OK;
    }
    if (!exact && wave_tc_applies(c)) {   // the windowed DFT on the tensor cores (k_wave_tc.cu)
        if ((rc = launch_wave_tc(c, d_audio, f_begin, f_end))) return rc;
        c->k_launches[PHN_K_WAVE] += 1;
        return PHN_OK;
    }
    switch (c->mt.logN) {
        case 8: rc = exact ? launch_wave_t<true, 8>(c, a) : launch_wave_t<false, 8>(c, a); break;
        case 9: rc = exact ? launch_wave_t<true, 9>(c, a) : launch_wave_t<false, 9>(c, a); break;
        case 10: rc = exact ? launch_wave_t<true, 10>(c, a) : launch_wave_t<false, 10>(c, a); break;
        default: return fail(c, PHN_ERR_UNSUPPORTED, "FFT size %d not instantiated (vector_size must be 129..1024)\n", c->mt.N);
    }
    if (rc) return rc;
    c->k_launches[PHN_K_WAVE] += 1;
    return PHN_OK;
}

}  // namespace phn
