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
#include "device_math.cuh"

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
            float o = ln_guarded(acc, s_logtab);
            if (a.frame_shift != 0.0f) o = __fadd_rn(o, a.frame_shift);           // srec.cpp:1594-1620
            if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
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
template <int LOGN, bool ALAW>
__global__ void __launch_bounds__(kWaveWarps * 32) k_wave_pair(WaveArgs a)
{
    __shared__ float s_lut[256];   // A-law byte -> sample, dc shift and scale applied in the reference's order (srec.cpp:768-788)
    constexpr int N = 1 << LOGN, N2 = N / 2;
    constexpr int OCT = N / 256;
    constexpr int L3 = LOGN - 6;
    constexpr int R3 = 1 << L3;
    constexpr int WORK = N + N / 32 + N2 / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using TW = float2;
    TW *s_tw = reinterpret_cast<TW *>(smem_raw);
    float *s_ham = reinterpret_cast<float *>(smem_raw + sizeof(double2) * N);
    float *s_coef = s_ham + N;
    int *s_bank = reinterpret_cast<int *>(s_coef + N2);
    double *s_logtab = reinterpret_cast<double *>(s_bank + N2);
    float2 *s_work = reinterpret_cast<float2 *>(s_logtab + 32);
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < N - 1; i += blockDim.x) {
        const double2 w = a.tw[i];
        s_tw[i] = make_float2((float)w.x, (float)w.y);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_ham[i] = i < a.vs ? a.hamming[i] : 0.0f;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) { s_coef[i] = a.coeffs[i]; s_bank[i] = a.banks[i]; }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = __fmul_rn(__fadd_rn(__fmul_rn(8.0f, (float)alaw_d5((unsigned)i)), a.dc_shift), a.scale);
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 *data = s_work + (size_t)warp * WORK;
    float *pwA = reinterpret_cast<float *>(data + N + N / 32);   // [N2] power spectrum of the first frame
    float *pwB = reinterpret_cast<float *>(data);                // [N2] second frame (aliases data[], free by then)
    constexpr int bps = ALAW ? 1 : 2;
    // work unit = frames (2i, 2i + 1) of ONE utterance (an odd last frame goes alone): which frames share an FFT does not
    // depend on what else is in the batch or on how the batch was cut into launches
    for (int64_t p = a.p_begin + (int64_t)blockIdx.x * kWaveWarps + warp; p < a.p_end; p += (int64_t)gridDim.x * kWaveWarps) {
        const int u = find_utt(a.pair_off, a.n_utt, p);
        const int64_t t0 = 2 * (p - a.pair_off[u]), T = a.frame_off[u + 1] - a.frame_off[u];
        const int64_t fA = a.frame_off[u] + t0, fB = fA + 1;
        const bool haveB = t0 + 1 < T;
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
        const float dc = a.dc_shift, sc = a.scale;
        auto decode = [&](const uint8_t *q) -> float {  // srec.cpp:742-743 / 768-769, dc shift, scale ((x + 0) * 1 == x)
            if (ALAW) return s_lut[q[0]];
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
            const float *hm = s_ham + ib;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                constexpr int dummy = 0; (void)dummy;
                const int off = (((r & 1) << 2) | (r & 2) | ((r >> 2) & 1)) << (LOGN - 3);   // brev3(r) << (LOGN-3)
                const float xa = ib + off < lim[0] ? decode(qa + bps * off) : 0.0f;
                const float xb = ib + off < lim[1] ? decode(qb + bps * off) : 0.0f;
                nzA |= xa != 0.0f; nzB |= xb != 0.0f;
                v[oc][r] = make_float2(xa * hm[off], xb * hm[off]);
            }
        }
        const bool liveA = __any_sync(0xffffffffu, nzA), liveB = __any_sync(0xffffffffu, nzB);
        // padded index of element e is e + (e >> 5); for the three access patterns the pad term is a lane-dependent
        // base plus a compile-time offset
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            fft_pass<false, 3, TW>(v[oc], s_tw, 1, 0);
            float2 *d1 = data + 8 * o + (o >> 2);                      // (8o + r) >> 5 == o >> 2
#pragma unroll
            for (int r = 0; r < 8; ++r) d1[r] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 2 (h = 8, 16, 32): elements low3 + 8r + 64*high
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
            const float2 *d2 = data + low3 + 66 * high;              // (low3 + 8r + 64 high) >> 5 == 2 high + (r >> 2)
#pragma unroll
            for (int r = 0; r < 8; ++r) v[oc][r] = d2[8 * r + (r >> 2)];
        }
        __syncwarp();
#pragma unroll
        for (int oc = 0; oc < OCT; ++oc) {
            const int o = lane + 32 * oc;
            const int low3 = o & 7, high = o >> 3;
            fft_pass<false, 3, TW>(v[oc], s_tw, 8, low3);
            float2 *d2 = data + low3 + 66 * high;
#pragma unroll
            for (int r = 0; r < 8; ++r) d2[8 * r + (r >> 2)] = v[oc][r];
        }
        __syncwarp();
        // ---- pass 3 (h = 64 .. N/2): elements low6 + 64r, two groups per lane; all N bins are needed now
        float2 w3[2][R3];
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const float2 *d3 = data + lane + 33 * gq;                  // (low6 + 64 r) >> 5 == 2 r + gq
#pragma unroll
            for (int r = 0; r < R3; ++r) w3[gq][r] = d3[66 * r];
        }
        __syncwarp();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const int low6 = lane + 32 * gq;
            fft_pass<false, L3, TW>(w3[gq], s_tw, 64, low6);
            float2 *d3 = data + lane + 33 * gq;
#pragma unroll
            for (int r = 0; r < R3; ++r) d3[66 * r] = w3[gq][r];
        }
        __syncwarp();
        // ---- separate the two spectra: bin k = low6 + 64r (r < R3/2) pairs with bin N - k
        float pa[2][R3 / 2], pb[2][R3 / 2];
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
            const int low6 = lane + 32 * gq;
#pragma unroll
            for (int r = 0; r < R3 / 2; ++r) {
                const int k = low6 + 64 * r;
                const float2 z = w3[gq][r];
                const float2 y = data[pad_idx((N - k) & (N - 1))];   // k = 0 pairs with itself
                const float sr = z.x + y.x, di = z.y - y.y, dr = z.x - y.x, si = z.y + y.y;
                pa[gq][r] = 0.25f * fmaf(sr, sr, di * di);           // cPower (dspc.h:141-146) of frame A
                pb[gq][r] = 0.25f * fmaf(dr, dr, si * si);           //                          frame B
            }
        }
        __syncwarp();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq)
#pragma unroll
            for (int r = 0; r < R3 / 2; ++r) {
                const int k = lane + 32 * gq + 64 * r;
                pwA[k] = pa[gq][r];
                pwB[k] = pb[gq][r];
            }
        __syncwarp();

        // ---- mel filterbank (dspc.cpp:236-269): bank b of frame A on lane b, of frame B on lane 16 + b when both fit
        // a warp, else one frame after the other
        const bool both = a.nbanks <= 16;
#pragma unroll 1
        for (int pass = 0; pass < (both ? 1 : 2); ++pass) {
            const int w = both ? (lane >> 4) : pass;
            const int bk = both ? (lane & 15) : lane;
            if (bk < a.nbanks && (w == 0 || haveB)) {
                const float *pw = w ? pwB : pwA;
                float acc = 0.0f;
                const int hi = a.khi[bk];
                for (int k = a.klo[bk]; k <= hi; ++k) {
                    const float pk = pw[k];
                    const float v2 = s_coef[k] * pk;
                    acc += s_bank[k] == bk ? pk - v2 : v2;
                }
                if (!(w ? liveB : liveA)) acc = 0.0f;                // digital silence stays exactly silent
                float o = ln_guarded(acc, s_logtab);
                if (a.frame_shift != 0.0f) o += a.frame_shift;                        // srec.cpp:1594-1620
                if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
                a.mel[(w ? fB : fA) * a.nbanks + bk] = o;
            }
        }
        __syncwarp();
    }
}

template <bool EXACT, int LOGN>
static int launch_wave_t(phn_ctx *c, const WaveArgs &a)
{
    constexpr int N = 1 << LOGN, N2 = N / 2;
    const size_t smem = sizeof(double2) * N + sizeof(float) * (N + N2) + sizeof(int) * N2 + sizeof(double) * 32 +
                        sizeof(float2) * (size_t)kWaveWarps * (N + N / 32 + N2 / 2);
    PHN_CUDA(c, cudaFuncSetAttribute(k_wave<EXACT, LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool pair = !EXACT && !a.z_mean && a.preem == 0.0f;
    const int64_t units = pair ? a.p_end - a.p_begin : a.f_end - a.f_begin;
    int64_t blocks = (units + kWaveWarps - 1) / kWaveWarps;
    const int64_t cap = (int64_t)c->num_sms * 6;
    if (blocks > cap) blocks = cap;
    if (pair && a.fmt == PHN_WAVE_ALAW) {
        PHN_CUDA(c, cudaFuncSetAttribute(k_wave_pair<LOGN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_wave_pair<LOGN, true><<<(unsigned)blocks, kWaveWarps * 32, smem, c->stream>>>(a);
    } else if (pair) {
        PHN_CUDA(c, cudaFuncSetAttribute(k_wave_pair<LOGN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_wave_pair<LOGN, false><<<(unsigned)blocks, kWaveWarps * 32, smem, c->stream>>>(a);
    } else
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
    a.hamming = c->tab.hamming; a.coeffs = c->tab.coeffs; a.banks = c->tab.banks;
    a.klo = c->tab.bank_klo; a.khi = c->tab.bank_khi; a.tw = c->tab.tw;
    a.mel = (float *)c->d_mel.p;
    // The exact instantiation always serves the stage-wise API (phn_mel: what `-t par` saves must be
    // the reference's bits); the fused tensor-core pipeline takes the fp32 one.
    const bool exact = c->mlp_mode != PHN_MLP_TC_F16 || c->force_exact_wave;
    int rc;
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
