// k_wave.cu — K-wave: waveform decode + framing + Hamming + FFT + power + mel filterbank + guarded ln.
//
// Replaces ConvertWaveformFormat (srec.cpp:709-791), MelBanks::GetFeatures/ProcessFrame
// (melbanks.cpp:111-204), cFour1 / _mbApply (dspc.cpp:24-78, 236-269), cPower / sLn (dspc.h:141-160)
// and FrameBasedNormalization (srec.cpp:1594-1620) for a ragged batch of utterances.
//
// One warp owns one frame at a time; the frame lives in that warp's shared-memory slice
// (re[N], im[N]).  The FFT is the reference's radix-2 decimation-in-time with the reference's
// double-precision twiddles (built on the host by the same recurrence) and the reference's
// rounding points, so mel values are bit-identical to the CPU implementation.
#include "internal.h"
#include "device_math.cuh"

namespace phn {

struct WaveArgs {
    const uint8_t *audio;
    const int64_t *byte_off, *frame_off;
    int n_utt;
    int64_t total_frames;
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

__global__ void __launch_bounds__(kWaveWarps * 32) k_wave(WaveArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N, N2 = N / 2;
    double2 *s_tw = reinterpret_cast<double2 *>(smem_raw);          // [N-1] (+1 pad)
    float *s_ham = reinterpret_cast<float *>(s_tw + N);             // [vs rounded to N]
    float *s_coef = s_ham + N;                                      // [N2]
    int *s_bank = reinterpret_cast<int *>(s_coef + N2);             // [N2]
    double *s_logtab = reinterpret_cast<double *>(s_bank + N2);     // [32] glibc logf table
    float *s_work = reinterpret_cast<float *>(s_logtab + 32);       // per warp: re[N] im[N] pw[N2]
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);

    for (int i = threadIdx.x; i < N - 1; i += blockDim.x) s_tw[i] = a.tw[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_ham[i] = i < a.vs ? a.hamming[i] : 0.0f;
    for (int i = threadIdx.x; i < N2; i += blockDim.x) { s_coef[i] = a.coeffs[i]; s_bank[i] = a.banks[i]; }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *re = s_work + (size_t)warp * (2 * N + N2);
    float *im = re + N;
    float *pw = im + N;
    const int bps = a.fmt == PHN_WAVE_LIN16 ? 2 : 1;

    for (int64_t f = (int64_t)blockIdx.x * kWaveWarps + warp; f < a.total_frames; f += (int64_t)gridDim.x * kWaveWarps) {
        const int u = find_utt(a.frame_off, a.n_utt, f);
        const int64_t t = f - a.frame_off[u];
        const int64_t b0 = a.byte_off[u];
        const int64_t len = (a.byte_off[u + 1] - b0) / bps;  // samples in this utterance
        const int64_t s0 = t * a.step;

        // ---- decode (srec.cpp:742-743 / 768-769), dc shift, scale; zero beyond the signal
        for (int i = lane; i < N; i += 32) {
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
            if (a.z_mean || a.preem != 0.0f) im[i] = x;  // staging for the optional per-frame ops
            else re[__brev((unsigned)i) >> (32 - a.logN)] = i < a.vs ? __fmul_rn(x, s_ham[i]) : 0.0f;
        }
        if (a.z_mean || a.preem != 0.0f) {
            // optional sSubtractAverage / sPreemphasis (dspc.h:63-84); not used by the shipped systems
            __syncwarp();
            float avg = 0.0f;
            if (a.z_mean) {
                if (lane == 0) {
                    float s = 0.0f;
                    for (int i = 0; i < a.vs; ++i) s = __fadd_rn(s, im[i]);
                    pw[0] = __fdiv_rn(s, (float)a.vs);
                }
                __syncwarp();
                avg = pw[0];
                __syncwarp();
            }
            for (int i = lane; i < N; i += 32) {
                float x = 0.0f;
                if (i < a.vs) {
                    x = a.z_mean ? __fsub_rn(im[i], avg) : im[i];
                    if (a.preem != 0.0f) {
                        if (i == 0) x = __fmul_rn(x, __fsub_rn(1.0f, a.preem));
                        else {
                            const float xp = a.z_mean ? __fsub_rn(im[i - 1], avg) : im[i - 1];
                            x = __fsub_rn(x, __fmul_rn(a.preem, xp));
                        }
                    }
                    x = __fmul_rn(x, s_ham[i]);
                }
                re[__brev((unsigned)i) >> (32 - a.logN)] = x;
            }
            __syncwarp();
        }
        for (int i = lane; i < N; i += 32) im[i] = 0.0f;
        __syncwarp();

        // ---- radix-2 DIT butterflies, stage half-size h = 1 .. N/2 (dspc.cpp:55-76)
        for (int h = 1; h < N; h <<= 1) {
            for (int j = lane; j < N2; j += 32) {
                const int m = j & (h - 1);
                const int i0 = ((j - m) << 1) + m;
                const int k0 = i0 + h;
                const double2 w = s_tw[h - 1 + m];
                const double kr = (double)re[k0], ki = (double)im[k0];
                const float tr = __double2float_rn(__dsub_rn(__dmul_rn(w.x, kr), __dmul_rn(w.y, ki)));
                const float ti = __double2float_rn(__dadd_rn(__dmul_rn(w.x, ki), __dmul_rn(w.y, kr)));
                const float ir = re[i0], ii = im[i0];
                re[k0] = __fsub_rn(ir, tr);
                im[k0] = __fsub_rn(ii, ti);
                re[i0] = __fadd_rn(ir, tr);
                im[i0] = __fadd_rn(ii, ti);
            }
            __syncwarp();
        }

        // ---- power spectrum (dspc.h:141-146), bins 0 .. N/2-1
        for (int k = lane; k < N2; k += 32) pw[k] = __fadd_rn(__fmul_rn(re[k], re[k]), __fmul_rn(im[k], im[k]));
        __syncwarp();

        // ---- mel filterbank (dspc.cpp:236-269): lane b accumulates bank b in ascending bin order
        if (lane < a.nbanks) {
            float acc = 0.0f;
            const int hi = a.khi[lane];
            for (int k = a.klo[lane]; k <= hi; ++k) {
                const float p = pw[k];
                const float v = __fmul_rn(s_coef[k], p);
                acc = __fadd_rn(acc, s_bank[k] == lane ? __fsub_rn(p, v) : v);
            }
            float o = ln_guarded(acc, s_logtab);
            if (a.frame_shift != 0.0f) o = __fadd_rn(o, a.frame_shift);           // srec.cpp:1594-1620
            if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
            a.mel[f * a.nbanks + lane] = o;
        }
        __syncwarp();
    }
}

int launch_wave(phn_ctx *c, const void *d_audio)
{
    if (c->total_frames == 0) return PHN_OK;
    WaveArgs a;
    a.audio = (const uint8_t *)d_audio;
    a.byte_off = (const int64_t *)c->d_byte_off.p;
    a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt;
    a.total_frames = c->total_frames;
    a.fmt = c->fmt; a.vs = c->vs; a.step = c->step; a.N = c->mt.N; a.logN = c->mt.logN; a.nbanks = c->nbanks;
    a.scale = c->scale; a.dc_shift = c->dc_shift; a.frame_shift = c->frame_shift; a.frame_floor = c->frame_floor;
    a.preem = c->preem; a.z_mean = c->z_mean;
    a.hamming = c->tab.hamming; a.coeffs = c->tab.coeffs; a.banks = c->tab.banks;
    a.klo = c->tab.bank_klo; a.khi = c->tab.bank_khi; a.tw = c->tab.tw;
    a.mel = (float *)c->d_mel.p;
    const int N = a.N, N2 = N / 2;
    const size_t smem = sizeof(double2) * N + sizeof(float) * (N + N2) + sizeof(int) * N2 + sizeof(double) * 32 +
                        sizeof(float) * (size_t)kWaveWarps * (2 * N + N2);
    PHN_CUDA(c, cudaFuncSetAttribute(k_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (c->total_frames + kWaveWarps - 1) / kWaveWarps;
    const int64_t cap = (int64_t)c->num_sms * 8;
    if (blocks > cap) blocks = cap;
    k_wave<<<(unsigned)blocks, kWaveWarps * 32, smem, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_WAVE] += 1;
    return PHN_OK;
}

}  // namespace phn
