// k_synth.cu — deterministic synthetic speech-like audio generated on the device (bench / tests).
// Not part of the reference: SURVEY §8(d) asks for synthetic audio of the named rate and duration.
// Utterance u: three formant-like sinusoids whose frequencies are redrawn every 100 ms, white
// noise at about -20 dB, a 4 Hz syllabic envelope, and one 0.5-1.5 s stretch of digital silence
// (exercises the sLn zero guard and long pause segments).  Output: int16 PCM or G.711 A-law bytes.
#include "internal.h"

namespace phn {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

__device__ __forceinline__ unsigned char lin2alaw(int pcm16)
{
    int pcm = pcm16 >> 3;
    int mask;
    if (pcm >= 0) mask = 0xD5; else { mask = 0x55; pcm = -pcm - 1; }
    int seg = 0;
    while (seg < 8 && pcm > ((0x20 << seg) - 1)) ++seg;
    if (seg >= 8) return (unsigned char)(0x7F ^ mask);
    int aval = seg << 4;
    aval |= seg < 2 ? (pcm >> 1) & 0xF : (pcm >> seg) & 0xF;
    return (unsigned char)(aval ^ mask);
}

__global__ void k_synth(unsigned char *out, int64_t bytes_per_utt, int n_utt, int fmt, int fs, uint64_t seed)
{
    const int bps = fmt == PHN_WAVE_LIN16 ? 2 : 1;
    const int64_t n_per = bytes_per_utt / bps;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_per * n_utt) return;
    const int u = (int)(idx / n_per);
    const int64_t n = idx - (int64_t)u * n_per;
    const uint64_t us = mix64(seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(u + 1)));
    const float t = (float)n / (float)fs;
    // silence stretch
    const float dur = 0.5f + u01(mix64(us ^ 1));
    const float total = (float)n_per / (float)fs;
    const float s0 = u01(mix64(us ^ 2)) * fmaxf(total - dur, 0.0f);
    float x = 0.0f;
    if (!(t >= s0 && t < s0 + dur)) {
        const int seg = (int)(n / (fs / 10));
        const uint64_t hs = mix64(us ^ (0x1000ull + (uint64_t)seg));
        const float f1 = 200.0f + 700.0f * u01(hs);
        const float f2 = 900.0f + 1500.0f * u01(mix64(hs ^ 11));
        const float f3 = 2400.0f + 1000.0f * u01(mix64(hs ^ 23));
        const float env = 0.5f * (1.0f - cospif(2.0f * 4.0f * t + u01(mix64(us ^ 3))));
        const float voiced = sinpif(2.0f * f1 * t) + 0.6f * sinpif(2.0f * f2 * t) + 0.3f * sinpif(2.0f * f3 * t);
        const float noise = 2.0f * u01(mix64(us ^ (0x5000000ull + (uint64_t)n))) - 1.0f;
        x = 3500.0f * env * voiced + 400.0f * noise;
    }
    int v = __float2int_rn(x);
    v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
    if (fmt == PHN_WAVE_LIN16) {
        out[2 * idx] = (unsigned char)(v & 0xff);
        out[2 * idx + 1] = (unsigned char)((v >> 8) & 0xff);
    } else {
        out[idx] = lin2alaw(v);
    }
}

int launch_synth(phn_ctx *c, void *d_audio, int64_t bytes_per_utt, int n_utt, uint64_t seed)
{
    const int bps = c->fmt == PHN_WAVE_LIN16 ? 2 : 1;
    const int64_t total = bytes_per_utt / bps * n_utt;
    if (total == 0) return PHN_OK;
    k_synth<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>((unsigned char *)d_audio, bytes_per_utt, n_utt, c->fmt,
                                                                    c->fs, seed);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

}  // namespace phn
