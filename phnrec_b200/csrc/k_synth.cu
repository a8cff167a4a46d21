// k_synth.cu — deterministic synthetic speech-like audio generated on the device (bench / tests).
// Not part of the reference: SURVEY §8(d) asks for synthetic audio of the named rate and duration.
// Utterance u: three formant-like sinusoids whose frequencies are redrawn every 100 ms, white
// noise at about -20 dB, a 4 Hz syllabic envelope, and one 0.5-1.5 s stretch of digital silence
// (exercises the sLn zero guard and long pause segments).  Output: int16 PCM or G.711 A-law bytes.
// Every operation is a single correctly rounded fp32 / integer operation (this file is compiled without FMA contraction,
// the sinusoids are a fixed parabola pair, no libm), so tools/synth_host.py reproduces the bytes exactly on the host: the
// CPU reference arm of bench.py is fed the SAME utterances the GPU arm recognises.
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

// sin(2 pi y) for y >= 0 from the fractional part of y: z = 2 frac(y) - 1 in [-1, 1), sin(2 pi frac) = -sin(pi z),
// sin(pi z) ~ p (0.775 + 0.225 |p|), p = 4 z (1 - |z|)   (max error 1e-3: irrelevant for synthetic audio)
__device__ __forceinline__ float sin_turns(float y)
{
    const float fr = __fsub_rn(y, floorf(y));
    const float z = __fsub_rn(__fmul_rn(2.0f, fr), 1.0f);
    const float p = __fmul_rn(__fmul_rn(4.0f, z), __fsub_rn(1.0f, fabsf(z)));
    return -__fmul_rn(p, __fadd_rn(0.775f, __fmul_rn(0.225f, fabsf(p))));
}

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
    const float t = __fdiv_rn((float)n, (float)fs);
    // silence stretch
    const float dur = __fadd_rn(0.5f, u01(mix64(us ^ 1)));
    const float total = __fdiv_rn((float)n_per, (float)fs);
    const float s0 = __fmul_rn(u01(mix64(us ^ 2)), fmaxf(__fsub_rn(total, dur), 0.0f));
    float x = 0.0f;
    if (!(t >= s0 && t < __fadd_rn(s0, dur))) {
        const int seg = (int)(n / (fs / 10));
        const uint64_t hs = mix64(us ^ (0x1000ull + (uint64_t)seg));
        const float f1 = __fadd_rn(200.0f, __fmul_rn(700.0f, u01(hs)));
        const float f2 = __fadd_rn(900.0f, __fmul_rn(1500.0f, u01(mix64(hs ^ 11))));
        const float f3 = __fadd_rn(2400.0f, __fmul_rn(1000.0f, u01(mix64(hs ^ 23))));
        // 4 Hz syllabic envelope 0.5 (1 - cos(2 pi (4 t + phase))) = 0.5 (1 - sin(2 pi (4 t + phase + 1/4)))
        const float env = __fmul_rn(0.5f, __fsub_rn(1.0f, sin_turns(__fadd_rn(__fadd_rn(__fmul_rn(4.0f, t), u01(mix64(us ^ 3))), 0.25f))));
        const float voiced = __fadd_rn(__fadd_rn(sin_turns(__fmul_rn(f1, t)), __fmul_rn(0.6f, sin_turns(__fmul_rn(f2, t)))),
                                       __fmul_rn(0.3f, sin_turns(__fmul_rn(f3, t))));
        const float noise = __fsub_rn(__fmul_rn(2.0f, u01(mix64(us ^ (0x5000000ull + (uint64_t)n)))), 1.0f);
        x = __fadd_rn(__fmul_rn(__fmul_rn(3500.0f, env), voiced), __fmul_rn(400.0f, noise));
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
