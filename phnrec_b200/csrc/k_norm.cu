// k_norm.cu — K-mean and K-norm.
//
// K-mean: per-utterance, per-band mean of the log mel energies, SentenceBasedNormalization with
// sent_mean_norm=true (srec.cpp:1492-1516; Mat::sumColumns matrix.h:2101-2116): a sequential fp32
// sum over the utterance's frames times 1.0f/T.  One thread per (utterance, band) keeps the
// reference's summation order, so the mean is bit-identical; the subtraction itself is fused into
// K-stc's loads.
//
// K-norm: the online normaliser's arithmetic (Normalization::ProcessFrame norm.cpp:216-234,
// ChannelNormParams::{Accum,Update,Norm} norm.cpp:92-148).  Not on the offline path (only
// ProcessOnline calls it, srec.cpp:806); provided for completeness of SURVEY §8(a) row N2.
#include "internal.h"

namespace phn {

__global__ void k_sentence_mean(const float *__restrict__ mel, const int64_t *__restrict__ frame_off, int n_utt, int nb,
                                float *__restrict__ mean)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_utt * nb) return;
    const int u = idx / nb, b = idx - u * nb;
    const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0;
    const float *p = mel + f0 * nb + b;
    float s = 0.0f;
    int64_t t = 0;
    for (; t + 8 <= T; t += 8) {  // loads are independent of the running sum: issue 8, then add in order
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = p[(t + k) * nb];
#pragma unroll
        for (int k = 0; k < 8; ++k) s = __fadd_rn(s, v[k]);
    }
    for (; t < T; ++t) s = __fadd_rn(s, p[t * nb]);
    mean[idx] = T > 0 ? __fmul_rn(s, __fdiv_rn(1.0f, (float)T)) : 0.0f;  // Mat::div(v) == mul(1/v), matrix.h:245
}

// Tensor-core pipeline: the same mean with a fixed-shape parallel sum (deterministic and batch-invariant, but not the
// reference's sequential order): one CTA per utterance; thread t owns band t % nb and every (240 / nb)-th frame, so the
// CTA reads the utterance's [T][nb] block as one coalesced stream; partial sums meet in shared memory in a fixed order.
__global__ void __launch_bounds__(256) k_sentence_mean_fast(const float *__restrict__ mel, const int64_t *__restrict__ frame_off, int nb,
                                                            float *__restrict__ mean)
{
    __shared__ float s_part[256];
    const int u = blockIdx.x;
    const int64_t f0 = frame_off[u], T = frame_off[u + 1] - f0;
    const int rows = 256 / nb, used = rows * nb;          // frames in flight per pass, active threads
    const int b = threadIdx.x % nb, r = threadIdx.x / nb;
    float s = 0.0f;
    if (threadIdx.x < used) {
        const float *p = mel + f0 * nb;
        for (int64_t t = r; t < T; t += rows) s += p[t * nb + b];
    }
    s_part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < nb) {
        float tot = 0.0f;
        for (int k = 0; k < rows; ++k) tot += s_part[k * nb + threadIdx.x];
        mean[(size_t)u * nb + threadIdx.x] = T > 0 ? tot * (1.0f / (float)T) : 0.0f;
    }
}

int launch_sentence_mean(phn_ctx *c, int u0, int u1)
{
    if (u1 < 0) u1 = c->n_utt;
    const int nu = u1 - u0, n = nu * c->nbanks;
    if (n <= 0) return PHN_OK;
    float *mean = (float *)c->d_mean.p + (size_t)u0 * c->nbanks;
    const int64_t *foff = (const int64_t *)c->d_frame_off.p + u0;
    if (!c->sent_mean_norm) {  // EN system: nothing is subtracted (x - 0.0f == x)
        PHN_CUDA(c, cudaMemsetAsync(mean, 0, sizeof(float) * n, c->stream));
        return PHN_OK;
    }
    if (c->mlp_mode == PHN_MLP_TC_F16 && !c->force_exact_wave && c->fast_front)
        k_sentence_mean_fast<<<nu, 256, 0, c->stream>>>((const float *)c->d_mel.p, foff, c->nbanks, mean);
    else
    k_sentence_mean<<<(n + 127) / 128, 128, 0, c->stream>>>((const float *)c->d_mel.p, foff, nu, c->nbanks, mean);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_MEAN] += 1;
    return PHN_OK;
}

// One thread per band, frames in order: the estimator's sums are the reference's sequential fp32 sums.
// Normalization::ProcessFrame (norm.cpp:216-234): Accum for the first `interval` frames, Update() when the count reaches
// `interval` - before Norm(), so frame interval-1 is the first one normalised - then `if(mean) x -= mean; if(var) x *= inv`
// (ChannelNormParams::Norm, norm.cpp:112-137).  Earlier frames see Null()'s mean 0 / inverse std 1: unchanged.
__global__ void k_online_norm(float *x, int64_t T, int nb, int interval, int mean_norm, int var_norm)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    float s = 0.0f, s2 = 0.0f;
    for (int t = 0; t < interval; ++t) {
        const float v = x[(int64_t)t * nb + b];
        s = __fadd_rn(s, v);
        s2 = __fadd_rn(s2, __fmul_rn(v, v));
    }
    const float mean = __fdiv_rn(s, (float)interval);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fsub_rn(__fdiv_rn(s2, (float)interval), __fmul_rn(mean, mean))));
    for (int64_t t = interval - 1; t < T; ++t) {
        float v = x[t * nb + b];
        if (mean_norm) v = __fsub_rn(v, mean);
        if (var_norm) v = __fmul_rn(v, inv);
        x[t * nb + b] = v;
    }
}

int launch_online_norm(phn_ctx *c, float *d_x, int64_t frames, int nb, int interval, int mean_norm, int var_norm)
{
    // ChannelNormParams::SetNorm (norm.cpp:150-155) asserts that variance normalisation comes with mean normalisation
    if (var_norm && !mean_norm) return fail(c, PHN_ERR_ARG, "online normalisation: var_norm without mean_norm (the reference asserts, norm.cpp:152)\n");
    if (interval <= 0 || frames < interval) return PHN_OK;
    k_online_norm<<<(nb + 31) / 32, 32, 0, c->stream>>>(d_x, frames, nb, interval, mean_norm, var_norm);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

}  // namespace phn
