// k_stc.cu — K-stc: split temporal context features + NN input normalisation.
//
// Replaces the 31-frame FIFO (Traps::AddVectorToBEMatrix traps.cpp:180-219 driven by
// srec.cpp:1035-1059), Traps::CalcInputFeaturesForBandNets case stlcrc (traps.cpp:285-342),
// CalcC0 / sDCT (dspc.h:206-233) and NeuralNet::Normalize (nn.cpp:702-716) of the two band nets.
// Net effect of the FIFO + warm-up + tail replication: output row r sees mel frames
// clamp(r-15 .. r+15, 0, T-1) of its own utterance, minus the sentence mean.
//
// One thread per (frame, side, band): 16 windowed samples in registers, 11 outputs
// (C0 + 10 DCT coefficients), sums in the reference's order with separately rounded mul/add.
#include "internal.h"

namespace phn {

struct StcArgs {
    const float *mel, *mean;
    const int64_t *frame_off;
    int n_utt, nb, ncoef;
    int64_t f0, nf;   // frame chunk [f0, f0+nf)
    const float *win, *dct;
    const float *nmean0, *ndev0, *nmean1, *ndev1;
    float normc;
    float *x0, *x1;   // fp32 outputs [nf][ld32]  (exact mode)   or nullptr
    uint8_t *x0h, *x1h; // fp16 outputs as shared-memory images (tensor-core mode, see k_mlp_tc.cu) or nullptr
    int ld32, kb1;      // fp32 row stride; 64-column blocks per row of the fp16 image
};

__device__ __forceinline__ int stc_find_utt(const int64_t *off, int n, int64_t f)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_stc(StcArgs a)
{
    __shared__ float s_win[32];
    __shared__ float s_dct[160];
    if (threadIdx.x < 32) s_win[threadIdx.x] = a.win[threadIdx.x];
    if (threadIdx.x < 160) s_dct[threadIdx.x] = a.dct[threadIdx.x];
    __syncthreads();

    const int per_frame = 2 * a.nb;
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= a.nf * per_frame) return;
    const int64_t fl = item / per_frame;          // frame within the chunk
    const int rem = (int)(item - fl * per_frame);
    const int side = rem / a.nb, b = rem - side * a.nb;
    const int64_t f = a.f0 + fl;
    const int u = stc_find_utt(a.frame_off, a.n_utt, f);
    const int64_t u0 = a.frame_off[u], T = a.frame_off[u + 1] - u0, r = f - u0;
    const float mu = a.mean[u * a.nb + b];

    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int64_t t = r - 15 + j + (side ? 15 : 0);
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        const float v = __fsub_rn(a.mel[(u0 + t) * a.nb + b], mu);   // sentence mean normalisation
        x[j] = __fmul_rn(v, s_win[side * 16 + j]);                   // traps.cpp:300-313
    }
    const float *nm = side ? a.nmean1 : a.nmean0;
    const float *nd = side ? a.ndev1 : a.ndev0;
    const int col0 = b * a.ncoef;
    float *o32 = side ? a.x1 : a.x0;
    uint8_t *o16 = side ? a.x1h : a.x0h;

    for (int k = 0; k < a.ncoef; ++k) {
        float s = 0.0f;
        if (k == 0) {  // CalcC0 (dspc.h:223-233): plain sum, no 1/sqrt2
#pragma unroll
            for (int j = 0; j < 16; ++j) s = __fadd_rn(s, x[j]);
        } else {       // sDCT (dspc.h:206-221): cos table row k-1
#pragma unroll
            for (int j = 0; j < 16; ++j) s = __fadd_rn(s, __fmul_rn(x[j], s_dct[(k - 1) * 16 + j]));
        }
        s = __fmul_rn(s, a.normc);
        const int col = col0 + k;
        const float xn = __fmul_rn(__fsub_rn(s, nm[col]), nd[col]);  // NeuralNet::Normalize nn.cpp:702-716
        if (o32) o32[fl * a.ld32 + col] = xn;
        if (o16) {  // [tile of 128 frames][64-column block][128 rows x 128 B, 16-byte chunks XOR-swizzled by row%8]
            const int r = (int)(fl & 127), cc = col & 63;
            uint8_t *blk = o16 + ((size_t)(fl >> 7) * a.kb1 + (col >> 6)) * 16384;
            *reinterpret_cast<__half *>(blk + r * 128 + ((((unsigned)cc >> 3) ^ ((unsigned)r & 7u)) << 4) + (cc & 7) * 2) =
                __float2half_rn(xn);
        }
    }
}

int launch_stc(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    StcArgs a;
    a.mel = (const float *)c->d_mel.p;
    a.mean = (const float *)c->d_mean.p;
    a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt; a.nb = c->nbanks; a.ncoef = c->ncoef;
    a.f0 = f0; a.nf = nf;
    a.win = c->tab.win; a.dct = c->tab.dct;
    a.nmean0 = c->net[0].mean; a.ndev0 = c->net[0].dev;
    a.nmean1 = c->net[1].mean; a.ndev1 = c->net[1].dev;
    a.normc = sqrtf(2.0f / 16.0f);
    const bool tc = c->mlp_mode == PHN_MLP_TC_F16;
    a.x0 = tc ? nullptr : (float *)c->d_x0.p;
    a.x1 = tc ? nullptr : (float *)c->d_x1.p;
    a.x0h = tc ? (uint8_t *)c->d_x0h.p : nullptr;
    a.x1h = tc ? (uint8_t *)c->d_x1h.p : nullptr;
    a.ld32 = c->net[0].kp;
    a.kb1 = c->net[0].k1P / 64;
    const int64_t items = nf * 2 * c->nbanks;
    k_stc<<<(unsigned)((items + 255) / 256), 256, 0, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_STC] += 1;
    return PHN_OK;
}

}  // namespace phn
