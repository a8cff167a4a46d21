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
    int64_t total_frames;
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

constexpr int STC_F = 32;  // frames per CTA

// One CTA = 32 consecutive frames of the chunk, both context sides.  Phase 1: one thread per
// (group of 4 frames, side, band) keeps the 4 x 16 windowed samples in registers and walks the
// 11 output coefficients, so every DCT-table word fetched from shared memory feeds 4 products.
// EXACT: separately rounded multiply / add in the reference's order (bit-identical features);
// otherwise FMAs (the tensor-core MLP rounds its inputs to fp16 anyway).
// Phase 2: the staging tile [side][frame][nin] is written out coalesced - fp32 rows for the exact
// MLP, or 16-byte chunks of the fp16 shared-memory image the tensor-core MLP loads by TMA.
template <bool EXACT, int NB>   // NB = banks when known at compile time (15, 23), 0 = read it from the arguments
__global__ void __launch_bounds__(256) k_stc(StcArgs a_)
{
    StcArgs a = a_;
    if (NB) { a.nb = NB; }
    a.ncoef = 11;   // enforced at phn_create (add_c0 + 10 DCT coefficients)
    extern __shared__ float s_val[];          // [2][STC_F][nin]
    __shared__ float s_win[32];
    __shared__ float s_dct[160];
    __shared__ int64_t s_u0[STC_F];
    __shared__ int s_T[STC_F], s_u[STC_F];
    __shared__ float s_mel[(STC_F + 30) * 32];   // mel rows G0-15 .. G0+STC_F+14 (nb <= 32)
    const int nin = a.nb * a.ncoef;
    const int64_t fl0 = (int64_t)blockIdx.x * STC_F;
    // A clamped context index always lies between the frame itself and the unclamped index, so every
    // row any of the 32 frames needs is inside this window of the global frame axis.
    const int64_t G0 = a.f0 + fl0 - 15;
    for (int q = threadIdx.x; q < (STC_F + 30) * a.nb; q += blockDim.x) {
        const int64_t g = G0 + q / a.nb;
        s_mel[q] = (g >= 0 && g < a.total_frames) ? a.mel[G0 * a.nb + q] : 0.0f;
    }
    if (threadIdx.x < 32) s_win[threadIdx.x] = a.win[threadIdx.x];
    if (threadIdx.x < 160) s_dct[threadIdx.x] = a.dct[threadIdx.x];
    if (threadIdx.x < STC_F) {
        const int64_t fl = fl0 + threadIdx.x;
        if (fl < a.nf) {
            const int u = stc_find_utt(a.frame_off, a.n_utt, a.f0 + fl);
            s_u[threadIdx.x] = u;
            s_u0[threadIdx.x] = a.frame_off[u];
            s_T[threadIdx.x] = (int)(a.frame_off[u + 1] - a.frame_off[u]);
        } else {
            s_u[threadIdx.x] = -1;
        }
    }
    __syncthreads();

    const int per_group = 2 * a.nb;
    for (int item = threadIdx.x; item < (STC_F / 4) * per_group; item += blockDim.x) {
        const int grp = item / per_group;
        const int rem = item - grp * per_group;
        const int side = rem / a.nb, b = rem - side * a.nb;
        float x[4][16];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int fi = grp * 4 + f;
            const int u = s_u[fi];
            const int64_t u0 = u < 0 ? 0 : s_u0[fi];
            const int T = u < 0 ? 1 : s_T[fi];
            const int r = (int)(a.f0 + fl0 + fi - u0);
            const float mu = u < 0 ? 0.0f : a.mean[u * a.nb + b];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                int t = r - 15 + j + (side ? 15 : 0);
                t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
                const float v = u < 0 ? 0.0f : __fsub_rn(s_mel[(int)(u0 + t - G0) * a.nb + b], mu);  // sentence mean normalisation
                x[f][j] = __fmul_rn(v, s_win[side * 16 + j]);                             // traps.cpp:300-313
            }
        }
        const float *nm = side ? a.nmean1 : a.nmean0;
        const float *nd = side ? a.ndev1 : a.ndev0;
        float *out = s_val + ((size_t)side * STC_F + grp * 4) * nin + b * a.ncoef;
        for (int k = 0; k < a.ncoef; ++k) {
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (k == 0) {  // CalcC0 (dspc.h:223-233): plain sum, no 1/sqrt2
#pragma unroll
                for (int j = 0; j < 16; ++j)
#pragma unroll
                    for (int f = 0; f < 4; ++f) s[f] = __fadd_rn(s[f], x[f][j]);
            } else {       // sDCT (dspc.h:206-221): cos table row k-1
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float cs = s_dct[(k - 1) * 16 + j];
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                        s[f] = EXACT ? __fadd_rn(s[f], __fmul_rn(x[f][j], cs)) : fmaf(x[f][j], cs, s[f]);
                }
            }
            const int col = b * a.ncoef + k;
            const float m = nm[col], d = nd[col];
#pragma unroll
            for (int f = 0; f < 4; ++f)  // NeuralNet::Normalize nn.cpp:702-716
                out[(size_t)f * nin + k] = __fmul_rn(__fsub_rn(__fmul_rn(s[f], a.normc), m), d);
        }
    }
    __syncthreads();

    if (a.x0) {   // fp32 rows [fl][ld32]; padding columns stay zero from allocation
        for (int q = threadIdx.x; q < 2 * STC_F * nin; q += blockDim.x) {
            const int side = q / (STC_F * nin);
            const int rem = q - side * STC_F * nin;
            const int fi = rem / nin, col = rem - fi * nin;
            if (fl0 + fi < a.nf) (side ? a.x1 : a.x0)[(fl0 + fi) * a.ld32 + col] = s_val[q];
        }
    } else {      // fp16 image: [tile of 128 frames][64-column block][128 rows x 128 B], chunk index XOR row%8
        const int chunks_per_row = a.kb1 * 8;
        for (int q = threadIdx.x; q < 2 * STC_F * chunks_per_row; q += blockDim.x) {
            const int side = q / (STC_F * chunks_per_row);
            const int rem = q - side * STC_F * chunks_per_row;
            const int fi = rem / chunks_per_row, ch = rem - fi * chunks_per_row;
            const int64_t fl = fl0 + fi;
            if (fl >= a.nf) continue;
            const float *src = s_val + ((size_t)side * STC_F + fi) * nin;
            unsigned h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c0 = ch * 8 + 2 * i;
                // columns nin, nin+1: the constant 1.0 that multiplies the two bias columns of the layer-1 weight image
                const __half2 p = __floats2half2_rn(c0 < nin ? src[c0] : (c0 < nin + 2 ? 1.0f : 0.0f),
                                                    c0 + 1 < nin ? src[c0 + 1] : (c0 + 1 < nin + 2 ? 1.0f : 0.0f));
                h[i] = *reinterpret_cast<const unsigned *>(&p);
            }
            const int r = (int)(fl & 127);
            uint8_t *blk = (side ? a.x1h : a.x0h) + ((size_t)(fl >> 7) * a.kb1 + (ch >> 3)) * 16384;
            *reinterpret_cast<uint4 *>(blk + r * 128 + ((((unsigned)ch & 7u) ^ ((unsigned)r & 7u)) << 4)) =
                make_uint4(h[0], h[1], h[2], h[3]);
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
    a.f0 = f0; a.nf = nf; a.total_frames = c->total_frames;
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
    const size_t smem = sizeof(float) * 2 * STC_F * (size_t)c->nbanks * c->ncoef;
    const unsigned grid = (unsigned)((nf + STC_F - 1) / STC_F);
#define PHN_STC_LAUNCH(EX, NBV)                                                                                     \
    do {                                                                                                            \
        PHN_CUDA(c, cudaFuncSetAttribute(k_stc<EX, NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k_stc<EX, NBV><<<grid, 256, smem, c->stream>>>(a);                                                          \
    } while (0)
    const int nbv = c->nbanks == 15 ? 15 : (c->nbanks == 23 ? 23 : 0);
    if (tc) { if (nbv == 15) PHN_STC_LAUNCH(false, 15); else if (nbv == 23) PHN_STC_LAUNCH(false, 23); else PHN_STC_LAUNCH(false, 0); }
    else    { if (nbv == 15) PHN_STC_LAUNCH(true, 15);  else if (nbv == 23) PHN_STC_LAUNCH(true, 23);  else PHN_STC_LAUNCH(true, 0); }
#undef PHN_STC_LAUNCH
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_STC] += 1;
    return PHN_OK;
}

}  // namespace phn
