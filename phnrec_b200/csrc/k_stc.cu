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

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

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


// ------------------------------------------------------------------------------------------------
// Tensor-core pipeline: the same features as 16 x 16 x 16 matrix products on mma.sync.
// x[frame][11 b + n] = sum_j (mel[c(frame, j)][b] - mean[b]) * Bm[s][b][j][n] - bias,
//     Bm = window[s][j] * basis[n][j] * sqrt(2/16) * dev[s][11 b + n],  bias = nn_mean * dev
// (window, C0/DCT basis, the sqrt(2/16) factor and NeuralNet::Normalize folded into one constant matrix per
// (side, band)).  A = 16 frames x 16 taps, B = 16 taps x 16 (11 used) coefficients, both split into fp16 hi + lo,
// three products (hi.hi + lo.hi + hi.lo) accumulated in fp32: 2^-22 relative, i.e. fp32-grade features for the
// price of 6 MMAs per (16 frames, side, band) instead of 2816 FMAs.
// CTA = one 128-frame tile of the fp16 activation images (8 warps x 16 frames), persistent over tiles.
struct StcMmaArgs {
    const float *mel, *mean;
    const int64_t *frame_off;
    int n_utt, nb;
    int64_t f0, nf, total_frames;
    const uint4 *btab;     // [2][nb][2 n-tiles][32 lanes] B fragments {hi k0-7, hi k8-15, lo k0-7, lo k8-15}
    const float *bias;     // [2][nb * 11]
    uint8_t *x0h, *x1h;
    int kb1, n_tiles, tile_lo;   // tiles [tile_lo, tile_lo + n_tiles) of the pass
    int64_t row_lo, row_hi;      // only frames row_lo <= f0 + fl < row_hi are produced (a group of whole utterances)
};

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_half2(float x, float y, uint32_t &hi, uint32_t &lo)
{
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

constexpr int STCM_F = 128;     // frames per CTA pass = one image tile

template <int NB>
__global__ void __launch_bounds__(256, 2) k_stc_mma(StcMmaArgs a)
{
    constexpr int NIN = NB * 11;
    constexpr int COLS = (NIN + 2 + 63) / 64 * 64;   // image columns (64-column blocks), = 64 * kb1
    constexpr int STCM_LD = COLS + 8;                // staging row stride in halves (16-byte aligned rows, conflict-free)
    extern __shared__ __align__(16) uint8_t stc_smem[];
    uint4 *s_b = reinterpret_cast<uint4 *>(stc_smem);                                   // [2][NB][2][32]
    float *s_bias = reinterpret_cast<float *>(s_b + 2 * NB * 2 * 32);                   // [2][NIN]
    float *s_mel = s_bias + 2 * NIN;                                                    // [(128 + 30)][NB]
    __half *s_out = reinterpret_cast<__half *>(s_mel + (STCM_F + 30) * NB + ((STCM_F + 30) * NB & 1) + 2);   // [8][16][STCM_LD]
    __shared__ int64_t s_u0[STCM_F];
    __shared__ int s_T[STCM_F], s_u[STCM_F];
    s_out = reinterpret_cast<__half *>((reinterpret_cast<uintptr_t>(s_out) + 15) & ~(uintptr_t)15);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 2 * NB * 2 * 32; i += blockDim.x) s_b[i] = a.btab[i];
    for (int i = threadIdx.x; i < 2 * NIN; i += blockDim.x) s_bias[i] = a.bias[i];
    // image columns beyond the features: the two constant-1 inputs that multiply the bias columns of the layer-1
    // weight image (k_mlp_tc.cu), then zeros; never overwritten below
    __half *my_out = s_out + (size_t)warp * 16 * STCM_LD;
    for (int i = lane; i < 16 * (COLS - NIN); i += 32) {
        const int r = i / (COLS - NIN), cidx = NIN + i % (COLS - NIN);
        my_out[r * STCM_LD + cidx] = __float2half_rn(cidx < NIN + 2 ? 1.0f : 0.0f);
    }
    for (int tile = a.tile_lo + blockIdx.x; tile < a.tile_lo + a.n_tiles; tile += gridDim.x) {
        const int64_t fl0 = (int64_t)tile * STCM_F;
        const int64_t G0 = a.f0 + fl0 - 15;
        __syncthreads();   // (previous tile's readers of s_mel / s_u are done)
        for (int q = threadIdx.x; q < (STCM_F + 30) * NB; q += blockDim.x) {
            const int64_t gr = G0 + q / NB;
            s_mel[q] = (gr >= 0 && gr < a.total_frames) ? a.mel[G0 * NB + q] : 0.0f;
        }
        if (threadIdx.x < STCM_F) {
            const int64_t fl = fl0 + threadIdx.x;
            if (fl < a.nf && a.f0 + fl >= a.row_lo && a.f0 + fl < a.row_hi) {
                const int u = stc_find_utt(a.frame_off, a.n_utt, a.f0 + fl);
                s_u[threadIdx.x] = u;
                s_u0[threadIdx.x] = a.frame_off[u];
                s_T[threadIdx.x] = (int)(a.frame_off[u + 1] - a.frame_off[u]);
            } else {
                s_u[threadIdx.x] = -1;
            }
        }
        __syncthreads();
        // this thread's two frames (fragment rows g and g + 8 of the warp's 16)
        const int fa = warp * 16 + g, fb = fa + 8;
        const int ua = s_u[fa], ub = s_u[fb];
        const int Ta = ua < 0 ? 1 : s_T[fa], Tb = ub < 0 ? 1 : s_T[fb];
        const int64_t u0a = ua < 0 ? 0 : s_u0[fa], u0b = ub < 0 ? 0 : s_u0[fb];
        const int ra = (int)(a.f0 + fl0 + fa - u0a), rb = (int)(a.f0 + fl0 + fb - u0b);
        const float *mean_a = a.mean + (size_t)(ua < 0 ? 0 : ua) * NB, *mean_b = a.mean + (size_t)(ub < 0 ? 0 : ub) * NB;
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            // shared-memory row offsets of the 4 taps (2t, 2t+1, 2t+8, 2t+9) of both frames; clamped inside the utterance
            int oa[4], ob[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 2 * t + (i & 1) + (i >> 1) * 8;
                int ta = ra - 15 + j + side * 15, tb = rb - 15 + j + side * 15;
                ta = ta < 0 ? 0 : (ta > Ta - 1 ? Ta - 1 : ta);
                tb = tb < 0 ? 0 : (tb > Tb - 1 ? Tb - 1 : tb);
                oa[i] = ua < 0 ? 0 : (int)(u0a + ta - G0) * NB;
                ob[i] = ub < 0 ? 0 : (int)(u0b + tb - G0) * NB;
            }
            const uint4 *bt = s_b + (size_t)side * NB * 64 + lane;
            const float *bias = s_bias + side * NIN;
#pragma unroll 3
            for (int b = 0; b < NB; ++b) {
                // (rows outside the pass or outside this launch's row range read window row 0 and utterance 0's mean: finite
                // numbers that are never written out)
                const float ma = __ldg(mean_a + b), mb = __ldg(mean_b + b);
                uint32_t ah[4], al[4];   // A fragment: {row g k 2t..}, {row g+8 k 2t..}, {row g k 2t+8..}, {row g+8 k 2t+8..}
                split_half2(s_mel[oa[0] + b] - ma, s_mel[oa[1] + b] - ma, ah[0], al[0]);
                split_half2(s_mel[ob[0] + b] - mb, s_mel[ob[1] + b] - mb, ah[1], al[1]);
                split_half2(s_mel[oa[2] + b] - ma, s_mel[oa[3] + b] - ma, ah[2], al[2]);
                split_half2(s_mel[ob[2] + b] - mb, s_mel[ob[3] + b] - mb, ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const uint4 bf = bt[(b * 2 + nt) * 32];
                    const int n = nt * 8 + 2 * t;    // coefficient of d[0] / d[2]; d[1] / d[3] are n + 1
                    const int col = b * 11 + n;
                    // the accumulator starts at -mean * dev of its column (NeuralNet::Normalize folded in)
                    const float nb0 = n < 11 ? -bias[col] : 0.0f, nb1 = n + 1 < 11 ? -bias[col + 1] : 0.0f;
                    float d[4] = {nb0, nb1, nb0, nb1};
                    mma16816(d, al, bf.x, bf.y);     // small terms first
                    mma16816(d, ah, bf.z, bf.w);
                    mma16816(d, ah, bf.x, bf.y);
                    if (n < 11) {
                        my_out[g * STCM_LD + col] = __float2half_rn(d[0]);
                        my_out[(g + 8) * STCM_LD + col] = __float2half_rn(d[2]);
                        if (n + 1 < 11) {
                            my_out[g * STCM_LD + col + 1] = __float2half_rn(d[1]);
                            my_out[(g + 8) * STCM_LD + col + 1] = __float2half_rn(d[3]);
                        }
                    }
                }
            }
            __syncwarp();
            // the warp's 16 rows x 24 chunks of 16 bytes -> image (K-major SW128 blocks, k_mlp_tc.cu)
            uint8_t *img = (side ? a.x1h : a.x0h) + (size_t)tile * a.kb1 * 16384;
#pragma unroll 4
            for (int q = lane; q < 16 * (COLS / 8); q += 32) {
                const int r16 = q / (COLS / 8), ch = q - r16 * (COLS / 8);   // (division by a compile-time constant)
                const int r = warp * 16 + r16;
                if (s_u[r] >= 0)   // (inside the pass and inside this launch's row range)
                    *reinterpret_cast<uint4 *>(img + (size_t)(ch >> 3) * 16384 + r * 128 + ((((unsigned)ch & 7u) ^ ((unsigned)r & 7u)) << 4)) =
                        *reinterpret_cast<const uint4 *>(my_out + r16 * STCM_LD + ch * 8);
            }
            __syncwarp();
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Tensor-core pipeline, fp32 form (default): the same features on the CUDA cores with packed fp32 pairs (sm_100 FFMA2).
//     x[frame][11 b + n] = (sum_j (mel[c(frame, j)][b] - mean[b]) * Cf[s][n][j]) * dev[s][11 b + n] - nn_mean * dev
//     Cf[s][n][j] = window[s][j] * basis[n][j] * sqrt(2/16)        (one table for all bands)
// One thread = (4 consecutive frames, 2 adjacent bands, one side): the 19 log-mel rows its frames' 16-tap windows cover
// sit in registers as (band b, band b+1) pairs - adjacent floats of a row of the shared-memory tile, one 64-bit load
// each - and every FFMA2 does two bands' worth of one tap against a scalar (broadcast) coefficient.  704 FFMA2 per thread
// and side against ~300 instructions of loads, index arithmetic, conversions and stores: ~130 warp-instructions per frame
// where the mma.sync formulation above (three hi/lo products + the splits of A) needs ~330.  fp32 throughout: the only
// rounding to fp16 is the final one into the activation image.
// CTA = one 128-frame image tile, persistent; the staging tile [128 rows][COLS] goes out as 16-byte chunks of the
// K-major SWIZZLE_128B image exactly like the mma.sync kernel's.
struct StcF2Args {
    const float *mel, *mean;
    const int64_t *frame_off;
    int n_utt;
    int64_t f0, nf, total_frames;
    const float *cf;       // [2][11][16]
    const float4 *sb;      // [2][11][NBP] {dev(b0), dev(b1), -mean*dev(b0), -mean*dev(b1)}
    uint8_t *x0h, *x1h;
    int kb1, n_tiles, tile_lo;
    int64_t row_lo, row_hi;
};

struct sf2 { float x, y; };
__device__ __forceinline__ uint64_t s_pk2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ sf2 s_up2(uint64_t v)
{
    sf2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
#ifndef PHN_STC_SCALAR      // (kernel development: 1 = two scalar FFMAs in place of every FFMA2)
#define PHN_STC_SCALAR 0
#endif
__device__ __forceinline__ sf2 s_ffma2(sf2 a, sf2 b, sf2 c)
{
#if PHN_STC_SCALAR
    return sf2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)};
#endif
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(s_pk2(a.x, a.y)), "l"(s_pk2(b.x, b.y)), "l"(s_pk2(c.x, c.y)));
    return s_up2(d);
}
__device__ __forceinline__ sf2 s_fadd2(sf2 a, sf2 b)
{
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(s_pk2(a.x, a.y)), "l"(s_pk2(b.x, b.y)));
    return s_up2(d);
}
__device__ __forceinline__ uint32_t s_h2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// one frame of a unit that is not 4 frames of one utterance (utterance boundary inside it, end of the batch): rare
template <int PITCH, int NBP>
__device__ __noinline__ void stc_f2_one_frame(const float *s_mel, const float *cf, const float4 *sb, sf2 nmean, int base, int t0, int T,
                                              int bp, bool has_b1, __half *orow)
{
    sf2 w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int t = t0 + j;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        const float2 v = *reinterpret_cast<const float2 *>(s_mel + (base + t) * PITCH + 2 * bp);
        w[j] = s_fadd2(sf2{v.x, v.y}, nmean);
    }
#pragma unroll 1
    for (int n = 0; n < 11; ++n) {
        sf2 acc = {0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 16; ++j) { const float c = cf[n * 16 + j]; acc = s_ffma2(w[j], sf2{c, c}, acc); }
        const float4 k4 = sb[n * NBP];
        const sf2 x = s_ffma2(acc, sf2{k4.x, k4.y}, sf2{k4.z, k4.w});
        orow[n] = __float2half_rn(x.x);
        if (has_b1) orow[11 + n] = __float2half_rn(x.y);
    }
}

// Resident CTAs per SM the register budget is sized for: 3 (85 registers, a few spilled bytes) where three tiles' shared
// memory fits (15 banks: 68 KB per CTA; measured 0.444 against 0.468 ms), 2 (124 registers) for the 23-bank system, whose
// 91 KB per CTA allow two anyway (0.629 against 0.652 ms with the tighter budget).
template <int NB>
__global__ void __launch_bounds__(256, NB <= 15 ? 3 : 2) k_stc_f2(StcF2Args a)
{
    constexpr int NBP = (NB + 1) / 2;                 // band pairs (the last one is half empty when NB is odd)
    constexpr int NIN = NB * 11;
    constexpr int COLS = (NIN + 2 + 63) / 64 * 64;    // image columns = 64 * kb1
    constexpr int LD = COLS + 8;                      // staging row stride in halves (16-byte aligned rows)
    constexpr int PITCH = (2 * NBP + 7) / 8 * 8 + 4;  // floats per log-mel row in shared memory: 4 rows = 64 B mod 128 B
    constexpr int ROWS = STCM_F + 30;
    constexpr int UNITS = 32 * NBP;                   // (frame group, band pair) units of one side
    extern __shared__ __align__(16) uint8_t stc_smem[];
    float *s_cf = reinterpret_cast<float *>(stc_smem);                          // [2][11][16]
    float4 *s_sb = reinterpret_cast<float4 *>(s_cf + 2 * 11 * 16);              // [2][11][NBP]
    float *s_mel = reinterpret_cast<float *>(s_sb + 2 * 11 * NBP);              // [ROWS][PITCH]
    __half *s_out = reinterpret_cast<__half *>(s_mel + ROWS * PITCH);           // [128][LD]
    __shared__ int64_t s_u0[STCM_F];
    __shared__ int s_T[STCM_F], s_u[STCM_F];
    for (int i = threadIdx.x; i < 2 * 11 * 16; i += blockDim.x) s_cf[i] = a.cf[i];
    for (int i = threadIdx.x; i < 2 * 11 * NBP; i += blockDim.x) s_sb[i] = a.sb[i];
    // image columns beyond the features: the two constant-1 inputs that multiply the bias columns of the layer-1 weight
    // image (k_mlp_tc.cu), then zeros; only column NIN is ever rewritten below (with the same 1.0)
    for (int i = threadIdx.x; i < STCM_F * (COLS - NIN); i += blockDim.x) {
        const int r = i / (COLS - NIN), cidx = NIN + i % (COLS - NIN);
        s_out[r * LD + cidx] = __float2half_rn(cidx < NIN + 2 ? 1.0f : 0.0f);
    }
    for (int i = threadIdx.x; i < ROWS * (PITCH - NB); i += blockDim.x)   // padding columns of the tile (the half-empty pair)
        s_mel[(i / (PITCH - NB)) * PITCH + NB + i % (PITCH - NB)] = 0.0f;
    for (int tile = a.tile_lo + blockIdx.x; tile < a.tile_lo + a.n_tiles; tile += gridDim.x) {
        const int64_t fl0 = (int64_t)tile * STCM_F;
        const int64_t G0 = a.f0 + fl0 - 15;
        __syncthreads();   // (previous tile's readers of s_mel / s_u / s_out are done)
        for (int q = threadIdx.x; q < ROWS * NB; q += blockDim.x) {
            const int r = q / NB, b = q - r * NB;
            const int64_t gr = G0 + r;
            s_mel[r * PITCH + b] = (gr >= 0 && gr < a.total_frames) ? a.mel[G0 * NB + q] : 0.0f;
        }
        if (threadIdx.x < STCM_F) {
            const int64_t fl = fl0 + threadIdx.x;
            if (fl < a.nf && a.f0 + fl >= a.row_lo && a.f0 + fl < a.row_hi) {
                const int u = stc_find_utt(a.frame_off, a.n_utt, a.f0 + fl);
                s_u[threadIdx.x] = u;
                s_u0[threadIdx.x] = a.frame_off[u];
                s_T[threadIdx.x] = (int)(a.frame_off[u + 1] - a.frame_off[u]);
            } else {
                s_u[threadIdx.x] = -1;
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int side = 0; side < 2; ++side) {
            const float *cf = s_cf + side * 176;
#pragma unroll 1
            for (int q = threadIdx.x; q < UNITS; q += blockDim.x) {
                const int fg = q / NBP, bp = q - fg * NBP;
                const int fr = 4 * fg;                          // first of the unit's 4 tile rows
                const int u = s_u[fr];
                const bool has_b1 = 2 * bp + 1 < NB;
                const float4 *sb = s_sb + side * 11 * NBP + bp;
                __half *orow = s_out + (size_t)fr * LD + 22 * bp;
                if (u >= 0 && s_u[fr + 3] == u) {               // 4 valid frames of one utterance (the common case)
                    const int T = s_T[fr];
                    const int base = (int)(s_u0[fr] - G0);
                    const int t0 = (int)(a.f0 + fl0 + fr - s_u0[fr]) - 15 + 15 * side;
                    const float *mp = a.mean + (size_t)u * NB + 2 * bp;
                    const sf2 nmean = {-__ldg(mp), has_b1 ? -__ldg(mp + 1) : 0.0f};
                    // the 19 rows the four 16-tap windows cover, clamped inside the utterance (S1: c(t, j) = clamp(t - 15 + j))
                    sf2 w[19];
#pragma unroll
                    for (int i = 0; i < 19; ++i) {
                        int t = t0 + i;
                        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
                        const float2 v = *reinterpret_cast<const float2 *>(s_mel + (base + t) * PITCH + 2 * bp);
                        w[i] = s_fadd2(sf2{v.x, v.y}, nmean);
                    }
                    // Coefficient n of both bands for the 4 frames.  The unit's 22 image columns go out as 11 pairs of halves:
                    // band b0 = 2 bp starts at an even column, so its pairs are (0,1) .. (8,9) and (10 | b1's 0); b1's are
                    // (1,2) .. (9,10).  When b1 does not exist (NB odd, last pair) its first column is the constant 1.0.
                    float holdA[4], holdB[4], holdB0[4];
#pragma unroll
                    for (int n = 0; n < 11; ++n) {
                        float c[16];
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 cc = *reinterpret_cast<const float4 *>(cf + n * 16 + 4 * j4);
                            c[4 * j4] = cc.x; c[4 * j4 + 1] = cc.y; c[4 * j4 + 2] = cc.z; c[4 * j4 + 3] = cc.w;
                        }
                        sf2 acc[4] = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}};
#pragma unroll
                        for (int j = 0; j < 16; ++j)
#pragma unroll
                            for (int d = 0; d < 4; ++d) acc[d] = s_ffma2(w[d + j], sf2{c[j], c[j]}, acc[d]);
                        const float4 k4 = sb[n * NBP];
#pragma unroll
                        for (int d = 0; d < 4; ++d) {
                            const sf2 x = s_ffma2(acc[d], sf2{k4.x, k4.y}, sf2{k4.z, k4.w});   // NeuralNet::Normalize folded in
                            uint32_t *o32 = reinterpret_cast<uint32_t *>(orow + d * LD);
                            if (n == 10) o32[5] = s_h2(x.x, has_b1 ? holdB0[d] : 1.0f);
                            else if ((n & 1) == 0) holdA[d] = x.x;
                            else o32[(n - 1) / 2] = s_h2(holdA[d], x.x);
                            if (n == 0) holdB0[d] = x.y;
                            else if (n & 1) holdB[d] = x.y;
                            else if (has_b1) o32[(10 + n) / 2] = s_h2(holdB[d], x.y);
                        }
                    }
                } else {
#pragma unroll 1
                    for (int d = 0; d < 4; ++d) {
                        const int ud = s_u[fr + d];
                        if (ud < 0) continue;                   // (outside the pass or this launch's row range: never written out)
                        const float *mp = a.mean + (size_t)ud * NB + 2 * bp;
                        const sf2 nmean = {-__ldg(mp), has_b1 ? -__ldg(mp + 1) : 0.0f};
                        stc_f2_one_frame<PITCH, NBP>(s_mel, cf, sb, nmean, (int)(s_u0[fr + d] - G0),
                                                     (int)(a.f0 + fl0 + fr + d - s_u0[fr + d]) - 15 + 15 * side, s_T[fr + d], bp, has_b1,
                                                     orow + d * LD);
                    }
                }
            }
            __syncthreads();
            // the tile's 128 rows x COLS / 8 chunks of 16 bytes -> image (K-major SW128 blocks, k_mlp_tc.cu)
            uint8_t *img = (side ? a.x1h : a.x0h) + (size_t)tile * a.kb1 * 16384;
#pragma unroll 4
            for (int q = threadIdx.x; q < STCM_F * (COLS / 8); q += blockDim.x) {
                const int r = q / (COLS / 8), ch = q - r * (COLS / 8);   // (division by a compile-time constant)
                if (s_u[r] >= 0)   // (inside the pass and inside this launch's row range)
                    *reinterpret_cast<uint4 *>(img + (size_t)(ch >> 3) * 16384 + r * 128 + ((((unsigned)ch & 7u) ^ ((unsigned)r & 7u)) << 4)) =
                        *reinterpret_cast<const uint4 *>(s_out + r * LD + ch * 8);
            }
            __syncthreads();
        }
    }
}

// tables of the fp32 form, built once per context on the host
static int stc_f2_prepare(phn_ctx *c)
{
    if (c->stc_cf) return PHN_OK;
    const int nb = c->nbanks, nbp = (nb + 1) / 2, nin = nb * 11;
    std::vector<float> cf((size_t)2 * 11 * 16);
    std::vector<float> sb((size_t)2 * 11 * nbp * 4);
    const double normc = (double)sqrtf(2.0f / 16.0f);
    const float PiByN = (float)M_PI / 16.0f;
    for (int s = 0; s < 2; ++s)
        for (int n = 0; n < 11; ++n) {
            for (int j = 0; j < 16; ++j) {
                const double basis = n == 0 ? 1.0 : (double)cosf(PiByN * (float)n * ((float)j + 0.5f));   // CalcC0 / sDCT, dspc.h:206-233
                cf[((size_t)s * 11 + n) * 16 + j] = (float)((double)c->win[s * 16 + j] * basis * normc);
            }
            for (int bp = 0; bp < nbp; ++bp) {
                float *o = &sb[(((size_t)s * 11 + n) * nbp + bp) * 4];
                for (int h = 0; h < 2; ++h) {
                    const int b = 2 * bp + h;
                    const float dev = b < nb ? c->hnet[s].dev[b * 11 + n] : 0.0f, mean = b < nb ? c->hnet[s].mean[b * 11 + n] : 0.0f;
                    o[h] = dev;
                    o[2 + h] = -mean * dev;
                }
            }
        }
    (void)nin;
    PHN_CUDA(c, cudaMalloc(&c->stc_cf, cf.size() * 4));
    PHN_CUDA(c, cudaMalloc(&c->stc_sb, sb.size() * 4));
    PHN_CUDA(c, cudaMemcpy(c->stc_cf, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice));
    PHN_CUDA(c, cudaMemcpy(c->stc_sb, sb.data(), sb.size() * 4, cudaMemcpyHostToDevice));
    return PHN_OK;
}

template <int NB>
static int launch_stc_f2_t(phn_ctx *c, const StcF2Args &a)
{
    constexpr int NBP = (NB + 1) / 2, NIN = NB * 11, COLS = (NIN + 2 + 63) / 64 * 64, LD = COLS + 8, PITCH = (2 * NBP + 7) / 8 * 8 + 4;
    const size_t smem = sizeof(float) * 2 * 11 * 16 + sizeof(float4) * 2 * 11 * NBP + sizeof(float) * (STCM_F + 30) * PITCH + (size_t)STCM_F * LD * 2;
    PHN_CUDA(c, cudaFuncSetAttribute(k_stc_f2<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // (grid = a few times the resident count: see launch_wave_pair_k)
    static const int oversub = getenv("PHNREC_FRONT_OVERSUB") ? atoi(getenv("PHNREC_FRONT_OVERSUB")) : 4;
    const int gcap = (NB <= 15 ? 3 : 2) * c->num_sms * (oversub > 0 ? oversub : 1);
    int grid = a.n_tiles < gcap ? a.n_tiles : gcap;
    k_stc_f2<NB><<<grid, 256, smem, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

static int launch_stc_f2(phn_ctx *c, int64_t f0, int64_t nf, int64_t row_lo, int64_t row_hi)
{
    int rc;
    if ((rc = stc_f2_prepare(c))) return rc;
    StcF2Args a;
    a.mel = (const float *)c->d_mel.p;
    a.mean = (const float *)c->d_mean.p;
    a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt;
    a.f0 = f0; a.nf = nf; a.total_frames = c->total_frames;
    a.cf = (const float *)c->stc_cf; a.sb = (const float4 *)c->stc_sb;
    a.x0h = (uint8_t *)c->d_x0h.p; a.x1h = (uint8_t *)c->d_x1h.p;
    a.kb1 = c->net[0].k1P / 64;
    if (row_lo < f0) row_lo = f0;
    if (row_hi > f0 + nf) row_hi = f0 + nf;
    if (row_hi <= row_lo) return PHN_OK;
    a.row_lo = row_lo; a.row_hi = row_hi;
    a.tile_lo = (int)((row_lo - f0) / STCM_F);
    a.n_tiles = (int)((row_hi - f0 + STCM_F - 1) / STCM_F) - a.tile_lo;
    rc = c->nbanks == 15 ? launch_stc_f2_t<15>(c, a) : launch_stc_f2_t<23>(c, a);
    if (rc) return rc;
    c->k_launches[PHN_K_STC] += 1;
    return PHN_OK;
}

// constant matrices of the tensor-core formulation, built once per context on the host (fp16 hi + lo fragments)
static int stc_mma_prepare(phn_ctx *c)
{
    if (c->stc_btab) return PHN_OK;
    const int nb = c->nbanks, nin = nb * 11;
    std::vector<uint32_t> tab((size_t)2 * nb * 2 * 32 * 4);
    std::vector<float> bias((size_t)2 * nin);
    const double normc = (double)sqrtf(2.0f / 16.0f);
    const float PiByN = (float)M_PI / 16.0f;
    auto coef = [&](int s, int b, int j, int n) -> double {   // Bm[s][b][j][n]
        if (n >= 11) return 0.0;
        const double basis = n == 0 ? 1.0 : (double)cosf(PiByN * (float)n * ((float)j + 0.5f));   // CalcC0 / sDCT, dspc.h:206-233
        return (double)c->win[s * 16 + j] * basis * normc * (double)c->hnet[s].dev[b * 11 + n];
    };
    auto pack = [&](double v0, double v1, uint32_t &hi, uint32_t &lo) {
        const __half h0 = __float2half_rn((float)v0), h1 = __float2half_rn((float)v1);
        const __half l0 = __float2half_rn((float)(v0 - (double)__half2float(h0))), l1 = __float2half_rn((float)(v1 - (double)__half2float(h1)));
        hi = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        lo = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    };
    for (int s = 0; s < 2; ++s)
        for (int b = 0; b < nb; ++b) {
            for (int nt = 0; nt < 2; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
                    uint32_t *o = &tab[((((size_t)s * nb + b) * 2 + nt) * 32 + lane) * 4];
                    pack(coef(s, b, 2 * t, n), coef(s, b, 2 * t + 1, n), o[0], o[2]);          // k = 2t, 2t+1
                    pack(coef(s, b, 2 * t + 8, n), coef(s, b, 2 * t + 9, n), o[1], o[3]);      // k = 2t+8, 2t+9
                }
            for (int n = 0; n < 11; ++n) bias[(size_t)s * nin + b * 11 + n] = c->hnet[s].mean[b * 11 + n] * c->hnet[s].dev[b * 11 + n];
        }
    PHN_CUDA(c, cudaMalloc(&c->stc_btab, tab.size() * 4));
    PHN_CUDA(c, cudaMalloc(&c->stc_bias, bias.size() * 4));
    PHN_CUDA(c, cudaMemcpy(c->stc_btab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    PHN_CUDA(c, cudaMemcpy(c->stc_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    return PHN_OK;
}

template <int NB>
static int launch_stc_mma_t(phn_ctx *c, const StcMmaArgs &a)
{
    constexpr int NIN = NB * 11, COLS = (NIN + 2 + 63) / 64 * 64;
    const size_t smem = (size_t)2 * NB * 2 * 32 * 16 + sizeof(float) * 2 * NIN + sizeof(float) * ((STCM_F + 30) * NB + 4) +
                        (size_t)8 * 16 * (COLS + 8) * 2 + 32;
    PHN_CUDA(c, cudaFuncSetAttribute(k_stc_mma<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = a.n_tiles < 2 * c->num_sms ? a.n_tiles : 2 * c->num_sms;
    k_stc_mma<NB><<<grid, 256, smem, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

static int launch_stc_mma(phn_ctx *c, int64_t f0, int64_t nf, int64_t row_lo, int64_t row_hi)
{
    int rc;
    if ((rc = stc_mma_prepare(c))) return rc;
    StcMmaArgs a;
    a.mel = (const float *)c->d_mel.p;
    a.mean = (const float *)c->d_mean.p;
    a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt; a.nb = c->nbanks;
    a.f0 = f0; a.nf = nf; a.total_frames = c->total_frames;
    a.btab = (const uint4 *)c->stc_btab; a.bias = (const float *)c->stc_bias;
    a.x0h = (uint8_t *)c->d_x0h.p; a.x1h = (uint8_t *)c->d_x1h.p;
    a.kb1 = c->net[0].k1P / 64;
    if (row_lo < f0) row_lo = f0;
    if (row_hi > f0 + nf) row_hi = f0 + nf;
    if (row_hi <= row_lo) return PHN_OK;
    a.row_lo = row_lo; a.row_hi = row_hi;
    a.tile_lo = (int)((row_lo - f0) / STCM_F);
    a.n_tiles = (int)((row_hi - f0 + STCM_F - 1) / STCM_F) - a.tile_lo;
    rc = c->nbanks == 15 ? launch_stc_mma_t<15>(c, a) : launch_stc_mma_t<23>(c, a);
    if (rc) return rc;
    c->k_launches[PHN_K_STC] += 1;
    return PHN_OK;
}

int launch_stc(phn_ctx *c, int64_t f0, int64_t nf, int64_t row_lo, int64_t row_hi)
{
    if (nf == 0) return PHN_OK;
    if (row_hi < 0) row_hi = f0 + nf;
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
    if (tc && (c->nbanks == 15 || c->nbanks == 23) && c->net[0].k1P / 64 == (c->nbanks * 11 + 2 + 63) / 64) {
        static const bool use_mma = getenv("PHNREC_STC") && !strcmp(getenv("PHNREC_STC"), "mma");   // (kernel development: the mma.sync form)
        return use_mma ? launch_stc_mma(c, f0, nf, row_lo, row_hi) : launch_stc_f2(c, f0, nf, row_lo, row_hi);
    }
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
