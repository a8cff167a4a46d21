// k_mlp_exact.cu — K-mlp, fp32 "exact" mode: the three MLPs on CUDA cores with the reference's
// own arithmetic, so posteriors are bit-identical to the reference CPU build (no-BLAS path).
//
// Replaces NeuralNet::Forward / ForwardPass1Bunch / MatrixMultiplyAndAdd (nn.cpp:771-793,
// 872-950), PrepareBiases (nn.cpp:857-870), fexp_sigmoid / fexp_softmax_v (fexp.h:33-78) and
// Traps::CalcInputFeaturesForMerger case stlcrc (traps.cpp:435-461).
//
// Why it is exact: the reference computes every pre-activation as  acc = bias; for k ascending:
// acc += x[k]*w[k]  with one rounding per multiply and per add (x86-64 SSE2, no FMA).  Both
// kernels below keep one accumulator per output in a register, start it at the bias and walk k
// in ascending order through the shared-memory tiles using __fmul_rn/__fadd_rn (which nvcc never
// fuses).  Padding adds exact zeros.  Sigmoid and softmax use the canonical Quicknet bit-trick
// exponential in double precision, the softmax denominator is a sequential fp32 sum.
//
//   k_l1_exact : H = fsig(b1 + Xn W1^T)            tile 128 frames x 128 hidden units, 8x8 / thread
//   k_l2_exact : P = fsoftmax(b2 + H W2^T)         tile  32 frames x all outputs (<= 192)
//                band nets: Xm[:, side*nout + i] = (sLn(P) - mean_m) * dev_m   (merger input)
//                merger   : post = P
#include "internal.h"
#include "device_math.cuh"

#include <cfloat>

namespace phn {

constexpr int L1_BM = 128, L1_BN = 128, L1_BK = 16, L1_LD = 132;

__global__ void __launch_bounds__(256, 2)
k_l1_exact(const float *__restrict__ X, int ldx, const float *__restrict__ W, int ldw, const float *__restrict__ bias,
           float *__restrict__ H, int ldh, int64_t nf, int nhid, int nhid4, int K4, int Kp)
{
    __shared__ __align__(16) float As[L1_BK][L1_LD];
    __shared__ __align__(16) float Bs[L1_BK][L1_LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * L1_BM;
    const int n0 = blockIdx.x * L1_BN;

    // global -> register staging: each thread moves 2 float4 of A and 2 of B per k-tile
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    float4 ga[2], gb[2];
    auto gload = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t r = m0 + lrow + 64 * h;
            const int n = n0 + lrow + 64 * h;
            const int k = k0 + lk;
            ga[h] = (r < nf && k < ldx) ? *reinterpret_cast<const float4 *>(X + r * ldx + k) : make_float4(0, 0, 0, 0);
            gb[h] = (n < nhid4 && k < K4) ? *reinterpret_cast<const float4 *>(W + (int64_t)n * ldw + k) : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lrow + 64 * h;
            As[lk + 0][r] = ga[h].x; As[lk + 1][r] = ga[h].y; As[lk + 2][r] = ga[h].z; As[lk + 3][r] = ga[h].w;
            Bs[lk + 0][r] = gb[h].x; Bs[lk + 1][r] = gb[h].y; Bs[lk + 2][r] = gb[h].z; Bs[lk + 3][r] = gb[h].w;
        }
    };

    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        const float b = n < nhid4 ? bias[n] : 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i][j] = b;
    }

    gload(0);
    sstore();
    __syncthreads();
    for (int k0 = 0; k0 < Kp; k0 += L1_BK) {
        const bool more = k0 + L1_BK < Kp;
        if (more) gload(k0 + L1_BK);
#pragma unroll
        for (int kk = 0; kk < L1_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
        }
        __syncthreads();
        if (more) {
            sstore();
            __syncthreads();
        }
    }

    // epilogue: canonical fast sigmoid; columns >= nhid are the reference's zeroed padding
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= nf) continue;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int n = n0 + g * 64 + tx * 4;
            if (n >= ldh) continue;
            float4 o;
            o.x = n + 0 < nhid ? fsigmoid_exact(acc[i][g * 4 + 0]) : 0.0f;
            o.y = n + 1 < nhid ? fsigmoid_exact(acc[i][g * 4 + 1]) : 0.0f;
            o.z = n + 2 < nhid ? fsigmoid_exact(acc[i][g * 4 + 2]) : 0.0f;
            o.w = n + 3 < nhid ? fsigmoid_exact(acc[i][g * 4 + 3]) : 0.0f;
            *reinterpret_cast<float4 *>(H + r * ldh + n) = o;
        }
    }
}

constexpr int L2_BM = 32, L2_BN = 192, L2_BK = 16, L2_LDA = 36, L2_LDB = 196, L2_LDO = 193;

struct L2Args {
    const float *H; int ldh;
    const float *W; int ldw;
    const float *bias;
    int64_t nf;
    int nout, nout4, K4, Kp;
    // band nets: merger input
    float *xm; int ldxm; int xm_col0; const float *mmean, *mdev;
    // merger: posteriors
    float *post; int ldpost;
    int negate;   // merger input = -sLn(p) (the 1BT / 3BT systems)
};

__global__ void __launch_bounds__(256) k_l2_exact(L2Args a)
{
    __shared__ __align__(16) float As[L2_BK][L2_LDA];
    __shared__ __align__(16) float Bs[L2_BK][L2_LDB];
    __shared__ float So[L2_BM][L2_LDO];
    __shared__ float s_sc[L2_BM];
    __shared__ double s_logtab[32];
    logf_table_to_smem(s_logtab, threadIdx.x, 256);
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * L2_BM;

    float4 ga, gb[3];
    const int arow = tid >> 2, ak = (tid & 3) * 4;  // threads 0..127 carry the A tile
    auto gload = [&](int k0) {
        if (tid < 128) {
            const int64_t r = m0 + arow;
            const int k = k0 + ak;
            ga = (r < a.nf && k < a.ldh) ? *reinterpret_cast<const float4 *>(a.H + r * a.ldh + k) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const int q = tid + 256 * h;  // 0..767
            const int n = q >> 2, k = k0 + (q & 3) * 4;
            gb[h] = (n < a.nout4 && k < a.K4) ? *reinterpret_cast<const float4 *>(a.W + (int64_t)n * a.ldw + k)
                                              : make_float4(0, 0, 0, 0);
        }
    };
    auto sstore = [&]() {
        if (tid < 128) {
            As[ak + 0][arow] = ga.x; As[ak + 1][arow] = ga.y; As[ak + 2][arow] = ga.z; As[ak + 3][arow] = ga.w;
        }
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const int q = tid + 256 * h;
            const int n = q >> 2, k = (q & 3) * 4;
            Bs[k + 0][n] = gb[h].x; Bs[k + 1][n] = gb[h].y; Bs[k + 2][n] = gb[h].z; Bs[k + 3][n] = gb[h].w;
        }
    };

    float acc[4][6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int n = tx + 32 * j;
        const float b = n < a.nout4 ? a.bias[n] : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = b;
    }
    gload(0);
    sstore();
    __syncthreads();
    for (int k0 = 0; k0 < a.Kp; k0 += L2_BK) {
        const bool more = k0 + L2_BK < a.Kp;
        if (more) gload(k0 + L2_BK);
#pragma unroll
        for (int kk = 0; kk < L2_BK; ++kk) {
            const float4 av4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float av[4] = {av4.x, av4.y, av4.z, av4.w};
            float bv[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) bv[j] = Bs[kk][tx + 32 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(av[i], bv[j]));
        }
        __syncthreads();
        if (more) {
            sstore();
            __syncthreads();
        }
    }

    // ---- softmax over the first nout columns (fexp.h:49-78)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) So[ty * 4 + i][tx + 32 * j] = acc[i][j];
    __syncthreads();
    for (int r = ty * 4; r < ty * 4 + 4; ++r) {  // warp ty owns rows ty*4 .. ty*4+3
        float m = -FLT_MAX;
        for (int n = tx; n < a.nout; n += 32) m = fmaxf(m, So[r][n]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        for (int n = tx; n < a.nout; n += 32)
            So[r][n] = __double2float_rn(fexp_canonical((double)__fsub_rn(So[r][n], m)));
    }
    __syncthreads();
    if (tid < L2_BM) {  // one lane per row: the reference's sequential fp32 sum
        float s = 0.0f;
        for (int n = 0; n < a.nout; ++n) s = __fadd_rn(s, So[tid][n]);
        s_sc[tid] = __fdiv_rn(1.0f, s);
    }
    __syncthreads();
    for (int q = tid; q < L2_BM * a.nout; q += 256) {
        const int r = q / a.nout, n = q - r * a.nout;
        const int64_t f = m0 + r;
        if (f >= a.nf) continue;
        const float p = __fmul_rn(So[r][n], s_sc[r]);
        if (a.post) {
            a.post[f * a.ldpost + n] = p;
        } else {  // merger input: sLn then the merger's own input normalisation (traps.cpp:459, nn.cpp:702-716)
            const int c = a.xm_col0 + n;
            float v = ln_guarded(p, s_logtab);
            if (a.negate) v = __fmul_rn(v, -1.0f);   // sMultiplication(.., -1), traps.cpp:427
            a.xm[f * a.ldxm + c] = __fmul_rn(__fsub_rn(v, a.mmean[c]), a.mdev[c]);
        }
    }
}

int run_net_exact(phn_ctx *c, const DevNet &n, const float *x, int ldx, int64_t nf, float *post, int ldpost, int xm_col0, int negate)
{
    float *H = (float *)c->d_h.p;
    if (n.nout > L2_BN) return fail(c, PHN_ERR_UNSUPPORTED, "more than %d network outputs\n", L2_BN);
    dim3 g1((n.ldh + L1_BN - 1) / L1_BN, (unsigned)((nf + L1_BM - 1) / L1_BM));
    k_l1_exact<<<g1, 256, 0, c->stream>>>(x, ldx, n.w1, n.nin4, n.b1, H, n.ldh, nf, n.nhid, n.nhid4, n.nin4, n.kp);
    PHN_CUDA(c, cudaGetLastError());
    L2Args a{};
    a.H = H; a.ldh = n.ldh; a.W = n.w2; a.ldw = n.nhid4; a.bias = n.b2; a.nf = nf;
    a.nout = n.nout; a.nout4 = n.nout4; a.K4 = n.nhid4; a.Kp = n.ldh;
    a.negate = negate;
    if (!post) {
        a.xm = (float *)c->d_xm.p; a.ldxm = c->net[2].kp; a.xm_col0 = xm_col0;
        a.mmean = c->net[2].mean; a.mdev = c->net[2].dev;
    } else {
        a.post = post; a.ldpost = ldpost;
    }
    k_l2_exact<<<(unsigned)((nf + L2_BM - 1) / L2_BM), 256, 0, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_MLP] += 2;
    return PHN_OK;
}

int launch_mlp_exact(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    int rc;
    if ((rc = run_net_exact(c, c->net[0], (const float *)c->d_x0.p, c->net[0].kp, nf, nullptr, 0, 0, 0))) return rc;
    if ((rc = run_net_exact(c, c->net[1], (const float *)c->d_x1.p, c->net[1].kp, nf, nullptr, 0, c->net[0].nout, 0))) return rc;
    return run_net_exact(c, c->net[2], (const float *)c->d_xm.p, c->net[2].kp, nf, (float *)c->d_post.p + f0 * c->ldp, c->ldp, 0, 0);
}

// 1BT / 3BT: one net per band into the merger's input matrix (as -sLn), then the merger; 1BT_DCT: the merger only
int launch_mlp_trap(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    int rc, col = 0;
    for (size_t b = 0; b < c->dband.size(); ++b) {
        const DevNet &n = c->dband[b];
        if ((rc = run_net_exact(c, n, (const float *)c->d_xb.p + b * (size_t)c->chunk_frames * n.kp, n.kp, nf, nullptr, 0, col, 1))) return rc;
        col += n.nout;
    }
    return run_net_exact(c, c->net[2], (const float *)c->d_xm.p, c->net[2].kp, nf, (float *)c->d_post.p + f0 * c->ldp, c->ldp, 0, 0);
}

}  // namespace phn
