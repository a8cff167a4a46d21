// k_trap.cu — the TRAPS systems no shipped model uses: posteriors/system = 1BT, 3BT, 1BT_DCT.
//
// Replaces Traps::CalcInputFeaturesForBandNets / ForwardPassBandNets / CalcInputFeaturesForMerger for those systems
// (traps.cpp:222-283, 344-361, 405-433) in the exact fp32 mode: every band's L-point trajectory is the FIFO's content
// (traps.cpp:180-219) = frames clamp(r - S .. r + S, 0, T - 1), S = (L - 1) / 2, of the sentence-normalised log mel-banks,
// times the Hamming window when posteriors/hamming is set (traps.cpp:232-241).
//   1BT / 3BT   k_trap_traj: the trajectory of band b, normalised with band net b's input norms (nn.cpp:702-716), as row of
//               that net's input matrix; the nets run in k_mlp_exact.cu, their outputs go to the merger's input matrix as
//               -sLn(p) (traps.cpp:425-427) with the merger's own input normalisation.  3BT as the reference executes it:
//               the first nb - 2 bands, one band per net.
//   1BT_DCT     k_trap_dct: per band [C0, DCT_1 .. DCT_{shift-1}] (add_c0) or [DCT_1 .. DCT_shift] (dspc.h:206-233: sequential
//               fp32 sums, cosines from the host libm) straight into the merger's input matrix.
// One rounding per reference operation, in its order: the posteriors are the reference binary's bits
// (tests/test_gpu_trap_systems.py against tests/golden/ref_trap_systems.npz).
#include "internal.h"

namespace phn {

struct TrapArgs {
    const float *mel, *mean;          // [F][nb] un-normalised log mel, [U][nb] sentence means (0 when switched off)
    const int64_t *frame_off;
    int n_utt, nb, tb, L, S;
    int64_t f0, nf;                   // frames [f0, f0 + nf) of the batch = rows [0, nf) of the pass's matrices
    const float *ham;                 // [L] or nullptr
    // 1BT / 3BT
    float *xb; int ldx; int64_t band_stride;       // band b's input matrix at xb + b * band_stride, rows of ldx floats
    const float *const *nmean, *const *ndev;       // [tb] device pointers to the band nets' input norms
    // 1BT_DCT
    float *xm; int ldxm; const float *mmean, *mdev; const float *cos_tab;   // cos_tab [nd][L]
    int nd, add_c0, shift; float normc;
};

__device__ __forceinline__ int trap_find_utt(const int64_t *off, int n, int64_t f)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= f) lo = mid; else hi = mid - 1;
    }
    while (lo + 1 < n && off[lo + 1] <= f) ++lo;   // (empty utterances share an offset)
    return lo;
}

// thread = (frame, band); the L values of the trajectory are written one by one
__global__ void __launch_bounds__(256) k_trap_traj(TrapArgs a)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.nf * a.tb) return;
    const int64_t fl = idx / a.tb;
    const int b = (int)(idx - fl * a.tb);
    const int64_t f = a.f0 + fl;
    const int u = trap_find_utt(a.frame_off, a.n_utt, f);
    const int64_t u0 = a.frame_off[u];
    const int T = (int)(a.frame_off[u + 1] - u0), r = (int)(f - u0);
    const float mu = a.mean[(size_t)u * a.nb + b];
    const float *nm = a.nmean[b], *nd = a.ndev[b];
    float *o = a.xb + (size_t)b * a.band_stride + fl * a.ldx;
    for (int j = 0; j < a.L; ++j) {
        int t = r - a.S + j;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        float x = __fsub_rn(a.mel[(u0 + t) * a.nb + b], mu);
        if (a.ham) x = __fmul_rn(x, a.ham[j]);
        o[j] = __fmul_rn(__fsub_rn(x, nm[j]), nd[j]);
    }
    for (int j = a.L; j < a.ldx; ++j) o[j] = 0.0f;
}

// thread = (frame, band): C0 and the DCT of the band's trajectory, sequential sums in the reference's order
__global__ void __launch_bounds__(128) k_trap_dct(TrapArgs a)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.nf * a.tb) return;
    const int64_t fl = idx / a.tb;
    const int b = (int)(idx - fl * a.tb);
    const int64_t f = a.f0 + fl;
    const int u = trap_find_utt(a.frame_off, a.n_utt, f);
    const int64_t u0 = a.frame_off[u];
    const int T = (int)(a.frame_off[u + 1] - u0), r = (int)(f - u0);
    const float mu = a.mean[(size_t)u * a.nb + b];
    float x[256];
    for (int j = 0; j < a.L; ++j) {
        int t = r - a.S + j;
        t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
        float v = __fsub_rn(a.mel[(u0 + t) * a.nb + b], mu);
        if (a.ham) v = __fmul_rn(v, a.ham[j]);
        x[j] = v;
    }
    float *o = a.xm + fl * a.ldxm;
    int cidx = b * a.shift;
    if (a.add_c0) {   // CalcC0, dspc.h:223-233
        float s = 0.0f;
        for (int j = 0; j < a.L; ++j) s = __fadd_rn(s, x[j]);
        s = __fmul_rn(s, a.normc);
        o[cidx] = __fmul_rn(__fsub_rn(s, a.mmean[cidx]), a.mdev[cidx]);
        ++cidx;
    }
    for (int k = 0; k < a.nd; ++k, ++cidx) {   // sDCT, dspc.h:206-221
        const float *ct = a.cos_tab + (size_t)k * a.L;
        float s = 0.0f;
        for (int j = 0; j < a.L; ++j) s = __fadd_rn(s, __fmul_rn(x[j], ct[j]));
        s = __fmul_rn(s, a.normc);
        o[cidx] = __fmul_rn(__fsub_rn(s, a.mmean[cidx]), a.mdev[cidx]);
    }
}

// tables of the system: Hamming window (dspc.h:162-167 applied to ones), DCT cosines (dspc.h:206-221), both with the host libm
int trap_prepare(phn_ctx *c)
{
    if (c->trap_ready) return PHN_OK;
    const int L = c->trap_len;
    std::vector<float> ham((size_t)L), ct;
    for (int i = 0; i < L; ++i) ham[i] = 1.0f * (0.54f - 0.46f * cosf(2.0f * (float)M_PI * i / (L - 1)));
    PHN_CUDA(c, cudaMalloc(&c->d_trap_ham, sizeof(float) * L));
    PHN_CUDA(c, cudaMemcpy(c->d_trap_ham, ham.data(), sizeof(float) * L, cudaMemcpyHostToDevice));
    if (c->system == PHN_SYS_1BT_DCT) {
        const int nd = c->add_c0 ? c->trap_shift_out - 1 : c->trap_shift_out;
        ct.resize((size_t)(nd > 0 ? nd : 1) * L);
        const float PiByN = (float)M_PI / (float)L;
        for (int k = 0; k < nd; ++k) {
            const float v = PiByN * (float)(k + 1);
            for (int j = 0; j < L; ++j) ct[(size_t)k * L + j] = cosf(v * ((float)j + 0.5f));
        }
        PHN_CUDA(c, cudaMalloc(&c->d_trap_cos, sizeof(float) * ct.size()));
        PHN_CUDA(c, cudaMemcpy(c->d_trap_cos, ct.data(), sizeof(float) * ct.size(), cudaMemcpyHostToDevice));
    } else {
        std::vector<const float *> pm, pd;
        for (auto &d : c->dband) { pm.push_back(d.mean); pd.push_back(d.dev); }
        PHN_CUDA(c, cudaMalloc(&c->d_trap_pm, sizeof(float *) * pm.size()));
        PHN_CUDA(c, cudaMalloc(&c->d_trap_pd, sizeof(float *) * pd.size()));
        PHN_CUDA(c, cudaMemcpy(c->d_trap_pm, pm.data(), sizeof(float *) * pm.size(), cudaMemcpyHostToDevice));
        PHN_CUDA(c, cudaMemcpy(c->d_trap_pd, pd.data(), sizeof(float *) * pd.size(), cudaMemcpyHostToDevice));
    }
    c->trap_ready = 1;
    return PHN_OK;
}

// frames [f0, f0 + nf) -> the pass's band-net input matrices (1BT / 3BT) or the merger's input matrix (1BT_DCT)
int launch_trap(phn_ctx *c, int64_t f0, int64_t nf)
{
    if (nf == 0) return PHN_OK;
    int rc;
    if ((rc = trap_prepare(c))) return rc;
    TrapArgs a{};
    a.mel = (const float *)c->d_mel.p; a.mean = (const float *)c->d_mean.p; a.frame_off = (const int64_t *)c->d_frame_off.p;
    a.n_utt = c->n_utt; a.nb = c->nbanks; a.tb = c->trap_bands; a.L = c->trap_len; a.S = c->tshift;
    a.f0 = f0; a.nf = nf;
    a.ham = c->use_hamming ? (const float *)c->d_trap_ham : nullptr;
    const int64_t n = nf * a.tb;
    if (c->system == PHN_SYS_1BT_DCT) {
        a.xm = (float *)c->d_xm.p; a.ldxm = c->net[2].kp; a.mmean = c->net[2].mean; a.mdev = c->net[2].dev;
        a.cos_tab = (const float *)c->d_trap_cos;
        a.shift = c->trap_shift_out; a.add_c0 = c->add_c0; a.nd = c->add_c0 ? a.shift - 1 : a.shift;
        a.normc = sqrtf(2.0f / (float)a.L);
        PHN_CUDA(c, cudaMemsetAsync(a.xm, 0, sizeof(float) * (size_t)nf * a.ldxm, c->stream));   // (padding columns)
        k_trap_dct<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(a);
    } else {
        a.xb = (float *)c->d_xb.p; a.ldx = c->dband[0].kp; a.band_stride = c->chunk_frames * a.ldx;
        a.nmean = (const float *const *)c->d_trap_pm; a.ndev = (const float *const *)c->d_trap_pd;
        k_trap_traj<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a);
    }
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_STC] += 1;
    return PHN_OK;
}

}  // namespace phn
