// k_plp.cu — K-plp: PLP coefficients (params/kind = plp), what `-t par` saves for that parameterisation.
//
// Replaces PLPCoefs::ProcessFrame (plp.cpp:91-141) behind MelBanks::ProcessFrame with the logarithm switched off: floor the bank
// energies at 1, equal-loudness curve, power-law compression, duplicated edge values, IDFT to autocorrelation coefficients,
// Durbin's recursion (dspc.cpp:275-308), LPC -> cepstrum (dspc.cpp:310-324), C0 = -logf(1 / gain), liftering, scaling, then
// FrameBasedNormalization (srec.cpp:1594-1620).  One thread per frame, one fp32 rounding per reference operation in the
// reference's order; the tables (equal-loudness curve, IDFT matrix, liftering window) come from the host libm like the
// reference's.  The one operation that is NOT the reference's bits is powf: the device evaluates pow in double and rounds once,
// glibc's powf is accurate to 0.8 ulp, so a coefficient may differ in its last bits (measured bound in
// tests/test_gpu_plp.py); logf is the glibc port of device_math.cuh.
// The reference compiles PLP out of its PHNREC_ONLY build (srec.cpp:563-583): the phnrec binary itself rejects kind = plp.  Here it
// is a parameterisation for `-t par` / phn_mel only - the TRAPS nets take mel-bank energies.
#include "internal.h"
#include "device_math.cuh"

namespace phn {

struct PlpArgs {
    const float *energy;   // [F][nb] bank energies (K-wave, raw_energy)
    float *out;            // [F][nparams]
    int64_t frames;
    int nb, order, add_c0, nparams;
    float compress, lifter, scale, frame_shift, frame_floor;
    const float *eql, *idft, *lift;   // [nb], [order + 1][nb + 2], [order]
};

__global__ void __launch_bounds__(128) k_plp(PlpArgs a)
{
    __shared__ double s_logtab[32];
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.frames) return;
    const int nb = a.nb, dim = nb + 2, P = a.order;
    float en[34], ac[33], lp[32], tmp[32], cep[33];
    for (int i = 0; i < nb; ++i) {
        float v = a.energy[f * nb + i];
        if (v < 1.0f) v = 1.0f;
        v = __fmul_rn(v, a.eql[i]);
        en[i + 1] = (float)pow((double)v, (double)a.compress);
    }
    en[0] = en[1];
    en[nb + 1] = en[nb];
    for (int i = 0; i <= P; ++i) {
        float s = 0.0f;
        for (int j = 0; j < dim; ++j) s = __fadd_rn(s, __fmul_rn(en[j], a.idft[i * dim + j]));
        ac[i] = s;
    }
    float E = ac[0];
    for (int i = 0; i < P; ++i) {
        float ki = ac[i + 1];
        for (int j = 0; j < i; ++j) ki = __fadd_rn(ki, __fmul_rn(lp[j], ac[i - j]));
        ki = __fdiv_rn(ki, E);
        E = __fmul_rn(E, __fsub_rn(1.0f, __fmul_rn(ki, ki)));
        tmp[i] = -ki;
        for (int j = 0; j < i; ++j) tmp[j] = __fsub_rn(lp[j], __fmul_rn(ki, lp[i - j - 1]));
        for (int j = 0; j <= i; ++j) lp[j] = tmp[j];
    }
    for (int i = 0; i < P; ++i) {
        float sum = 0.0f;
        for (int j = 0; j < i; ++j) sum = __fadd_rn(sum, __fmul_rn(__fmul_rn((float)(i - j), lp[j]), cep[i - j - 1]));
        cep[i] = __fsub_rn(-lp[i], __fdiv_rn(sum, (float)(i + 1)));
    }
    cep[P] = -logf_glibc(__fdiv_rn(1.0f, E), s_logtab);
    if (a.lifter != 0.0f) for (int i = 0; i < P; ++i) cep[i] = __fmul_rn(cep[i], a.lift[i]);
    if (a.scale != 1.0f) for (int i = 0; i <= P; ++i) cep[i] = __fmul_rn(cep[i], a.scale);
    for (int i = 0; i < a.nparams; ++i) {
        float o = cep[i];
        if (a.frame_shift != 0.0f) o = __fadd_rn(o, a.frame_shift);
        if (a.frame_floor != -9999.9f && o < a.frame_floor) o = a.frame_floor;
        a.out[f * a.nparams + i] = o;
    }
}

int launch_plp(phn_ctx *c)
{
    const int64_t F = c->total_frames;
    int rc;
    if ((rc = ensure(c, c->d_par, sizeof(float) * (size_t)(F + 1) * c->nparams))) return rc;
    if (F == 0) return PHN_OK;
    PlpArgs a{(const float *)c->d_mel.p, (float *)c->d_par.p, F, c->nbanks, c->plp_order, c->plp_add_c0, c->nparams,
              c->plp_compress, c->plp_lifter, c->plp_scale, c->frame_shift, c->frame_floor, c->d_plp_eql, c->d_plp_idft, c->d_plp_lift};
    k_plp<<<(unsigned)((F + 127) / 128), 128, 0, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_WAVE] += 1;
    return PHN_OK;
}

}  // namespace phn
