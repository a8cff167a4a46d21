// device_math.cuh — bit-faithful device restatements of the scalar functions the reference's
// results depend on.  Every operation is an explicitly rounded IEEE op (__f*_rn / __d*_rn) so that
// nvcc can neither contract a*b+c into an FMA nor reassociate: the reference binary is x86-64
// SSE2 code without FMA, one rounding per written operation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phn {

// A-law byte -> ALawTableD5 value (alaw.cpp:14-48): G.711 expansion >> 3.
__device__ __forceinline__ int alaw_d5(unsigned a)
{
    a ^= 0x55u;
    int t = (int)(a & 0x0fu) << 4;
    const int seg = (int)(a & 0x70u) >> 4;
    t = seg == 0 ? t + 8 : (t + 0x108) << (seg - 1);
    t >>= 3;
    return (a & 0x80u) ? t : -t;
}

// glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c): 16-entry {1/c, log c} table,
// cubic in r = z/c - 1, all in double, one final rounding to float.  The reference's SoftLog
// (srec.h:192-195) and sLn (dspc.h:155-160) are calls to this libm function.
__constant__ double kLogfTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

// The table is indexed per lane; __constant__ memory serialises divergent indices, so kernels on a
// hot path copy it to shared memory once (logf_table_to_smem) and pass that pointer instead.
__device__ __forceinline__ void logf_table_to_smem(double *dst /*[32]*/, int tid, int nthreads)
{
    for (int i = tid; i < 32; i += nthreads) dst[i] = kLogfTab[i >> 1][i & 1];
}

__device__ __forceinline__ float logf_glibc(float x, const double *tab /* [16][2] */)
{
    const double Ln2 = 0x1.62e42fefa39efp-1;
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return __uint_as_float(0xff800000u);              // log(+-0) = -inf
        if (ix == 0x7f800000u) return x;                                    // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __uint_as_float(0x7fc00000u);  // NaN
        ix = __float_as_uint(__fmul_rn(x, 8388608.0f));                     // subnormal: scale by 2^23
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int32_t)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double r = __dsub_rn(__dmul_rn(z, tab[2 * i]), 1.0);
    const double y0 = __dadd_rn(tab[2 * i + 1], __dmul_rn((double)k, Ln2));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// sLn (dspc.h:155-160): guarded log, digital silence maps to 0.0 (not -inf)
__device__ __forceinline__ float ln_guarded(float v, const double *tab) { return v > 0.0f ? logf_glibc(v, tab) : 0.0f; }

// Canonical Quicknet/Schraudolph exp (fexp.h:14-21, low word := 0): a double whose high word is
// trunc(2^20/ln2 * y) + (1023*2^20 - 60801).
__device__ __forceinline__ double fexp_canonical(double y)
{
    const double A = 1048576 / 0.69314718055994530942;
    const int hi = __double2int_rz(__dmul_rn(A, y)) + (1072693248 - 60801);
    return __hiloint2double(hi, 0);
}

// fexp_sigmoid (fexp.h:33-38): 1.0f/(1.0f + D(-x)), evaluated in double, one rounding to float
__device__ __forceinline__ float fsigmoid_exact(float x)
{
    const double d = fexp_canonical((double)(-x));
    return __double2float_rn(__ddiv_rn(1.0, __dadd_rn(1.0, d)));
}

// Fast fp32 form of the same function for the tensor-core mode (hidden activations are rounded to
// fp16 right after): D(-x) built with integer ops, 1+D and the reciprocal in fp32.
__device__ __forceinline__ float fsigmoid_fast(float x)
{
    float y = fminf(fmaxf(-x, -87.0f), 87.0f);
    // hi = trunc(A*y) + C ; the fp32 product may differ from the double one by 1 unit of hi
    // (2^-20 relative in D) for |y| > 11 - far below fp16 resolution.
    const int hi = __float2int_rz(y * 1512775.395f) + (1072693248 - 60801);
    // repack the double's high word (1+11+20 bits) as a float (1+8+23 bits)
    const int e = (hi >> 20) - 1023 + 127;
    const float d = __int_as_float((e << 23) | ((hi & 0xFFFFF) << 3));
    return __frcp_rn(1.0f + d);
}

}  // namespace phn
