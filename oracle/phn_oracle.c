/*
 * phn_oracle.c — CPU restatement of the PhnRec recognition hot path.
 *
 * >>> TEST INFRASTRUCTURE ONLY. <<<
 * This file is the parity CHECKER for the CUDA path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load libphn_oracle.so. Nothing in phnrec_b200/ links or calls it; the
 * product fails loudly when its CUDA library is missing.
 *
 * It is a plain-C restatement (own structure, no code copied) of the
 * reference's single-threaded algorithm; each function cites the reference
 * file:line it follows (paths relative to the reference checkout).
 *
 * Pinning: tests/test_oracle_pinned.py checks this restatement
 *   (a) bit-for-bit against outputs of the reference's own sources compiled
 *       by oracle/Makefile (oracle/_ref/phnrec_ref: mel `-t par`, posteriors
 *       `-t post`, `.rec` labels incl. the %f scores), and
 *   (b) against the 7 golden label files the reference ships
 *       (tests/golden/ref_labels.json).
 *
 * Floating point: compiled with -O2 -ffp-contract=off so every + and * below
 * is a separately rounded IEEE operation, in the order written — that is what
 * the reference's g++ -O2 x86-64 (SSE2, no FMA) build executes.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif
#ifndef M_LN2
#define M_LN2 0.69314718055994530942
#endif

#define ORC_MB_VECTORSIZE 200 /* config.h:21 — the 8 kHz constant used for the short-signal pad */

/* ------------------------------------------------------------------------- */
/* W2: waveform decode.  srec.cpp:709-791, alaw.cpp:14-48                    */
/* ------------------------------------------------------------------------- */

/* G.711 A-law expansion divided by 8 — reproduces ALawTableD5[256]
 * (alaw.cpp:14-48); the table is regenerated, not copied. */
static short alaw_d5(unsigned a)
{
    a ^= 0x55u;
    int t = (int)(a & 0x0fu) << 4;
    int seg = (int)(a & 0x70u) >> 4;
    if (seg == 0)
        t += 8;
    else
        t = (t + 0x108) << (seg - 1);
    t >>= 3;
    return (short)((a & 0x80u) ? t : -t);
}

void orc_alaw_table(short *out256)
{
    for (int i = 0; i < 256; ++i) out256[i] = alaw_d5((unsigned)i);
}

/* fmt: 0 = lin16 (host-endian short), 1 = alaw.  Returns the sample count.
 * `out` must hold max(n, 200) floats: the first MB_VECTORSIZE entries are
 * zeroed first so that a too-short signal still gives one (zero padded) frame
 * (srec.cpp:739-740, 766-767). */
int orc_wave_to_float(int fmt, const void *sig, int nbytes, float scale, float dc_shift, float *out)
{
    int n = (fmt == 0) ? nbytes / 2 : nbytes;
    for (int i = 0; i < ORC_MB_VECTORSIZE; ++i) out[i] = 0.0f;
    if (fmt == 0) {
        const short *s = (const short *)sig;
        for (int i = 0; i < n; ++i) out[i] = (float)s[i]; /* srec.cpp:742-743 */
    } else {
        const unsigned char *s = (const unsigned char *)sig;
        for (int i = 0; i < n; ++i) out[i] = 8.0f * (float)alaw_d5(s[i]); /* srec.cpp:768-769 */
    }
    if (dc_shift != 0.0f) /* srec.cpp:780-781 */
        for (int i = 0; i < n; ++i) out[i] = out[i] + dc_shift;
    if (scale != 1.0f) /* srec.cpp:784-785 */
        for (int i = 0; i < n; ++i) out[i] = out[i] * scale;
    return n;
}

/* ------------------------------------------------------------------------- */
/* M1/M2: mel-bank front end.  melbanks.cpp:111-204, dspc.cpp:24-269          */
/* ------------------------------------------------------------------------- */

typedef struct {
    int nbanks, vs, step, fs, N, N2;
    float preem;
    int z_mean;
    int fftlo, ffthi;
    float *hamming; /* [vs]  dspc.h:162-167 */
    float *coeffs;  /* [N2]  dspc.cpp:200-217 */
    short *banks;   /* [N2]  dspc.cpp:183-197 */
    float *fft;     /* [2N+1] NR 1-based work array, melbanks.cpp:62 */
    float *frame;   /* [vs] */
    float *pw;      /* [N2] */
    float *f0;      /* [nbanks + 1] centres of the banks in Hz (dspc.cpp:156-162), used by PLP */
    /* PLP (plp.cpp; params/kind = plp): order 0 = plain mel-banks */
    int plp_order, plp_add_c0;
    float plp_compress, plp_lifter, plp_scale;
    float *plp_eql, *plp_idft, *plp_lift;   /* [nbanks], [order+1][nbanks+2], [order] */
} orc_mel;

static float scale_mel(float f) { return 1127.0f * logf(1.0f + f / 700.0f); } /* dspc.h:174-177 */

orc_mel *orc_mel_create(int nbanks, int vs, int step, int fs, float lo, float hi, float preem, int z_mean)
{
    orc_mel *m = (orc_mel *)calloc(1, sizeof(orc_mel));
    m->nbanks = nbanks; m->vs = vs; m->step = step; m->fs = fs;
    m->preem = preem; m->z_mean = z_mean;
    int N = 1;
    while (N < vs) N *= 2; /* melbanks.cpp:44-46 */
    m->N = N; m->N2 = N / 2;
    m->hamming = (float *)malloc(sizeof(float) * vs);
    for (int i = 0; i < vs; ++i) /* dspc.h:162-167 applied to a vector of ones */
        m->hamming[i] = 1.0f * (0.54f - 0.46f * cosf(2.0f * (float)M_PI * i / (vs - 1)));
    m->coeffs = (float *)calloc(m->N2, sizeof(float));
    m->banks = (short *)malloc(sizeof(short) * m->N2);
    m->fft = (float *)malloc(sizeof(float) * (2 * N + 1));
    m->frame = (float *)malloc(sizeof(float) * vs);
    m->pw = (float *)malloc(sizeof(float) * m->N2);

    /* _mbInit, dspc.cpp:80-225 */
    if (lo < 0.0f) lo = 0.0f;
    if (hi > (float)fs / 2.0f) hi = (float)fs / 2.0f;
    float *f0m = (float *)malloc(sizeof(float) * (nbanks + 2));
    f0m[nbanks + 1] = INFINITY; /* the reference's loop may peek one past the end (dspc.cpp:175) */
    float bf = (float)fs / (float)N;
    float mlo = scale_mel(lo), mhi = scale_mel(hi);
    int fftlo = (int)(lo / bf + 1.5f);
    int ffthi = (int)(hi / bf - 0.5f);
    if (fftlo < 1) fftlo = 1;
    if (ffthi >= m->N2) ffthi = m->N2 - 1;
    m->fftlo = fftlo; m->ffthi = ffthi;
    float delta = (mhi - mlo) / (nbanks + 1);
    float mf = mlo;
    m->f0 = (float *)malloc(sizeof(float) * (nbanks + 1));
    for (int i = 0; i <= nbanks; ++i) { /* dspc.cpp:156-162 (repeated float addition) */
        mf = mf + delta;
        f0m[i] = mf;
        m->f0[i] = 700.0f * (expf(mf / 1127.0f) - 1.0f); /* Scale_MelToLinear, dspc.h:169-172 */
    }
    int ch = 0;
    for (int i = 0; i < m->N2; ++i) { /* dspc.cpp:165-180 */
        if (i < fftlo || i > ffthi) {
            m->banks[i] = -1;
        } else {
            float x = scale_mel((float)i * bf);
            while (x > f0m[ch] && ch <= nbanks) ++ch;
            m->banks[i] = (short)ch;
        }
    }
    for (int i = 0; i < m->N2; ++i) { /* dspc.cpp:183-199 */
        int c = m->banks[i];
        if (i < fftlo || i > ffthi)
            m->coeffs[i] = 0;
        else if (c == 0)
            m->coeffs[i] = (f0m[0] - scale_mel((float)i * bf)) / (f0m[0] - mlo);
        else
            m->coeffs[i] = (f0m[c] - scale_mel((float)i * bf)) / (f0m[c] - f0m[c - 1]);
    }
    free(f0m);
    return m;
}

void orc_mel_destroy(orc_mel *m)
{
    if (!m) return;
    free(m->hamming); free(m->coeffs); free(m->banks); free(m->fft); free(m->frame); free(m->pw);
    free(m);
}

int orc_mel_fft_size(const orc_mel *m) { return m->N; }
void orc_mel_tables(const orc_mel *m, float *hamming, float *coeffs, short *banks, int *fftlo, int *ffthi)
{
    memcpy(hamming, m->hamming, sizeof(float) * m->vs);
    memcpy(coeffs, m->coeffs, sizeof(float) * m->N2);
    memcpy(banks, m->banks, sizeof(short) * m->N2);
    *fftlo = m->fftlo; *ffthi = m->ffthi;
}

/* srec.cpp:945 */
int orc_num_frames(int len, int vs, int step) { return len > vs ? (len - vs) / step + 1 : 1; }

/* Numerical-Recipes radix-2 complex FFT, 1-based interleaved array, forward
 * transform (isign = -1), float data with a double twiddle recurrence.
 * dspc.cpp:24-78. */
static void four1_forward(float *data, unsigned nn)
{
    unsigned n = nn << 1, j = 1;
    for (unsigned i = 1; i < n; i += 2) {
        if (j > i) {
            float t = data[j]; data[j] = data[i]; data[i] = t;
            t = data[j + 1]; data[j + 1] = data[i + 1]; data[i + 1] = t;
        }
        unsigned mm = n >> 1;
        while (mm >= 2 && j > mm) { j -= mm; mm >>= 1; }
        j += mm;
    }
    unsigned mmax = 2;
    while (n > mmax) {
        unsigned istep = mmax << 1;
        double theta = -1 * (6.28318530717959 / mmax);
        double wtemp = sin(0.5 * theta);
        double wpr = -2.0 * wtemp * wtemp;
        double wpi = sin(theta);
        double wr = 1.0, wi = 0.0;
        for (unsigned mi = 1; mi < mmax; mi += 2) {
            for (unsigned i = mi; i <= n; i += istep) {
                unsigned k = i + mmax;
                /* float = (float)(double*float - double*float): computed in
                 * double, rounded once on assignment (dspc.cpp:62-63) */
                float tempr = (float)(wr * data[k] - wi * data[k + 1]);
                float tempi = (float)(wr * data[k + 1] + wi * data[k]);
                data[k] = data[i] - tempr;
                data[k + 1] = data[i + 1] - tempi;
                data[i] += tempr;
                data[i + 1] += tempi;
            }
            wtemp = wr;
            wr = wtemp * wpr - wi * wpi + wr;
            wi = wi * wpr + wtemp * wpi + wi;
        }
        mmax = istep;
    }
}

/* One frame: melbanks.cpp:111-149 */
static void mel_frame(orc_mel *m, float *fr, float *out)
{
    int vs = m->vs, N = m->N, N2 = m->N2, nb = m->nbanks;
    if (m->z_mean) { /* dspc.h:63-75 */
        float avg = 0.0f;
        for (int i = 0; i < vs; ++i) avg += fr[i];
        avg /= (float)vs;
        for (int i = 0; i < vs; ++i) fr[i] -= avg;
    }
    if (m->preem != 0.0f) { /* dspc.h:77-84 */
        for (int i = vs - 1; i > 0; --i) fr[i] -= m->preem * fr[i - 1];
        fr[0] *= (1.0f - m->preem);
    }
    for (int i = 0; i < vs; ++i) fr[i] *= m->hamming[i];
    float *d = m->fft;
    d[0] = 0.0f;
    for (int i = 0; i < N; ++i) {
        d[1 + 2 * i] = i < vs ? fr[i] : 0.0f;
        d[2 + 2 * i] = 0.0f;
    }
    four1_forward(d, (unsigned)N);
    for (int k = 0; k < N2; ++k) { /* dspc.h:141-146 */
        float re = d[1 + 2 * k], im = d[2 + 2 * k];
        m->pw[k] = (re * re + im * im);
    }
    for (int b = 0; b < nb; ++b) out[b] = 0.0f;
    for (int k = m->fftlo; k <= m->ffthi; ++k) { /* dspc.cpp:236-269 */
        float v = m->coeffs[k] * m->pw[k];
        int B = m->banks[k];
        if (B > 0) out[B - 1] += v;
        if (B < nb) out[B] += (m->pw[k] - v);
    }
    if (m->plp_order > 0) return; /* PLPCoefs switches the logarithm off (plp.cpp:70, SetTakeLog(false)) */
    for (int b = 0; b < nb; ++b) out[b] = (out[b] > 0.0f ? logf(out[b]) : 0.0f); /* dspc.h:155-160 */
}

/* PLP coefficients (params/kind = plp).  PLPCoefs::Init / ProcessFrame (plp.cpp:38-141), CreateIDFTMatrix (plp.cpp:143-165),
 * sEqualLaudnessCurve / sPower / sLowerFloor (dspc.h:235-265), sDurbin / sLPC2Cepstrum / sLifteringWindow (dspc.cpp:275-335).
 * Compiled out of the reference's PHNREC_ONLY build (srec.cpp:563-583); pinned against the class itself through
 * oracle/_ref/online_ref plp (tests/golden/ref_plp.npz). */
void orc_mel_set_plp(orc_mel *m, int order, float compress, float lifter, float scale, int add_c0)
{
    const int nb = m->nbanks, dim = nb + 2;
    m->plp_order = order; m->plp_compress = compress; m->plp_lifter = lifter; m->plp_scale = scale; m->plp_add_c0 = add_c0;
    m->plp_eql = (float *)malloc(sizeof(float) * nb);
    for (int i = 0; i < nb; ++i) {
        float fsq = (m->f0[i] * m->f0[i]);
        float fsub = fsq / (fsq + 1.6e5f);
        m->plp_eql[i] = fsub * fsub * ((fsq + 1.44e6f) / (fsq + 9.61e6f));
    }
    m->plp_idft = (float *)malloc(sizeof(float) * (order + 1) * dim);
    float angle = M_PI / (float)(dim - 1);
    float scl = 1.0f / (2.0f * (dim - 1));
    for (int i = 0; i <= order; ++i) {
        float *row = m->plp_idft + (size_t)i * dim;
        row[0] = 1.0f * scl;
        /* C++: cos(float) is the float overload, i.e. cosf; `2.0 * scale * ...` is then evaluated in double and rounded once */
        for (int j = 1; j < dim - 1; ++j) row[j] = (float)(2.0 * scl * cosf(angle * (float)i * (float)j));
        row[dim - 1] = scl * cosf(angle * (float)i * (float)(dim - 1));
    }
    m->plp_lift = (float *)malloc(sizeof(float) * (order > 0 ? order : 1));
    int Q = (int)lifter;
    for (int i = 0; i < order; ++i) m->plp_lift[i] = 1.0f + 0.5f * Q * sinf(M_PI * (float)(i + 1) / (float)Q);
}

int orc_mel_nparams(const orc_mel *m) { return m->plp_order > 0 ? m->plp_order + (m->plp_add_c0 ? 1 : 0) : m->nbanks; }

/* energies [nbanks] (no logarithm) -> coefficients [nparams] */
static void plp_frame(const orc_mel *m, const float *en_in, float *out)
{
    const int nb = m->nbanks, dim = nb + 2, P = m->plp_order;
    float en[64], ac[64], lp[64], tmp[64], cep[64];
    for (int i = 0; i < nb; ++i) {
        float v = en_in[i];
        if (v < 1.0f) v = 1.0f;                 /* sLowerFloor(.., 1.0f) */
        v = v * m->plp_eql[i];                  /* equal loudness */
        en[i + 1] = powf(v, m->plp_compress);   /* sPower; stored shifted right by one */
    }
    en[0] = en[1];                              /* duplicate the first and the last value */
    en[nb + 1] = en[nb];
    for (int i = 0; i <= P; ++i) {              /* ApplyLinTransform */
        float s = 0.0f;
        for (int j = 0; j < dim; ++j) s += en[j] * m->plp_idft[(size_t)i * dim + j];
        ac[i] = s;
    }
    float E = ac[0];                            /* sDurbin */
    for (int i = 0; i < P; ++i) {
        float ki = ac[i + 1];
        for (int j = 0; j < i; ++j) ki = ki + lp[j] * ac[i - j];
        ki = ki / E;
        E *= 1 - ki * ki;
        tmp[i] = -ki;
        for (int j = 0; j < i; ++j) tmp[j] = lp[j] - ki * lp[i - j - 1];
        for (int j = 0; j <= i; ++j) lp[j] = tmp[j];
    }
    for (int i = 0; i < P; ++i) {               /* sLPC2Cepstrum */
        float sum = 0.0f;
        for (int j = 0; j < i; ++j) sum += (float)(i - j) * lp[j] * cep[i - j - 1];
        cep[i] = -lp[i] - sum / (float)(i + 1);
    }
    cep[P] = -logf(1.0f / E);                   /* C0 */
    if (m->plp_lifter != 0.0f) for (int i = 0; i < P; ++i) cep[i] *= m->plp_lift[i];
    if (m->plp_scale != 1.0f) for (int i = 0; i <= P; ++i) cep[i] *= m->plp_scale;
    for (int i = 0; i < (m->plp_add_c0 ? P + 1 : P); ++i) out[i] = cep[i];
}

/* Whole utterance (framing of melbanks.cpp:151-204 + srec.cpp:945,965-971):
 * frame t = samples [t*step, t*step+vs); tail samples dropped.  `wav` must be
 * readable for max(len, vs) floats (zero padded when len < vs).
 * frame_shift/frame_floor: srec.cpp:1594-1620 (disabled when 0 / -9999.9). */
int orc_mel_compute(orc_mel *m, const float *wav, int len, float frame_shift, float frame_floor, float *mel)
{
    int T = orc_num_frames(len, m->vs, m->step);
    for (int t = 0; t < T; ++t) {
        for (int i = 0; i < m->vs; ++i) {
            int idx = t * m->step + i;
            m->frame[i] = (idx < len || idx < ORC_MB_VECTORSIZE) ? wav[idx] : 0.0f;
        }
        const int np = orc_mel_nparams(m);
        float *o = mel + (size_t)t * np;
        if (m->plp_order > 0) {
            float en[64];
            mel_frame(m, m->frame, en);
            plp_frame(m, en, o);
        } else {
            mel_frame(m, m->frame, o);
        }
        if (frame_shift != 0.0f)
            for (int b = 0; b < np; ++b) o[b] += frame_shift;
        if (frame_floor != -9999.9f)
            for (int b = 0; b < np; ++b)
                if (o[b] < frame_floor) o[b] = frame_floor;
    }
    return T;
}

/* ------------------------------------------------------------------------- */
/* N1: sentence mean normalisation.  srec.cpp:1492-1516, matrix.h:2101-2116   */
/* ------------------------------------------------------------------------- */
void orc_sentence_mean_norm(float *mel, int T, int nb)
{
    for (int b = 0; b < nb; ++b) {
        float sum = 0.0f;
        for (int t = 0; t < T; ++t) sum += mel[(size_t)t * nb + b];
        /* Mat::div(v) is mul(1.0f/v) (matrix.h:245): multiply by the reciprocal */
        float mean = sum * (1.0f / (float)T);
        for (int t = 0; t < T; ++t) mel[(size_t)t * nb + b] -= mean;
    }
}

/* N2: online normaliser arithmetic (norm.cpp:92-148, 216-234): estimate
 * mean / inverse std over the first `interval` frames, apply to all frames
 * from the moment the estimate exists.  Restated for completeness; the
 * offline hot path never calls it (srec.cpp:806 is ProcessOnline only). */
void orc_online_norm(float *x, int T, int nb, int interval, int mean_norm, int var_norm)
{
    /* ProcessFrame (norm.cpp:216-234) per frame: Accum while fewer than `interval` frames were seen
     * (norm.cpp:92-110); Update() the moment the count reaches `interval` (norm.cpp:139-148) - BEFORE Norm(), so the
     * frame that completes the estimate (index interval-1) is already normalised with it; Norm() on every frame
     * (norm.cpp:112-137): `if(mean) x -= mean; if(var) x *= invstd;`.  Until the estimate exists the parameters are
     * Null()'s mean 0 / inverse std 1 (norm.cpp:59-70), i.e. the earlier frames pass through unchanged.
     * SetNorm asserts that var_norm implies mean_norm (norm.cpp:150-155); callers reject that combination. */
    if (interval <= 0 || T < interval) return;
    for (int b = 0; b < nb; ++b) {
        float s = 0.0f, s2 = 0.0f;
        for (int t = 0; t < interval; ++t) {
            float v = x[(size_t)t * nb + b];
            s += v;
            s2 += v * v;
        }
        float mean = s / (float)interval;
        float inv = 1.0f / sqrtf(s2 / (float)interval - mean * mean);
        for (int t = interval - 1; t < T; ++t) {
            float v = x[(size_t)t * nb + b];
            if (mean_norm) v -= mean;
            if (var_norm) v *= inv;
            x[(size_t)t * nb + b] = v;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* S1/S2: split temporal context features.  traps.cpp:180-219, 285-342,       */
/* dspc.h:206-233; driver srec.cpp:1035-1059                                  */
/* ------------------------------------------------------------------------- */
/* Net effect of the 31-slot FIFO + warm-up + tail replication: output row r
 * sees mel frames clamp(r-15 .. r+15, 0, T-1).  The FIFO
 * itself is restated in orc_stc_fifo() below and the two are cross-checked in
 * tests. win = 32 floats: windows/band0.window then windows/band1.window. */
void orc_stc(const float *mel, int T, int nb, const float *win, int ncoef /*11*/, float *XL, float *XR)
{
    const int H = 16;
    float NormC = sqrtf(2.0f / (float)H);
    float PiByN = (float)M_PI / (float)H;
    float lc[16], rc[16];
    int nOut = ncoef - 1;
    for (int r = 0; r < T; ++r) {
        for (int b = 0; b < nb; ++b) {
            for (int j = 0; j < H; ++j) {
                int tl = r - 15 + j, tr = r + j;
                if (tl < 0) tl = 0;
                if (tl > T - 1) tl = T - 1;
                if (tr > T - 1) tr = T - 1;
                lc[j] = mel[(size_t)tl * nb + b] * win[j];
                rc[j] = mel[(size_t)tr * nb + b] * win[H + j];
            }
            float *ol = XL + ((size_t)r * nb + b) * ncoef;
            float *orr = XR + ((size_t)r * nb + b) * ncoef;
            float s = 0.0f;
            for (int j = 0; j < H; ++j) s += lc[j];
            s *= NormC;
            ol[0] = s; /* CalcC0 dspc.h:223-233 */
            s = 0.0f;
            for (int j = 0; j < H; ++j) s += rc[j];
            s *= NormC;
            orr[0] = s;
            for (int k = 0; k < nOut; ++k) { /* sDCT dspc.h:206-221 */
                float v = PiByN * (float)(k + 1);
                float a = 0, c = 0;
                for (int j = 0; j < H; ++j) {
                    float cs = cosf(v * ((float)j + 0.5f));
                    a += lc[j] * cs;
                    c += rc[j] * cs;
                }
                a *= NormC;
                c *= NormC;
                ol[1 + k] = a;
                orr[1 + k] = c;
            }
        }
    }
}

/* The cos table the DCT uses, for upload by tests that want to look at it. */
void orc_dct_table(float *tab /* [10][16] */)
{
    float PiByN = (float)M_PI / 16.0f;
    for (int k = 0; k < 10; ++k) {
        float v = PiByN * (float)(k + 1);
        for (int j = 0; j < 16; ++j) tab[k * 16 + j] = cosf(v * ((float)j + 0.5f));
    }
}

/* Literal FIFO restatement (traps.cpp:180-219 + srec.cpp:1035-1059) producing
 * the 31-frame context of every output row; used only to validate the clamp
 * formula above.  ctx: [T][nb][31]. */
void orc_stc_fifo(const float *mel, int T, int nb, float *ctx)
{
    const int L = 31, shift = 15;
    float *be = (float *)malloc(sizeof(float) * (nb * L + 1));
    int init = 0, outrow = 0;
    /* sequence of fed frames: all T frames (padded to 15 with the last one if
     * T < 15), then the last frame min(T,15) more times; the first 15 feeds
     * produce no output. */
    int nfeed_first = T >= shift ? T : shift;
    int ntail = T > shift ? shift : T;
    int total = nfeed_first + ntail;
    for (int f = 0; f < total; ++f) {
        int src = f < T ? f : T - 1;
        const float *e = mel + (size_t)src * nb;
        if (!init) {
            init = 1;
            for (int i = 0; i < nb; ++i)
                for (int j = 0; j < L; ++j) be[i * L + j] = e[i];
        } else {
            for (int i = 0; i < nb * L; ++i) be[i] = (i + 1 < nb * L) ? be[i + 1] : 0.0f;
            for (int i = 0; i < nb; ++i) be[i * L + L - 1] = e[i];
        }
        if (f >= shift && outrow < T) {
            memcpy(ctx + (size_t)outrow * nb * L, be, sizeof(float) * nb * L);
            ++outrow;
        }
    }
    free(be);
}

/* ------------------------------------------------------------------------- */
/* N0/N3/N4/N5: the MLP.  nn.cpp:464-531, 633-682, 702-950, fexp.h            */
/* ------------------------------------------------------------------------- */
typedef struct {
    int nin, nhid, nout, nin4, nhid4, nout4;
    float *w1; /* [nhid4][nin4]  */
    float *w2; /* [nout4][nhid4] */
    float *b1, *b2, *mean, *dev;
} orc_nn;

static int align4(int n) { return (n + 3) & ~3; } /* Align16 in bytes, nn.cpp:633-640 */

orc_nn *orc_nn_load(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4 || hdr[0] != 2) { fclose(f); return NULL; }
    orc_nn *n = (orc_nn *)calloc(1, sizeof(orc_nn));
    n->nin = hdr[1]; n->nhid = hdr[2]; n->nout = hdr[3];
    n->nin4 = align4(n->nin); n->nhid4 = align4(n->nhid); n->nout4 = align4(n->nout);
    size_t s1 = (size_t)n->nhid4 * n->nin4, s2 = (size_t)n->nout4 * n->nhid4;
    n->w1 = (float *)malloc(4 * s1); n->w2 = (float *)malloc(4 * s2);
    n->b1 = (float *)malloc(4 * n->nhid4); n->b2 = (float *)malloc(4 * n->nout4);
    n->mean = (float *)malloc(4 * n->nin4); n->dev = (float *)malloc(4 * n->nin4);
    int ok = fread(n->w1, 4, s1, f) == s1 && fread(n->w2, 4, s2, f) == s2 &&
             fread(n->b1, 4, n->nhid4, f) == (size_t)n->nhid4 && fread(n->b2, 4, n->nout4, f) == (size_t)n->nout4 &&
             fread(n->mean, 4, n->nin4, f) == (size_t)n->nin4 && fread(n->dev, 4, n->nin4, f) == (size_t)n->nin4;
    fclose(f);
    if (!ok) { free(n); return NULL; }
    return n;
}

void orc_nn_destroy(orc_nn *n)
{
    if (!n) return;
    free(n->w1); free(n->w2); free(n->b1); free(n->b2); free(n->mean); free(n->dev); free(n);
}
void orc_nn_dims(const orc_nn *n, int *d) { d[0] = n->nin; d[1] = n->nhid; d[2] = n->nout; }

/* Canonical Schraudolph/Quicknet exp: fexp.h:14-21 with the low word := 0. */
static double fexp_d(double y)
{
    const double A = 1048576 / M_LN2;
    union { double d; struct { int32_t j, i; } n; } u;
    u.n.j = 0;
    u.n.i = (int)(A * y) + (1072693248 - 60801);
    return u.d;
}
double orc_fexp(double y) { return fexp_d(y); }
float orc_fsigmoid(float x) { return 1.0f / (1.0f + fexp_d(-x)); } /* fexp.h:33-38 (double arithmetic) */

void orc_fsoftmax(int n, float *p) /* fexp.h:49-78 */
{
    float mx = -FLT_MAX;
    for (int i = 0; i < n; ++i)
        if (p[i] > mx) mx = p[i];
    float sum = 0.0f;
    for (int i = 0; i < n; ++i) {
        p[i] = fexp_d(p[i] - mx);
        sum += p[i];
    }
    float sc = 1.0f / sum;
    for (int i = 0; i < n; ++i) p[i] *= sc;
}

/* One frame through the 3-layer net, naive (no-BLAS) summation order:
 * nn.cpp:872-899 + 771-793. */
static void nn_forward1(const orc_nn *n, const float *x, float *out, float *xin, float *hid, float *o)
{
    for (int i = 0; i < n->nin4; ++i) {
        float v = i < n->nin ? x[i] : 0.0f;
        v -= n->mean[i];
        v *= n->dev[i];
        xin[i] = v;
    }
    for (int j = 0; j < n->nhid4; ++j) {
        float acc = n->b1[j];
        const float *w = n->w1 + (size_t)j * n->nin4;
        for (int k = 0; k < n->nin4; ++k) acc += xin[k] * w[k];
        hid[j] = acc;
    }
    for (int j = 0; j < n->nhid; ++j) hid[j] = orc_fsigmoid(hid[j]);
    for (int j = n->nhid; j < n->nhid4; ++j) hid[j] = 0.0f;
    for (int j = 0; j < n->nout4; ++j) {
        float acc = n->b2[j];
        const float *w = n->w2 + (size_t)j * n->nhid4;
        for (int k = 0; k < n->nhid4; ++k) acc += hid[k] * w[k];
        o[j] = acc;
    }
    orc_fsoftmax(n->nout, o);
    memcpy(out, o, sizeof(float) * n->nout);
}

void orc_nn_forward(const orc_nn *n, const float *in, float *out, int nvec)
{
    float *xin = (float *)malloc(4 * n->nin4), *hid = (float *)malloc(4 * n->nhid4), *o = (float *)malloc(4 * n->nout4);
    for (int v = 0; v < nvec; ++v) nn_forward1(n, in + (size_t)v * n->nin, out + (size_t)v * n->nout, xin, hid, o);
    free(xin); free(hid); free(o);
}

/* Hidden-layer activations of one net (debug aid for kernel tests). */
void orc_nn_hidden(const orc_nn *n, const float *in, float *hid_out, int nvec)
{
    float *xin = (float *)malloc(4 * n->nin4);
    for (int v = 0; v < nvec; ++v) {
        const float *x = in + (size_t)v * n->nin;
        for (int i = 0; i < n->nin4; ++i) {
            float t = i < n->nin ? x[i] : 0.0f;
            t -= n->mean[i]; t *= n->dev[i]; xin[i] = t;
        }
        for (int j = 0; j < n->nhid; ++j) {
            float acc = n->b1[j];
            const float *w = n->w1 + (size_t)j * n->nin4;
            for (int k = 0; k < n->nin4; ++k) acc += xin[k] * w[k];
            hid_out[(size_t)v * n->nhid + j] = orc_fsigmoid(acc);
        }
    }
    free(xin);
}

/* S3 + merger: traps.cpp:435-461, 465.  mel must already be sentence
 * normalised.  post: [T][nout(merger)]. */
void orc_posteriors_from_normed_mel(const float *mel, int T, int nb, const float *win,
                                    const orc_nn *b0, const orc_nn *b1, const orc_nn *mg, float *post)
{
    int ncoef = b0->nin / nb; /* 11 (add_c0) */
    float *XL = (float *)malloc(sizeof(float) * (size_t)T * b0->nin);
    float *XR = (float *)malloc(sizeof(float) * (size_t)T * b1->nin);
    orc_stc(mel, T, nb, win, ncoef, XL, XR);
    float *pl = (float *)malloc(4 * (size_t)b0->nout), *pr = (float *)malloc(4 * (size_t)b1->nout);
    float *mi = (float *)malloc(4 * (size_t)mg->nin);
    for (int t = 0; t < T; ++t) {
        orc_nn_forward(b0, XL + (size_t)t * b0->nin, pl, 1);
        orc_nn_forward(b1, XR + (size_t)t * b1->nin, pr, 1);
        memcpy(mi, pl, 4 * (size_t)b0->nout);
        memcpy(mi + b0->nout, pr, 4 * (size_t)b1->nout);
        for (int i = 0; i < mg->nin; ++i) mi[i] = (mi[i] > 0.0f ? logf(mi[i]) : 0.0f); /* sLn, traps.cpp:459 */
        orc_nn_forward(mg, mi, post + (size_t)t * mg->nout, 1);
    }
    free(XL); free(XR); free(pl); free(pr); free(mi);
}

/* The other TRAPS systems: posteriors/system = 1BT, 3BT, 1BT_DCT
 * (traps.cpp:222-283 CalcInputFeaturesForBandNets, 344-361 ForwardPassBandNets, 405-433 CalcInputFeaturesForMerger).
 * Every band's trajectory is the FIFO's content = frames clamp(r - S .. r + S, 0, T-1), S = (L-1)/2 (same FIFO and
 * warm-up / tail driver as LCRC), times the Hamming window when posteriors/hamming is set (traps.cpp:232-241,
 * sWindow_Hamming dspc.h:162-167 over a vector of ones).
 *   1BT     one net per band on its L-point trajectory; merger input = -sLn(concatenated band outputs).
 *   3BT     as executed by the reference: the same, over the first nb - 2 bands (the code copies ONE band's L values per
 *           net, traps.cpp:249-261; nets with 3 L inputs would read uninitialised memory - refused here).
 *   1BT_DCT no band nets: merger input = per band [C0, DCT_1 .. DCT_{shift-1}] (add_c0) or [DCT_1 .. DCT_shift],
 *           shift = merger inputs / bands (traps.cpp:263-283, 170).
 * system: 1 = 1BT, 2 = 1BT_DCT, 3 = 3BT.  mel must already be (sentence) normalised.  Returns 0, or -1 when the nets do
 * not fit the system the way the reference needs them. */
int orc_posteriors_trap_system(const float *mel, int T, int nb, int system, int L, int use_hamming, int add_c0,
                               orc_nn *const *bands, const orc_nn *mg, float *post)
{
    const int S = (L - 1) / 2;
    const int tb = system == 3 ? nb - 2 : nb;   /* trap_bands, traps.cpp:95-97 */
    if (L < 3 || L > 255 || tb < 1) return -1;
    float *ham = (float *)malloc(sizeof(float) * L);
    for (int i = 0; i < L; ++i) ham[i] = 1.0f * (0.54f - 0.46f * cosf(2.0f * (float)M_PI * i / (L - 1)));
    float *x = (float *)malloc(sizeof(float) * L);
    float *mi = (float *)calloc((size_t)mg->nin + 4, sizeof(float));
    int shift = 0, rc = 0;
    if (system == 2) {
        shift = mg->nin / tb;
        if (shift * tb != mg->nin || shift < 1 || (add_c0 ? shift - 1 : shift) > L) rc = -1;
    } else {
        int tot = 0;
        for (int i = 0; i < tb; ++i) { if (!bands[i] || bands[i]->nin != L) rc = -1; else tot += bands[i]->nout; }
        if (tot != mg->nin) rc = -1;
    }
    const float NormC = sqrtf(2.0f / (float)L), PiByN = (float)M_PI / (float)L;
    for (int r = 0; r < T && rc == 0; ++r) {
        float *o = mi;
        for (int b = 0; b < tb; ++b) {
            for (int j = 0; j < L; ++j) {
                int t = r - S + j;
                if (t < 0) t = 0;
                if (t > T - 1) t = T - 1;
                x[j] = mel[(size_t)t * nb + b];
                if (use_hamming) x[j] = x[j] * ham[j];
            }
            if (system == 2) {
                int nd = add_c0 ? shift - 1 : shift;
                if (add_c0) {   /* CalcC0, dspc.h:223-233 */
                    float sum = 0.0f;
                    for (int j = 0; j < L; ++j) sum += x[j];
                    sum *= NormC;
                    *o++ = sum;
                }
                for (int k = 0; k < nd; ++k) {   /* sDCT, dspc.h:206-221 */
                    float acc = 0, v = PiByN * (float)(k + 1);
                    for (int j = 0; j < L; ++j) acc += x[j] * cosf(v * ((float)j + 0.5f));
                    acc *= NormC;
                    *o++ = acc;
                }
            } else {
                orc_nn_forward(bands[b], x, o, 1);
                o += bands[b]->nout;
            }
        }
        if (system != 2)   /* sLn, then times -1 (traps.cpp:425-427) */
            for (int i = 0; i < mg->nin; ++i) mi[i] = (mi[i] > 0.0f ? logf(mi[i]) : 0.0f) * -1;
        orc_nn_forward(mg, mi, post + (size_t)r * mg->nout, 1);
    }
    free(ham); free(x); free(mi);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* P1: decoder soft function = glibc logf.  srec.h:192-195, srec.cpp:1088-97  */
/* ------------------------------------------------------------------------- */
void orc_log_inplace(float *p, size_t n)
{
    for (size_t i = 0; i < n; ++i) p[i] = logf(p[i]);
}

/* Restatement of glibc's logf algorithm (sysdeps/ieee754/flt-32/e_logf.c,
 * table e_logf_data.c; glibc 2.39) — SURVEY Appendix C.  The CUDA device
 * logf is a port of THIS function; tests check port == libm logf. */
static const double LOGF_T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

float orc_logf_port(float x)
{
    const double Ln2 = 0x1.62e42fefa39efp-1;
    const double A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    uint32_t ix;
    memcpy(&ix, &x, 4);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
        float xs = x * 0x1p23f; /* subnormal */
        memcpy(&ix, &xs, 4);
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> 19) & 15;
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    float zf;
    memcpy(&zf, &iz, 4);
    double z = (double)zf;
    double r = z * LOGF_T[i][0] - 1;
    double y0 = LOGF_T[i][1] + (double)k * Ln2;
    double r2 = r * r;
    double y = A1 * r + A2;
    y = A0 * r2 + y;
    y = y * r2 + (y0 + r);
    return (float)y;
}

/* count mismatches port vs libm over the bit-pattern range [lo, hi) */
uint64_t orc_logf_port_mismatches(uint32_t lo, uint32_t hi, uint32_t stride)
{
    uint64_t bad = 0;
    for (uint64_t b = lo; b < hi; b += stride) {
        uint32_t u = (uint32_t)b;
        float x;
        memcpy(&x, &u, 4);
        float a = logf(x), c = orc_logf_port(x);
        if (memcmp(&a, &c, 4) != 0 && !(isnan(a) && isnan(c))) ++bad;
    }
    return bad;
}

/* ------------------------------------------------------------------------- */
/* V1-V6: phone-loop token passing decoder.  phndec.cpp:44-302                */
/* ------------------------------------------------------------------------- */
typedef struct { int32_t phn, start, end; float like; } orc_label;

#define ORC_LN05 (-0.69314718055994530941723212145818f) /* phndec.cpp:9 */

/* logpost: [T][ncols] log-posteriors (state (i,j) reads column 3i+j-1, j=1..3,
 * phndec.cpp:352-368).  P phonemes, S = 3 states, hist = time_pruning.
 * Streams exactly like PhnDec::Init/ProcessFrame/Done with the (hist+1)-slot
 * shift register.  Returns the number of labels (written up to cap). */
int orc_decode(const float *logpost, int T, int ncols, int P, int S, int hist, float wp, orc_label *out, int cap)
{
    int cols = hist + 1, W = S + 1, nl = 0;
    float *alpha = (float *)malloc(sizeof(float) * P * W);
    int *prev = (int *)malloc(sizeof(int) * P * W), *len = (int *)malloc(sizeof(int) * P * W);
    int *hphn = (int *)malloc(sizeof(int) * cols), *hlen = (int *)malloc(sizeof(int) * cols);
    float *halpha = (float *)malloc(sizeof(float) * cols);
    for (int i = 0; i < P * W; ++i) { alpha[i] = -FLT_MAX; prev[i] = -1; len[i] = 0; }
    for (int i = 0; i < cols; ++i) { hphn[i] = -1; hlen[i] = -1; halpha[i] = -1.0f; }
    for (int i = 0; i < P; ++i) { alpha[i * W] = wp; prev[i * W] = -1; len[i * W] = 0; }
    int nframes = 0;
    float prev_alpha = 0.0f;

    for (int t = 0; t < T; ++t) {
        const float *fr = logpost + (size_t)t * ncols;
        /* PropagateInModels, phndec.cpp:96-119 */
        for (int i = 0; i < P; ++i)
            for (int j = S; j > 0; --j) {
                float cur = alpha[i * W + j] + ORC_LN05;
                float pv = alpha[i * W + j - 1] + ORC_LN05;
                float obs = fr[i * S + (j - 1)];
                if (cur > pv) {
                    alpha[i * W + j] = cur + obs;
                    len[i * W + j] += 1;
                } else {
                    alpha[i * W + j] = pv + obs;
                    prev[i * W + j] = prev[i * W + j - 1];
                    len[i * W + j] = len[i * W + j - 1] + 1;
                }
            }
        /* PropagateInNetwork + AddHistory, phndec.cpp:121-158 */
        float mx = -FLT_MAX;
        int mi = 0;
        for (int i = 0; i < P; ++i)
            if (alpha[i * W + S] > mx) { mx = alpha[i * W + S]; mi = i; }
        for (int i = 1; i < cols; ++i) { hphn[i - 1] = hphn[i]; hlen[i - 1] = hlen[i]; halpha[i - 1] = halpha[i]; }
        hphn[cols - 1] = prev[mi * W + S]; hlen[cols - 1] = len[mi * W + S]; halpha[cols - 1] = mx;
        for (int i = 0; i < P; ++i) { alpha[i * W] = mx + wp; prev[i * W] = mi; len[i * W] = 0; }
        ++nframes;
        /* TimePruning + GetBestToken, phndec.cpp:169-234 */
        if (nframes >= cols) {
            float bm = -FLT_MAX;
            int bl = 1, bp = 0;
            for (int i = 0; i < P; ++i)
                for (int j = 1; j <= S; ++j)
                    if (alpha[i * W + j] > bm) { bm = alpha[i * W + j]; bl = len[i * W + j]; bp = prev[i * W + j]; }
            int offs = cols - 1 - bl;
            while (offs > 0) {
                int l = hlen[offs];
                bp = hphn[offs];
                offs -= l;
            }
            if (offs == 0) {
                int end = nframes - cols + 1, start = end - hlen[0];
                float like = halpha[0] - prev_alpha;
                prev_alpha = halpha[0];
                if (nl < cap) { out[nl].phn = bp; out[nl].start = start; out[nl].end = end; out[nl].like = like; }
                ++nl;
            }
        }
    }
    /* Done, phndec.cpp:236-302 */
    {
        int offs = cols - 1, end = nframes, phn = prev[0];
        int cnt = 0;
        orc_label *tmp = (orc_label *)malloc(sizeof(orc_label) * (cols + 1));
        while (offs > 0 && phn != -1) {
            int l = hlen[offs], start = end - l;
            float a = halpha[offs];
            int pp = hphn[offs];
            offs -= l;
            float like = offs > 0 ? a - halpha[offs] : a - prev_alpha;
            tmp[cnt].phn = phn; tmp[cnt].start = start; tmp[cnt].end = end; tmp[cnt].like = like;
            ++cnt;
            end = start;
            phn = pp;
        }
        for (int k = cnt - 1; k >= 0; --k) {
            if (nl < cap) out[nl] = tmp[k];
            ++nl;
        }
        free(tmp);
    }
    free(alpha); free(prev); free(len); free(hphn); free(hlen); free(halpha);
    return nl;
}

/* ------------------------------------------------------------------------- */
/* Whole system: config dir -> model; audio -> labels.  srec.cpp:235-707,     */
/* 929-1111                                                                   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int fs, nbanks, vs, step, fmt, sent_mean_norm, hist, S, P;
    float lo, hi, preem, wpenalty, frame_shift, frame_floor, scale, dc_shift;
    int z_mean;
    int on_interval, on_mean, on_var;   /* [onlinenorm] estim_interval, mean_norm, var_norm */
    int bunch;                          /* [posteriors] bunch_size */
    int plp, plp_order, plp_add_c0;     /* [params] kind = plp, [plp] order / add_c0 */
    float plp_compress, plp_lifter, plp_scale;
    int system, trap_len, hamming, add_c0;   /* [posteriors] system (0 LCRC, 1 1BT, 2 1BT_DCT, 3 3BT), length, hamming, add_c0 */
    orc_nn *bands[32];                  /* 1BT / 3BT: one net per band */
    float win[32];
    orc_nn *b0, *b1, *mg;
    orc_mel *mel;
    char phn[128][32];
} orc_model;

static void cfg_get(const char *txt, const char *sec, const char *var, char *out, const char *def)
{
    /* minimal reader for the INI dialect of configz.cpp:102-165 */
    strcpy(out, def);
    char cur[256] = "";
    const char *p = txt;
    while (*p) {
        char line[1024];
        int n = 0;
        while (*p && *p != '\n' && *p != '\r' && n < 1023) line[n++] = *p++;
        line[n] = 0;
        while (*p == '\n' || *p == '\r') ++p;
        if (n > 1 && line[0] == '[') {
            strcpy(cur, line + 1);
            cur[strlen(cur) - 1] = 0;
        } else if (n == 0 || line[0] == '#') {
        } else if (strcmp(cur, sec) == 0) {
            char *eq = strchr(line, '=');
            if (!eq) continue;
            *eq = 0;
            if (strcmp(line, var) == 0) {
                char *v = eq + 1, *h = strchr(v, '#');
                if (h) *h = 0;
                strcpy(out, v);
            }
        }
    }
}

orc_model *orc_model_load(const char *dir)
{
    char path[2048], buf[1024];
    snprintf(path, sizeof path, "%s/config", dir);
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    char *txt = (char *)calloc(1, 1 << 16);
    size_t rd = fread(txt, 1, (1 << 16) - 1, f);
    (void)rd;
    fclose(f);
    orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
    cfg_get(txt, "source", "sample_freq", buf, "8000"); m->fs = atoi(buf);
    cfg_get(txt, "source", "format", buf, "lin16"); m->fmt = strcmp(buf, "alaw") == 0;
    cfg_get(txt, "source", "scale", buf, "1.0"); sscanf(buf, "%f", &m->scale);
    cfg_get(txt, "source", "dc_shift", buf, "0.0"); sscanf(buf, "%f", &m->dc_shift);
    cfg_get(txt, "melbanks", "nbanks", buf, "15"); m->nbanks = atoi(buf);
    cfg_get(txt, "melbanks", "vector_size", buf, "200"); m->vs = atoi(buf);
    cfg_get(txt, "melbanks", "vector_step", buf, "80"); m->step = atoi(buf);
    cfg_get(txt, "melbanks", "lower_freq", buf, "0"); sscanf(buf, "%f", &m->lo);
    cfg_get(txt, "melbanks", "higher_freq", buf, "4000"); sscanf(buf, "%f", &m->hi);
    cfg_get(txt, "melbanks", "preem_coef", buf, "0.0"); sscanf(buf, "%f", &m->preem);
    cfg_get(txt, "melbanks", "z_mean_source", buf, "false"); m->z_mean = strcmp(buf, "true") == 0;
    cfg_get(txt, "offlinenorm", "sent_mean_norm", buf, "false"); m->sent_mean_norm = strcmp(buf, "true") == 0;
    cfg_get(txt, "framenorm", "shift", buf, "0"); sscanf(buf, "%f", &m->frame_shift);
    cfg_get(txt, "framenorm", "min_floor", buf, "-9999.9"); sscanf(buf, "%f", &m->frame_floor);
    cfg_get(txt, "params", "kind", buf, "fbanks"); m->plp = strcmp(buf, "plp") == 0;     /* srec.cpp:41, 544-583 */
    cfg_get(txt, "plp", "order", buf, "12"); m->plp_order = atoi(buf);                   /* srec.cpp:51-55 */
    cfg_get(txt, "plp", "compress_fact", buf, "0.3333333"); sscanf(buf, "%f", &m->plp_compress);
    cfg_get(txt, "plp", "cep_lifter", buf, "22"); sscanf(buf, "%f", &m->plp_lifter);
    cfg_get(txt, "plp", "cep_scale", buf, "10"); sscanf(buf, "%f", &m->plp_scale);
    cfg_get(txt, "plp", "add_c0", buf, "false"); m->plp_add_c0 = strcmp(buf, "true") == 0;
    cfg_get(txt, "posteriors", "bunch_size", buf, "1"); m->bunch = atoi(buf) > 0 ? atoi(buf) : 1;
    cfg_get(txt, "posteriors", "system", buf, "1BT_DCT");
    m->system = !strcmp(buf, "LCRC") ? 0 : !strcmp(buf, "1BT") ? 1 : !strcmp(buf, "1BT_DCT") ? 2 : !strcmp(buf, "3BT") ? 3 : -1;
    cfg_get(txt, "posteriors", "length", buf, "31"); m->trap_len = atoi(buf);
    cfg_get(txt, "posteriors", "hamming", buf, "false"); m->hamming = strcmp(buf, "true") == 0;
    cfg_get(txt, "posteriors", "add_c0", buf, "true"); m->add_c0 = strcmp(buf, "true") == 0;
    cfg_get(txt, "onlinenorm", "estim_interval", buf, "0"); m->on_interval = atoi(buf);          /* srec.cpp:56-61, 594-601 */
    cfg_get(txt, "onlinenorm", "mean_norm", buf, "false"); m->on_mean = strcmp(buf, "true") == 0;
    cfg_get(txt, "onlinenorm", "var_norm", buf, "false"); m->on_var = strcmp(buf, "true") == 0;
    cfg_get(txt, "decoder", "wpenalty", buf, "-2.0"); sscanf(buf, "%f", &m->wpenalty);
    cfg_get(txt, "decoder", "time_pruning", buf, "40"); m->hist = atoi(buf);
    cfg_get(txt, "decoder", "num_states_per_phn", buf, "1"); m->S = atoi(buf);
    free(txt);
    snprintf(path, sizeof path, "%s/weights/merger.nbin", dir); m->mg = orc_nn_load(path);
    if (m->system != 0) {   /* the other TRAPS systems: one net per band (1BT, 3BT) or none (1BT_DCT); no windows */
        const int tb = m->system == 3 ? m->nbanks - 2 : m->nbanks;
        if (m->system < 0 || !m->mg || tb > 32) { free(m); return NULL; }
        for (int i = 0; i < tb && m->system != 2; ++i) {
            snprintf(path, sizeof path, "%s/weights/band%d.nbin", dir, i);
            m->bands[i] = orc_nn_load(path);
            if (!m->bands[i]) { free(m); return NULL; }
        }
    } else {
    snprintf(path, sizeof path, "%s/weights/band0.nbin", dir); m->b0 = orc_nn_load(path);
    snprintf(path, sizeof path, "%s/weights/band1.nbin", dir); m->b1 = orc_nn_load(path);
    if (!m->b0 || !m->b1 || !m->mg) { free(m); return NULL; }
    }
    for (int w = 0; w < 2 && m->system == 0; ++w) { /* traps.cpp:549-570 */
        snprintf(path, sizeof path, "%s/windows/band%d.window", dir, w);
        f = fopen(path, "r");
        if (!f) { free(m); return NULL; }
        for (int i = 0; i < 16; ++i)
            if (fscanf(f, "%f", &m->win[w * 16 + i]) != 1) { fclose(f); free(m); return NULL; }
        fclose(f);
    }
    snprintf(path, sizeof path, "%s/dicts/phonemes", dir); /* phndec.cpp:305-350 */
    f = fopen(path, "r");
    if (!f) { free(m); return NULL; }
    m->P = 0;
    while (fgets(buf, 255, f) && m->P < 128) {
        for (char *c = buf; *c; ++c)
            if (*c == '\n' || *c == '\r') *c = 0;
        strncpy(m->phn[m->P], buf, 31);
        ++m->P;
    }
    fclose(f);
    m->mel = orc_mel_create(m->nbanks, m->vs, m->step, m->fs, m->lo, m->hi, m->preem, m->z_mean);
    if (m->plp) orc_mel_set_plp(m->mel, m->plp_order, m->plp_compress, m->plp_lifter, m->plp_scale, m->plp_add_c0);
    return m;
}

void orc_model_destroy(orc_model *m)
{
    if (!m) return;
    orc_nn_destroy(m->b0); orc_nn_destroy(m->b1); orc_nn_destroy(m->mg); orc_mel_destroy(m->mel);
    for (int i = 0; i < 32; ++i) orc_nn_destroy(m->bands[i]);
    free(m);
}

int orc_model_nparams(const orc_model *m) { return orc_mel_nparams(m->mel); }   /* columns of the `-t par` matrix */

/* info: fs, nbanks, vs, step, fmt, sent_mean_norm, hist, S, P, nout, nin_band, nhid */
void orc_model_info(const orc_model *m, int *info, float *wpenalty)
{
    info[0] = m->fs; info[1] = m->nbanks; info[2] = m->vs; info[3] = m->step; info[4] = m->fmt;
    info[5] = m->sent_mean_norm; info[6] = m->hist; info[7] = m->S; info[8] = m->P;
    info[9] = m->mg->nout; info[10] = m->b0 ? m->b0->nin : (m->bands[0] ? m->bands[0]->nin : 0); info[11] = m->b0 ? m->b0->nhid : (m->bands[0] ? m->bands[0]->nhid : 0);
    *wpenalty = m->wpenalty;
}
const char *orc_model_phoneme(const orc_model *m, int i) { return m->phn[i]; }
orc_nn *orc_model_net(orc_model *m, int which) { return which == 0 ? m->b0 : which == 1 ? m->b1 : m->mg; }
const float *orc_model_windows(const orc_model *m) { return m->win; }
orc_mel *orc_model_mel(orc_model *m) { return m->mel; }

int orc_model_num_frames(const orc_model *m, int nbytes, int fmt)
{
    if (fmt < 0) fmt = m->fmt;
    int n = fmt == 0 ? nbytes / 2 : nbytes;
    return orc_num_frames(n, m->vs, m->step);
}

/* audio -> un-normalised log mel [T][nbanks] (what `-t par` saves) */
int orc_model_mel_from_audio(orc_model *m, const void *audio, int nbytes, int fmt, float *mel)
{
    if (fmt < 0) fmt = m->fmt;
    int n = fmt == 0 ? nbytes / 2 : nbytes;
    int cap = n > m->vs ? n : m->vs;
    if (cap < ORC_MB_VECTORSIZE) cap = ORC_MB_VECTORSIZE;
    float *wav = (float *)calloc((size_t)cap + 1, sizeof(float));
    orc_wave_to_float(fmt, audio, nbytes, m->scale, m->dc_shift, wav);
    int T = orc_mel_compute(m->mel, wav, n, m->frame_shift, m->frame_floor, mel);
    free(wav);
    return T;
}

/* un-normalised mel -> linear posteriors [T][nout] (what `-t post` saves) */
void orc_model_posteriors(orc_model *m, const float *mel, int T, float *post)
{
    float *mm = (float *)malloc(sizeof(float) * (size_t)T * m->nbanks);
    memcpy(mm, mel, sizeof(float) * (size_t)T * m->nbanks);
    if (m->sent_mean_norm) orc_sentence_mean_norm(mm, T, m->nbanks);
    if (m->system == 0) orc_posteriors_from_normed_mel(mm, T, m->nbanks, m->win, m->b0, m->b1, m->mg, post);
    else orc_posteriors_trap_system(mm, T, m->nbanks, m->system, m->trap_len, m->hamming, m->add_c0, m->bands, m->mg, post);
    free(mm);
}

/* linear posteriors -> labels (decSoftFunc = logf, then PhnDec) */
int orc_model_decode(orc_model *m, const float *post, int T, float wp, orc_label *out, int cap)
{
    int nc = m->mg->nout;
    float *lp = (float *)malloc(sizeof(float) * (size_t)T * nc);
    memcpy(lp, post, sizeof(float) * (size_t)T * nc);
    orc_log_inplace(lp, (size_t)T * nc);
    int n = orc_decode(lp, T, nc, m->P, m->S, m->hist, wp, out, cap);
    free(lp);
    return n;
}

/* The ONLINE path, SpeechRec::ProcessOnline + ProcessTail (srec.cpp:793-927) fed block by block: the streaming objects carry
 * their state across blocks (MelBanks frame buffer, melbanks.cpp:151-204; the 31-frame FIFO of Traps, traps.cpp:180-219; the
 * decoder), so the result does not depend on the block size and equals this whole-signal restatement:
 *   frames t = samples [t step, t step + vs) for every full frame (no frame at all below vs samples);
 *   FrameBasedNormalization, then Normalization::ProcessFrame (orc_online_norm) - NO sentence normalisation;
 *   the FIFO's warm-up / tail = clamped context (orc_stc); soft functions; PhnDec.
 * Which posterior rows reach the decoder (srec.cpp:826-831, 865-869, 904-908): a whole bunch is passed on when, AFTER it, the
 * FIFO's delay (= vectors fed - 1, traps.cpp:199-217) has reached the trap shift (15); the tail's 15 vectors always are.
 * With bunches of 5 that is exactly rows 0..T-1 for T >= 15; a SHORTER utterance gets max(T, 15) frames decoded - the tail
 * also carries 15 - T warm-up rows (virtual rows r < 0, context clamp(r-15..r+15, 0, T-1), i.e. rows of a signal with
 * copies of frame 0 in front), and the label times count those frames.  Bunch sizes that do not divide 15 add up to
 * bunch - 1 such rows in front of every utterance.
 * interval / mean_norm / var_norm < 0: the model's [onlinenorm] values.
 * Pinned against oracle/_ref/online_ref (the reference's own objects driven in blocks) by tests/test_oracle_pinned.py. */
int orc_model_recognize_online(orc_model *m, const void *audio, int nbytes, int fmt, float wp, int interval, int mean_norm,
                               int var_norm, orc_label *out, int cap)
{
    if (fmt < 0) fmt = m->fmt;
    if (interval < 0) interval = m->on_interval;
    if (mean_norm < 0) mean_norm = m->on_mean;
    if (var_norm < 0) var_norm = m->on_var;
    const int n = fmt == 0 ? nbytes / 2 : nbytes;
    if (n < m->vs) return 0;
    const int T = (n - m->vs) / m->step + 1;
    /* first row handed to the decoder: the first bunch whose last vector is number shift + 1 or later starts at vector a0
     * (shift = Traps::GetTrapShift = (length - 1) / 2 = 15 for the 31-frame systems) */
    const int shift = (m->trap_len - 1) / 2;
    int a0 = -1;
    for (int a = 1; a <= T; a += m->bunch) {
        const int b = a + m->bunch - 1 < T ? a + m->bunch - 1 : T;
        if (b >= shift + 1) { a0 = a; break; }
    }
    const int r_first = a0 > 0 ? a0 - (shift + 1) : T - shift;
    const int P = r_first < 0 ? -r_first : 0;
    float *mel = (float *)malloc(sizeof(float) * (size_t)(T + P) * m->nbanks);
    float *post = (float *)malloc(sizeof(float) * (size_t)(T + P) * m->mg->nout);
    orc_model_mel_from_audio(m, audio, nbytes, fmt, mel + (size_t)P * m->nbanks);
    orc_online_norm(mel + (size_t)P * m->nbanks, T, m->nbanks, interval, mean_norm, var_norm);
    for (int p = 0; p < P; ++p) memcpy(mel + (size_t)p * m->nbanks, mel + (size_t)P * m->nbanks, sizeof(float) * m->nbanks);
    if (m->system == 0) orc_posteriors_from_normed_mel(mel, T + P, m->nbanks, m->win, m->b0, m->b1, m->mg, post);
    else orc_posteriors_trap_system(mel, T + P, m->nbanks, m->system, m->trap_len, m->hamming, m->add_c0, m->bands, m->mg, post);
    const int k = orc_model_decode(m, post, T + P, wp, out, cap);
    free(mel);
    free(post);
    return k;
}

int orc_model_recognize(orc_model *m, const void *audio, int nbytes, int fmt, float wp, orc_label *out, int cap)
{
    int T = orc_model_num_frames(m, nbytes, fmt);
    float *mel = (float *)malloc(sizeof(float) * (size_t)T * m->nbanks);
    float *post = (float *)malloc(sizeof(float) * (size_t)T * m->mg->nout);
    orc_model_mel_from_audio(m, audio, nbytes, fmt, mel);
    orc_model_posteriors(m, mel, T, post);
    int n = orc_model_decode(m, post, T, wp, out, cap);
    free(mel); free(post);
    return n;
}
