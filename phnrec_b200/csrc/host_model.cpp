// host_model.cpp — model artefacts read from a PHN_* directory and the front-end tables.
// Everything here runs once at phn_create(); tables are built with the same single-precision
// expressions (and the same libm) the reference uses, so the device kernels that consume them
// reproduce the reference's numbers bit for bit.  Compile with -ffp-contract=off.
#include "internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace phn {

// .nbin layout (nn.cpp:464-531; padding rule nn.cpp:633-682): int32 nlayers(=2), nIn, nHid, nOut;
// then W1 [nHid4][nIn4], W2 [nOut4][nHid4], b1, b2, mean, dev with X4 = X rounded up to 4 floats.
int HostNet::load(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return PHN_ERR_NN_FILE;
    int32_t hdr[4];
    if (fread(hdr, sizeof(int32_t), 4, f) != 4) { fclose(f); return PHN_ERR_NN_FILE; }
    if (hdr[0] != 2 || hdr[1] <= 0 || hdr[2] <= 0 || hdr[3] <= 0) { fclose(f); return PHN_ERR_NN_FORMAT; }
    // the header is not trusted beyond what the file can hold (a corrupt size must not turn into a huge allocation)
    if (hdr[1] > (1 << 20) || hdr[2] > (1 << 20) || hdr[3] > (1 << 20)) { fclose(f); return PHN_ERR_NN_FORMAT; }
    nin = hdr[1]; nhid = hdr[2]; nout = hdr[3];
    auto up4 = [](int n) { return (n + 3) / 4 * 4; };
    nin4 = up4(nin); nhid4 = up4(nhid); nout4 = up4(nout);
    {
        const long at = ftell(f);
        fseek(f, 0, SEEK_END);
        const long fsize = ftell(f);
        fseek(f, at, SEEK_SET);
        const unsigned long long need = 4ull * ((unsigned long long)nhid4 * nin4 + (unsigned long long)nout4 * nhid4 + nhid4 + nout4 + 2ull * nin4);
        if (at < 0 || fsize < 0 || (unsigned long long)(fsize - at) < need) { fclose(f); return PHN_ERR_NN_FILE; }   // NN_READERR: truncated
    }
    struct { std::vector<float> *v; size_t n; } parts[] = {
        {&w1, (size_t)nhid4 * nin4}, {&w2, (size_t)nout4 * nhid4}, {&b1, (size_t)nhid4},
        {&b2, (size_t)nout4},        {&mean, (size_t)nin4},        {&dev, (size_t)nin4}};
    for (auto &p : parts) {
        p.v->resize(p.n);
        if (fread(p.v->data(), sizeof(float), p.n, f) != p.n) { fclose(f); return PHN_ERR_NN_FILE; }
    }
    fclose(f);
    return PHN_OK;
}

// ---- ASCII weights / norms (NeuralNet::LoadAscii, GetInfo, ParseWeights, ParseNorms: nn.cpp:113-462)
namespace {
struct Tok {   // whitespace-separated tokens, numbers read with sscanf("%e") / ("%d") like GetFloatValue / GetIntValue (nn.cpp:952-982)
    const char *p;
    explicit Tok(const char *s) : p(s) {}
    bool next(char *buf, size_t cap)
    {
        while (*p && strchr(" \t\n\r", *p)) ++p;
        if (!*p) return false;
        size_t i = 0;
        while (*p && !strchr(" \t\n\r", *p)) { if (i + 1 < cap) buf[i++] = *p; ++p; }
        buf[i] = 0;
        return true;
    }
    bool word(const char *w) { char b[100]; return next(b, sizeof b) && !strcmp(b, w); }
    bool integer(int *v) { char b[100]; return next(b, sizeof b) && sscanf(b, "%d", v) == 1; }
    bool real(float *v) { char b[100]; return next(b, sizeof b) && sscanf(b, "%e", v) == 1; }
};
bool read_text(const std::string &path, std::string &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}
}  // namespace

int HostNet::load_ascii(const std::string &weights_path, const std::string &norms_path)
{
    std::string txt;
    if (!read_text(weights_path, txt)) return PHN_ERR_NN_FILE;
    // sizes (GetInfo): nOut / nHid from the bias vectors, nIn = |W1| / nHid
    std::vector<float> v[4];
    const char *tags[4] = {"weigvec", "weigvec", "biasvec", "biasvec"};
    Tok t(txt.c_str());
    for (int k = 0; k < 4; ++k) {
        int n = 0;
        if (!t.word(tags[k]) || !t.integer(&n) || n <= 0) return PHN_ERR_NN_FORMAT;
        v[k].resize((size_t)n);
        for (int i = 0; i < n; ++i)
            if (!t.real(&v[k][i])) return PHN_ERR_NN_FORMAT;
    }
    nhid = (int)v[2].size(); nout = (int)v[3].size();
    nin = (int)(v[0].size() / (size_t)nhid);
    if (nin <= 0 || (size_t)nin * nhid != v[0].size() || (size_t)nhid * nout != v[1].size()) return PHN_ERR_NN_FORMAT;
    auto up4 = [](int n) { return (n + 3) / 4 * 4; };
    nin4 = up4(nin); nhid4 = up4(nhid); nout4 = up4(nout);
    w1.assign((size_t)nhid4 * nin4, 0.0f); w2.assign((size_t)nout4 * nhid4, 0.0f);
    b1.assign((size_t)nhid4, 0.0f); b2.assign((size_t)nout4, 0.0f);
    mean.assign((size_t)nin4, 0.0f); dev.assign((size_t)nin4, 1.0f);
    for (int j = 0; j < nhid; ++j) memcpy(&w1[(size_t)j * nin4], &v[0][(size_t)j * nin], sizeof(float) * nin);     // row = hidden unit
    for (int k = 0; k < nout; ++k) memcpy(&w2[(size_t)k * nhid4], &v[1][(size_t)k * nhid], sizeof(float) * nhid);  // row = output
    memcpy(b1.data(), v[2].data(), sizeof(float) * nhid);
    memcpy(b2.data(), v[3].data(), sizeof(float) * nout);
    if (!norms_path.empty()) {
        if (!read_text(norms_path, txt)) return PHN_ERR_NN_FILE;
        Tok u(txt.c_str());
        for (int k = 0; k < 2; ++k) {
            int n = 0;
            if (!u.word("vec") || !u.integer(&n)) return PHN_ERR_NN_FORMAT;
            float *dst = k ? dev.data() : mean.data();
            for (int i = 0; i < nin; ++i)
                if (!u.real(&dst[i])) return PHN_ERR_NN_FORMAT;
        }
    }
    return PHN_OK;
}

int HostNet::save_nbin(const std::string &path) const
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return PHN_ERR_NN_FILE;
    const int32_t hdr[4] = {2, nin, nhid, nout};
    bool ok = fwrite(hdr, sizeof(int32_t), 4, f) == 4;
    const std::vector<float> *parts[] = {&w1, &w2, &b1, &b2, &mean, &dev};
    for (auto *p : parts) ok = ok && fwrite(p->data(), sizeof(float), p->size(), f) == p->size();
    ok = (fclose(f) == 0) && ok;
    return ok ? PHN_OK : PHN_ERR_NN_FILE;
}

static float mel_scale(float hz) { return 1127.0f * logf(1.0f + hz / 700.0f); }  // dspc.h:169-177

void MelTables::build(int nbanks_, int vs_, int step_, int fs_, float lo, float hi)
{
    nbanks = nbanks_; vs = vs_; step = step_; fs = fs_;
    N = 1; logN = 0;
    while (N < vs) { N <<= 1; ++logN; }  // melbanks.cpp:44-46
    N2 = N / 2;

    // Hamming window as the reference obtains it: sWindow_Hamming applied to ones (dspc.h:162-167)
    hamming.resize(vs);
    for (int i = 0; i < vs; ++i) hamming[i] = 1.0f * (0.54f - 0.46f * cosf(2.0f * (float)M_PI * i / (vs - 1)));

    // triangular filters on the mel axis (dspc.cpp:80-225)
    if (lo < 0.0f) lo = 0.0f;
    if (hi > (float)fs / 2.0f) hi = (float)fs / 2.0f;
    const float bin_hz = (float)fs / (float)N;
    const float mlo = mel_scale(lo), mhi = mel_scale(hi);
    fftlo = (int)(lo / bin_hz + 1.5f);
    ffthi = (int)(hi / bin_hz - 0.5f);
    if (fftlo < 1) fftlo = 1;
    if (ffthi >= N2) ffthi = N2 - 1;
    std::vector<float> centre(nbanks + 2);
    const float dmel = (mhi - mlo) / (nbanks + 1);
    float acc = mlo;
    f0.assign((size_t)nbanks + 1, 0.0f);
    for (int i = 0; i <= nbanks; ++i) { acc = acc + dmel; centre[i] = acc; f0[i] = 700.0f * (expf(acc / 1127.0f) - 1.0f); }  // repeated addition, dspc.cpp:156-162; Scale_MelToLinear
    centre[nbanks + 1] = INFINITY;
    banks.assign(N2, -1);
    coeffs.assign(N2, 0.0f);
    int ch = 0;
    for (int k = fftlo; k <= ffthi; ++k) {
        const float m = mel_scale((float)k * bin_hz);
        while (m > centre[ch] && ch <= nbanks) ++ch;
        banks[k] = ch;
        const float below = ch == 0 ? mlo : centre[ch - 1];
        coeffs[k] = (centre[ch] - m) / (centre[ch] - below);
    }
    // bank b collects (1-c)P[k] from bins assigned to b and cP[k] from bins assigned to b+1, in
    // ascending k (dspc.cpp:236-269); Banks[] is non-decreasing so that is one contiguous range.
    bank_klo.assign(nbanks, 0);
    bank_khi.assign(nbanks, -1);
    for (int b = 0; b < nbanks; ++b) {
        int lo_k = -1, hi_k = -2;
        for (int k = fftlo; k <= ffthi; ++k)
            if (banks[k] == b || banks[k] == b + 1) {
                if (lo_k < 0) lo_k = k;
                hi_k = k;
            }
        bank_klo[b] = lo_k < 0 ? 0 : lo_k;
        bank_khi[b] = hi_k;
    }

    // FFT twiddles exactly as the Numerical-Recipes recurrence in cFour1 produces them
    // (dspc.cpp:24-78, isign = -1): per stage a double-precision rotation started at (1, 0).
    tw.assign((size_t)2 * (N - 1), 0.0);
    for (int h = 1; h < N; h <<= 1) {
        const unsigned mmax = 2u * (unsigned)h;  // NR counts floats
        const double theta = -1 * (6.28318530717959 / mmax);
        double wtemp = sin(0.5 * theta);
        const double wpr = -2.0 * wtemp * wtemp;
        const double wpi = sin(theta);
        double wr = 1.0, wi = 0.0;
        for (int m = 0; m < h; ++m) {
            tw[2 * (size_t)(h - 1 + m)] = wr;
            tw[2 * (size_t)(h - 1 + m) + 1] = wi;
            wtemp = wr;
            wr = wtemp * wpr - wi * wpi + wr;
            wi = wi * wpr + wtemp * wpi + wi;
        }
    }
}

}  // namespace phn
