// ref_online.cpp — a small DRIVER over the reference's own objects (TEST INFRASTRUCTURE ONLY).
//
// The reference reaches its online path (SpeechRec::ProcessOnline, srec.cpp:793-927; Normalization::ProcessFrame,
// norm.cpp:216-234) only from a sound card (RunLive, srec.cpp:1438-1490).  This driver feeds FILES through the very
// same objects so that their output can pin the restatement in oracle/phn_oracle.c and the CUDA streaming path.
// It contains no reference code: it only calls the reference's public classes, compiled where they lie under
// /root/reference by oracle/Makefile (target `ref`), and is linked with those objects into oracle/_ref/online_ref.
//
//   online_ref norm  <in.f32> <rows> <cols> <interval> <mean 0|1> <var 0|1> <out.f32>
//        Normalization::{StartEstimation, SetMeanNorm, SetVarNorm, ProcessFrame} row by row over a raw float32 matrix.
//   online_ref plp <audio.lin16> <out.f32> <fs> <vs> <step> <nbanks> <lo> <hi> <preem> <zmean 0|1> <order> <compress> <lifter> <scale> <add_c0 0|1>
//        PLPCoefs (plp.cpp:91-141 - compiled out of the reference's PHNREC_ONLY build, srec.cpp:563-583, so the phnrec binary cannot
//        produce it): AddWaveform + GetFeatures over a 16-bit file; rows of order (+1) float32 coefficients.
//   online_ref stream <config_dir> <audio> <block_bytes> <lin16|alaw> <penalty|-> <out.rec>
//        SpeechRec::Init + PhnDec::Init(out.rec) + ProcessOnline in blocks of <block_bytes> (last block flagged), Done.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

#include <cstdarg>

#include "srec.h"
#include "norm.h"
#include "plp.h"

// Linked with -Wl,--wrap=sprintf (oracle/Makefile): Normalization::Save (norm.cpp:339,353) prints " %e" into a char[10];
// for that format at most 9 characters are stored, every other call is the plain sprintf.
extern "C" int __wrap_sprintf(char *dst, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    int n;
    if (!strcmp(fmt, " %e")) {
        char tmp[64];
        n = vsnprintf(tmp, sizeof tmp, fmt, ap);
        strncpy(dst, tmp, 9);
        dst[9] = 0;
    } else {
        n = vsprintf(dst, fmt, ap);
    }
    va_end(ap);
    return n;
}

static int run_norm(int argc, char **argv)
{
    if (argc != 9) return 2;
    const int rows = atoi(argv[3]), cols = atoi(argv[4]), interval = atoi(argv[5]);
    const bool mean = atoi(argv[6]) != 0, var = atoi(argv[7]) != 0;
    std::vector<float> m((size_t)rows * cols);
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(m.data(), sizeof(float), m.size(), f) != m.size()) return 3;
    fclose(f);
    FILE *fo = fopen(argv[8], "wb");   // (opened before the chdir below: the path may be relative)
    if (!fo) return 5;
    // (Normalization::Save writes its XML to the configured file name when the estimate completes - default "none",
    // norm.cpp:216-234,309; run from a scratch directory so that side effect lands there)
    char tmpl[] = "/tmp/online_ref.XXXXXX";
    const char *td = mkdtemp(tmpl);
    if (!td || chdir(td) != 0) return 4;
    {
        Normalization N;
        if (interval != 0) N.StartEstimation(interval);   // srec.cpp:594-601
        N.SetSignalEstimEnd(false);
        N.SetMeanNorm(mean);
        N.SetVarNorm(var);
        N.SetScaleToGVar(false);
        for (int r = 0; r < rows; ++r) N.ProcessFrame(cols, m.data() + (size_t)r * cols);
    }
    unlink("none");
    rmdir(td);
    if (fwrite(m.data(), sizeof(float), m.size(), fo) != m.size()) return 5;
    fclose(fo);
    return 0;
}

static std::string abs_path(const char *p)
{
    if (p[0] == '/') return p;
    char cwd[2048];
    if (!getcwd(cwd, sizeof cwd)) return p;
    return std::string(cwd) + "/" + p;
}

static int run_stream(int argc, char **argv)
{
    if (argc != 8) return 2;
    // Normalization::SetFile (norm.cpp:248-259, called from SpeechRec::Init with the config's default file name "none")
    // LOADS a file of that name if the working directory has one, and Normalization::Save writes one there the moment an
    // estimate completes: run from a fresh scratch directory so that no run sees a previous run's estimate.
    const std::string cfg_dir = abs_path(argv[2]), audio = abs_path(argv[3]), out_rec = abs_path(argv[7]);
    char tmpl[] = "/tmp/online_ref.XXXXXX";
    const char *td = mkdtemp(tmpl);
    if (!td || chdir(td) != 0) return 9;
    struct Cleanup { std::string d; ~Cleanup() { unlink((d + "/none").c_str()); rmdir(d.c_str()); } } cleanup{td};
    SpeechRec SR;
    char cfg[2048];
    snprintf(cfg, sizeof cfg, "%s/config", cfg_dir.c_str());
    if (!SR.Init(cfg)) return 3;
    if (strcmp(argv[6], "-") != 0) SR.DE->SetWPenalty((float)atof(argv[6]));   // phnrec.cpp:212-221
    const SpeechRec::wave_format wf = SR.Str2WaveFormat(argv[5]);
    if (wf == SpeechRec::wfUnknown) return 4;
    SR.SetWaveFormat(wf);
    void *wave = 0;
    int nbytes = 0;
    if (!SR.LoadWaveform((char *)audio.c_str(), &wave, &nbytes)) return 5;
    const int block = atoi(argv[4]);
    if (block <= 0) return 6;
    // what RunLive does around ProcessOnline (srec.cpp:1438-1490): reset the streaming objects, open the decoder
    SR.MB.Reset();
    SR.TR.Reset();
    SR.ResetBunchBuff();
    if (SR.DE->Init((char *)out_rec.c_str()) != DECERR_NONE) return 7;
    int pos = 0;
    do {
        const int n = nbytes - pos < block ? nbytes - pos : block;
        const bool last = pos + n >= nbytes;
        if (!SR.ProcessOnline((char *)wave + pos, n, last)) return 8;
        pos += n;
    } while (pos < nbytes);
    SR.DE->Done();
    free(wave);
    return 0;
}

static int run_plp(int argc, char **argv)
{
    if (argc != 17) return 2;
    FILE *f = fopen(argv[2], "rb");
    if (!f) return 3;
    std::vector<short> raw;
    short buf[4096];
    size_t n;
    while ((n = fread(buf, sizeof(short), 4096, f)) > 0) raw.insert(raw.end(), buf, buf + n);
    fclose(f);
    std::vector<float> wav(raw.size() + 8);
    for (size_t i = 0; i < raw.size(); ++i) wav[i] = (float)raw[i];   // srec.cpp:742-743
    PLPCoefs P;
    P.SetSampleFreq(atoi(argv[4])); P.SetVectorSize(atoi(argv[5])); P.SetStep(atoi(argv[6]));
    P.SetBanksNum(atoi(argv[7])); P.SetLowFreq((float)atof(argv[8])); P.SetHighFreq((float)atof(argv[9]));
    P.SetPreemCoef((float)atof(argv[10])); P.SetZMeanSource(atoi(argv[11]) != 0);
    P.SetLPCOrder(atoi(argv[12])); P.SetCompressFactor((float)atof(argv[13])); P.SetCepstralLifter((float)atof(argv[14]));
    P.SetCepstralScale((float)atof(argv[15])); P.SetAddC0(atoi(argv[16]) != 0);
    if ((int)raw.size() < atoi(argv[5])) return 4;
    FILE *fo = fopen(argv[3], "wb");
    if (!fo) return 5;
    P.AddWaveform(wav.data(), (int)raw.size());
    std::vector<float> out(64);
    const int np = P.GetNParams();
    while (P.GetFeatures(out.data())) fwrite(out.data(), sizeof(float), np, fo);
    fclose(fo);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "norm")) return run_norm(argc, argv);
    if (argc >= 2 && !strcmp(argv[1], "plp")) return run_plp(argc, argv);
    if (argc >= 2 && !strcmp(argv[1], "stream")) return run_stream(argc, argv);
    fprintf(stderr, "usage: online_ref norm|stream ...\n");
    return 2;
}
