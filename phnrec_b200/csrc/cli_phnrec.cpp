// cli_phnrec.cpp — the `phnrec` command, drop-in for the reference CLI on the offline path.
//
// Same switches, same config/model directory, same output files (HTK .rec / MLF labels, HTK
// parameter and posterior matrices) as phnrec.cpp:113-299 + SpeechRec::ProcessFile*
// (srec.cpp:1113-1291); the work itself is done by libphnrec_b200 through its C ABI, batching
// the lines of a file list into ragged GPU batches.  `-a` (live input) takes raw samples from stdin in
// the reference's 125 ms blocks through the streaming API and prints the committed labels in the
// reference's live formats (-f str | strlen | lab).
#include <cctype>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

#include "phnrec_b200.h"

namespace {

enum DataFmt { dfWaveform = 0, dfParams = 1, dfPosteriors = 2, dfStrings = 3, dfUnknown = 4 };  // srec.h order

bool g_verbose = false;

// PHNREC_CLI_TIMING=1: wall-clock of the process's phases on stderr (development aid)
double wall_ms()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
const bool g_timing = getenv("PHNREC_CLI_TIMING") != nullptr;
const double g_t0 = wall_ms();
void phase(const char *what) { if (g_timing) fprintf(stderr, "[phnrec timing] %8.1f ms  %s\n", wall_ms() - g_t0, what); }

[[noreturn]] void die(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void die(const char *fmt, ...)
{
    // SpeechRec::MError (srec.cpp:118-122): "ERROR: <msg>" on stderr, exit(1)
    va_list ap;
    va_start(ap, fmt);
    fputs("ERROR: ", stderr);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    exit(1);
}

void logmsg(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void logmsg(const char *fmt, ...)
{
    if (!g_verbose) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stdout, fmt, ap);
    va_end(ap);
}

void help()
{
#ifdef PHN_VADALIZE
    puts("\nUSAGE: vadalize [options]\n");   // vadalize.cpp:28
#else
    puts("\nUSAGE: phnrec [options]\n");
#endif
    puts(" -c dir             configuration directory");
    puts(" -l file            list of files");
    puts(" -i file            input file");
    puts(" -o file            output file");
    puts(" -m file            output MLF");
    puts(" -a                 live audio input");
    puts(" -s fmt [waveform]  source format (wf-waveform, par-parameters, post-posteriors)");
    puts(" -t fmt [strings]   target format (par-parameters, post-posteriors, str-strings)");
    puts(" -w fmt [lin16]     waveform format (lin16, alaw)");
    puts(" -f fmt [str]       live output format (str, strlen, lab)");
    puts(" -p num [-3.8]      phoneme insertion penalty");
    puts(" -v                 verbose\n");
}

DataFmt str2fmt(const char *s)
{
    if (!strcmp(s, "wf")) return dfWaveform;
    if (!strcmp(s, "par")) return dfParams;
    if (!strcmp(s, "post")) return dfPosteriors;
    if (!strcmp(s, "str")) return dfStrings;
    die("Invalid data format '%s'. Supported data formats are 'wf', 'mb', 'post' and 'str'.\n", s);
}

// ---- file-name helpers with the reference's rules (filename.cpp:30-114)
size_t last_sep(const std::string &s)
{
    size_t a = s.rfind('/'), b = s.rfind('\\');
    if (a == std::string::npos) return b;
    if (b == std::string::npos) return a;
    return a > b ? a : b;
}
std::string change_suffix(std::string f, const std::string &suf)
{
    size_t dot = f.rfind('.'), sep = last_sep(f);
    if (dot == std::string::npos || (sep != std::string::npos && sep > dot)) return f + "." + suf;
    return f.substr(0, dot + 1) + suf;
}
std::string change_path(const std::string &f, const std::string &path)
{
    size_t sep = last_sep(f);
    if (sep == std::string::npos) return f;  // only when the name has a directory
    return path + f.substr(sep);
}

// ---- HTK parameter files (Mat::saveHTK / loadHTK, matrix.h:2506-2573)
uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
bool load_htk(const std::string &path, std::vector<float> &m, int &rows, int &cols)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char h[12];
    if (fread(h, 1, 12, f) != 12) { fclose(f); return false; }
    rows = (int)((uint32_t)h[0] << 24 | (uint32_t)h[1] << 16 | (uint32_t)h[2] << 8 | h[3]);
    const int samp_size = (int)((uint32_t)h[8] << 8 | h[9]);
    cols = samp_size / 4;
    if (rows < 0 || cols <= 0) { fclose(f); return false; }
    {   // the header is not trusted beyond what the file holds (a corrupt row count must not become a huge allocation)
        struct stat sb;
        if (fstat(fileno(f), &sb) != 0 || !S_ISREG(sb.st_mode) || (unsigned long long)sb.st_size < 12ull + 4ull * (unsigned long long)rows * cols) { fclose(f); return false; }
    }
    m.resize((size_t)rows * cols);
    const size_t n = fread(m.data(), 4, m.size(), f);
    fclose(f);
    if (n != m.size()) return false;
    uint32_t *u = reinterpret_cast<uint32_t *>(m.data());
    for (size_t i = 0; i < m.size(); ++i) u[i] = bswap32(u[i]);
    return true;
}
bool save_htk(const std::string &path, const float *m, int rows, int cols)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t period = 100000;
    const uint16_t size = (uint16_t)(cols * 4), kind = 6;
    unsigned char h[12] = {(unsigned char)(rows >> 24), (unsigned char)(rows >> 16), (unsigned char)(rows >> 8), (unsigned char)rows,
                           (unsigned char)(period >> 24), (unsigned char)(period >> 16), (unsigned char)(period >> 8), (unsigned char)period,
                           (unsigned char)(size >> 8), (unsigned char)size, (unsigned char)(kind >> 8), (unsigned char)kind};
    bool ok = fwrite(h, 1, 12, f) == 12;
    std::vector<uint32_t> be((size_t)rows * cols);
    const uint32_t *u = reinterpret_cast<const uint32_t *>(m);
    for (size_t i = 0; i < be.size(); ++i) be[i] = bswap32(u[i]);
    ok = ok && fwrite(be.data(), 4, be.size(), f) == be.size();
    fclose(f);
    return ok;
}

bool load_bytes(const std::string &path, std::vector<unsigned char> &out)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    if (n < 0) { fclose(f); return false; }   // a directory, a FIFO: "Can not open waveform file" like the reference's failed read
    fseek(f, 0, SEEK_SET);
    const size_t at = out.size();
    out.resize(at + (size_t)n);
    const bool ok = n == 0 || fread(out.data() + at, 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

struct Job { std::string src, dst; };

// label text, as strings (the list pipeline formats on its worker threads and writes from one thread, in list order)
void append_rec(std::string &out, phn_ctx *ctx, const phn_label *l, int64_t n)
{
    char buf[256];
#ifdef PHN_VADALIZE
    for (int64_t i = 0; i < n; ++i) {   // phndecalize.cpp:227-239, 299-314
        const std::string ph = phn_phoneme(ctx, l[i].phn);
        if (ph == "pau" || ph == "int" || ph == "spk") continue;
        const float alizeStart = (float)l[i].start, alizeEnd = (float)l[i].end;
        snprintf(buf, sizeof buf, "%.2f %.2f speech\n", alizeStart / 100, alizeEnd / 100);
        out += buf;
    }
#else
    for (int64_t i = 0; i < n; ++i) {   // phndec.cpp:230,292
        snprintf(buf, sizeof buf, "%d00000 %d00000 %s %f\n", l[i].start, l[i].end, phn_phoneme(ctx, l[i].phn), l[i].like);
        out += buf;
    }
#endif
}
void append_mlf_lines(std::string &out, phn_ctx *ctx, const phn_label *l, int64_t n)
{
    char buf[256];
    for (int64_t i = 0; i < n; ++i) {   // SpeechRec::OnWordMLF, srec.cpp:137-161
        if (l[i].start == 0) out += "0"; else { snprintf(buf, sizeof buf, "%u00000", (unsigned)l[i].start); out += buf; }
        if (l[i].end == 0) out += " 0"; else { snprintf(buf, sizeof buf, " %u00000", (unsigned)l[i].end); out += buf; }
        snprintf(buf, sizeof buf, " %s %f\n", phn_phoneme(ctx, l[i].phn), l[i].like);
        out += buf;
    }
}

// ------------------------------------------------------------------------------------------------
// List mode, audio -> labels (SpeechRec::ProcessFileList, srec.cpp:1246-1291, for dfWaveform -> dfStrings).
// The reference walks the list one file after the other; utterances are independent (fresh decoder, per-utterance mean,
// clamped context: srec.cpp:1148-1167), so the list is cut into contiguous batches that are handed out to one host
// thread + one phn_ctx per GPU (PHNREC_DEVICES).  A worker keeps two batches in flight (phn_recognize_async / phn_wait):
// it reads the files of batch k+1 into page-locked memory while batch k is on its GPU.  Label text is formatted on the
// workers and written by the calling thread strictly in list order - the output (MLF or .rec files) is byte-identical
// whatever the number of GPUs.  No data moves between GPUs; the only gather is this ordered write on the host.
struct ListPipeline {
    struct Result { std::vector<std::string> text; int n_good = 0; std::string error; bool done = false; };
    std::vector<Job> jobs;
    std::vector<std::pair<size_t, size_t>> batches;
    std::vector<Result> results;
    std::vector<phn_ctx *> ctxs;
    bool to_mlf = false;
    int readers = 2;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<size_t> next{0};

    struct Stage {
        unsigned char *pinned = nullptr;
        size_t cap = 0;
        std::vector<int64_t> off;
        size_t b = 0;
        int n = 0;
        int64_t frames = 0;
        std::string error;
    };

    static void parallel_for(int n, int threads, const std::function<void(int)> &fn)
    {
        if (threads <= 1 || n < 2 * threads) { for (int i = 0; i < n; ++i) fn(i); return; }
        std::atomic<int> at{0};
        auto body = [&] { for (int i = at.fetch_add(16); i < n; i = at.fetch_add(16)) for (int k = i; k < n && k < i + 16; ++k) fn(k); };
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(body);
        body();
        for (auto &t : th) t.join();
    }

    // files of batch b -> one page-locked buffer (whole files as raw bytes, SpeechRec::LoadWaveform, srec.cpp:1384-1422);
    // the first file that cannot be read ends the batch there, like the reference stopping at it
    void load(Stage &st, size_t b, phn_ctx *ctx)
    {
        const size_t j0 = batches[b].first, n = batches[b].second - j0;
        st.b = b; st.error.clear();
        std::vector<int64_t> size(n, -1);
        parallel_for((int)n, readers, [&](int i) {
            struct stat sb;
            if (stat(jobs[j0 + i].src.c_str(), &sb) == 0 && S_ISREG(sb.st_mode)) size[i] = (int64_t)sb.st_size;
        });
        size_t good = n;
        for (size_t i = 0; i < n; ++i) if (size[i] < 0) { good = i; break; }
        st.off.assign(good + 1, 0);
        for (size_t i = 0; i < good; ++i) st.off[i + 1] = st.off[i] + size[i];
        const size_t total = (size_t)st.off[good];
        if (total + 16 > st.cap) {
            if (st.pinned) phn_host_free_pinned(st.pinned);
            st.cap = total + total / 2 + 4096;
            st.pinned = (unsigned char *)phn_host_alloc_pinned((int64_t)st.cap);
            if (!st.pinned) die("Can not allocate %zu bytes of page-locked memory\n", st.cap);
        }
        std::atomic<int> first_bad{(int)good};
        parallel_for((int)good, readers, [&](int i) {
            FILE *f = fopen(jobs[j0 + i].src.c_str(), "rb");
            bool ok = f != nullptr;
            if (ok) ok = size[i] == 0 || fread(st.pinned + st.off[i], 1, (size_t)size[i], f) == (size_t)size[i];
            if (f) fclose(f);
            if (!ok) { int cur = first_bad.load(); while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {} }
        });
        if ((size_t)first_bad.load() < good) good = (size_t)first_bad.load();
        if (good < n) st.error = "Can not open waveform file: " + jobs[j0 + good].src + "\n";
        st.off.resize(good + 1);
        st.n = (int)good;
        st.frames = 0;
        for (size_t i = 0; i < good; ++i) st.frames += phn_num_frames(ctx, st.off[i + 1] - st.off[i]);
    }

    void finish(Stage &st, phn_ctx *ctx, bool waited)
    {
        Result r;
        r.n_good = st.n;
        r.error = st.error;
        if (st.n > 0 && waited) {
            const int64_t cap = st.frames + (int64_t)48 * st.n;
            std::vector<phn_label> lab((size_t)cap);
            std::vector<int64_t> loff((size_t)st.n + 1);
            if (phn_wait(ctx, lab.data(), cap, loff.data(), nullptr)) die("%s", phn_last_error(ctx));
            r.text.resize((size_t)st.n);
            for (int i = 0; i < st.n; ++i) {
                if (to_mlf) append_mlf_lines(r.text[i], ctx, lab.data() + loff[i], loff[i + 1] - loff[i]);
                else if (jobs[batches[st.b].first + i].dst.empty()) append_mlf_lines(r.text[i], ctx, lab.data() + loff[i], loff[i + 1] - loff[i]);
                else append_rec(r.text[i], ctx, lab.data() + loff[i], loff[i + 1] - loff[i]);
            }
        }
        r.done = true;
        {
            std::lock_guard<std::mutex> g(mu);
            results[st.b] = std::move(r);
        }
        cv.notify_all();
    }

    void worker(phn_ctx *ctx)
    {
        Stage st[2];
        std::deque<int> inflight;   // stage indices, oldest first
        int k = 0;
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= batches.size()) break;
            Stage &s = st[k & 1];
            load(s, b, ctx);
            if (s.n > 0 && phn_recognize_async(ctx, s.pinned, s.off.data(), s.n)) die("%s", phn_last_error(ctx));
            inflight.push_back(k & 1);
            if (inflight.size() == 2) { Stage &o = st[inflight.front()]; inflight.pop_front(); finish(o, ctx, true); }
            ++k;
        }
        while (!inflight.empty()) { Stage &o = st[inflight.front()]; inflight.pop_front(); finish(o, ctx, true); }
        for (auto &s : st) if (s.pinned) phn_host_free_pinned(s.pinned);
    }

    // runs the pipeline; writes outputs in list order; returns the first error (empty: none)
    std::string run(FILE *mlf)
    {
        results.assign(batches.size(), Result());
        std::vector<std::thread> th;
        for (phn_ctx *c : ctxs) th.emplace_back([this, c] { worker(c); });
        std::string err;
        for (size_t b = 0; b < batches.size() && err.empty(); ++b) {
            Result r;
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return results[b].done; });
                r = std::move(results[b]);
            }
            const size_t j0 = batches[b].first;
            for (int i = 0; i < r.n_good; ++i) {
                const Job &j = jobs[j0 + i];
                logmsg("%s -> %s\n", j.src.c_str(), j.dst.c_str());
                if (mlf) { fprintf(mlf, "\"%s\"\n", j.dst.c_str()); fputs(r.text[i].c_str(), mlf); fputs(".\n", mlf); continue; }
                if (j.dst.empty()) { fputs(r.text[i].c_str(), stdout); continue; }
                FILE *f = fopen(j.dst.c_str(), "w");
                if (!f) { err = "Can not create the label file: " + j.dst + "\n"; break; }
                fputs(r.text[i].c_str(), f);
                fclose(f);
            }
            if (err.empty()) err = r.error;
        }
        if (!err.empty()) next.store(batches.size());   // stop handing out work
        for (auto &t : th) t.join();
        return err;
    }
};

struct Runner {
    phn_ctx *ctx = nullptr;
    phn_info info{};
    DataFmt inf = dfWaveform, outf = dfStrings;
    FILE *mlf = nullptr;
    float info_penalty = 0.f;   // the insertion penalty in force (config value or -p): further devices' contexts get the same

    void ck(int rc) { if (rc) die("%s", phn_last_error(ctx)); }

    // label writers: PhnDec's fprintf (phndec.cpp:230,292) and SpeechRec::OnWordMLF (srec.cpp:137-161)
    void write_rec(FILE *f, const phn_label *l, int64_t n)
    {
#ifdef PHN_VADALIZE
        // the `vadalize` personality (phndecalize.cpp:227-239, 299-314): one "start end speech" line, in seconds, for every
        // segment that is not pau / int / spk; float arithmetic as in the reference (float / int, printed with %.2f)
        for (int64_t i = 0; i < n; ++i) {
            const std::string ph = phn_phoneme(ctx, l[i].phn);
            if (ph == "pau" || ph == "int" || ph == "spk") continue;
            const float alizeStart = (float)l[i].start, alizeEnd = (float)l[i].end;
            fprintf(f, "%.2f %.2f speech\n", alizeStart / 100, alizeEnd / 100);
        }
#else
        for (int64_t i = 0; i < n; ++i) fprintf(f, "%d00000 %d00000 %s %f\n", l[i].start, l[i].end, phn_phoneme(ctx, l[i].phn), l[i].like);
#endif
    }
    void write_mlf(const std::string &name, const phn_label *l, int64_t n)
    {
        fprintf(mlf, "\"%s\"\n", name.c_str());
        for (int64_t i = 0; i < n; ++i) {
            if (l[i].start == 0) fprintf(mlf, "0"); else fprintf(mlf, "%u00000", (unsigned)l[i].start);
            if (l[i].end == 0) fprintf(mlf, " 0"); else fprintf(mlf, " %u00000", (unsigned)l[i].end);
            fprintf(mlf, " %s %f\n", phn_phoneme(ctx, l[i].phn), l[i].like);
        }
        fprintf(mlf, ".\n");
    }

    // One ragged batch through the GPU.  `pending_error` (if any) is raised after the batch's
    // outputs are written, like the reference stopping at the first bad list line.
    void run(const std::vector<Job> &jobs, const std::string &pending_error)
    {
        const int n = (int)jobs.size();
        std::vector<int64_t> off(n + 1, 0), foff(n + 1, 0);
        std::vector<unsigned char> audio;
        std::vector<float> mat;  // mel or posteriors, concatenated
        std::string late_error = pending_error;
        int good = n;
        const int want_cols = inf == dfParams ? info.nbanks : info.n_outputs;
        for (int i = 0; i < n; ++i) {
            logmsg("%s -> %s\n", jobs[i].src.c_str(), jobs[i].dst.c_str());
            if (inf == dfWaveform) {
                if (!load_bytes(jobs[i].src, audio)) { late_error = "Can not open waveform file: " + jobs[i].src + "\n"; good = i; break; }
                off[i + 1] = (int64_t)audio.size();
            } else {
                std::vector<float> m;
                int rows, cols;
                if (!load_htk(jobs[i].src, m, rows, cols)) { late_error = "Can not open file: " + jobs[i].src + "\n"; good = i; break; }
                if (cols < want_cols) { late_error = "Invalid dimensionality of parameter vectors\n"; good = i; break; }
                for (int r = 0; r < rows; ++r) mat.insert(mat.end(), m.begin() + (size_t)r * cols, m.begin() + (size_t)r * cols + want_cols);
                foff[i + 1] = foff[i] + rows;  // extra columns are dropped (srec.cpp:983-997)
            }
        }
        if (good > 0) process(jobs, good, audio, off, mat, foff);
        if (!late_error.empty()) die("%s", late_error.c_str());
    }

    void process(const std::vector<Job> &jobs, int n, std::vector<unsigned char> &audio, std::vector<int64_t> &off,
                 std::vector<float> &mat, std::vector<int64_t> &foff)
    {
        std::vector<float> mel, post;
        if (inf == dfWaveform) {
            if (outf == dfStrings) {  // the fused path: audio -> labels without leaving the GPU
                ck(phn_mel(ctx, audio.data(), off.data(), n, nullptr, foff.data()));
                const int64_t cap = phn_label_capacity(ctx, foff.data(), n);
                std::vector<phn_label> lab((size_t)cap);
                std::vector<int64_t> loff(n + 1);
                ck(phn_recognize(ctx, audio.data(), off.data(), n, lab.data(), cap, loff.data(), nullptr));
                emit_labels(jobs, n, lab, loff);
                return;
            }
            ck(phn_mel(ctx, audio.data(), off.data(), n, nullptr, foff.data()));
            mel.resize((size_t)foff[n] * info.n_params);
            ck(phn_mel(ctx, audio.data(), off.data(), n, mel.data(), foff.data()));
            if (outf == dfParams) { emit_matrix(jobs, n, mel, foff, info.n_params); return; }
        } else if (inf == dfParams) {
            mel.swap(mat);
        } else {
            post.swap(mat);
        }
        if (inf != dfPosteriors) {
            post.resize((size_t)foff[n] * info.n_outputs);
            ck(phn_posteriors(ctx, mel.data(), foff.data(), n, post.data()));
            if (outf == dfPosteriors) { emit_matrix(jobs, n, post, foff, info.n_outputs); return; }
        }
        const int64_t cap = phn_label_capacity(ctx, foff.data(), n);
        std::vector<phn_label> lab((size_t)cap);
        std::vector<int64_t> loff(n + 1);
        ck(phn_decode(ctx, post.data(), foff.data(), n, nullptr, 1, lab.data(), cap, loff.data()));
        emit_labels(jobs, n, lab, loff);
    }

    void emit_labels(const std::vector<Job> &jobs, int n, const std::vector<phn_label> &lab, const std::vector<int64_t> &loff)
    {
        for (int i = 0; i < n; ++i) {
            const phn_label *l = lab.data() + loff[i];
            const int64_t cnt = loff[i + 1] - loff[i];
            if (mlf) { write_mlf(jobs[i].dst, l, cnt); continue; }
            if (jobs[i].dst.empty()) {  // no target: the decoder's default callback prints MLF-style lines to stdout
                FILE *save = mlf; mlf = stdout;
                for (int64_t k = 0; k < cnt; ++k) {
                    if (l[k].start == 0) fprintf(mlf, "0"); else fprintf(mlf, "%u00000", (unsigned)l[k].start);
                    if (l[k].end == 0) fprintf(mlf, " 0"); else fprintf(mlf, " %u00000", (unsigned)l[k].end);
                    fprintf(mlf, " %s %f\n", phn_phoneme(ctx, l[k].phn), l[k].like);
                }
                mlf = save;
                continue;
            }
            FILE *f = fopen(jobs[i].dst.c_str(), "w");
            if (!f) die("Can not create the label file: %s\n", jobs[i].dst.c_str());
            write_rec(f, l, cnt);
            fclose(f);
        }
    }

    void emit_matrix(const std::vector<Job> &jobs, int n, const std::vector<float> &m, const std::vector<int64_t> &foff, int cols)
    {
        for (int i = 0; i < n; ++i)
            if (!save_htk(jobs[i].dst, m.data() + (size_t)foff[i] * cols, (int)(foff[i + 1] - foff[i]), cols))
                die("Can not create file: %s\n", jobs[i].dst.c_str());
    }
};

}  // namespace


int main(int argc, char *argv[])
{
    const char *config_dir = nullptr, *file_list = nullptr, *input_file = nullptr, *output_file = nullptr;
    const char *output_mlf = nullptr, *wpenalty = nullptr, *wformat = nullptr;
    bool live = false;
    enum { ofLab, ofStr, ofStrLen } live_fmt = ofStr;   // phnrec.cpp:43-58,123
    int mlp_mode = PHN_MLP_EXACT_FP32;
    Runner R;

    if (argc == 1) { help(); return 1; }
    // the reference's bundled getopt (getopt.cpp:21-41): "-x val" or "-xval", the next argv is
    // taken verbatim as the argument (so "-p -3.0" works), bare words are ignored
    const char *opts = "-c:l:i:o:m:as:t:w:f:p:v";
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!(a[0] == '-' && isalpha((unsigned char)a[1]))) continue;
        const char *o = strchr(opts, a[1]);
        if (!o) { fprintf(stderr, "ERROR: Error during command line parsing\n"); return 1; }
        const char *arg = nullptr;
        if (o[1] == ':') {
            if (a[2]) arg = a + 2;
            else if (++i == argc) { fprintf(stderr, "ERROR: Error during command line parsing\n"); return 1; }
            else arg = argv[i];
        }
        switch (a[1]) {
            case 'c': config_dir = arg; break;
            case 'l': file_list = arg; break;
            case 'i': input_file = arg; break;
            case 'o': output_file = arg; break;
            case 'm': output_mlf = arg; break;
            case 'a': live = true; break;
            case 's': R.inf = str2fmt(arg); break;
            case 't': R.outf = str2fmt(arg); break;
            case 'w':
                if (strcmp(arg, "lin16") && strcmp(arg, "alaw"))
                    die("Invalid waveform format '%s'. Supported data formats are 'lin16' and 'alaw'.\n", arg);
                wformat = arg;
                break;
            case 'p': wpenalty = arg; break;
            case 'f':
                if (strcmp(arg, "lab") && strcmp(arg, "str") && strcmp(arg, "strlen")) {
                    fprintf(stderr, "ERROR: Invalid output format: %s. (can be 'lab', 'str', 'strlen')\n", arg);
                    return 1;
                }
                live_fmt = !strcmp(arg, "lab") ? ofLab : !strcmp(arg, "str") ? ofStr : ofStrLen;
                break;
            case 'v': g_verbose = true; break;
        }
    }
    // PHNREC_MLP=tc selects the tensor-core posterior estimator (not a reference switch)
    if (const char *e = getenv("PHNREC_MLP")) mlp_mode = !strcmp(e, "tc") ? PHN_MLP_TC_F16 : PHN_MLP_EXACT_FP32;
    // PHNREC_DEVICES = "all" | "0-7" | "0,2,5" (list mode shards the file list over them); PHNREC_DEVICE = one device
    std::vector<int> devices;
    if (const char *e = getenv("PHNREC_DEVICES")) {
        const int have = phn_device_count();
        if (!strcmp(e, "all")) { for (int d = 0; d < have; ++d) devices.push_back(d); }
        else {
            const char *q = e;
            while (*q) {
                char *end;
                const long a = strtol(q, &end, 10);
                if (end == q) break;
                long b = a;
                if (*end == '-') { q = end + 1; b = strtol(q, &end, 10); }
                for (long d = a; d <= b; ++d) devices.push_back((int)d);
                q = *end == ',' ? end + 1 : end;
                if (*end && *end != ',') break;
            }
        }
        for (int d : devices) if (d < 0 || d >= have) die("PHNREC_DEVICES names device %d; this machine has %d\n", d, have);
    }
    if (devices.empty()) devices.push_back(getenv("PHNREC_DEVICE") ? atoi(getenv("PHNREC_DEVICE")) : 0);
    const int device = devices[0];

    if (!config_dir) { fprintf(stderr, "ERROR: Configuration directory is not set (-c)\n"); return 1; }
    logmsg("\nSystem initialization\n");
    phase("arguments parsed");
    // contexts on the further devices (list mode, audio -> labels) are created side by side with the first one
    std::vector<std::thread> more_threads;
    std::vector<phn_ctx *> more(devices.size(), nullptr);
    std::vector<std::string> more_err(devices.size());
    if (file_list && R.inf == dfWaveform && R.outf == dfStrings)
        for (size_t d = 1; d < devices.size(); ++d)
            more_threads.emplace_back([&, d] {
                if (phn_create(config_dir, devices[d], &more[d])) { more_err[d] = phn_last_error(nullptr); more[d] = nullptr; }
            });
    if (phn_create(config_dir, device, &R.ctx)) die("%s", phn_last_error(nullptr));
    phase("first context created (CUDA context, model load)");
    phn_get_info(R.ctx, &R.info);
    logmsg("  - mel-banks ...\n  - online normalization ...\n  - posteriors (loading NNs) ...\n  - decoder ...\n\n");
    logmsg("------------------- SUMMARY -------------------\n");
    logmsg("Dictionary:   %s\n", phn_config_get(R.ctx, "dicts", "phoneme_list"));
    logmsg("Network file: %s\n", phn_config_get(R.ctx, "networks", "default"));
    logmsg("HMM file:     %s\n", phn_config_get(R.ctx, "models", "hmm_defs"));
    logmsg("#States/Phn:  %d\n", atoi(phn_config_get(R.ctx, "models", "nstates")));
    logmsg("Time pruning: %d\n", R.info.time_pruning);
    logmsg("Word penalty: %f\n", R.info.wpenalty);
    logmsg("Soft func:    %s\n", phn_config_get(R.ctx, "decoder", "softening_func"));
    logmsg("-----------------------------------------------\n\n");
    R.ck(phn_set_mlp_mode(R.ctx, mlp_mode));
    R.info_penalty = R.info.wpenalty;

    if (wpenalty) {
        float v;
        if (sscanf(wpenalty, "%f", &v) != 1) {
            fprintf(stderr, "ERROR: Invalid argument for -p switch at command line: %s\n", wpenalty);
            return 1;
        }
        phn_set_penalty(R.ctx, v);
        R.info_penalty = v;
    }
    if (wformat) phn_set_wave_format(R.ctx, !strcmp(wformat, "alaw") ? PHN_WAVE_ALAW : PHN_WAVE_LIN16);
    if (output_file && !input_file) { fprintf(stderr, "ERROR: The input file is not specified (-i)\n"); return 1; }
    if (!((int)R.outf > (int)R.inf)) { fprintf(stderr, "ERROR: Unsupported data conversion (-s, -t)\n"); return 1; }

    const std::string params_suffix = phn_config_get(R.ctx, "params", "suffix");
    const std::string labels_suffix = phn_config_get(R.ctx, "labels", "suffix");
    const bool remove_path = !strcmp(phn_config_get(R.ctx, "labels", "remove_path"), "true");

    // one list line -> a job (SpeechRec::ProcessFileListLine, srec.cpp:1201-1244)
    auto parse_line = [&](const char *line, bool in_mlf, Job &job, std::string &err) {
        char f1[1024], sep[256], f2[1024];
        if (sscanf(line, "%1023[^ \n\r\t]%255[ \t]%1023[^ \n\r\t]", f1, sep, f2) == 3) { job.src = f1; job.dst = f2; return true; }
        if (sscanf(line, "%1023s", f1) != 1) { err = std::string("Invalid line in file list: ") + line + "\n"; return false; }
        job.src = f1;
        switch (R.outf) {
            case dfParams: job.dst = change_suffix(f1, params_suffix); break;
            case dfPosteriors:
                // the reference asks for the unknown variable traps/suffix here and aborts (srec.cpp:1224)
                err = "Posterior output needs an explicit target file ('source target' list lines or -i/-o)\n";
                return false;
            case dfStrings:
                if (in_mlf) {  // CreateLabelFileNameForMLF, srec.cpp:1424-1436
                    std::string s = f1;
                    for (char &ch : s) if (ch == '\\') ch = '/';
                    s = change_suffix(s, labels_suffix);
                    job.dst = remove_path ? change_path(s, "*") : s;
                } else {
                    job.dst = change_suffix(f1, labels_suffix);
                }
                break;
            default: job.dst = f1; break;
        }
        return true;
    };

    if (input_file) {  // -i [-o]: phnrec.cpp:241-252
        std::string line = input_file;
        if (output_file) line += std::string(" ") + output_file;
        Job j;
        std::string err;
        if (!parse_line(line.c_str(), false, j, err)) die("%s", err.c_str());
        R.run({j}, "");
    }

    if (file_list) {  // SpeechRec::ProcessFileList, srec.cpp:1246-1291
        FILE *fl = fopen(file_list, "r");
        if (!fl) die("Can not open the file list: %s\n", file_list);
        if (output_mlf) {
            R.mlf = fopen(output_mlf, "w");
            if (!R.mlf) die("Can not create the MLF: %s\n", output_mlf);
            fprintf(R.mlf, "#!MLF!#\n");
        }
        size_t max_batch = 2048;
        if (const char *e = getenv("PHNREC_BATCH")) max_batch = (size_t)atol(e) > 0 ? (size_t)atol(e) : max_batch;
        char line[1024];
        std::string err;
        if (R.inf == dfWaveform && R.outf == dfStrings) {
            // audio -> labels: the multi-GPU list pipeline (one host thread and context per device, ordered gather)
            ListPipeline P;
            P.to_mlf = R.mlf != nullptr;
            while (fgets(line, 1023, fl)) {
                Job j;
                if (!parse_line(line, R.mlf != nullptr, j, err)) break;   // (the reference stops at the first bad line)
                P.jobs.push_back(j);
            }
            P.ctxs.push_back(R.ctx);
            for (auto &t : more_threads) t.join();
            more_threads.clear();
            phase("further contexts created");
            for (size_t d = 1; d < devices.size(); ++d) {   // same model, same settings
                if (!more[d]) die("%s", more_err[d].empty() ? "Can not create a context on a further device\n" : more_err[d].c_str());
                phn_set_mlp_mode(more[d], mlp_mode);
                phn_set_penalty(more[d], R.info_penalty);
                if (wformat) phn_set_wave_format(more[d], !strcmp(wformat, "alaw") ? PHN_WAVE_ALAW : PHN_WAVE_LIN16);
                P.ctxs.push_back(more[d]);
            }
            const size_t nd = P.ctxs.size();
            size_t per = (P.jobs.size() + 4 * nd - 1) / (4 * nd);   // about four batches per device: file reading overlaps the GPU
            if (per < 64) per = 64;
            if (per > max_batch) per = max_batch;
            for (size_t j0 = 0; j0 < P.jobs.size(); j0 += per) P.batches.emplace_back(j0, std::min(P.jobs.size(), j0 + per));
            unsigned hw = std::thread::hardware_concurrency();
            P.readers = (int)std::max<size_t>(1, std::min<size_t>(8, (hw ? hw : 8) / nd));
            if (const char *e = getenv("PHNREC_READERS")) P.readers = atoi(e) > 0 ? atoi(e) : P.readers;
            const std::string perr = P.run(R.mlf);
            phase("list processed, outputs written");
            for (size_t d = 1; d < P.ctxs.size(); ++d) phn_destroy(P.ctxs[d]);
            phase("further contexts destroyed");
            if (!perr.empty()) { if (R.mlf) fflush(R.mlf); die("%s", perr.c_str()); }
            if (!err.empty()) { if (R.mlf) fflush(R.mlf); die("%s", err.c_str()); }
        } else {
        std::vector<Job> batch;
        while (fgets(line, 1023, fl)) {
            Job j;
            if (!parse_line(line, R.mlf != nullptr, j, err)) break;
            batch.push_back(j);
            if (batch.size() >= max_batch) { R.run(batch, ""); batch.clear(); }
        }
        if (!batch.empty() || !err.empty()) {
            if (!batch.empty()) R.run(batch, err);
            else die("%s", err.c_str());
        }
        }
        if (R.mlf) fclose(R.mlf);
        fclose(fl);
    }

    for (auto &t : more_threads) t.join();   // (contexts nobody used: -i only)

    // -a: live input (phnrec.cpp:261-296 + SpeechRec::RunLive, srec.cpp:1438-1490).  The reference reads a sound card
    // in blocks of sample_freq / 8 samples and hands each to ProcessOnline; the labels come out of the decoder's
    // callback as TimePruning commits them (live_callback, phnrec.cpp:71-110).  Here the waveform source is stdin
    // (raw samples in the context's wave format: `arecord -t raw -f S16_LE -r 8000 | phnrec -c cfg -a`), one stream
    // of the streaming API takes the blocks, and each push's committed labels are printed in the callback's formats.
    // End of input ends the utterance the way ProcessOnline(.., last = true) does (tail flush + Decoder::Done); the
    // reference's loop only stops on a signal and then calls Done() without the tail.
    if (live) {
        const int bps = (wformat ? !strcmp(wformat, "alaw") : R.info.wave_format == PHN_WAVE_ALAW) ? 1 : 2;
        const size_t blk = (size_t)(R.info.sample_freq / 8) * bps;
        if (blk == 0) die("source/sample_freq is too small for live input\n");
        R.ck(phn_stream_open(R.ctx, 1));
        if (atoi(phn_config_get(R.ctx, "onlinenorm", "estim_interval")) != 0)
            printf("Estimation of normalization parameters, please speak ...\n");
        std::vector<char> buf(blk);
        std::vector<phn_label> lab(4096);
        const int sid = 0;
        auto emit = [&](int64_t n) {
            for (int64_t i = 0; i < n; ++i) {
                const phn_label &l = lab[(size_t)i];
                const long long start = (long long)l.start * 100000ll, stop = (long long)l.end * 100000ll;
                const char *word = phn_phoneme(R.ctx, l.phn);
                switch (live_fmt) {
                    case ofLab: fprintf(stdout, "%lli %lli %s %f\n", start, stop, word, l.like); break;
                    case ofStr: fprintf(stdout, " %s\n", word); break;
                    case ofStrLen: fprintf(stdout, " %s(%d)\n", word, (int)((stop - start) / 100000 + 1)); break;
                }
            }
            if (n) fflush(stdout);   // (a live display: every committed label is visible at once)
        };
        auto push = [&](size_t nbytes, int last) {
            int64_t boff[2] = {0, (int64_t)nbytes}, loff[2] = {0, 0};
            int rc = phn_stream_push(R.ctx, &sid, 1, buf.data(), boff, &last, lab.data(), (int64_t)lab.size(), loff);
            if (rc) die("%s", phn_last_error(R.ctx));
            emit(loff[1]);
        };
        for (;;) {   // (a pipe may deliver short reads: fill the block like WFS.read does)
            size_t got = 0;
            while (got < blk) {
                const size_t r = fread(buf.data() + got, 1, blk - got, stdin);
                if (r == 0) break;
                got += r;
            }
            if (got < blk) { push(got, 1); break; }
            push(got, 0);
        }
    }
    phn_destroy(R.ctx);
    phase("done");
    return 0;
}
