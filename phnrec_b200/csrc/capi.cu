// capi.cu — the C ABI (include/phnrec_b200.h): context lifetime, batch planning, stage sequencing
// on one CUDA stream, host<->device copies.  No computation happens on the host.
#include "internal.h"
#include <algorithm>
#include <chrono>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/stat.h>

static std::string g_create_err;

namespace phn {

int fail(phn_ctx *c, int code, const char *fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

int ensure(phn_ctx *c, phn_ctx::Buf &b, size_t bytes)
{
    if (bytes <= b.cap) return PHN_OK;
    if (b.p) {
        PHN_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->vit_stream) PHN_CUDA(c, cudaStreamSynchronize(c->vit_stream));
        PHN_CUDA(c, cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(c, PHN_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s\n", want, cudaGetErrorString(e));
    }
    b.cap = want;
    PHN_CUDA(c, cudaMemsetAsync(b.p, 0, want, c->stream));  // padding columns rely on zero fill
    return PHN_OK;
}

// ---- PHNREC_TIMELINE (development aid): per batch, events 0 copy stream released, 1 first K-wave may start, 2 front end
// enqueued work done, 3 nets done, 4 decoder may start, 5 decoder done; host clock at 0 call entered, 1 call returned,
// 2 wait entered, 3 wait returned.
static double tl_now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void tl_begin(phn_ctx *c)
{
    if (c->tl_on < 0) c->tl_on = getenv("PHNREC_TIMELINE") != nullptr;
    c->tl_cur = -1;
    if (!c->tl_on || c->tl.size() >= 256) return;
    c->tl.emplace_back();
    c->tl_cur = (int)c->tl.size() - 1;
    c->tl[c->tl_cur].h[0] = tl_now();
}
static void tl_ev(phn_ctx *c, int i, cudaStream_t s)
{
    if (c->tl_cur < 0) return;
    cudaEvent_t &e = c->tl[c->tl_cur].e[i];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, s);
}
static void tl_host(phn_ctx *c, int row, int i) { if (row >= 0 && row < (int)c->tl.size()) c->tl[row].h[i] = tl_now(); }
static void tl_dump(phn_ctx *c)
{
    if (c->tl_on <= 0 || c->tl.empty()) return;
    FILE *f = fopen(getenv("PHNREC_TIMELINE"), "a");
    cudaEvent_t ref = nullptr;
    for (auto &r : c->tl) for (int i = 0; i < 6 && !ref; ++i) if (r.e[i]) ref = r.e[i];
    if (f) fprintf(f, "# batch | device ms since the first event: copy-go wave-go front-done nets-done vit-go vit-done copy0-done copyN-done | host ms since the first call: call-in call-out wait-in wait-out\n");
    const double h0 = c->tl[0].h[0];
    int k = 0;
    for (auto &r : c->tl) {
        if (f) fprintf(f, "%3d |", k++);
        for (int i = 0; i < 6; ++i) {
            float ms = -1.f;
            if (r.e[i] && ref && cudaEventElapsedTime(&ms, ref, r.e[i]) != cudaSuccess) { ms = -1.f; cudaGetLastError(); }
            if (f) fprintf(f, " %9.3f", ms);
        }
        for (int i = 0; i < 2; ++i) {
            float ms = -1.f;
            if (r.cg[i] && ref && cudaEventElapsedTime(&ms, ref, r.cg[i]) != cudaSuccess) { ms = -1.f; cudaGetLastError(); }
            if (f) fprintf(f, " %9.3f", ms);
        }
        if (f) fprintf(f, " |");
        for (int i = 0; i < 4; ++i) if (f) fprintf(f, " %9.3f", r.h[i] > 0 ? r.h[i] - h0 : -1.0);
        if (f) fprintf(f, "\n");
    }
    if (f) fclose(f);
    for (auto &r : c->tl) { for (auto &e : r.e) if (e) cudaEventDestroy(e); for (auto &e : r.cg) if (e) cudaEventDestroy(e); }
    c->tl.clear();
}

// Small per-batch uploads (offset tables, penalties) from pageable host memory: the driver stages them at call time (the
// host does not wait for the stream, measured: 0.2 ms per enqueued batch while the previous one runs) and they do not queue
// behind the audio chunks on the copy engine - page-locked staging + cudaMemcpyAsync was measured to do exactly that
// (K-wave of the first group then started only when the whole batch had landed).
int upload_small(phn_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    if (!bytes) return PHN_OK;
    PHN_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return PHN_OK;
}

template <typename T>
static int upload(phn_ctx *c, T **dst, const T *src, size_t n)
{
    PHN_CUDA(c, cudaMalloc((void **)dst, sizeof(T) * (n ? n : 1)));
    if (n) PHN_CUDA(c, cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    return PHN_OK;
}

static int up16(int n) { return (n + 15) / 16 * 16; }

static int upload_one_net(phn_ctx *c, const HostNet &h, DevNet &d, int kin)
{
    d.nin = h.nin; d.nhid = h.nhid; d.nout = h.nout; d.nin4 = h.nin4; d.nhid4 = h.nhid4; d.nout4 = h.nout4;
    d.kp = up16(h.nin4);
    d.ldh = up16(h.nhid4);
    d.w1h = nullptr; d.w2h = nullptr; d.nhidP = 0; d.noutP = 0;
    d.kin = kin;
    d.k1P = (kin + 2 + 63) / 64 * 64;   // + the two bias columns (k_mlp_tc.cu)
    int rc;
    if ((rc = upload(c, &d.w1, h.w1.data(), h.w1.size()))) return rc;
    if ((rc = upload(c, &d.w2, h.w2.data(), h.w2.size()))) return rc;
    if ((rc = upload(c, &d.b1, h.b1.data(), h.b1.size()))) return rc;
    if ((rc = upload(c, &d.b2, h.b2.data(), h.b2.size()))) return rc;
    if ((rc = upload(c, &d.mean, h.mean.data(), h.mean.size()))) return rc;
    if ((rc = upload(c, &d.dev, h.dev.data(), h.dev.size()))) return rc;
    return PHN_OK;
}

static int upload_net(phn_ctx *c, int which)
{
    const HostNet &h = c->hnet[which];
    // fp16 image width: the merger's image keeps its two halves 8-column aligned (k_mlp_tc.cu)
    const int kin = which == 2 && c->system == PHN_SYS_LCRC ? (c->hnet[0].nout + 7) / 8 * 8 + c->hnet[0].nout : h.nin;
    return upload_one_net(c, h, c->net[which], kin);
}

// frames of one utterance (srec.cpp:945)
static int64_t frames_of(const phn_ctx *c, int64_t nbytes)
{
    const int64_t n = c->fmt == PHN_WAVE_LIN16 ? nbytes / 2 : nbytes;
    if (c->stream_frames) return n >= c->vs ? (n - c->vs) / c->step + 1 : 0;   // MelBanks::GetFeatures fed in blocks: full frames only
    return n > c->vs ? (n - c->vs) / c->step + 1 : 1;
}

// ---- batch planning: offsets + device buffers for `n_pen` decoder passes
static int plan_frames(phn_ctx *c, const int64_t *frame_off, int n_utt, int n_pen)
{
    if (n_utt < 0 || n_pen < 1 || !frame_off) return fail(c, PHN_ERR_ARG, "invalid batch\n");
    c->post_valid = 0; c->logp_valid = 0; c->logp_layout = 0;   // (whatever an earlier batch left in d_post / d_logp is not this batch's)
    c->n_utt = n_utt; c->n_pen = n_pen;
    c->h_frame_off.assign(frame_off, frame_off + n_utt + 1);
    if (c->h_frame_off[0] != 0) return fail(c, PHN_ERR_ARG, "offsets must start at 0\n");
    for (int u = 0; u < n_utt; ++u)
        if (c->h_frame_off[u + 1] < c->h_frame_off[u]) return fail(c, PHN_ERR_ARG, "offsets must be non-decreasing\n");
    c->total_frames = c->h_frame_off[n_utt];
    const int nseg = n_utt * n_pen;
    c->h_lab_off.resize((size_t)nseg + 1);
    c->h_lab_off[0] = 0;
    for (int k = 0; k < n_pen; ++k)
        for (int u = 0; u < n_utt; ++u) {
            const int64_t T = c->h_frame_off[u + 1] - c->h_frame_off[u];
            c->h_lab_off[(size_t)k * n_utt + u + 1] = c->h_lab_off[(size_t)k * n_utt + u] + T + 48;
        }
    c->label_cap = c->h_lab_off[nseg];
    const int64_t F = c->total_frames;
    int rc;
    if ((rc = ensure(c, c->d_frame_off, sizeof(int64_t) * (n_utt + 1)))) return rc;
    if ((rc = ensure(c, c->d_lab_off, sizeof(int64_t) * (nseg + 1)))) return rc;
    if ((rc = ensure(c, c->d_mel, sizeof(float) * F * c->nbanks))) return rc;
    if ((rc = ensure(c, c->d_mean, sizeof(float) * (size_t)n_utt * c->nbanks))) return rc;
    if ((rc = ensure(c, c->d_post, sizeof(float) * F * c->ldp))) return rc;
    if ((rc = ensure(c, c->d_rec, (size_t)20 * F * n_pen))) return rc;
    if ((rc = ensure(c, c->d_pen, sizeof(float) * n_pen))) return rc;
    if ((rc = upload_small(c, c->d_frame_off.p, c->h_frame_off.data(), sizeof(int64_t) * (n_utt + 1), c->stream))) return rc;
    if ((rc = upload_small(c, c->d_lab_off.p, c->h_lab_off.data(), sizeof(int64_t) * (nseg + 1), c->stream))) return rc;
    return PHN_OK;
}

static int plan_audio(phn_ctx *c, const int64_t *byte_off, int n_utt)
{
    if (n_utt < 0 || !byte_off) return fail(c, PHN_ERR_ARG, "invalid batch\n");
    c->h_byte_off.assign(byte_off, byte_off + n_utt + 1);
    if (c->h_byte_off[0] != 0) return fail(c, PHN_ERR_ARG, "offsets must start at 0\n");
    std::vector<int64_t> fo((size_t)n_utt + 1, 0);
    for (int u = 0; u < n_utt; ++u) {
        const int64_t nb = byte_off[u + 1] - byte_off[u];
        if (nb < 0) return fail(c, PHN_ERR_ARG, "offsets must be non-decreasing\n");
        fo[u + 1] = fo[u] + frames_of(c, nb);
    }
    c->total_bytes = byte_off[n_utt];
    int rc;
    if ((rc = plan_frames(c, fo.data(), n_utt, 1))) return rc;
    if ((rc = ensure(c, c->d_byte_off, sizeof(int64_t) * (n_utt + 1)))) return rc;
    if ((rc = upload_small(c, c->d_byte_off.p, c->h_byte_off.data(), sizeof(int64_t) * (n_utt + 1), c->stream))) return rc;
    // frame pairs of the fp32 front end never straddle utterances (results must not depend on the batch around them)
    c->h_pair_off.assign((size_t)n_utt + 1, 0);
    for (int u = 0; u < n_utt; ++u) c->h_pair_off[u + 1] = c->h_pair_off[u] + (fo[u + 1] - fo[u] + 1) / 2;
    if ((rc = ensure(c, c->d_pair_off, sizeof(int64_t) * (n_utt + 1)))) return rc;
    if ((rc = upload_small(c, c->d_pair_off.p, c->h_pair_off.data(), sizeof(int64_t) * (n_utt + 1), c->stream))) return rc;
    return PHN_OK;
}

struct StageTimer {  // CUDA-event timing of one kernel family on the stream it is launched on (profiling only)
    phn_ctx *c; int fam; cudaStream_t s;
    StageTimer(phn_ctx *c_, int fam_, cudaStream_t s_ = nullptr) : c(c_), fam(fam_), s(s_ ? s_ : c_->stream) { if (c->profiling) cudaEventRecord(c->ev[2 * fam], s); }
    ~StageTimer()
    {
        if (!c->profiling) return;
        cudaEventRecord(c->ev[2 * fam + 1], s);
        cudaEventSynchronize(c->ev[2 * fam + 1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev[2 * fam], c->ev[2 * fam + 1]);
        c->k_ms[fam] += ms;
    }
};

static void reset_timing(phn_ctx *c)
{
    for (int i = 0; i < PHN_K_COUNT; ++i) { c->k_ms[i] = 0.f; c->k_launches[i] = 0; }
}

// mel (device, un-normalised) -> posteriors (device)
// workspace of the posterior estimator for the planned batch; returns the frames per MLP pass in c->chunk_frames
static int prepare_posteriors(phn_ctx *c)
{
    int rc;
    if (c->plp) return fail(c, PHN_ERR_UNSUPPORTED, "params/kind = plp is a parameterisation for `-t par` only: the TRAPS nets take mel-bank energies\n");
    const bool tc = c->mlp_mode == PHN_MLP_TC_F16;
    const int64_t F = c->total_frames;
    int64_t ch = tc ? (int64_t)1 << 20 : (int64_t)1 << 15;   // frames per pass of the MLP workspace (tensor-core: 1.4 KB per frame)
    if (const char *e = getenv("PHNREC_PASS_FRAMES")) {      // testing aid: force several passes on small batches (multiple of 128)
        const int64_t v = atoll(e) / 128 * 128;
        if (v >= 128) ch = v;
    }
    if (ch > F) ch = (F + 127) / 128 * 128;
    c->chunk_frames = ch;
    if (ch == 0) return PHN_OK;
    if (c->system != PHN_SYS_LCRC) {   // the other TRAPS systems: exact mode only
        if (tc) return fail(c, PHN_ERR_UNSUPPORTED, "the tensor-core mode implements the LCRC system only\n");
        int ldh = c->net[2].ldh;
        for (auto &d : c->dband) ldh = d.ldh > ldh ? d.ldh : ldh;
        if ((rc = ensure(c, c->d_xm, sizeof(float) * ch * c->net[2].kp))) return rc;
        if ((rc = ensure(c, c->d_h, sizeof(float) * ch * ldh))) return rc;
        if (!c->dband.empty() && (rc = ensure(c, c->d_xb, sizeof(float) * ch * c->dband[0].kp * c->dband.size()))) return rc;
        return PHN_OK;
    }
    if (tc) {
        if ((rc = mlp_tc_prepare(c))) return rc;
        if (c->fuse_logp && (rc = ensure(c, c->d_logp, sizeof(float) * (size_t)((F + 127) / 128 * 128 + 128) * c->ldp))) return rc;
        if ((rc = ensure(c, c->d_x0h, sizeof(__half) * ch * c->net[0].k1P))) return rc;
        if ((rc = ensure(c, c->d_x1h, sizeof(__half) * ch * c->net[1].k1P))) return rc;
        const void *xm_before = c->d_xmh.p;
        if ((rc = ensure(c, c->d_xmh, sizeof(__half) * ch * c->net[2].k1P))) return rc;
        if (c->d_xmh.p != xm_before && (rc = mlp_tc_fill_merger_bias(c, (int64_t)(c->d_xmh.cap / (sizeof(__half) * c->net[2].k1P)) / 128 * 128))) return rc;
    } else {
        if ((rc = ensure(c, c->d_x0, sizeof(float) * ch * c->net[0].kp))) return rc;
        if ((rc = ensure(c, c->d_x1, sizeof(float) * ch * c->net[1].kp))) return rc;
        if ((rc = ensure(c, c->d_xm, sizeof(float) * ch * c->net[2].kp))) return rc;
        if ((rc = ensure(c, c->d_h, sizeof(float) * ch * c->net[0].ldh))) return rc;
    }
    return PHN_OK;
}

// mel (device, un-normalised) -> posteriors (device).  front_done: the caller has already run K-mean and K-stc for
// the whole batch (phn_recognize does so group by group under the audio copy; only when the batch is one MLP pass).
static int run_posteriors(phn_ctx *c, bool front_done = false)
{
    int rc;
    if (!front_done) {
        if ((rc = prepare_posteriors(c))) return rc;
        StageTimer t(c, PHN_K_MEAN);
        if ((rc = launch_sentence_mean(c))) return rc;
    }
    const bool tc = c->mlp_mode == PHN_MLP_TC_F16;
    const int64_t F = c->total_frames, ch = c->chunk_frames;
    if (ch == 0) return PHN_OK;
    c->logp_valid = 0;
    if (c->vit_pending) {   // the previous batch's decoder (side stream) still owns d_logp and its offset copies
        PHN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_vit_done, 0));
        c->vit_pending = 0;
    }
    for (int64_t f0 = 0; f0 < F; f0 += ch) {
        const int64_t nf = F - f0 < ch ? F - f0 : ch;
        if (c->system != PHN_SYS_LCRC) {
            { StageTimer t(c, PHN_K_STC); if ((rc = launch_trap(c, f0, nf))) return rc; }
            { StageTimer t(c, PHN_K_MLP); if ((rc = launch_mlp_trap(c, f0, nf))) return rc; }
            continue;
        }
        if (!front_done) { StageTimer t(c, PHN_K_STC); if ((rc = launch_stc(c, f0, nf))) return rc; }
        { StageTimer t(c, PHN_K_MLP); if ((rc = tc ? launch_mlp_tc(c, f0, nf) : launch_mlp_exact(c, f0, nf))) return rc; }
    }
    c->logp_valid = tc && c->fuse_logp;
    c->post_valid = !c->logp_valid;
    return PHN_OK;
}

// Runs the decoder for the planned batch into the next result slot.
// side: on the context's decoder stream (audio -> labels path), so that the next call's front end may start while this
// decoder is still running; `stream` waits for ev_vit_done before its nets touch d_logp again (run_posteriors).  Everything
// the decoder reads that the NEXT batch's planning rewrites (frame and capacity offsets) is the slot's own copy.
static int run_decode(phn_ctx *c, const float *penalties, int n_pen, bool side = false)
{
    c->h_pen.resize(n_pen);  // member: stays alive until the (staged) copy has been consumed
    for (int k = 0; k < n_pen; ++k) c->h_pen[k] = penalties ? penalties[k] : c->wpenalty;
    if (c->vit_pending) {   // (a decode that does not follow run_posteriors, e.g. phn_decode_device after phn_recognize_device)
        PHN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_vit_done, 0));
        c->vit_pending = 0;
    }
    { int rc0; if ((rc0 = upload_small(c, c->d_pen.p, c->h_pen.data(), sizeof(float) * n_pen, c->stream))) return rc0; }
    if (getenv("PHNREC_VIT_INLINE")) side = false;   // (kernel development: decoder on the main stream)
    const int nseg = c->n_utt * n_pen;
    for (int i = 0; i < c->n_pend; ++i)
        if (c->pend[i] == c->slot_w) return fail(c, PHN_ERR_ARG, "both result slots hold asynchronous batches that have not been waited for (phn_wait)\n");
    phn_ctx::DecSlot &sl = c->slot[c->slot_w];
    c->slot_last = c->slot_w;
    c->slot_w ^= 1;
    int rc;
    if ((rc = ensure(c, sl.d_labels, sizeof(phn_label) * (size_t)c->label_cap))) return rc;
    if ((rc = ensure(c, sl.d_nlab, sizeof(int) * (size_t)nseg))) return rc;
    if ((rc = ensure(c, sl.d_frame_off, sizeof(int64_t) * (c->n_utt + 1)))) return rc;
    if ((rc = ensure(c, sl.d_lab_off, sizeof(int64_t) * (nseg + 1)))) return rc;
    if ((rc = ensure(c, sl.d_coff, sizeof(int64_t) * (nseg + 2)))) return rc;
    if ((rc = ensure(c, sl.d_labels_c, sizeof(phn_label) * (size_t)c->label_cap))) return rc;
    if (!sl.done) PHN_CUDA(c, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    sl.h_lab_off = c->h_lab_off; sl.h_frame_off = c->h_frame_off;
    sl.n_utt = c->n_utt; sl.n_pen = n_pen;
    PHN_CUDA(c, cudaMemcpyAsync(sl.d_frame_off.p, c->d_frame_off.p, sizeof(int64_t) * (c->n_utt + 1), cudaMemcpyDeviceToDevice, c->stream));
    PHN_CUDA(c, cudaMemcpyAsync(sl.d_lab_off.p, c->d_lab_off.p, sizeof(int64_t) * (nseg + 1), cudaMemcpyDeviceToDevice, c->stream));
    if (!side) {
        sl.s = c->stream;
        {
            StageTimer t(c, PHN_K_VIT);
            rc = launch_viterbi(c, (const float *)c->d_pen.p, n_pen, c->stream, sl);
            if (!rc) rc = launch_compact_labels(c, nseg, sl);
        }
        if (rc) return rc;
        PHN_CUDA(c, cudaEventRecord(sl.done, c->stream));
        return PHN_OK;
    }
    tl_ev(c, 3, c->stream);
    PHN_CUDA(c, cudaEventRecord(c->ev_mlp_done, c->stream));
    PHN_CUDA(c, cudaStreamWaitEvent(c->vit_stream, c->ev_mlp_done, 0));
    sl.s = c->vit_stream;
    tl_ev(c, 4, c->vit_stream);
    {
        StageTimer t(c, PHN_K_VIT, c->vit_stream);
        rc = launch_viterbi(c, (const float *)c->d_pen.p, n_pen, c->vit_stream, sl);
        if (!rc) rc = launch_compact_labels(c, nseg, sl);
    }
    if (rc) return rc;
    PHN_CUDA(c, cudaEventRecord(sl.done, c->vit_stream));
    tl_ev(c, 5, c->vit_stream);
    PHN_CUDA(c, cudaEventRecord(c->ev_vit_done, c->vit_stream));
    c->vit_pending = 1;
    return PHN_OK;
}

// Labels of one result slot -> host (penalty-major segments): the compact offsets, then the labels themselves - two D2H
// copies on the fetch stream behind the slot's `done` event (the decoder's stream may already hold the next batch).
static int fetch_slot(phn_ctx *c, phn_ctx::DecSlot &sl, phn_label *labels, int64_t label_cap, int64_t *label_off)
{
    const int nseg = sl.n_utt * sl.n_pen;
    cudaStream_t s = c->fetch_stream;
    std::vector<int64_t> off((size_t)nseg + 2, 0);
    PHN_CUDA(c, cudaStreamWaitEvent(s, sl.done, 0));
    PHN_CUDA(c, cudaMemcpyAsync(off.data(), sl.d_coff.p, sizeof(int64_t) * (nseg + 2), cudaMemcpyDeviceToHost, s));
    PHN_CUDA(c, cudaStreamSynchronize(s));
    if (off[(size_t)nseg + 1]) return fail(c, PHN_ERR_CAPACITY, "internal label capacity exceeded (segment %lld)\n", (long long)off[(size_t)nseg + 1] - 1);
    const int64_t total = off[nseg];
    if (label_off) memcpy(label_off, off.data(), sizeof(int64_t) * (nseg + 1));
    if (!labels) return PHN_OK;
    if (label_cap < total) return fail(c, PHN_ERR_CAPACITY, "label buffer too small: %lld needed\n", (long long)total);
    if (total == 0) return PHN_OK;
    PHN_CUDA(c, cudaMemcpyAsync(labels, sl.d_labels_c.p, sizeof(phn_label) * (size_t)total, cudaMemcpyDeviceToHost, s));
    PHN_CUDA(c, cudaStreamSynchronize(s));
    return PHN_OK;
}

}  // namespace phn

using namespace phn;

// ===================================================================== C ABI
extern "C" {

const char *phn_version(void) { return "phnrec_b200 0.1 (sm_100a)"; }

int phn_convert_weights(const char *weights_path, const char *norms_path, const char *nbin_out)
{
    if (!weights_path || !nbin_out) return PHN_ERR_ARG;
    HostNet n;
    int rc = n.load_ascii(weights_path, norms_path ? norms_path : "");
    if (rc != PHN_OK) return rc;
    return n.save_nbin(nbin_out);
}

int phn_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *phn_last_error(const phn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int phn_create(const char *cfg_dir, int device, phn_ctx **out)
{
    if (!cfg_dir || !out) return fail(nullptr, PHN_ERR_ARG, "phn_create: null argument\n");
    *out = nullptr;
    phn_ctx *c = new phn_ctx();
    auto bail = [&](int rc) {
        g_create_err = c->err;
        phn_destroy(c);
        return rc;
    };
    c->cfg_dir = cfg_dir;
    c->device = device;
    const std::string cfg_file = c->cfg_dir + "/config";
    int line = 0;
    int rc = c->cfg.load(cfg_file, &line);
    if (rc != PHN_OK) {  // messages of SpeechRec::Init, srec.cpp:248-262
        switch (rc) {
            case PHN_ERR_CFG_UNKVAR: fail(c, rc, "Unknown variable in configuration file '%s', line %d\n", cfg_file.c_str(), line); break;
            case PHN_ERR_CFG_BADVAL: fail(c, rc, "Invalid argument for a vatiable in configuration file '%s', line %d\n", cfg_file.c_str(), line); break;
            case PHN_ERR_CFG_FILE: fail(c, rc, "Can not open configuration file '%s'\n", cfg_file.c_str()); break;
            default: fail(c, rc, "Invalid notation of variable in configuration file '%s', line %d\n", cfg_file.c_str(), line); break;
        }
        return bail(rc);
    }
    {   // SubstVars (srec.cpp:219-233): $C = config dir, $T = tmp dir, for the path-like variables
        auto subst = [&](const char *sec, const char *var, const std::string &tdir) {
            std::string &v = c->cfg.kv[std::string(sec) + "/" + var];
            size_t p;
            while ((p = v.find("$C")) != std::string::npos) v.replace(p, 2, c->cfg_dir);
            while ((p = v.find("$T")) != std::string::npos) v.replace(p, 2, tdir);
        };
        subst("dirs", "tmp", "");
        const std::string tdir = c->cfg.str("dirs", "tmp");
        const char *paths[][2] = {{"models", "hmm_defs"}, {"dicts", "phoneme_list"}, {"networks", "default"},
                                  {"dicts", "lexicon1"}, {"dicts", "lexicon2"}, {"dicts", "keyword_list"},
                                  {"kws", "thresholds_file"}, {"gptransc", "rules"}, {"gptransc", "symbols"},
                                  {"onlinenorm", "file"}};
        for (auto &pv : paths) subst(pv[0], pv[1], tdir);
        mkdir(tdir.c_str(), 0777);  // attempted, failure ignored (srec.cpp:270-279)
    }
    const Config &C = c->cfg;
    // ---- what this hot path implements: fbanks -> LCRC -> phndec, offline
    if (C.str("params", "kind") == "plp") {   // srec.cpp:563-583 (compiled out of the reference's PHNREC_ONLY build)
        c->plp = 1;
        c->plp_order = C.i("plp", "order"); c->plp_compress = C.f("plp", "compress_fact"); c->plp_lifter = C.f("plp", "cep_lifter");
        c->plp_scale = C.f("plp", "cep_scale"); c->plp_add_c0 = C.b("plp", "add_c0");
        if (c->plp_order < 1 || c->plp_order > 31) return bail(fail(c, PHN_ERR_CFG_BADVAL, "plp/order out of range (1..31)\n"));
    } else if (C.str("params", "kind") != "fbanks")
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unknown parameterization (parameters/kind): '%s'\n", C.str("params", "kind").c_str()));
    if (C.i("melbanks", "nbanks_full") != -1 && C.i("melbanks", "nbanks_full") != C.i("melbanks", "nbanks"))
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "melbanks/nbanks_full other than nbanks is not supported\n"));
    {   // Traps::SetSystem (traps.cpp:572-585)
        const std::string &sy = C.str("posteriors", "system");
        if (sy == "LCRC") c->system = PHN_SYS_LCRC;
        else if (sy == "1BT") c->system = PHN_SYS_1BT;
        else if (sy == "1BT_DCT") c->system = PHN_SYS_1BT_DCT;
        else if (sy == "3BT") c->system = PHN_SYS_3BT;
        else return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unknown system, check configuration: %s\n", sy.c_str()));
    }
    if (C.str("decoder", "type") != "phndec") return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unsupported decoder type: %s (only phndec)\n", C.str("decoder", "type").c_str()));
    c->trap_len = C.i("posteriors", "length");
    c->tshift = (c->trap_len - 1) / 2;          // Traps::GetTrapShift, traps.h:67
    c->use_hamming = C.b("posteriors", "hamming");
    c->add_c0 = C.b("posteriors", "add_c0");
    if (c->system == PHN_SYS_LCRC && (c->trap_len != 31 || !c->add_c0 || c->use_hamming))
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unsupported LCRC set-up (needs length=31, add_c0=true, hamming=false)\n"));
    if (c->trap_len < 3 || c->trap_len > 255 || !(c->trap_len & 1))
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "posteriors/length must be odd and in 3..255\n"));
    if (C.str("posteriors", "softening_func").rfind("none", 0) != 0 || C.str("decoder", "softening_func").rfind("log", 0) != 0)
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unsupported softening functions (needs posteriors none, decoder log)\n"));
    if (C.b("offlinenorm", "sent_var_norm") || C.b("offlinenorm", "sent_max_norm") || C.b("offlinenorm", "sent_chmax_norm"))
        return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unsupported offline normalisation (only sent_mean_norm)\n"));
    if (C.f("source", "noise_level") != 0.0f) return bail(fail(c, PHN_ERR_UNSUPPORTED, "source/noise_level is not supported\n"));
    if (C.i("decoder", "num_states_per_phn") != 3) return bail(fail(c, PHN_ERR_UNSUPPORTED, "decoder/num_states_per_phn must be 3\n"));
    const std::string &fmt = C.str("source", "format");
    if (fmt == "lin16") c->fmt = PHN_WAVE_LIN16;
    else if (fmt == "alaw") c->fmt = PHN_WAVE_ALAW;
    else return bail(fail(c, PHN_ERR_UNSUPPORTED, "Unknown waveform format: %s\n", fmt.c_str()));
    c->fs = C.i("source", "sample_freq");
    c->scale = C.f("source", "scale");
    c->dc_shift = C.f("source", "dc_shift");
    c->nbanks = C.i("melbanks", "nbanks");
    c->vs = C.i("melbanks", "vector_size");
    c->step = C.i("melbanks", "vector_step");
    c->lo = C.f("melbanks", "lower_freq");
    c->hi = C.f("melbanks", "higher_freq");
    c->preem = C.f("melbanks", "preem_coef");
    c->z_mean = C.b("melbanks", "z_mean_source");
    c->sent_mean_norm = C.b("offlinenorm", "sent_mean_norm");
    c->frame_shift = C.f("framenorm", "shift");
    c->frame_floor = C.f("framenorm", "min_floor");
    c->wpenalty = C.f("decoder", "wpenalty");
    c->hist = C.i("decoder", "time_pruning");
    c->on_interval = C.i("onlinenorm", "estim_interval");   // srec.cpp:594-601 (used by the streaming path only)
    c->on_mean = C.b("onlinenorm", "mean_norm");
    c->on_var = C.b("onlinenorm", "var_norm");
    c->bunch = atoi(C.str("posteriors", "bunch_size").c_str());
    c->S = 3;
    if (c->nbanks < 1 || c->nbanks > 32 || c->vs < 2 || c->vs > 4096 || c->step < 1 || c->hist < 1)
        return bail(fail(c, PHN_ERR_CFG_BADVAL, "Front-end sizes out of range in '%s'\n", cfg_file.c_str()));
    // ---- nets (traps.cpp:139-166: .nbin tried first; it is the only weight source we read)
    auto load_net = [&](HostNet &hn, const std::string &name) -> int {
        // NeuralNet::Load (nn.cpp:594-621): the binary cache first, else the ASCII pair, then the cache is written
        const std::string p = c->cfg_dir + "/weights/" + name + ".nbin";
        int r = hn.load(p);
        if (r != PHN_OK) {
            const std::string w = c->cfg_dir + "/weights/" + name + ".weights", nr = c->cfg_dir + "/norms/" + name + ".norms";
            r = hn.load_ascii(w, nr);
            if (r == PHN_OK) hn.save_nbin(p);   // (failure to write the cache is ignored, as in the reference)
        }
        if (r != PHN_OK) return fail(c, r, "Can not load neural network: %s\n", p.c_str());
        return PHN_OK;
    };
    if (c->system == PHN_SYS_LCRC) {
    const char *names[3] = {"band0", "band1", "merger"};
    for (int i = 0; i < 3; ++i)
        if ((rc = load_net(c->hnet[i], names[i]))) return bail(rc);
    if (c->hnet[0].nin % c->nbanks || c->hnet[0].nin != c->hnet[1].nin || c->hnet[0].nout != c->hnet[1].nout ||
        c->hnet[2].nin != 2 * c->hnet[0].nout)
        return bail(fail(c, PHN_ERR_NN_FORMAT, "Inconsistent network sizes in %s/weights\n", cfg_dir));
    c->ldp = (c->hnet[2].nout + 3) / 4 * 4;  // posterior row stride on the device (16-byte aligned rows)
    c->ncoef = c->hnet[0].nin / c->nbanks;
    if (c->ncoef != 11) return bail(fail(c, PHN_ERR_UNSUPPORTED, "band nets must take 11 coefficients per band\n"));
    for (int w = 0; w < 2; ++w) {  // traps.cpp:549-570
        const std::string p = c->cfg_dir + "/windows/band" + std::to_string(w) + ".window";
        FILE *f = fopen(p.c_str(), "r");
        if (!f) return bail(fail(c, PHN_ERR_DEC_INPUT, "Can not open window file: %s\n", p.c_str()));
        for (int i = 0; i < 16; ++i)
            if (fscanf(f, "%f", &c->win[w * 16 + i]) != 1) { fclose(f); return bail(fail(c, PHN_ERR_DEC_INPUT, "Invalid window file: %s\n", p.c_str())); }
        fclose(f);
    }
    } else {
        // the other TRAPS systems (traps.cpp:86-171): one net per band (1BT; 3BT: the first nbanks - 2 bands) or none (1BT_DCT), no windows
        c->trap_bands = c->system == PHN_SYS_3BT ? c->nbanks - 2 : c->nbanks;
        if (c->trap_bands < 1) return bail(fail(c, PHN_ERR_CFG_BADVAL, "too few mel banks for system 3BT\n"));
        if ((rc = load_net(c->hnet[2], "merger"))) return bail(rc);
        c->ldp = (c->hnet[2].nout + 3) / 4 * 4;
        if (c->system == PHN_SYS_1BT_DCT) {
            c->trap_shift_out = c->hnet[2].nin / c->trap_bands;   // merger_input_shift, traps.cpp:170
            const int nd = c->add_c0 ? c->trap_shift_out - 1 : c->trap_shift_out;
            if (c->trap_shift_out * c->trap_bands != c->hnet[2].nin || nd < (c->add_c0 ? 0 : 1))
                return bail(fail(c, PHN_ERR_NN_FORMAT, "merger inputs (%d) are not a whole number of coefficients per band (%d bands)\n", c->hnet[2].nin, c->trap_bands));
        } else {
            c->hband.resize((size_t)c->trap_bands);
            int tot = 0;
            for (int i = 0; i < c->trap_bands; ++i) {
                if ((rc = load_net(c->hband[i], "band" + std::to_string(i)))) return bail(rc);
                // (the reference copies ONE band's `length` values per net, traps.cpp:249-261: a net with any other input size
                // would read uninitialised memory there)
                if (c->hband[i].nin != c->trap_len)
                    return bail(fail(c, PHN_ERR_NN_FORMAT, "band net %d takes %d inputs; the trajectory has %d points\n", i, c->hband[i].nin, c->trap_len));
                tot += c->hband[i].nout;
            }
            if (tot != c->hnet[2].nin) return bail(fail(c, PHN_ERR_NN_FORMAT, "Inconsistent network sizes in %s/weights\n", cfg_dir));
        }
    }
    {   // phoneme list (phndec.cpp:305-350); $C substitution as srec.cpp:219-233
        const std::string p = C.str("dicts", "phoneme_list");
        FILE *f = fopen(p.c_str(), "r");
        if (!f) return bail(fail(c, PHN_ERR_DEC_INPUT, "Can not open the phoneme list: %s\n", p.c_str()));
        char buf[256];
        while (fgets(buf, 255, f)) {
            buf[strcspn(buf, "\r\n")] = 0;
            c->phonemes.push_back(buf);
        }
        fclose(f);
        c->P = (int)c->phonemes.size();
        if (c->P < 1 || c->P * 3 > c->hnet[2].nout)
            return bail(fail(c, PHN_ERR_DEC_INPUT, "Phoneme list %s does not fit the %d network outputs\n", p.c_str(), c->hnet[2].nout));
    }
    c->mt.build(c->nbanks, c->vs, c->step, c->fs, c->lo, c->hi);

    // ---- device
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return bail(fail(c, PHN_ERR_CUDA, "no CUDA device available (this library has no CPU path)\n"));
    if (device < 0 || device >= ndev) return bail(fail(c, PHN_ERR_ARG, "device %d out of range (0..%d)\n", device, ndev - 1));
    auto cu = [&](cudaError_t e, const char *what) -> int {
        if (e == cudaSuccess) return PHN_OK;
        return fail(c, PHN_ERR_CUDA, "CUDA failure in %s: %s\n", what, cudaGetErrorString(e));
    };
    if ((rc = cu(cudaSetDevice(device), "cudaSetDevice"))) return bail(rc);
    cudaDeviceProp prop;
    if ((rc = cu(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties"))) return bail(rc);
    if (prop.major != 10)
        return bail(fail(c, PHN_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only\n", device, prop.major, prop.minor));
    c->num_sms = prop.multiProcessorCount;
    if ((rc = cu(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"))) return bail(rc);
    if ((rc = cu(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate"))) return bail(rc);
    for (auto &e : c->ev_copy)
        if ((rc = cu(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"))) return bail(rc);
    if ((rc = cu(cudaEventCreateWithFlags(&c->ev_free, cudaEventDisableTiming), "cudaEventCreate"))) return bail(rc);
    if ((rc = cu(cudaStreamCreateWithFlags(&c->vit_stream, cudaStreamNonBlocking), "cudaStreamCreate"))) return bail(rc);
    if ((rc = cu(cudaStreamCreateWithFlags(&c->fetch_stream, cudaStreamNonBlocking), "cudaStreamCreate"))) return bail(rc);
    if ((rc = cu(cudaEventCreateWithFlags(&c->ev_mlp_done, cudaEventDisableTiming), "cudaEventCreate"))) return bail(rc);
    if ((rc = cu(cudaEventCreateWithFlags(&c->ev_vit_done, cudaEventDisableTiming), "cudaEventCreate"))) return bail(rc);
    if ((rc = cu(cudaEventCreateWithFlags(&c->ev_audio_free, cudaEventDisableTiming), "cudaEventCreate"))) return bail(rc);
    for (int i = 0; i < 2 * PHN_K_COUNT; ++i)
        if ((rc = cu(cudaEventCreate(&c->ev[i]), "cudaEventCreate"))) return bail(rc);
    for (int i = c->system == PHN_SYS_LCRC ? 0 : 2; i < 3; ++i)
        if ((rc = upload_net(c, i))) return bail(rc);
    c->dband.resize(c->hband.size());
    for (size_t i = 0; i < c->hband.size(); ++i)
        if ((rc = upload_one_net(c, c->hband[i], c->dband[i], c->hband[i].nin))) return bail(rc);
    const MelTables &mt = c->mt;
    std::vector<double2> tw((size_t)mt.N);
    for (int i = 0; i < mt.N - 1; ++i) tw[i] = make_double2(mt.tw[2 * i], mt.tw[2 * i + 1]);
    std::vector<float> dct(160);
    {   // sDCT's cosine arguments (dspc.h:206-221), evaluated with the host libm like the reference
        const float PiByN = (float)M_PI / 16.0f;
        for (int k = 0; k < 10; ++k) {
            const float v = PiByN * (float)(k + 1);
            for (int j = 0; j < 16; ++j) dct[k * 16 + j] = cosf(v * ((float)j + 0.5f));
        }
    }
    c->nparams = c->plp ? c->plp_order + (c->plp_add_c0 ? 1 : 0) : c->nbanks;
    if (c->plp) {   // PLPCoefs::Init (plp.cpp:38-70): equal-loudness curve, IDFT matrix, liftering window - the reference's float expressions
        const int nb = c->nbanks, dim = nb + 2, P = c->plp_order;
        std::vector<float> eql((size_t)nb), idft((size_t)(P + 1) * dim), lift((size_t)P);
        for (int i = 0; i < nb; ++i) {   // sEqualLaudnessCurve, dspc.h:235-244
            const float fsq = (mt.f0[i] * mt.f0[i]);
            const float fsub = fsq / (fsq + 1.6e5f);
            eql[i] = fsub * fsub * ((fsq + 1.44e6f) / (fsq + 9.61e6f));
        }
        const float angle = M_PI / (float)(dim - 1);   // CreateIDFTMatrix, plp.cpp:143-165 (cos on a float is cosf in C++)
        const float scl = 1.0f / (2.0f * (dim - 1));
        for (int i = 0; i <= P; ++i) {
            float *row = idft.data() + (size_t)i * dim;
            row[0] = 1.0f * scl;
            for (int j = 1; j < dim - 1; ++j) row[j] = (float)(2.0 * scl * cosf(angle * (float)i * (float)j));
            row[dim - 1] = scl * cosf(angle * (float)i * (float)(dim - 1));
        }
        const int Q = (int)c->plp_lifter;   // sLifteringWindow, dspc.cpp:326-335
        for (int i = 0; i < P; ++i) lift[i] = 1.0f + 0.5f * Q * sinf(M_PI * (float)(i + 1) / (float)Q);
        if ((rc = upload(c, &c->d_plp_eql, eql.data(), eql.size()))) return bail(rc);
        if ((rc = upload(c, &c->d_plp_idft, idft.data(), idft.size()))) return bail(rc);
        if ((rc = upload(c, &c->d_plp_lift, lift.data(), lift.size()))) return bail(rc);
    }
    if ((rc = upload(c, &c->tab.hamming, mt.hamming.data(), mt.hamming.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.coeffs, mt.coeffs.data(), mt.coeffs.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.banks, mt.banks.data(), mt.banks.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.bank_klo, mt.bank_klo.data(), mt.bank_klo.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.bank_khi, mt.bank_khi.data(), mt.bank_khi.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.tw, tw.data(), tw.size()))) return bail(rc);
    if ((rc = upload(c, &c->tab.win, c->win, 32))) return bail(rc);
    if ((rc = upload(c, &c->tab.dct, dct.data(), dct.size()))) return bail(rc);
    *out = c;
    return PHN_OK;
}

void phn_destroy(phn_ctx *c)
{
    if (!c) return;
    if (c->stream) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->vit_stream) cudaStreamSynchronize(c->vit_stream);
        tl_dump(c);
    }
    phn_ctx::Buf *bufs[] = {&c->d_par, &c->d_xb, &c->d_st_hist, &c->d_st_norm, &c->d_st_cnt, &c->d_st_vit, &c->d_st_args, &c->d_win, &c->d_st_labels, &c->d_st_nlab, &c->d_audio, &c->d_byte_off, &c->d_frame_off, &c->d_lab_off, &c->d_mel, &c->d_mean, &c->d_post,
                            &c->d_rec, &c->d_pen, &c->d_x0, &c->d_x1, &c->d_h, &c->d_xm,
                            &c->d_x0h, &c->d_x1h, &c->d_xmh, &c->d_tile_ctr, &c->d_logp, &c->d_pair_off,
                            &c->slot[0].d_labels, &c->slot[0].d_nlab, &c->slot[0].d_lab_off, &c->slot[0].d_frame_off, &c->slot[0].d_coff, &c->slot[0].d_labels_c,
                            &c->slot[1].d_labels, &c->slot[1].d_nlab, &c->slot[1].d_lab_off, &c->slot[1].d_frame_off, &c->slot[1].d_coff, &c->slot[1].d_labels_c};
    for (auto *b : bufs)
        if (b->p) cudaFree(b->p);
    mlp_tc_release(c);
    wave_tc_release(c);
    wave_tc16_release(c);
    if (c->tc_dbg) cudaFree(c->tc_dbg);
    if (c->stc_btab) cudaFree(c->stc_btab);
    if (c->stc_bias) cudaFree(c->stc_bias);
    if (c->stc_cf) cudaFree(c->stc_cf);
    if (c->stc_sb) cudaFree(c->stc_sb);
    for (auto &d : c->dband) {
        void *ps[] = {d.w1, d.w2, d.b1, d.b2, d.mean, d.dev};
        for (void *q : ps) if (q) cudaFree(q);
    }
    if (c->d_plp_eql) cudaFree(c->d_plp_eql);
    if (c->d_plp_idft) cudaFree(c->d_plp_idft);
    if (c->d_plp_lift) cudaFree(c->d_plp_lift);
    if (c->d_trap_ham) cudaFree(c->d_trap_ham);
    if (c->d_trap_cos) cudaFree(c->d_trap_cos);
    if (c->d_trap_pm) cudaFree(c->d_trap_pm);
    if (c->d_trap_pd) cudaFree(c->d_trap_pd);
    for (int i = 0; i < 3; ++i) {
        DevNet &d = c->net[i];
        void *ps[] = {d.w1, d.w2, d.b1, d.b2, d.mean, d.dev, d.w1h, d.w2h};
        for (void *p : ps)
            if (p) cudaFree(p);
    }
    void *ts[] = {c->tab.hamming, c->tab.coeffs, c->tab.banks, c->tab.bank_klo, c->tab.bank_khi, c->tab.tw, c->tab.win, c->tab.dct};
    for (void *p : ts)
        if (p) cudaFree(p);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_copy)
        if (e) cudaEventDestroy(e);
    if (c->ev_free) cudaEventDestroy(c->ev_free);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_mlp_done) cudaEventDestroy(c->ev_mlp_done);
    if (c->ev_vit_done) cudaEventDestroy(c->ev_vit_done);
    if (c->ev_audio_free) cudaEventDestroy(c->ev_audio_free);
    if (c->vit_stream) cudaStreamDestroy(c->vit_stream);
    if (c->fetch_stream) cudaStreamDestroy(c->fetch_stream);
    for (auto &sl : c->slot) if (sl.done) cudaEventDestroy(sl.done);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int phn_get_info(const phn_ctx *c, phn_info *o)
{
    if (!c || !o) return PHN_ERR_ARG;
    o->sample_freq = c->fs; o->wave_format = c->fmt; o->nbanks = c->nbanks; o->vector_size = c->vs;
    o->vector_step = c->step; o->fft_size = c->mt.N; o->n_phonemes = c->P; o->n_states = c->S;
    o->n_outputs = c->hnet[2].nout; o->band_inputs = c->system == PHN_SYS_LCRC ? c->hnet[0].nin : (c->hband.empty() ? 0 : c->hband[0].nin); o->merger_inputs = c->hnet[2].nin;
    o->hidden = c->system == PHN_SYS_LCRC ? c->hnet[0].nhid : c->hnet[2].nhid; o->sent_mean_norm = c->sent_mean_norm; o->time_pruning = c->hist;
    o->mlp_mode = c->mlp_mode; o->device = c->device; o->wpenalty = c->wpenalty; o->n_params = c->nparams;
    return PHN_OK;
}

const char *phn_phoneme(const phn_ctx *c, int i) { return (c && i >= 0 && i < c->P) ? c->phonemes[i].c_str() : nullptr; }

const char *phn_config_get(const phn_ctx *c, const char *sec, const char *var)
{
    if (!c || !sec || !var) return nullptr;
    auto it = c->cfg.kv.find(std::string(sec) + "/" + var);
    return it == c->cfg.kv.end() ? nullptr : it->second.c_str();
}

int phn_set_penalty(phn_ctx *c, float wp)
{
    if (!c) return PHN_ERR_ARG;
    c->wpenalty = wp;
    return PHN_OK;
}
int phn_set_wave_format(phn_ctx *c, int fmt)
{
    if (!c) return PHN_ERR_ARG;
    if (fmt != PHN_WAVE_LIN16 && fmt != PHN_WAVE_ALAW) return fail(c, PHN_ERR_ARG, "Unknown waveform format\n");
    c->fmt = fmt;
    return PHN_OK;
}
int phn_set_mlp_mode(phn_ctx *c, int mode)
{
    if (!c) return PHN_ERR_ARG;
    if (mode != PHN_MLP_EXACT_FP32 && mode != PHN_MLP_TC_F16) return fail(c, PHN_ERR_ARG, "Unknown MLP mode\n");
    if (mode == PHN_MLP_TC_F16 && c->system != PHN_SYS_LCRC) return fail(c, PHN_ERR_UNSUPPORTED, "the tensor-core mode implements the LCRC system only\n");
    c->mlp_mode = mode;
    return PHN_OK;
}
int phn_set_profiling(phn_ctx *c, int on)
{
    if (!c) return PHN_ERR_ARG;
    c->profiling = on;
    return PHN_OK;
}
int phn_last_timing(phn_ctx *c, float ms[PHN_K_COUNT], int64_t launches[PHN_K_COUNT])
{
    if (!c) return PHN_ERR_ARG;
    for (int i = 0; i < PHN_K_COUNT; ++i) {
        if (ms) ms[i] = c->k_ms[i];
        if (launches) launches[i] = c->k_launches[i];
    }
    return PHN_OK;
}

int64_t phn_num_frames(const phn_ctx *c, int64_t nbytes) { return c ? frames_of(c, nbytes) : -1; }

int64_t phn_label_capacity(const phn_ctx *c, const int64_t *frame_off, int n_utt)
{
    if (!c || !frame_off || n_utt < 0) return -1;
    return frame_off[n_utt] + (int64_t)48 * n_utt;
}

void *phn_stream(phn_ctx *c) { return c ? (void *)c->stream : nullptr; }
int phn_sync(phn_ctx *c)
{
    if (!c) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->vit_stream));
    return PHN_OK;
}
void *phn_device_alloc(phn_ctx *c, int64_t n)
{
    void *p = nullptr;
    if (!c || cudaSetDevice(c->device) != cudaSuccess || cudaMalloc(&p, (size_t)(n > 0 ? n : 1)) != cudaSuccess) return nullptr;
    return p;
}
void phn_device_free(phn_ctx *c, void *p)
{
    if (c && p) { cudaSetDevice(c->device); cudaFree(p); }
}
void *phn_host_alloc_pinned(int64_t n)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, (size_t)(n > 0 ? n : 1)) != cudaSuccess) return nullptr;
    return p;
}
void phn_host_free_pinned(void *p) { if (p) cudaFreeHost(p); }
int phn_memcpy_h2d(phn_ctx *c, void *dst, const void *src, int64_t n)
{
    if (!c) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    PHN_CUDA(c, cudaMemcpyAsync(dst, src, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return PHN_OK;
}
int phn_memcpy_d2h(phn_ctx *c, void *dst, const void *src, int64_t n)
{
    if (!c) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    PHN_CUDA(c, cudaMemcpyAsync(dst, src, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHN_OK;
}
int phn_synth_audio_device(phn_ctx *c, void *d_audio, int64_t bytes_per_utt, int n_utt, uint64_t seed)
{
    if (!c || !d_audio) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    return launch_synth(c, d_audio, bytes_per_utt, n_utt, seed);
}

// --------------------------------------------------------------------- stages
// mel (device) -> labels (device): everything behind K-wave on the audio -> labels path
static int recognize_after_wave(phn_ctx *c, bool front_done = false)
{
    c->fuse_logp = 1;
    c->fast_front = 1;
    int rc = run_posteriors(c, front_done);
    c->fuse_logp = 0;
    c->fast_front = 0;
    if (rc) return rc;
    rc = run_decode(c, nullptr, 1, true);
    c->logp_valid = 0;
    return rc;
}

int phn_recognize_device(phn_ctx *c, const void *d_audio, const int64_t *byte_off, int n_utt)
{
    if (!c || !d_audio) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    reset_timing(c);
    int rc;
    tl_begin(c);
    if ((rc = plan_audio(c, byte_off, n_utt))) return rc;
    tl_ev(c, 1, c->stream);
    { StageTimer t(c, PHN_K_WAVE); if ((rc = launch_wave(c, d_audio))) return rc; }
    rc = recognize_after_wave(c);
    tl_host(c, c->tl_cur, 1);
    return rc;
}

int phn_fetch_mel(phn_ctx *c, float *mel_out)
{
    if (!c || !mel_out) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    PHN_CUDA(c, cudaMemcpyAsync(mel_out, c->plp ? c->d_par.p : c->d_mel.p, sizeof(float) * c->total_frames * c->nparams, cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHN_OK;
}
int phn_fetch_posteriors(phn_ctx *c, float *post_out)
{
    if (!c || !post_out) return PHN_ERR_ARG;
    if (!c->post_valid)
        return fail(c, PHN_ERR_ARG, "no posteriors to fetch: the last call was the fused audio -> labels path of the tensor-core mode "
                                    "(it hands ln p straight to the decoder); use phn_posteriors()\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    const size_t rowb = sizeof(float) * c->net[2].nout;
    if (c->total_frames)
        PHN_CUDA(c, cudaMemcpy2DAsync(post_out, rowb, c->d_post.p, sizeof(float) * c->ldp, rowb, (size_t)c->total_frames, cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHN_OK;
}

// ln p exactly as the decoder of the last call consumed it (the first 3P columns of every frame), whichever kernel
// produced it: K-log (glibc logf of the linear posteriors) or the tensor-core merger's epilogue (tiled in HBM).
int phn_fetch_logp(phn_ctx *c, float *logp_out)
{
    if (!c || !logp_out) return PHN_ERR_ARG;
    if (!c->logp_layout) return fail(c, PHN_ERR_ARG, "no log-posteriors to fetch: no decode has run on the current batch\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    const int64_t F = c->total_frames;
    const int nc = 3 * c->P, ld = c->ldp;
    if (F == 0) return PHN_OK;
    if (c->logp_layout == 1) {
        PHN_CUDA(c, cudaMemcpy2DAsync(logp_out, sizeof(float) * nc, c->d_logp.p, sizeof(float) * ld, sizeof(float) * nc, (size_t)F, cudaMemcpyDeviceToHost, c->stream));
        PHN_CUDA(c, cudaStreamSynchronize(c->stream));
        return PHN_OK;
    }
    const int64_t tiles = (F + 127) / 128;
    std::vector<float> raw((size_t)tiles * ld * 128);
    PHN_CUDA(c, cudaMemcpyAsync(raw.data(), c->d_logp.p, sizeof(float) * raw.size(), cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int64_t f = 0; f < F; ++f) {
        const float *src = raw.data() + ((f >> 7) * ld) * 128 + (f & 127);
        for (int k = 0; k < nc; ++k) logp_out[f * nc + k] = src[(size_t)k * 128];
    }
    return PHN_OK;
}

int phn_fetch_labels(phn_ctx *c, phn_label *labels, int64_t label_cap, int64_t *label_off)
{
    if (!c) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    if (c->slot_last < 0) {   // nothing decoded yet: an empty result
        if (label_off) label_off[0] = 0;
        return PHN_OK;
    }
    return fetch_slot(c, c->slot[c->slot_last], labels, label_cap, label_off);
}

int phn_mel(phn_ctx *c, const void *audio, const int64_t *byte_off, int n_utt, float *mel_out, int64_t *frame_off)
{
    if (!c) return PHN_ERR_ARG;
    if (!byte_off || n_utt < 0) return fail(c, PHN_ERR_ARG, "invalid batch\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    reset_timing(c);
    int rc;
    if ((rc = plan_audio(c, byte_off, n_utt))) return rc;
    if (!audio && mel_out && c->total_bytes > 0) return fail(c, PHN_ERR_ARG, "null audio buffer\n");
    if (frame_off) memcpy(frame_off, c->h_frame_off.data(), sizeof(int64_t) * (n_utt + 1));
    if (!mel_out) return PHN_OK;
    if ((rc = ensure(c, c->d_audio, (size_t)c->total_bytes + 16))) return rc;
    if (c->total_bytes)
        PHN_CUDA(c, cudaMemcpyAsync(c->d_audio.p, audio, (size_t)c->total_bytes, cudaMemcpyHostToDevice, c->stream));
    c->force_exact_wave = 1;
    { StageTimer t(c, PHN_K_WAVE); rc = launch_wave(c, c->d_audio.p); if (!rc && c->plp) rc = launch_plp(c); }
    c->force_exact_wave = 0;
    if (rc) return rc;
    return phn_fetch_mel(c, mel_out);
}

int phn_posteriors(phn_ctx *c, const float *mel, const int64_t *frame_off, int n_utt, float *post_out)
{
    if (!c || !mel || !frame_off || !post_out) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    reset_timing(c);
    int rc;
    if ((rc = plan_frames(c, frame_off, n_utt, 1))) return rc;
    PHN_CUDA(c, cudaMemcpyAsync(c->d_mel.p, mel, sizeof(float) * c->total_frames * c->nbanks, cudaMemcpyHostToDevice, c->stream));
    if ((rc = run_posteriors(c))) return rc;
    return phn_fetch_posteriors(c, post_out);
}

int phn_decode(phn_ctx *c, const float *post, const int64_t *frame_off, int n_utt, const float *penalties, int n_pen,
               phn_label *labels, int64_t label_cap, int64_t *label_off)
{
    if (!c || !post || !frame_off) return PHN_ERR_ARG;
    if (!penalties) n_pen = 1;
    PHN_CUDA(c, cudaSetDevice(c->device));
    reset_timing(c);
    int rc;
    if ((rc = plan_frames(c, frame_off, n_utt, n_pen))) return rc;
    const size_t rowb = sizeof(float) * c->net[2].nout;
    if (c->total_frames)
        PHN_CUDA(c, cudaMemcpy2DAsync(c->d_post.p, sizeof(float) * c->ldp, post, rowb, rowb, (size_t)c->total_frames, cudaMemcpyHostToDevice, c->stream));
    c->logp_valid = 0;
    c->post_valid = 1;
    if ((rc = run_decode(c, penalties, n_pen))) return rc;
    return phn_fetch_labels(c, labels, label_cap, label_off);
}

// The penalty sweep on posteriors that are already in HBM (left there by phn_posteriors / a staged call): K-log once,
// then one decoder pass per penalty; labels stay on the device until phn_fetch_labels.
int phn_decode_device(phn_ctx *c, const float *penalties, int n_pen)
{
    if (!c) return PHN_ERR_ARG;
    if (!penalties) n_pen = 1;
    if (!c->post_valid) return fail(c, PHN_ERR_ARG, "no posteriors resident on the device (call phn_posteriors first)\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    reset_timing(c);
    int rc;
    const std::vector<int64_t> foff = c->h_frame_off;   // (plan_frames reassigns the member)
    if ((rc = plan_frames(c, foff.data(), c->n_utt, n_pen))) return rc;
    c->post_valid = 1;                                   // same frames, same buffer: nothing moved
    return run_decode(c, penalties, n_pen);
}

// audio (host) -> labels, everything enqueued, nothing waited for.  wait = false: the result slot is queued for phn_wait.
static int recognize_host(phn_ctx *c, const void *audio, const int64_t *byte_off, int n_utt, phn_label *labels, int64_t label_cap,
                          int64_t *label_off, int64_t *frame_off_out, bool wait)
{
    if (!c) return PHN_ERR_ARG;
    if (!byte_off || n_utt < 0) return fail(c, PHN_ERR_ARG, "invalid batch\n");
    if (wait && c->n_pend) return fail(c, PHN_ERR_ARG, "asynchronous batches are in flight: phn_wait for them first\n");
    if (!wait && c->n_pend >= 2) return fail(c, PHN_ERR_ARG, "two asynchronous batches are already in flight: phn_wait first\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    int rc;
    const bool trace = wait && getenv("PHNREC_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tt0 = now();
    struct TraceEvents {   // (destroyed on every exit path)
        cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
        bool on;
        explicit TraceEvents(bool on_) : on(on_) { if (on) for (auto &x : e) cudaEventCreate(&x); }
        ~TraceEvents() { if (on) for (auto &x : e) if (x) cudaEventDestroy(x); }
    } tev_holder(trace);
    cudaEvent_t *tev = tev_holder.e;
    reset_timing(c);
    tl_begin(c);
    // offsets are validated (start at 0, non-decreasing) before anything is sized by them or read through them
    if ((rc = plan_audio(c, byte_off, n_utt))) return rc;
    if (!audio && c->total_bytes > 0) return fail(c, PHN_ERR_ARG, "null audio buffer\n");
    const void *audio_before = c->d_audio.p;
    if ((rc = ensure(c, c->d_audio, (size_t)c->total_bytes + 16))) return rc;
    const double tt1 = now();
    if (trace) cudaEventRecord(tev[0], c->stream);
    // The audio goes up in groups of whole utterances on a copy stream; K-wave of group g starts as soon as the
    // group has landed and runs under the copy of group g+1, so only the first group's transfer is exposed.
    const int64_t total = byte_off[n_utt];
    int64_t group_mb = 8;
    if (const char *e = getenv("PHNREC_COPY_GROUP_MB")) group_mb = atoll(e) > 0 ? atoll(e) : group_mb;   // (kernel development)
    int ng = (int)(total / (group_mb << 20));
    ng = ng < 1 ? 1 : (ng > 16 ? 16 : ng);
    // A batch enqueued behind one that is still on the device: its audio lands while that batch's nets run, long before
    // its own front end gets the SMs - one launch per kernel over the whole batch instead of sixteen small ones.
    if (!wait && c->n_pend >= 1 && !getenv("PHNREC_ASYNC_GROUPS")) ng = 1;
    // When the batch is one pass of the tensor-core MLP, the sentence mean and the STC features of a group follow its
    // K-wave at once, so the whole front end runs under the copy and the MLP starts when the last group has landed.
    c->fuse_logp = 1; c->fast_front = 1;
    rc = prepare_posteriors(c);
    const bool front = rc == PHN_OK && c->mlp_mode == PHN_MLP_TC_F16 && c->chunk_frames >= c->total_frames && c->total_frames > 0 &&
                       (c->nbanks == 15 || c->nbanks == 23);   // (the row-range form of K-stc exists for the shipped bank counts)
    c->fuse_logp = 0; c->fast_front = 0;
    if (rc) return rc;
    // The copy may start as soon as the previous call's K-wave has read the buffer - while that call's nets are still
    // running (asynchronous calls) - unless the buffer was just regrown (its zero fill is work of `stream`).
    if (c->audio_free_valid && c->d_audio.p == audio_before) {
        PHN_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_audio_free, 0));
    } else {
        PHN_CUDA(c, cudaEventRecord(c->ev_free, c->stream));
        PHN_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_free, 0));
    }
    tl_ev(c, 0, c->copy_stream);
    {
        StageTimer t(c, PHN_K_WAVE);
        int u0 = 0;
        for (int g = 0; g < ng; ++g) {
            int u1 = u0;
            const int64_t want = total * (g + 1) / ng;
            while (u1 < n_utt && (byte_off[u1 + 1] <= want || g == ng - 1)) ++u1;
            if (g == ng - 1) u1 = n_utt;
            if (u1 == u0) continue;
            const int64_t b0 = byte_off[u0], nb = byte_off[u1] - b0;
            if (nb)
                PHN_CUDA(c, cudaMemcpyAsync((uint8_t *)c->d_audio.p + b0, (const uint8_t *)audio + b0, (size_t)nb, cudaMemcpyHostToDevice, c->copy_stream));
            PHN_CUDA(c, cudaEventRecord(c->ev_copy[g], c->copy_stream));
            if (c->tl_cur >= 0 && (u0 == 0 || g == ng - 1)) {
                cudaEvent_t &e = c->tl[c->tl_cur].cg[u0 == 0 ? 0 : 1];
                if (!e) cudaEventCreate(&e);
                cudaEventRecord(e, c->copy_stream);
            }
            PHN_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy[g], 0));
            if (u0 == 0) tl_ev(c, 1, c->stream);
            if ((rc = launch_wave(c, c->d_audio.p, u0, u1))) return rc;
            if (g == ng - 1) {   // the audio buffer has been consumed
                PHN_CUDA(c, cudaEventRecord(c->ev_audio_free, c->stream));
                c->audio_free_valid = 1;
            }
            if (front) {
                c->fast_front = 1;
                rc = launch_sentence_mean(c, u0, u1);
                if (!rc) rc = launch_stc(c, 0, c->total_frames, c->h_frame_off[u0], c->h_frame_off[u1]);
                c->fast_front = 0;
                if (rc) return rc;
            }
            u0 = u1;
        }
    }
    const double tt2 = now();
    if (trace) cudaEventRecord(tev[1], c->stream);
    tl_ev(c, 2, c->stream);
    if ((rc = recognize_after_wave(c, front))) return rc;
    tl_host(c, c->tl_cur, 1);
    if (!wait) {
        c->pend_tl[c->n_pend] = c->tl_cur;
        c->pend[c->n_pend++] = c->slot_last;
        return PHN_OK;
    }
    const double tt3 = now();
    if (trace) { cudaEventRecord(tev[2], c->stream); cudaStreamSynchronize(c->stream); }
    const double tt4 = now();
    if (frame_off_out) memcpy(frame_off_out, c->h_frame_off.data(), sizeof(int64_t) * (n_utt + 1));
    rc = phn_fetch_labels(c, labels, label_cap, label_off);
    const double tt5 = now();
    if (trace) {
        float e01 = 0, e12 = 0, ec = 0;
        cudaEventElapsedTime(&e01, tev[0], tev[1]); cudaEventElapsedTime(&e12, tev[1], tev[2]);
        (void)ng;
        FILE *tf = fopen(getenv("PHNREC_TRACE"), "a");
        if (tf) fprintf(tf, "[trace] n_utt %d host: plan %.3f enqueue-front %.3f enqueue-rest %.3f wait-gpu %.3f fetch %.3f total %.3f | gpu: front-end span %.3f (last copy landed at %.3f) mlp(+vit on its own stream) %.3f\n",
                n_utt, tt1 - tt0, tt2 - tt1, tt3 - tt2, tt4 - tt3, tt5 - tt4, tt5 - tt0, e01, ec, e12);
        if (tf) fclose(tf);
    }
    return rc;
}

int phn_recognize(phn_ctx *c, const void *audio, const int64_t *byte_off, int n_utt, phn_label *labels, int64_t label_cap,
                  int64_t *label_off, int64_t *frame_off_out)
{
    return recognize_host(c, audio, byte_off, n_utt, labels, label_cap, label_off, frame_off_out, true);
}

int phn_recognize_async(phn_ctx *c, const void *audio, const int64_t *byte_off, int n_utt)
{
    return recognize_host(c, audio, byte_off, n_utt, nullptr, 0, nullptr, nullptr, false);
}

int phn_wait(phn_ctx *c, phn_label *labels, int64_t label_cap, int64_t *label_off, int64_t *frame_off_out)
{
    if (!c) return PHN_ERR_ARG;
    if (!c->n_pend) return fail(c, PHN_ERR_ARG, "phn_wait: no asynchronous batch in flight\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    phn_ctx::DecSlot &sl = c->slot[c->pend[0]];
    if (frame_off_out) memcpy(frame_off_out, sl.h_frame_off.data(), sizeof(int64_t) * (sl.n_utt + 1));
    tl_host(c, c->pend_tl[0], 2);
    const int rc = fetch_slot(c, sl, labels, label_cap, label_off);
    tl_host(c, c->pend_tl[0], 3);
    // a capacity error leaves the batch queued (the caller may come back with a larger buffer); anything else retires it
    if (rc != PHN_ERR_CAPACITY || !labels) {
        c->pend[0] = c->pend[1]; c->pend_tl[0] = c->pend_tl[1];
        --c->n_pend;
    }
    return rc;
}

int phn_pending(const phn_ctx *c) { return c ? c->n_pend : 0; }

// ===================================================================== streaming (online) path
// SpeechRec::ProcessOnline / ProcessTail (srec.cpp:793-927) for many concurrent streams: every push hands one block of audio
// per stream to the GPU; the streams' state lives in the context between pushes (k_stream.cu).
int phn_stream_open(phn_ctx *c, int n_streams)
{
    if (!c || n_streams < 0) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    if (c->bunch < 1 || c->tshift % c->bunch != 0)
        return fail(c, PHN_ERR_UNSUPPORTED, "streaming needs a posteriors/bunch_size that divides the trap shift %d; it is %d: the reference's online path "
                                            "then hands warm-up rows to the decoder depending on where bunches fall\n", c->tshift, c->bunch);
    if (c->hist + 1 > 64) return fail(c, PHN_ERR_UNSUPPORTED, "streaming keeps a 64-slot decoder history; decoder/time_pruning is %d\n", c->hist);
    if (c->plp) return fail(c, PHN_ERR_UNSUPPORTED, "params/kind = plp is a parameterisation for `-t par` only\n");
    if (c->on_var && !c->on_mean) return fail(c, PHN_ERR_ARG, "online normalisation: var_norm without mean_norm (the reference asserts, norm.cpp:152)\n");
    if (c->cfg.str("onlinenorm", "file") != "none") return fail(c, PHN_ERR_UNSUPPORTED, "onlinenorm/file (XML persistence of the estimates) is not supported\n");
    if (c->cfg.b("onlinenorm", "scale_to_gvar")) return fail(c, PHN_ERR_UNSUPPORTED, "onlinenorm/scale_to_gvar is not supported\n");
    int rc;
    const size_t n = (size_t)n_streams;
    if ((rc = ensure(c, c->d_st_hist, sizeof(float) * n * 2 * c->tshift * c->nbanks))) return rc;
    if ((rc = ensure(c, c->d_st_norm, sizeof(float) * n * 4 * c->nbanks))) return rc;
    if ((rc = ensure(c, c->d_st_cnt, sizeof(unsigned) * n))) return rc;
    if ((rc = ensure(c, c->d_st_vit, sizeof(VitStreamState) * n))) return rc;
    c->streams.assign(n, phn_ctx::StreamHost());
    for (int s = 0; s < n_streams; ++s)
        if ((rc = phn_stream_reset(c, s))) return rc;
    return PHN_OK;
}

int phn_stream_count(const phn_ctx *c) { return c ? (int)c->streams.size() : 0; }

// A stream back to its initial state, the live normaliser included (a fresh Normalization::StartEstimation).
int phn_stream_reset(phn_ctx *c, int sid)
{
    if (!c) return PHN_ERR_ARG;
    if (sid < 0 || sid >= (int)c->streams.size()) return fail(c, PHN_ERR_ARG, "phn_stream_reset: no stream %d\n", sid);
    PHN_CUDA(c, cudaSetDevice(c->device));
    c->streams[sid] = phn_ctx::StreamHost();
    // ChannelNormParams::Null (norm.cpp:59-70): sums 0, mean 0, inverse std 1; nothing accumulated
    std::vector<float> z((size_t)4 * c->nbanks, 0.0f);
    for (int b = 0; b < c->nbanks; ++b) z[(size_t)3 * c->nbanks + b] = 1.0f;
    PHN_CUDA(c, cudaMemcpyAsync((float *)c->d_st_norm.p + (size_t)sid * 4 * c->nbanks, z.data(), sizeof(float) * z.size(), cudaMemcpyHostToDevice, c->stream));
    PHN_CUDA(c, cudaMemsetAsync((unsigned *)c->d_st_cnt.p + sid, 0, sizeof(unsigned), c->stream));
    return PHN_OK;
}

int phn_stream_push(phn_ctx *c, const int *sids, int n, const void *audio, const int64_t *byte_off, const int *last, phn_label *labels,
                    int64_t label_cap, int64_t *label_off)
{
    if (!c) return PHN_ERR_ARG;
    if (n < 0 || (n > 0 && (!sids || !byte_off || !label_off))) return fail(c, PHN_ERR_ARG, "phn_stream_push: invalid arguments\n");
    if (c->n_pend) return fail(c, PHN_ERR_ARG, "asynchronous batches are in flight: phn_wait for them first\n");
    PHN_CUDA(c, cudaSetDevice(c->device));
    if (label_off) label_off[0] = 0;
    if (n == 0) return PHN_OK;
    const int bps = c->fmt == PHN_WAVE_LIN16 ? 2 : 1;
    const int nstr = (int)c->streams.size();
    {   // every stream at most once per push
        std::vector<char> seen((size_t)nstr, 0);
        if (byte_off[0] != 0) return fail(c, PHN_ERR_ARG, "offsets must start at 0\n");
        for (int i = 0; i < n; ++i) {
            if (sids[i] < 0 || sids[i] >= nstr) return fail(c, PHN_ERR_ARG, "phn_stream_push: no stream %d (phn_stream_open)\n", sids[i]);
            if (seen[sids[i]]) return fail(c, PHN_ERR_ARG, "phn_stream_push: stream %d appears twice in one push\n", sids[i]);
            seen[sids[i]] = 1;
            if (byte_off[i + 1] < byte_off[i]) return fail(c, PHN_ERR_ARG, "offsets must be non-decreasing\n");
        }
        if (!audio && byte_off[n] > 0) return fail(c, PHN_ERR_ARG, "null audio buffer\n");
    }
    // ---- host: [tail | block] per stream (a lin16 block contributes floor(bytes / 2) samples taken from its start, like
    //      ConvertWaveformFormat called per block, srec.cpp:709-743), new frame counts, windows, decoder row ranges
    std::vector<int64_t> boff((size_t)n + 1, 0), noff((size_t)n + 1, 0), woff((size_t)n + 1, 0), row0((size_t)n), loff((size_t)n + 1, 0);
    std::vector<int> pad((size_t)n), hist((size_t)n), cnt((size_t)n), fresh((size_t)n), lastv((size_t)n);
    c->h_stage.clear();
    const int TS = c->tshift;   // the trap shift: rows wait for this many frames of right context (15 for the 31-frame systems)
    for (int i = 0; i < n; ++i) {
        phn_ctx::StreamHost &S = c->streams[sids[i]];
        const int64_t nb = (byte_off[i + 1] - byte_off[i]) / bps * bps;
        c->h_stage.insert(c->h_stage.end(), S.tail.begin(), S.tail.end());
        const uint8_t *src = (const uint8_t *)audio + byte_off[i];
        c->h_stage.insert(c->h_stage.end(), src, src + nb);
        boff[i + 1] = (int64_t)c->h_stage.size();
        const int64_t samples = (boff[i + 1] - boff[i]) / bps;
        const int64_t nf = samples >= c->vs ? (samples - c->vs) / c->step + 1 : 0;
        noff[i + 1] = noff[i] + nf;
        // what stays for the next push: everything from the first sample of the next frame on
        const int64_t used = nf * c->step * bps;
        S.tail.assign(c->h_stage.begin() + boff[i] + used, c->h_stage.begin() + boff[i + 1]);
        const bool is_last = last && last[i];
        const int64_t T = S.frames + nf;                       // frames of the stream so far
        const int64_t w0 = S.frames - S.hist;                  // global index of the window's first real frame
        const int64_t avail = is_last ? T : std::max<int64_t>(S.rows_done, T - TS);
        int64_t r_lo = S.rows_done;
        pad[i] = 0;
        if (is_last && T > 0 && T < TS && S.rows_done == 0) { pad[i] = (int)(TS - T); r_lo = -(int64_t)pad[i]; }   // the tail's warm-up rows
        hist[i] = S.hist;
        woff[i + 1] = woff[i] + pad[i] + S.hist + nf;
        row0[i] = woff[i] + pad[i] + (r_lo - w0);
        cnt[i] = (int)(avail - r_lo);
        fresh[i] = S.fed ? 0 : 1;
        lastv[i] = is_last ? 1 : 0;
        loff[i + 1] = loff[i] + cnt[i] + 48;
        S.fed = S.fed || cnt[i] > 0;
        S.frames = T;
        S.rows_done = avail;
        S.hist = (int)std::min<int64_t>(2 * TS, S.hist + nf);
        if (is_last) {   // the next block starts a new utterance on this stream; the normaliser's estimate stays (it outlives
            const std::vector<uint8_t> none;   // MelBanks / Traps / decoder resets in the reference as well)
            S = phn_ctx::StreamHost();
        }
    }
    // ---- audio -> new log-mel frames (streaming frame count), live normaliser
    int rc;
    reset_timing(c);
    c->stream_frames = 1;
    rc = plan_audio(c, boff.data(), n);
    c->stream_frames = 0;
    if (rc) return rc;
    if ((rc = ensure(c, c->d_audio, (size_t)c->total_bytes + 16))) return rc;
    if (c->total_bytes)
        PHN_CUDA(c, cudaMemcpyAsync(c->d_audio.p, c->h_stage.data(), (size_t)c->total_bytes, cudaMemcpyHostToDevice, c->stream));
    const bool tc = c->mlp_mode == PHN_MLP_TC_F16;
    c->force_exact_wave = tc ? 0 : 1;
    c->fast_front = tc ? 1 : 0;
    { StageTimer t(c, PHN_K_WAVE); rc = launch_wave(c, c->d_audio.p); }
    c->force_exact_wave = 0;
    c->fast_front = 0;
    if (rc) return rc;
    // per-push argument block: [sids | pad | hist | cnt | fresh | last] ints, then [noff | woff | row0 | loff] int64
    const size_t ni = (size_t)n;
    std::vector<int> ia(6 * ni);
    for (size_t i = 0; i < ni; ++i) { ia[i] = sids[i]; ia[ni + i] = pad[i]; ia[2 * ni + i] = hist[i]; ia[3 * ni + i] = cnt[i]; ia[4 * ni + i] = fresh[i]; ia[5 * ni + i] = lastv[i]; }
    const size_t ib = (sizeof(int) * 6 * ni + 15) / 16 * 16;
    std::vector<int64_t> la;
    la.insert(la.end(), noff.begin(), noff.end());
    la.insert(la.end(), woff.begin(), woff.end());
    la.insert(la.end(), row0.begin(), row0.end());
    la.insert(la.end(), loff.begin(), loff.end());
    if ((rc = ensure(c, c->d_st_args, ib + sizeof(int64_t) * la.size()))) return rc;
    uint8_t *dargs = (uint8_t *)c->d_st_args.p;
    if ((rc = upload_small(c, dargs, ia.data(), sizeof(int) * ia.size(), c->stream))) return rc;
    if ((rc = upload_small(c, dargs + ib, la.data(), sizeof(int64_t) * la.size(), c->stream))) return rc;
    const int *d_sid = (const int *)dargs, *d_pad = d_sid + ni, *d_hist = d_sid + 2 * ni, *d_cnt = d_sid + 3 * ni, *d_fresh = d_sid + 4 * ni, *d_last = d_sid + 5 * ni;
    const int64_t *d_noff = (const int64_t *)(dargs + ib), *d_woff = d_noff + (ni + 1), *d_row0 = d_woff + (ni + 1), *d_loff = d_row0 + ni;
    if ((rc = launch_stream_norm(c, n, d_sid, (float *)c->d_st_norm.p, (unsigned *)c->d_st_cnt.p, c->on_interval, c->on_mean, c->on_var))) return rc;
    // ---- windows [warm-up copies | history | new frames] -> the batch posterior estimator, no sentence normalisation
    const int64_t W = woff[n];
    if ((rc = ensure(c, c->d_win, sizeof(float) * (size_t)(W + 1) * c->nbanks))) return rc;
    if ((rc = launch_stream_assemble(c, n, d_noff, d_woff, d_sid, d_pad, d_hist, (float *)c->d_win.p, (float *)c->d_st_hist.p, 2 * TS))) return rc;
    if ((rc = plan_frames(c, woff.data(), n, 1))) return rc;
    if (W) PHN_CUDA(c, cudaMemcpyAsync(c->d_mel.p, c->d_win.p, sizeof(float) * (size_t)W * c->nbanks, cudaMemcpyDeviceToDevice, c->stream));
    const int smn = c->sent_mean_norm;
    c->sent_mean_norm = 0;
    rc = run_posteriors(c);
    c->sent_mean_norm = smn;
    if (rc) return rc;
    // ---- decoder: ln p, then every stream's machine over its new rows; labels committed by TimePruning (and Done) come back
    if ((rc = ensure(c, c->d_st_labels, sizeof(phn_label) * (size_t)loff[n]))) return rc;
    if ((rc = ensure(c, c->d_st_nlab, sizeof(int) * ni))) return rc;
    {
        StageTimer t(c, PHN_K_VIT);
        rc = launch_stream_decode(c, n, W, d_row0, d_cnt, d_sid, d_fresh, d_last, (VitStreamState *)c->d_st_vit.p, (phn_label *)c->d_st_labels.p, d_loff,
                                  (int *)c->d_st_nlab.p);
    }
    if (rc) return rc;
    std::vector<int> nl(ni);
    PHN_CUDA(c, cudaMemcpyAsync(nl.data(), c->d_st_nlab.p, sizeof(int) * ni, cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    int64_t total = 0;
    for (size_t i = 0; i < ni; ++i) {
        if (nl[i] > cnt[i] + 48) return fail(c, PHN_ERR_CAPACITY, "internal label capacity exceeded (stream %d)\n", sids[i]);
        total += nl[i];
        label_off[i + 1] = total;
    }
    if (!labels && total) return fail(c, PHN_ERR_CAPACITY, "label buffer too small: %lld needed\n", (long long)total);
    if (label_cap < total) return fail(c, PHN_ERR_CAPACITY, "label buffer too small: %lld needed\n", (long long)total);
    for (size_t i = 0; i < ni; ++i)
        if (nl[i])
            PHN_CUDA(c, cudaMemcpyAsync(labels + label_off[i], (const phn_label *)c->d_st_labels.p + loff[i], sizeof(phn_label) * (size_t)nl[i],
                                        cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHN_OK;
}

// Debug aid (not part of the stable ABI surface used by the reference binding): the next tensor-core launches of
// net `which` record clock64() timestamps of CTA 0's second tile into a 16 x 16 table; which < 0 reads it back.
int phn_debug_tc_timeline(phn_ctx *c, int which, long long *out /*[256]*/)
{
    if (!c) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    if (which >= 0) {
        if (!c->tc_dbg) PHN_CUDA(c, cudaMalloc(&c->tc_dbg, 256 * sizeof(long long)));
        PHN_CUDA(c, cudaMemsetAsync(c->tc_dbg, 0, 256 * sizeof(long long), c->stream));
        c->tc_dbg_net = which;
        return PHN_OK;
    }
    c->tc_dbg_net = -1;
    if (!c->tc_dbg || !out) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaMemcpyAsync(out, c->tc_dbg, 256 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PHN_CUDA(c, cudaStreamSynchronize(c->stream));
    return PHN_OK;
}

// Verification aid: the device's logf (the glibc port every ln on the exact path goes through) on the n consecutive float
// bit patterns starting at first_bits.
int phn_debug_logf(phn_ctx *c, uint32_t first_bits, int64_t n, float *out)
{
    if (!c || !out || n < 0) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    int rc;
    phn_ctx::Buf tmp;
    if ((rc = ensure(c, tmp, sizeof(float) * (size_t)n))) return rc;
    rc = launch_logf_range(c, first_bits, n, (float *)tmp.p);
    if (rc == PHN_OK && n) {
        cudaMemcpyAsync(out, tmp.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, PHN_ERR_CUDA, "CUDA failure in phn_debug_logf\n");
    }
    cudaFree(tmp.p);
    return rc;
}

int phn_online_norm(phn_ctx *c, float *x, int64_t frames, int nbanks, int interval, int mean_norm, int var_norm)
{
    if (!c || !x || frames < 0 || nbanks < 1) return PHN_ERR_ARG;
    PHN_CUDA(c, cudaSetDevice(c->device));
    int rc;
    phn_ctx::Buf tmp;
    if ((rc = ensure(c, tmp, sizeof(float) * frames * nbanks))) return rc;
    PHN_CUDA(c, cudaMemcpyAsync(tmp.p, x, sizeof(float) * frames * nbanks, cudaMemcpyHostToDevice, c->stream));
    rc = launch_online_norm(c, (float *)tmp.p, frames, nbanks, interval, mean_norm, var_norm);
    if (rc == PHN_OK) {
        cudaMemcpyAsync(x, tmp.p, sizeof(float) * frames * nbanks, cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    cudaFree(tmp.p);
    return rc;
}

}  // extern "C"
