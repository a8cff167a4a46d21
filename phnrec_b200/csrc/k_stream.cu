// k_stream.cu — the state-carrying kernels of the streaming (online) path, many concurrent streams per launch.
//
// Replaces, for one pushed block per stream (SpeechRec::ProcessOnline / ProcessTail, srec.cpp:793-927):
//   * Normalization::ProcessFrame (norm.cpp:216-234) as a LIVE estimator: sums, count and the finished estimate live in the
//     stream's state between pushes (k_stream_norm);
//   * the 31-frame FIFO of Traps (traps.cpp:180-219): the last 30 normalised log-mel frames of a stream are its state; a push
//     assembles [copies of frame 0 for the virtual warm-up rows of a short utterance | history | new frames] into one
//     window per stream, on which the batch K-stc / K-mlp kernels run unchanged - a window row's clamped context is the
//     FIFO's content (k_stream_assemble);
//   * PhnDec (phndec.cpp:44-302) as a resumable machine: alphas, back pointers, lengths, the history ring (64 slots, the
//     reference keeps 41), mPrevAlpha and the frame counter are loaded, advanced over the block's new posterior rows with
//     TimePruning committing labels as they become final, and stored back; Done runs when the stream ends
//     (k_viterbi_stream).  Arithmetic and tie breaking are those of k_vit.cu (bit-identical labels and scores).
#include "internal.h"
#include "device_math.cuh"

#include <cfloat>

namespace phn {

// ------------------------------------------------------------------------------------------------ live normaliser
// state per stream: [4][nb] = running sum, running sum of squares, mean, inverse std; cnt = frames accumulated (UINT_MAX once
// the estimate has been made, like ChannelNormParams::Update, norm.cpp:139-148)
__global__ void k_stream_norm(float *__restrict__ mel, const int64_t *__restrict__ frame_off, const int *__restrict__ sid, int nb,
                              float *__restrict__ st, unsigned *__restrict__ cnt, unsigned interval, int mean_norm, int var_norm)
{
    const int i = blockIdx.x, b = threadIdx.x;
    const int s = sid[i];
    const int64_t f0 = frame_off[i], T = frame_off[i + 1] - f0;
    unsigned n = cnt[s];
    __syncthreads();   // (every thread has read the count before thread 0 rewrites it)
    if (b < nb) {
        float *q = st + (size_t)s * 4 * nb;
        float sx = q[b], sx2 = q[nb + b], mean = q[2 * nb + b], inv = q[3 * nb + b];
        for (int64_t t = 0; t < T; ++t) {
            float v = mel[(f0 + t) * nb + b];
            if (n < interval) {   // Accum (norm.cpp:92-110)
                sx = __fadd_rn(sx, v);
                sx2 = __fadd_rn(sx2, __fmul_rn(v, v));
                ++n;
            }
            if (interval != 0 && n == interval) {   // Update, before Norm: this frame is already normalised
                mean = __fdiv_rn(sx, (float)interval);
                inv = __fdiv_rn(1.0f, __fsqrt_rn(__fsub_rn(__fdiv_rn(sx2, (float)interval), __fmul_rn(mean, mean))));
                n = 0xffffffffu;
            }
            if (mean_norm) v = __fsub_rn(v, mean);   // Norm (norm.cpp:112-137); mean 0 / inverse std 1 until the estimate exists
            if (var_norm) v = __fmul_rn(v, inv);
            mel[(f0 + t) * nb + b] = v;
        }
        q[b] = sx; q[nb + b] = sx2; q[2 * nb + b] = mean; q[3 * nb + b] = inv;
    }
    if (b == 0) cnt[s] = n;
}

// ------------------------------------------------------------------------------------------------ windows
struct AsmArgs {
    const float *mel;            // new frames of this push, [sum new][nb]
    const int64_t *new_off;      // [n + 1]
    const int64_t *win_off;      // [n + 1] window offsets (rows) in `win`
    const int *sid, *pad, *hist; // per pushed stream: state slot, virtual warm-up rows in front, frames held in the history
    float *win;                  // [sum window rows][nb]
    float *st_hist;              // [n_streams][keep][nb]
    int nb, keep;                // keep = 2 x trap shift: the frames a later row's context may reach back to (30)
};

__global__ void __launch_bounds__(128) k_stream_assemble(AsmArgs a)
{
    const int i = blockIdx.x, nb = a.nb;
    const int s = a.sid[i], P = a.pad[i], H = a.hist[i];
    const int64_t n0 = a.new_off[i], N = a.new_off[i + 1] - n0;
    float *w = a.win + a.win_off[i] * nb;
    float *h = a.st_hist + (size_t)s * a.keep * nb;
    for (int64_t k = threadIdx.x; k < (int64_t)H * nb; k += blockDim.x) w[(int64_t)P * nb + k] = h[k];
    for (int64_t k = threadIdx.x; k < N * nb; k += blockDim.x) w[((int64_t)P + H) * nb + k] = a.mel[n0 * nb + k];
    __syncthreads();
    // the virtual warm-up rows of a short utterance: copies of its first frame (the FIFO starts filled with it, traps.cpp:182-199)
    for (int64_t k = threadIdx.x; k < (int64_t)P * nb; k += blockDim.x) w[k] = w[(int64_t)P * nb + k % nb];
    // new history: the last (up to) 30 real frames
    const int64_t R = H + N, keep = R < a.keep ? R : a.keep;
    __syncthreads();
    for (int64_t k = threadIdx.x; k < keep * nb; k += blockDim.x) h[k] = w[((int64_t)P + R - keep) * nb + k];
}

int launch_stream_norm(phn_ctx *c, int n, const int *d_sid, float *d_state, unsigned *d_cnt, int interval, int mean_norm, int var_norm)
{
    if (n == 0) return PHN_OK;
    k_stream_norm<<<n, 32, 0, c->stream>>>((float *)c->d_mel.p, (const int64_t *)c->d_frame_off.p, d_sid, c->nbanks, d_state, d_cnt,
                                          (unsigned)interval, mean_norm, var_norm);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

int launch_stream_assemble(phn_ctx *c, int n, const int64_t *d_new_off, const int64_t *d_win_off, const int *d_sid, const int *d_pad,
                           const int *d_hist, float *d_win, float *d_st_hist, int keep)
{
    if (n == 0) return PHN_OK;
    AsmArgs a{(const float *)c->d_mel.p, d_new_off, d_win_off, d_sid, d_pad, d_hist, d_win, d_st_hist, c->nbanks, keep};
    k_stream_assemble<<<n, 128, 0, c->stream>>>(a);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

// ------------------------------------------------------------------------------------------------ decoder
#define PHN_LN05 (-0.69314718055994530941723212145818f) /* phndec.cpp:9 */

__device__ __forceinline__ unsigned s_f2ord(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float s_ord2f(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct VitStreamArgs {
    const float *logp;         // row-major [rows][ld]: ln of the windows' posteriors (K-log)
    int ld, P, H;
    const int64_t *row0;       // [n] first row (in logp) this push feeds to the stream's decoder
    const int *count;          // [n] rows to feed
    const int *sid, *fresh, *last;   // [n] state slot; 1: the decoder starts here (PhnDec::Init); 1: the stream ends (Done)
    float wp;
    VitStreamState *st;
    phn_label *labels;
    const int64_t *lab_off;    // [n + 1] capacity offsets
    int *nlab;                 // [n]
};

template <int PPL>
__global__ void __launch_bounds__(32) k_viterbi_stream(VitStreamArgs a)
{
    const int i = blockIdx.x, lane = threadIdx.x;
    VitStreamState &S = a.st[a.sid[i]];
    const bool fresh = a.fresh[i] != 0;
    const int H = a.H;
    const float wp = a.wp;
    const unsigned FULL = 0xffffffffu;
    const unsigned ORD_FLOOR = s_f2ord(-FLT_MAX);
    __shared__ int s_hphn[64], s_hlen[64];
    __shared__ float s_halpha[64];
    __shared__ int s_bp, s_bl;

    float al[PPL][4];
    int pv[PPL][4], ln[PPL][4];
    bool valid[PPL];
#pragma unroll
    for (int r = 0; r < PPL; ++r) {
        const int ph = lane + 32 * r;
        valid[r] = ph < a.P;
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // PhnDec::Init (phndec.cpp:44-94) or the stored machine
            al[r][j] = fresh || !valid[r] ? (j == 0 ? wp : -FLT_MAX) : S.al[ph * 4 + j];
            pv[r][j] = fresh || !valid[r] ? -1 : S.pv[ph * 4 + j];
            ln[r][j] = fresh || !valid[r] ? 0 : S.ln[ph * 4 + j];
        }
    }
    for (int k = lane; k < 64; k += 32) {
        s_hphn[k] = fresh ? -1 : S.hphn[k];
        s_hlen[k] = fresh ? -1 : S.hlen[k];
        s_halpha[k] = fresh ? -1.0f : S.halpha[k];
    }
    int n = fresh ? 0 : S.n;                  // frames consumed so far (mNFrames)
    int last_mi = fresh ? -1 : S.last_mi;
    float prev_alpha = fresh ? 0.0f : S.prev_alpha;   // mPrevAlpha
    __syncwarp();
    // record of frame number m (1-based) lives in slot m & 63; m <= 0 is the initial shift-register content (phndec.cpp:71-78)
    auto Hphn = [&](int m) { return m >= 1 ? s_hphn[m & 63] : -1; };
    auto Hlen = [&](int m) { return m >= 1 ? s_hlen[m & 63] : -1; };
    auto Halpha = [&](int m) { return m >= 1 ? s_halpha[m & 63] : -1.0f; };

    phn_label *out = a.labels + a.lab_off[i];
    const int cap = (int)(a.lab_off[i + 1] - a.lab_off[i]);
    int count = 0;
    const int T = a.count[i];
    const float *rowp = a.logp + a.row0[i] * a.ld;
    for (int t = 0; t < T; ++t, rowp += a.ld) {
        float obs[PPL][3];
#pragma unroll
        for (int r = 0; r < PPL; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) obs[r][j] = valid[r] ? rowp[3 * (lane + 32 * r) + j] : 0.0f;
        // ---- PropagateInModels (phndec.cpp:96-119)
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            if (!valid[r]) continue;
#pragma unroll
            for (int j = 3; j > 0; --j) {
                const float cur = __fadd_rn(al[r][j], PHN_LN05);
                const float prv = __fadd_rn(al[r][j - 1], PHN_LN05);
                if (cur > prv) {
                    al[r][j] = __fadd_rn(cur, obs[r][j - 1]);
                    ln[r][j] += 1;
                } else {
                    al[r][j] = __fadd_rn(prv, obs[r][j - 1]);
                    pv[r][j] = pv[r][j - 1];
                    ln[r][j] = ln[r][j - 1] + 1;
                }
            }
        }
        // ---- PropagateInNetwork + AddHistory (phndec.cpp:121-158)
        float bv = -FLT_MAX;
        int bi = 0, b_prev = pv[0][3], b_len = ln[0][3];
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            const float v = al[r][3];
            if (valid[r] && v > bv) { bv = v; bi = lane + 32 * r; b_prev = pv[r][3]; b_len = ln[r][3]; }
        }
        const unsigned bo = bv > -FLT_MAX ? s_f2ord(bv + 0.0f) : ORD_FLOOR;
        const unsigned mo = __reduce_max_sync(FULL, bo);
        const int mi = (int)__reduce_min_sync(FULL, bo == mo ? (unsigned)bi : 0x7fffffffu);
        const float mx = mo == ORD_FLOOR ? -FLT_MAX : s_ord2f(mo);
        ++n;
        if ((mi & 31) == lane) { s_hphn[n & 63] = b_prev; s_hlen[n & 63] = b_len; s_halpha[n & 63] = mx; }
        const float entry = __fadd_rn(mx, wp);
#pragma unroll
        for (int r = 0; r < PPL; ++r) { al[r][0] = entry; pv[r][0] = mi; ln[r][0] = 0; }
        last_mi = mi;
        // ---- TimePruning (phndec.cpp:169-234) once the shift register is full
        if (n >= H + 1) {
            float cv = -FLT_MAX;
            int ci = 0x7fffffff, c_prev = 0, c_len = 1;
#pragma unroll
            for (int r = 0; r < PPL; ++r)
#pragma unroll
                for (int j = 1; j <= 3; ++j) {
                    const float v = al[r][j];
                    if (valid[r] && v > cv) { cv = v; ci = (lane + 32 * r) * 3 + (j - 1); c_prev = pv[r][j]; c_len = ln[r][j]; }
                }
            const unsigned co = cv > -FLT_MAX ? s_f2ord(cv + 0.0f) : ORD_FLOOR;
            const unsigned to = __reduce_max_sync(FULL, co);
            const int ti = (int)__reduce_min_sync(FULL, (co == to && ci != 0x7fffffff) ? (unsigned)ci : 0x7fffffffu);
            if (ti == 0x7fffffff) { if (lane == 0) { s_bl = 1; s_bp = 0; } }   // nothing above -FLT_MAX: the scan's initial (len 1, prev 0)
            else if (ci == ti) { s_bl = c_len; s_bp = c_prev; }
            __syncwarp();
            int o = H - s_bl, q = s_bp;
            while (o > 0) {
                const int uu = n - H + o;
                q = Hphn(uu);
                o -= Hlen(uu);
            }
            if (o == 0) {
                const int end = n - H;
                const float av = Halpha(end);
                if (lane == 0 && count < cap) {
                    out[count].phn = q;
                    out[count].start = end - Hlen(end);
                    out[count].end = end;
                    out[count].like = __fsub_rn(av, prev_alpha);
                }
                prev_alpha = av;
                ++count;
            }
            __syncwarp();   // (the walk's reads of s_bl / s_bp / the ring precede the next frame's writes)
        } else {
            __syncwarp();
        }
    }
    __syncwarp();
    if (a.last[i]) {   // Done (phndec.cpp:236-302): final traceback, at most H labels, emitted in time order
        const int Tn = n;
        int cnt = 0;
        {
            int o = H, q = Tn > 0 ? last_mi : -1;
            while (o > 0 && q != -1) {
                const int uu = Tn - H + o;
                q = Hphn(uu);
                o -= Hlen(uu);
                ++cnt;
            }
        }
        if (lane == 0) {
            int o = H, end = Tn, q = Tn > 0 ? last_mi : -1, idx = 0;
            while (o > 0 && q != -1) {
                const int uu = Tn - H + o;
                const int l = Hlen(uu);
                const float av = Halpha(uu);
                const int qq = Hphn(uu);
                o -= l;
                const float like = o > 0 ? __fsub_rn(av, Halpha(Tn - H + o)) : __fsub_rn(av, prev_alpha);
                const int pos = count + (cnt - 1 - idx);
                if (pos < cap) { out[pos].phn = q; out[pos].start = end - l; out[pos].end = end; out[pos].like = like; }
                end -= l;
                q = qq;
                ++idx;
            }
        }
        count += cnt;
    } else {
#pragma unroll
        for (int r = 0; r < PPL; ++r) {
            const int ph = lane + 32 * r;
            if (!valid[r]) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) { S.al[ph * 4 + j] = al[r][j]; S.pv[ph * 4 + j] = pv[r][j]; S.ln[ph * 4 + j] = ln[r][j]; }
        }
        for (int k = lane; k < 64; k += 32) { S.hphn[k] = s_hphn[k]; S.hlen[k] = s_hlen[k]; S.halpha[k] = s_halpha[k]; }
        if (lane == 0) { S.n = n; S.last_mi = last_mi; S.prev_alpha = prev_alpha; }
    }
    if (lane == 0) a.nlab[i] = count;
}

// ln of the windows' posteriors (the decoder soft function, srec.cpp:1088-1097), then the resumable decoder
int launch_stream_decode(phn_ctx *c, int n, int64_t rows, const int64_t *d_row0, const int *d_count, const int *d_sid, const int *d_fresh,
                         const int *d_last, VitStreamState *d_st, phn_label *d_labels, const int64_t *d_lab_off, int *d_nlab)
{
    if (n == 0) return PHN_OK;
    int rc = ensure(c, c->d_logp, sizeof(float) * (size_t)((rows + 127) / 128 * 128 + 128) * c->ldp);
    if (rc) return rc;
    if ((rc = launch_log_post(c, rows))) return rc;
    VitStreamArgs a;
    a.logp = (const float *)c->d_logp.p; a.ld = c->ldp; a.P = c->P; a.H = c->hist;
    a.row0 = d_row0; a.count = d_count; a.sid = d_sid; a.fresh = d_fresh; a.last = d_last;
    a.wp = c->wpenalty; a.st = d_st; a.labels = d_labels; a.lab_off = d_lab_off; a.nlab = d_nlab;
    switch ((c->P + 31) / 32) {
        case 1: k_viterbi_stream<1><<<n, 32, 0, c->stream>>>(a); break;
        case 2: k_viterbi_stream<2><<<n, 32, 0, c->stream>>>(a); break;
        case 3: k_viterbi_stream<3><<<n, 32, 0, c->stream>>>(a); break;
        case 4: k_viterbi_stream<4><<<n, 32, 0, c->stream>>>(a); break;
        default: return fail(c, PHN_ERR_UNSUPPORTED, "more than 128 phonemes\n");
    }
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_VIT] += 1;
    return PHN_OK;
}

}  // namespace phn
