// k_vit.cu — K-vit + K-trace: the phone-loop token-passing decoder, one warp (one CTA) per
// (utterance, penalty).
//
// Replaces the decoder soft function decSoftFunc=log (srec.h:192-195, srec.cpp:1088-1097) and
// PhnDec::{Init, PropagateInModels, PropagateInNetwork, AddHistory, GetBestToken, TimePruning,
// Done} (phndec.cpp:44-302).  All comparisons are the reference's strict fp32 `>` with
// first-index tie breaking; additions are single fp32 adds in the written order; log is the
// glibc logf port, so labels, boundaries AND scores are bit-identical to the reference when both
// are fed the same posteriors.
//
// Phase 1 (sequential in time): lane l owns phones l, l+32, ... with their 3 emitting states in
//   registers.  Per frame: state update, then redux.sync max/min over the warp for the best
//   phone-end token (mx, mi) and for the best token overall; the winners' (prev, len) go to a
//   per-frame record in global memory instead of the reference's 41-slot shift register.
// Phase 2 (parallel over frames): the partial traceback of TimePruning is an independent walk
//   for every frame n > H over those records (SURVEY §8a "verified restatement notes"); emitted
//   labels are compacted in time order with ballots, scores are differences of consecutive
//   commit alphas; lane 0 appends the final traceback of Done.
#include "internal.h"
#include "device_math.cuh"

#include <cfloat>

namespace phn {

struct VitArgs {
    const float *logp;       // log-posteriors, row-major [total_frames][ld] (K-log) or TILED [frame / 128][ld][128] (written
    int ld;                  // by the tensor-core merger's epilogue); the decoder reads the first 3P columns of a row
    const int64_t *frame_off;
    int n_utt, P, H;
    int64_t total_frames;
    const float *pen;        // [n_pen]
    int *r_hphn, *r_hlen, *r_bp, *r_bl;  // [n_pen * total_frames]
    float *r_halpha;
    phn_label *labels;
    const int64_t *lab_off;  // [n_pen*n_utt + 1] capacity offsets
    int *nlab;               // [n_pen*n_utt]
};

#define PHN_LN05 (-0.69314718055994530941723212145818f) /* phndec.cpp:9 */

__device__ __forceinline__ unsigned f2ord(float v)
{
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// K-log: the decoder soft function SoftLog (srec.h:192-195, applied srec.cpp:1088-1097) = glibc logf,
// no guard, on the 3P columns the decoder reads.  Fully parallel, so the sequential kernel below is
// pure token passing; a penalty sweep reuses the same log-posteriors.
__global__ void __launch_bounds__(256) k_log_post(const float *__restrict__ post, int ldp, int ncols, int64_t frames,
                                                 float *__restrict__ logp)
{
    __shared__ double s_logtab[32];
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int64_t f = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); f < frames; f += (int64_t)gridDim.x * 8)   // one warp per row
        for (int cidx = lane; cidx < ncols; cidx += 32) logp[f * ldp + cidx] = logf_glibc(post[f * ldp + cidx], s_logtab);
}

// Verification aid (phn_debug_logf): the device logf over a run of consecutive float bit patterns, so that a test can compare
// it with the host's libm over every float in (0, 1].
__global__ void __launch_bounds__(256) k_logf_range(uint32_t first_bits, int64_t n, float *__restrict__ out)
{
    __shared__ double s_logtab[32];
    logf_table_to_smem(s_logtab, threadIdx.x, blockDim.x);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = logf_glibc(__uint_as_float(first_bits + (uint32_t)i), s_logtab);
}

int launch_logf_range(phn_ctx *c, uint32_t first_bits, int64_t n, float *d_out)
{
    if (n <= 0) return PHN_OK;
    k_logf_range<<<c->num_sms * 8, 256, 0, c->stream>>>(first_bits, n, d_out);
    PHN_CUDA(c, cudaGetLastError());
    return PHN_OK;
}

// Register budget (kernel development switch): the decoder of batch k shares the SMs with the front end of batch k+1 (it runs
// on its own stream).  Capping it at 64 registers (PHN_VIT_MINB=32) so that its 7 warps per SM leave room for three K-wave
// CTAs was measured and dropped: the decoder alone goes from 0.52 to 0.80 ms (spills, less latency hiding) and the step
// from 4.74 to 5.02 ms.
#ifndef PHN_VIT_MINB
#define PHN_VIT_MINB 1
#endif
int launch_log_post(phn_ctx *c, int64_t rows)
{
    if (rows <= 0) return PHN_OK;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > (int64_t)c->num_sms * 16) blocks = (int64_t)c->num_sms * 16;
    k_log_post<<<(unsigned)blocks, 256, 0, c->stream>>>((const float *)c->d_post.p, c->ldp, 3 * c->P, rows, (float *)c->d_logp.p);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_VIT] += 1;
    return PHN_OK;
}

// LAYOUT of ln p: 0 row-major [frame][ld]; 1 the tensor-core merger's tiles (column-major inside 128-frame tiles) staged through
// shared-memory panels (default); 2 the same tiles read directly - no shared memory, at most 96 registers (PPL <= 2): a decoder
// whose one-warp CTAs fit beside the next batch's persistent MLP kernel (measured and not adopted, see launch_viterbi).
template <int PPL, int LAYOUT>
__global__ void __launch_bounds__(32, PPL <= 2 ? (LAYOUT == 2 ? 21 : PHN_VIT_MINB) : 1) k_viterbi(VitArgs a)
{
    const int seg = blockIdx.x;
    const int kpen = seg / a.n_utt, u = seg - kpen * a.n_utt;
    const int lane = threadIdx.x;
    const int64_t f0 = a.frame_off[u];
    const int T = (int)(a.frame_off[u + 1] - f0);
    const float wp = a.pen[kpen];
    const int H = a.H;
    const int64_t rbase = (int64_t)kpen * a.total_frames + f0;
    int *hphn = a.r_hphn + rbase, *hlen = a.r_hlen + rbase, *rbp = a.r_bp + rbase, *rbl = a.r_bl + rbase;
    float *halpha = a.r_halpha + rbase;
    const unsigned FULL = 0xffffffffu;
    const unsigned ORD_FLOOR = f2ord(-FLT_MAX);

    float al[PPL][4];
    int pv[PPL][4], ln[PPL][4];
    bool valid[PPL];
#pragma unroll
    for (int r = 0; r < PPL; ++r) {
        valid[r] = lane + 32 * r < a.P;
#pragma unroll
        for (int j = 0; j < 4; ++j) { al[r][j] = -FLT_MAX; pv[r][j] = -1; ln[r][j] = 0; }
        al[r][0] = wp;
    }
    int last_mi = -1;  // prev[0][0] after the last frame (phndec.cpp:240)

    constexpr bool TILED = LAYOUT == 1;
    constexpr int FB = 8;  // frames whose observations are fetched ahead of the recurrence
    // TILED input (column-major inside 128-frame tiles): a warp's row read would touch one 128-byte line per
    // column, so 16-frame panels [16 frames][3P columns] are staged through shared memory with 4-byte cp.async
    // (two columns x 16 consecutive frames = 2 x 64 contiguous bytes per instruction), a ring of three panels
    // indexed by global frame number, the next panel in flight while the recurrence consumes the current ones.
    extern __shared__ float s_panel[];                 // [48][pstride]
    const int ncol = 3 * a.P;
    const int pstride = ncol | 1;                      // odd row stride: conflict-free panel writes
    int64_t next_blk = f0 >> 4;                        // next 16-frame panel (global numbering) to request
    const int64_t last_blk = T > 0 ? (f0 + T - 1) >> 4 : -1;
    int next_slot = (int)(next_blk % 3);                // ring slot of that panel (= panel number mod 3)
    int row0 = (int)(f0 % 48);                          // ring row of the utterance's frame tb (kept incrementally)
    auto request_panel = [&](int64_t blk) {
        const int64_t F = blk * 16 + (lane & 15);
        const float *src = a.logp + ((F >> 7) * a.ld) * 128 + (F & 127);
        float *dst = s_panel + (size_t)(next_slot * 16 + (lane & 15)) * pstride;
        if (++next_slot == 3) next_slot = 0;
        for (int cc = lane >> 4; cc < ncol; cc += 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst + cc)), "l"(__cvta_generic_to_global(src + (size_t)cc * 128)) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // LAYOUT 2: groups start at multiples of 8 of the GLOBAL frame number (the first one may begin before the utterance), so the
    // eight frames of a column are one aligned 32-byte run inside a tile: two 16-byte loads, every sector fetched exactly once
    for (int tb = LAYOUT == 2 ? -(int)(f0 & 7) : 0; tb < T; tb += FB) {
        float obs[FB][PPL][3];
        if (LAYOUT == 2) {
            const int64_t F0 = f0 + tb;
            const float *base = a.logp + ((F0 >> 7) * a.ld) * 128 + (F0 & 127);
#pragma unroll
            for (int r = 0; r < PPL; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (valid[r]) {
                        const float4 *p = reinterpret_cast<const float4 *>(base + (size_t)(3 * (lane + 32 * r) + j) * 128);
                        v0 = __ldg(p); v1 = __ldg(p + 1);
                    }
                    obs[0][r][j] = v0.x; obs[1][r][j] = v0.y; obs[2][r][j] = v0.z; obs[3][r][j] = v0.w;
                    obs[4][r][j] = v1.x; obs[5][r][j] = v1.y; obs[6][r][j] = v1.z; obs[7][r][j] = v1.w;
                }
        }
        if (TILED) {
            int64_t need = (f0 + tb + FB - 1) >> 4;    // last panel this group of frames reads
            if (need > last_blk) need = last_blk;
            while (next_blk <= need) request_panel(next_blk++);
            if (next_blk <= last_blk && next_blk == need + 1) {   // one panel ahead
                request_panel(next_blk++);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else if (next_blk == need + 2) {
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
        }
        if (LAYOUT != 2)
#pragma unroll
        for (int q = 0; q < FB; ++q)
#pragma unroll
            for (int r = 0; r < PPL; ++r)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int i = lane + 32 * r;
                    const int64_t f = f0 + tb + q;
                    int rr = row0 + q;
                    if (rr >= 48) rr -= 48;
                    if (TILED) obs[q][r][j] = (valid[r] && tb + q < T) ? s_panel[rr * pstride + 3 * i + j] : 0.0f;
                    else obs[q][r][j] = (valid[r] && tb + q < T) ? a.logp[f * a.ld + 3 * i + j] : 0.0f;
                }
        if (TILED) __syncwarp();   // (all lanes have their observations before a later request reuses a panel slot)
        row0 += FB;
        if (row0 >= 48) row0 -= 48;

#pragma unroll
        for (int q = 0; q < FB; ++q) {
            const int t = tb + q;
            if (t < T && (LAYOUT != 2 || t >= 0)) {   // predicated, not `break`: keeps obs[][][] in registers (fully unrolled indices)
            // ---- PropagateInModels (phndec.cpp:96-119): descending j, in place
#pragma unroll
            for (int r = 0; r < PPL; ++r) {
                if (!valid[r]) continue;
#pragma unroll
                for (int j = 3; j > 0; --j) {
                    const float cur = __fadd_rn(al[r][j], PHN_LN05);
                    const float prv = __fadd_rn(al[r][j - 1], PHN_LN05);
                    if (cur > prv) {
                        al[r][j] = __fadd_rn(cur, obs[q][r][j - 1]);
                        ln[r][j] += 1;
                    } else {
                        al[r][j] = __fadd_rn(prv, obs[q][r][j - 1]);
                        pv[r][j] = pv[r][j - 1];
                        ln[r][j] = ln[r][j - 1] + 1;
                    }
                }
            }
            // ---- PropagateInNetwork (phndec.cpp:121-144): first strict max of alpha[i][3] from (-FLT_MAX, 0)
            // (inside a lane the candidates are visited in index order, so the reference's strict `>` on floats picks the
            // lane's first maximum; one order-preserving key per lane then goes into the warp reduction)
            float bv = -FLT_MAX;
            int bi = 0, b_prev = pv[0][3], b_len = ln[0][3];   // payload of the local winner travels with it
#pragma unroll
            for (int r = 0; r < PPL; ++r) {
                const float v = al[r][3];
                if (valid[r] && v > bv) { bv = v; bi = lane + 32 * r; b_prev = pv[r][3]; b_len = ln[r][3]; }
            }
            const unsigned bo = bv > -FLT_MAX ? f2ord(bv + 0.0f) : ORD_FLOOR;
            const unsigned mo = __reduce_max_sync(FULL, bo);
            const int mi = (int)__reduce_min_sync(FULL, bo == mo ? (unsigned)bi : 0x7fffffffu);
            const float mx = mo == ORD_FLOOR ? -FLT_MAX : ord2f(mo);
            // AddHistory (phndec.cpp:146-158): record of frame t+1, written by the lane that owns phone mi
            if ((mi & 31) == lane) { hphn[t] = b_prev; hlen[t] = b_len; halpha[t] = mx; }
            const float entry = __fadd_rn(mx, wp);
#pragma unroll
            for (int r = 0; r < PPL; ++r) { al[r][0] = entry; pv[r][0] = mi; ln[r][0] = 0; }
            last_mi = mi;
            // ---- GetBestToken (phndec.cpp:169-189), only consumed by TimePruning when n >= H+1
            if (t + 1 >= H + 1) {
                float cv = -FLT_MAX;
                int ci = 0x7fffffff, c_prev = 0, c_len = 1;
#pragma unroll
                for (int r = 0; r < PPL; ++r)
#pragma unroll
                    for (int j = 1; j <= 3; ++j) {
                        const float v = al[r][j];
                        if (valid[r] && v > cv) { cv = v; ci = (lane + 32 * r) * 3 + (j - 1); c_prev = pv[r][j]; c_len = ln[r][j]; }
                    }
                const unsigned co = cv > -FLT_MAX ? f2ord(cv + 0.0f) : ORD_FLOOR;
                const unsigned to = __reduce_max_sync(FULL, co);
                const int ti = (int)__reduce_min_sync(FULL, (co == to && ci != 0x7fffffff) ? (unsigned)ci : 0x7fffffffu);
                if (ti == 0x7fffffff) {  // nothing above -FLT_MAX: the scan's initial (len 1, prev 0)
                    if (lane == 0) { rbl[t] = 1; rbp[t] = 0; }
                } else if (ci == ti) {   // exactly one lane holds the winning (phone, state)
                    rbl[t] = c_len; rbp[t] = c_prev;
                }
            }
            }
        }
    }
    __syncwarp();
    __threadfence_block();

    // ------------------------------------------------------------------ phase 2: traceback
    // record of frame number n (1-based) lives at index n-1; n <= 0 is the initial shift-register
    // content (phndec.cpp:71-78): phn -1, len -1, alpha -1.0f
    auto Hphn = [&](int n) { return n >= 1 ? hphn[n - 1] : -1; };
    auto Hlen = [&](int n) { return n >= 1 ? hlen[n - 1] : -1; };
    auto Halpha = [&](int n) { return n >= 1 ? halpha[n - 1] : -1.0f; };

    phn_label *out = a.labels + a.lab_off[seg];
    const int cap = (int)(a.lab_off[seg + 1] - a.lab_off[seg]);
    int count = 0;
    // TimePruning (phndec.cpp:191-234) for every n in [H+1, T]
    for (int nb = H + 1; nb <= T; nb += 32) {
        const int n = nb + lane;
        bool emit = false;
        int q = 0;
        if (n <= T) {
            int o = H - rbl[n - 1];
            q = rbp[n - 1];
            while (o > 0) {
                const int uu = n - H + o;
                q = Hphn(uu);
                o -= Hlen(uu);
            }
            emit = o == 0;
        }
        const unsigned m = __ballot_sync(FULL, emit);
        if (emit) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                const int end = n - H;
                out[pos].phn = q;
                out[pos].start = end - Hlen(end);
                out[pos].end = end;
                out[pos].like = Halpha(end);  // commit alpha; turned into a score below
            }
        }
        count += __popc(m);
    }
    __syncwarp();
    const int n_commit = count < cap ? count : cap;
    // like_k = alpha_k - alpha_{k-1} (mPrevAlpha chain, phndec.cpp:226-232), alpha_{-1} = 0.
    // Chunks are rewritten from the end so that alpha_{k-1} is still intact when chunk k reads it.
    float prev_alpha = n_commit > 0 ? out[n_commit - 1].like : 0.0f;
    for (int kb = ((n_commit - 1) / 32) * 32; kb >= 0 && n_commit > 0; kb -= 32) {
        const int k = kb + lane;
        float ak = 0.0f, ap = 0.0f;
        if (k < n_commit) { ak = out[k].like; ap = k > 0 ? out[k - 1].like : 0.0f; }
        __syncwarp();
        if (k < n_commit) out[k].like = __fsub_rn(ak, ap);
        __syncwarp();
    }

    // Done (phndec.cpp:236-302): final traceback, lane 0, at most H labels
    if (lane == 0) {
        int cnt = 0;
        {
            int o = H, q = T > 0 ? last_mi : -1;
            while (o > 0 && q != -1) {
                const int uu = T - H + o;
                q = Hphn(uu);
                o -= Hlen(uu);
                ++cnt;
            }
        }
        int o = H, end = T, q = T > 0 ? last_mi : -1, idx = 0;
        while (o > 0 && q != -1) {
            const int uu = T - H + o;
            const int l = Hlen(uu);
            const float av = Halpha(uu);
            const int qq = Hphn(uu);
            o -= l;
            const float like = o > 0 ? __fsub_rn(av, Halpha(T - H + o)) : __fsub_rn(av, prev_alpha);
            const int pos = count + (cnt - 1 - idx);
            if (pos < cap) { out[pos].phn = q; out[pos].start = end - l; out[pos].end = end; out[pos].like = like; }
            end -= l;
            q = qq;
            ++idx;
        }
        a.nlab[seg] = count + cnt;
    }
}

int launch_viterbi(phn_ctx *c, const float *d_pen, int n_pen, cudaStream_t s, phn_ctx::DecSlot &sl)
{
    const int nseg = c->n_utt * n_pen;
    if (nseg == 0) return PHN_OK;
    const int ncols = 3 * c->P;
    int rc = ensure(c, c->d_logp, sizeof(float) * (size_t)((c->total_frames + 127) / 128 * 128 + 128) * c->ldp);
    if (rc) return rc;
    if (c->total_frames && !c->logp_valid) {   // (the tensor-core merger writes ln p itself when it feeds the decoder directly)
        int64_t blocks = (c->total_frames + 7) / 8;
        if (blocks > (int64_t)c->num_sms * 16) blocks = (int64_t)c->num_sms * 16;
        k_log_post<<<(unsigned)blocks, 256, 0, s>>>((const float *)c->d_post.p, c->ldp, ncols, c->total_frames, (float *)c->d_logp.p);
        PHN_CUDA(c, cudaGetLastError());
        c->k_launches[PHN_K_VIT] += 1;
    }
    VitArgs a;
    a.logp = (const float *)c->d_logp.p;
    a.ld = c->ldp;
    a.frame_off = (const int64_t *)sl.d_frame_off.p;
    a.n_utt = c->n_utt; a.P = c->P; a.H = c->hist;
    a.total_frames = c->total_frames;
    a.pen = d_pen;
    const size_t R = (size_t)n_pen * (size_t)c->total_frames;
    int *rec = (int *)c->d_rec.p;
    a.r_hphn = rec; a.r_hlen = rec + R; a.r_bp = rec + 2 * R; a.r_bl = rec + 3 * R;
    a.r_halpha = (float *)(rec + 4 * R);
    a.labels = (phn_label *)sl.d_labels.p;
    a.lab_off = (const int64_t *)sl.d_lab_off.p;
    a.nlab = (int *)sl.d_nlab.p;
    const int ppl = (c->P + 31) / 32;
    const bool tiled = c->logp_valid != 0;   // ln p came from the tensor-core merger's epilogue (tiled layout)
    c->logp_layout = tiled ? 2 : 1;
    const size_t vsmem = sizeof(float) * 48 * (size_t)((3 * c->P) | 1);
    // (the panel ring of the tiled form exceeds the 48 KB default from 86 phonemes on)
    // PHNREC_VIT_DIRECT=1 / 0: the tiled decoder without / with shared-memory panels (CZ: 0.55 against 0.53 ms alone; first built for
    // an experiment in which it ran beside the next batch's MLP kernel - tools/experiments/README.md).
    // Unset: panels, unless their shared memory keeps the batch from being resident at once while the panel-free form is - one
    // warp per utterance is a latency-bound chain, a second wave doubles the time (HU, 61 phonemes: 35 KB of panels = 6 CTAs per SM
    // = 888 of 1000 utterances: 0.92 ms; panel-free 0.56 ms).
    static const int direct_env = getenv("PHNREC_VIT_DIRECT") ? (atoi(getenv("PHNREC_VIT_DIRECT")) != 0 ? 1 : 0) : -1;
    bool panels = direct_env != 1;
#define PHN_VIT(N)                                                          \
    do {                                                                    \
        if (tiled && direct_env < 0) {                                      \
            int occ_p = 0, occ_d = 0;                                       \
            PHN_CUDA(c, cudaFuncSetAttribute(k_viterbi<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem)); \
            PHN_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, k_viterbi<N, 1>, 32, vsmem)); \
            PHN_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_d, k_viterbi<N, 2>, 32, 0)); \
            panels = !((int64_t)nseg > (int64_t)occ_p * c->num_sms && (int64_t)nseg <= (int64_t)occ_d * c->num_sms); \
        }                                                                   \
        if (tiled && panels) {                                              \
            PHN_CUDA(c, cudaFuncSetAttribute(k_viterbi<N, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vsmem)); \
            k_viterbi<N, 1><<<nseg, 32, vsmem, s>>>(a);                     \
        } else if (tiled) k_viterbi<N, 2><<<nseg, 32, 0, s>>>(a);           \
        else k_viterbi<N, 0><<<nseg, 32, 0, s>>>(a);                        \
    } while (0)
    switch (ppl) {
        case 1: PHN_VIT(1); break;
        case 2: PHN_VIT(2); break;
        case 3: PHN_VIT(3); break;
        case 4: PHN_VIT(4); break;
        default: return fail(c, PHN_ERR_UNSUPPORTED, "more than 128 phonemes\n");
    }
#undef PHN_VIT
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_VIT] += 1;
    return PHN_OK;
}

// Label counts -> compact offsets: out_off[0..nseg] = exclusive prefix sums of nlab, out_off[nseg + 1] = 1 + the first segment
// whose count exceeds its capacity (0: none).  One block; runs on the decoder's stream right behind K-vit, so that fetching
// the labels is two plain D2H copies and needs no kernel of its own (the next batch's persistent kernels own the SMs by then).
__global__ void __launch_bounds__(1024) k_label_scan(const int *__restrict__ nlab, const int64_t *__restrict__ cap_off, int nseg,
                                                     int64_t *__restrict__ out_off)
{
    __shared__ int64_t s_sum[1024];
    __shared__ int s_bad;
    const int t = threadIdx.x;
    if (t == 0) s_bad = 0x7fffffff;
    __syncthreads();
    const int per = (nseg + 1023) / 1024;
    const int lo = t * per, hi = min(nseg, lo + per);
    int64_t sum = 0;
    for (int k = lo; k < hi; ++k) {
        const int n = nlab[k];
        if ((int64_t)n > cap_off[k + 1] - cap_off[k]) atomicMin(&s_bad, k);
        sum += n;
    }
    s_sum[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {   // inclusive scan of the per-thread sums
        const int64_t v = t >= d ? s_sum[t - d] : 0;
        __syncthreads();
        s_sum[t] += v;
        __syncthreads();
    }
    int64_t run = s_sum[t] - sum;
    for (int k = lo; k < hi; ++k) { out_off[k] = run; run += nlab[k]; }
    if (t == 1023) { out_off[nseg] = s_sum[1023]; out_off[nseg + 1] = s_bad == 0x7fffffff ? 0 : (int64_t)s_bad + 1; }
}

// Gather each segment's labels (stored at capacity offsets) into one contiguous run.
__global__ void k_compact_labels(const phn_label *__restrict__ src, const int64_t *__restrict__ cap_off,
                                 const int64_t *__restrict__ out_off, phn_label *__restrict__ dst)
{
    const int seg = blockIdx.x;
    const int64_t s0 = cap_off[seg], d0 = out_off[seg];
    int64_t n = out_off[seg + 1] - d0;
    if (n > cap_off[seg + 1] - s0) n = cap_off[seg + 1] - s0;   // (an overflowing segment is reported by the scan; never read past its run)
    const int4 *s = reinterpret_cast<const int4 *>(src + s0);
    int4 *d = reinterpret_cast<int4 *>(dst + d0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

// scan + gather on the decoder's stream (sl.s), right behind the decoder
int launch_compact_labels(phn_ctx *c, int nseg, phn_ctx::DecSlot &sl)
{
    k_label_scan<<<1, 1024, 0, sl.s>>>((const int *)sl.d_nlab.p, (const int64_t *)sl.d_lab_off.p, nseg, (int64_t *)sl.d_coff.p);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_VIT] += 1;
    if (nseg == 0) return PHN_OK;
    k_compact_labels<<<nseg, 64, 0, sl.s>>>((const phn_label *)sl.d_labels.p, (const int64_t *)sl.d_lab_off.p,
                                            (const int64_t *)sl.d_coff.p, (phn_label *)sl.d_labels_c.p);
    PHN_CUDA(c, cudaGetLastError());
    c->k_launches[PHN_K_VIT] += 1;
    return PHN_OK;
}

}  // namespace phn
