// internal.h — shared declarations of libphnrec_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/phnrec_b200.h"

namespace phn {

enum { PHN_SYS_LCRC = 0, PHN_SYS_1BT = 1, PHN_SYS_1BT_DCT = 2, PHN_SYS_3BT = 3 };

// ---------------------------------------------------------------- host: config
// The INI dialect of configz.cpp:102-165 with the typed variable table of srec.cpp:34-110.
struct Config {
    std::map<std::string, std::string> kv;  // "section/variable" -> value
    int load(const std::string &file, int *err_line);  // PHN_OK or PHN_ERR_CFG_*
    const std::string &str(const char *sec, const char *var) const;
    int i(const char *sec, const char *var) const;
    float f(const char *sec, const char *var) const;
    bool b(const char *sec, const char *var) const;
};

// ---------------------------------------------------------------- host: model
struct HostNet {  // .nbin image, nn.cpp:464-531
    int nin = 0, nhid = 0, nout = 0, nin4 = 0, nhid4 = 0, nout4 = 0;
    std::vector<float> w1, w2, b1, b2, mean, dev;
    int load(const std::string &path);  // PHN_OK / PHN_ERR_NN_*
    // ASCII model files (nn.cpp:199-462): `weigvec N` x2, `biasvec N` x2; norms `vec N` x2 (means, inverse std devs;
    // norms_path empty -> means 0, devs 1).  save_nbin writes the reference's binary cache (nn.cpp:533-592).
    int load_ascii(const std::string &weights_path, const std::string &norms_path);
    int save_nbin(const std::string &path) const;
};

struct MelTables {  // melbanks.cpp:38-70, dspc.cpp:80-225, dspc.h:162-167
    int nbanks, vs, step, fs, N, logN, N2, fftlo, ffthi;
    std::vector<float> hamming, coeffs;
    std::vector<int> banks;
    std::vector<int> bank_klo, bank_khi;  // per bank: contiguous bin range that feeds it
    std::vector<double> tw;               // (wr, wi) pairs: stage h=1,2,4..N/2, index m<h at [2*(h-1+m)]
    std::vector<float> f0;                // centres of the banks in Hz (dspc.cpp:156-162), used by PLP's equal-loudness curve
    void build(int nbanks, int vs, int step, int fs, float lo, float hi);
};

// ---------------------------------------------------------------- device: model
struct DevNet {
    int nin, nhid, nout, nin4, nhid4, nout4;
    int kp;    // layer-1 K rounded up to 16  (row stride of the fp32 input matrix)
    int ldh;   // hidden row stride: nhid4 rounded up to 16
    float *w1, *w2, *b1, *b2, *mean, *dev;  // fp32 images exactly as in the .nbin
    // fp16 tensor-core images (K-major, zero padded): w1h [nhidP][k1P], w2h [noutP][nhidP]
    __half *w1h, *w2h;
    int k1P, nhidP, noutP;
    int kin;   // data columns of the fp16 activation image; columns kin, kin+1 hold the constant 1.0 that multiplies b1
};

// Decoder of one stream between two pushes (k_stream.cu): PhnDec's members (phndec.cpp:44-94), the history as a 64-slot ring
struct VitStreamState {
    float al[128 * 4];
    int pv[128 * 4], ln[128 * 4];
    int hphn[64], hlen[64];
    float halpha[64];
    int n, last_mi;
    float prev_alpha;
    int pad_;
};

struct DevTables {
    float *hamming, *coeffs;
    int *banks, *bank_klo, *bank_khi;
    double2 *tw;
    float *win;   // [32] band0 then band1 window
    float *dct;   // [10][16] cosf table (dspc.h:206-221)
};

}  // namespace phn

// ---------------------------------------------------------------- the context
struct phn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // H2D of audio groups, overlapped with K-wave of the previous group (phn_recognize)
    cudaEvent_t ev_copy[16] = {nullptr};     // group g has landed
    cudaEvent_t ev_free = nullptr;           // the audio buffer's previous readers are done
    // The decoder of the audio -> labels path runs on its own stream: K-vit is one warp per utterance (latency bound, a
    // fraction of the SMs' issue slots), so the decoder of batch k runs under the front end (K-wave, K-stc) of batch k+1.
    cudaStream_t vit_stream = nullptr;
    cudaStream_t fetch_stream = nullptr;     // label read-back: two D2H copies behind the slot's `done` event, never behind a later batch's decoder
    cudaEvent_t ev_mlp_done = nullptr;       // ln p of the batch is complete (recorded on `stream`)
    cudaEvent_t ev_vit_done = nullptr;       // the decoder has consumed it (recorded on `vit_stream`)
    int vit_pending = 0;                     // a decoder launch on vit_stream has not been waited for by `stream` yet
    cudaEvent_t ev_audio_free = nullptr;     // K-wave of the last host-audio call has consumed d_audio (the next call's copy may start)
    int audio_free_valid = 0;
    std::string err;
    std::string cfg_dir;
    phn::Config cfg;
    // config-derived
    int fs = 8000, fmt = 0, nbanks = 15, vs = 200, step = 80, S = 3, P = 0, hist = 40;
    int sent_mean_norm = 0, z_mean = 0;
    float lo = 0, hi = 4000, preem = 0, wpenalty = -2.f, frame_shift = 0.f, frame_floor = -9999.9f, scale = 1.f, dc_shift = 0.f;
    // TRAPS system (posteriors/system): LCRC = the shipped systems; 1BT / 3BT / 1BT_DCT = k_trap.cu (exact mode only)
    int system = 0, trap_len = 31, tshift = 15, use_hamming = 0, add_c0 = 1, trap_bands = 0, trap_shift_out = 0;
    std::vector<phn::HostNet> hband;    // 1BT / 3BT: one net per band
    std::vector<phn::DevNet> dband;
    float *d_trap_ham = nullptr, *d_trap_cos = nullptr;
    void *d_trap_pm = nullptr, *d_trap_pd = nullptr;
    int trap_ready = 0;
    // params/kind = plp (k_plp.cu): `-t par` / phn_mel only
    int plp = 0, plp_order = 12, plp_add_c0 = 0, nparams = 0;
    float plp_compress = 0.3333333f, plp_lifter = 22.f, plp_scale = 10.f;
    float *d_plp_eql = nullptr, *d_plp_idft = nullptr, *d_plp_lift = nullptr;
    int mlp_mode = PHN_MLP_EXACT_FP32;
    void *tc = nullptr;  // tensor-core mode state (k_mlp_tc.cu)
    void *wave_tc = nullptr;  // tensor-core front end: DFT matrix image + filterbank tables (k_wave_tc.cu)
    void *wave_tc16 = nullptr;  // ... of the 16 kHz systems (k_wave_tc16.cu)
    void *stc_btab = nullptr, *stc_bias = nullptr;   // K-stc tensor-core formulation: constant matrices (k_stc.cu)
    void *stc_cf = nullptr, *stc_sb = nullptr;       // K-stc fp32 (FFMA2) form: window x basis table, per-column scale / bias pairs
    int force_exact_wave = 0;
    int fast_front = 0;   // audio -> labels path of the tensor-core mode: fp32-tolerance front end (K-wave pairs, parallel sentence mean)
    void *tc_dbg = nullptr;   // device buffer for the tensor-core kernel's debug timeline (phn_debug_tc_timeline)
    int tc_dbg_net = -1;
    int fuse_logp = 0;   // tensor-core merger also writes ln(posteriors) for the decoder (audio -> labels path)
    int post_valid = 0;  // d_post holds the linear posteriors of the current batch
    int logp_layout = 0; // what the last decode left in d_logp: 0 nothing, 1 row-major [F][ldp] (K-log), 2 tiled [F/128][ldp][128] (merger epilogue)
    int logp_valid = 0;  // d_logp already holds the decoder's input for the current batch  // phn_mel always returns the reference's bits, whatever the MLP mode
    std::vector<std::string> phonemes;
    float win[32];
    phn::HostNet hnet[3];
    phn::MelTables mt;
    phn::DevNet net[3];
    phn::DevTables tab{};
    int ncoef = 11;
    int ldp = 0;  // device row stride of the posterior matrix (n_outputs rounded up to 4)
    int num_sms = 148;

    // ---- PHNREC_TIMELINE=<file>: CUDA events + host clock per batch, dumped by phn_destroy (development aid)
    struct TlRow { cudaEvent_t e[6] = {nullptr}; cudaEvent_t cg[2] = {nullptr, nullptr}; double h[4] = {0, 0, 0, 0}; };
    std::vector<TlRow> tl;
    int tl_on = -1;       // -1: environment not read yet
    int tl_cur = -1;      // row of the batch being enqueued

    // ---- batch state (grow-only device buffers)
    int n_utt = 0, n_pen = 1;
    int64_t total_bytes = 0, total_frames = 0, label_cap = 0;
    std::vector<int64_t> h_byte_off, h_frame_off, h_lab_off;
    struct Buf { void *p = nullptr; size_t cap = 0; };
    Buf d_audio, d_byte_off, d_frame_off, d_lab_off, d_mel, d_mean, d_post, d_rec, d_pen;
    Buf d_x0, d_x1, d_h, d_xm, d_x0h, d_x1h, d_xmh;  // MLP workspace (per frame chunk)
    Buf d_xb;                                        // 1BT / 3BT: [bands][chunk][kp] band-net inputs
    Buf d_par;                                       // PLP coefficients [F][nparams]
    Buf d_tile_ctr, d_logp;
    // Results of one decoder launch.  Two slots, used alternately: the labels of batch k can be fetched (phn_wait) while
    // batch k+1 is already running (phn_recognize_async).  A slot carries its own copies of the frame / capacity offsets:
    // the next batch's planning rewrites d_frame_off / d_lab_off and the host vectors.
    struct DecSlot {
        Buf d_labels, d_nlab, d_lab_off, d_frame_off, d_coff, d_labels_c;
        std::vector<int64_t> h_lab_off, h_frame_off;
        std::vector<int32_t> h_nlab;
        int n_utt = 0, n_pen = 1;
        cudaStream_t s = nullptr;            // the stream the decoder ran on
        cudaEvent_t done = nullptr;          // decoder + label compaction finished (fetching waits for this, on fetch_stream)
    } slot[2];
    int slot_w = 0;                          // the slot the next decode writes
    int slot_last = -1;                      // the slot of the most recent decode (phn_fetch_labels)
    int pend[2] = {0, 0}, n_pend = 0;        // slots of asynchronous batches not yet waited for, oldest first
    int pend_tl[2] = {-1, -1};               // ... their PHNREC_TIMELINE rows
    Buf d_pair_off;                          // [n_utt + 1] prefix sums of ceil(T_u / 2): work units of the frame-pair K-wave
    std::vector<int64_t> h_pair_off;
    std::vector<float> h_pen;
    int64_t chunk_frames = 0;
    // ---- streaming (online) path: per-stream state, host side and device side (phn_stream_*)
    struct StreamHost {
        std::vector<uint8_t> tail;   // whole samples that have not completed a frame yet (raw bytes)
        int64_t frames = 0;          // log-mel frames produced so far
        int64_t rows_done = 0;       // posterior rows handed to the decoder so far
        int hist = 0;                // frames held in the device history (<= 30)
        bool fed = false;            // the decoder has consumed at least one row (its state on the device is live)
    };
    std::vector<StreamHost> streams;
    Buf d_st_hist, d_st_norm, d_st_cnt, d_st_vit;     // [n_streams] x {30 x nbanks floats, 4 x nbanks floats, unsigned, VitStreamState}
    Buf d_st_args, d_win, d_st_labels, d_st_nlab;     // per-push argument block, assembled windows, label output
    std::vector<uint8_t> h_stage;                     // [tail | block] of every pushed stream
    int on_interval = 0, on_mean = 0, on_var = 0, bunch = 5;   // [onlinenorm] estim_interval / mean_norm / var_norm, [posteriors] bunch_size
    int stream_frames = 0;                            // frames_of() counts streaming frames (none below vector_size samples)

    // profiling
    int profiling = 0;
    float k_ms[PHN_K_COUNT] = {0};
    int64_t k_launches[PHN_K_COUNT] = {0};
    cudaEvent_t ev[2 * PHN_K_COUNT] = {nullptr};
};

namespace phn {
// ---- error helpers
int fail(phn_ctx *c, int code, const char *fmt, ...);
#define PHN_CUDA(c, expr)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return phn::fail((c), PHN_ERR_CUDA, "CUDA failure: %s (%s) at %s:%d\n", cudaGetErrorString(e__), #expr, \
                             __FILE__, __LINE__);                                                \
    } while (0)

int ensure(phn_ctx *c, phn_ctx::Buf &b, size_t bytes);
int upload_small(phn_ctx *c, void *dst, const void *src, size_t bytes, cudaStream_t s);   // small per-batch tables (pageable source, staged by the driver)

// ---- kernel launchers (each in its own .cu)
int launch_wave(phn_ctx *c, const void *d_audio, int u0 = 0, int u1 = -1);   // k_wave.cu (utterance range)
int launch_sentence_mean(phn_ctx *c, int u0 = 0, int u1 = -1);             // k_norm.cu (utterance range)
int launch_online_norm(phn_ctx *c, float *d_x, int64_t frames, int nb, int interval, int mean_norm, int var_norm);
int launch_stc(phn_ctx *c, int64_t f0, int64_t nf, int64_t row_lo = 0, int64_t row_hi = -1);   // k_stc.cu: frames [f0, f0+nf) of the pass;
                                                                                                  // only rows row_lo <= frame < row_hi are produced
int launch_mlp_exact(phn_ctx *c, int64_t f0, int64_t nf);                  // k_mlp_exact.cu
// one net of the exact mode: outputs to `post` (linear posteriors) or, when post == nullptr, into the merger's input matrix at
// column xm_col0 as (+-)sLn(p) with the merger's input normalisation (negate: the 1BT / 3BT systems, traps.cpp:425-427)
int run_net_exact(phn_ctx *c, const phn::DevNet &n, const float *x, int ldx, int64_t nf, float *post, int ldpost, int xm_col0, int negate);
int launch_plp(phn_ctx *c);                                                // k_plp.cu: d_mel (raw bank energies) -> d_par
int launch_trap(phn_ctx *c, int64_t f0, int64_t nf);                        // k_trap.cu: 1BT / 3BT / 1BT_DCT front end
int launch_mlp_trap(phn_ctx *c, int64_t f0, int64_t nf);                    // k_mlp_exact.cu: their nets
int launch_mlp_tc(phn_ctx *c, int64_t f0, int64_t nf);                     // k_mlp_tc.cu
int launch_viterbi(phn_ctx *c, const float *d_pen, int n_pen, cudaStream_t s, phn_ctx::DecSlot &sl);   // k_vit.cu
int launch_compact_labels(phn_ctx *c, int nseg, phn_ctx::DecSlot &sl);     // k_vit.cu
int launch_logf_range(phn_ctx *c, uint32_t first_bits, int64_t n, float *d_out);   // k_vit.cu (verification aid)
int launch_log_post(phn_ctx *c, int64_t rows);                               // k_vit.cu: d_post -> d_logp (row-major), first `rows` rows
// k_stream.cu: the state-carrying kernels of the streaming path
int launch_stream_norm(phn_ctx *c, int n, const int *d_sid, float *d_state, unsigned *d_cnt, int interval, int mean_norm, int var_norm);
int launch_stream_assemble(phn_ctx *c, int n, const int64_t *d_new_off, const int64_t *d_win_off, const int *d_sid, const int *d_pad,
                           const int *d_hist, float *d_win, float *d_st_hist, int keep);
int launch_stream_decode(phn_ctx *c, int n, int64_t rows, const int64_t *d_row0, const int *d_count, const int *d_sid, const int *d_fresh,
                         const int *d_last, phn::VitStreamState *d_st, phn_label *d_labels, const int64_t *d_lab_off, int *d_nlab);
int launch_synth(phn_ctx *c, void *d_audio, int64_t bytes_per_utt, int n_utt, uint64_t seed);  // k_synth.cu
int mlp_tc_fill_merger_bias(phn_ctx *c, int64_t rows);                     // constant-1 bias columns of the merger image
int wave_tc_prepare(phn_ctx *c);                                           // k_wave_tc.cu
void wave_tc_release(phn_ctx *c);
bool wave_tc_applies(phn_ctx *c);
int launch_wave_tc(phn_ctx *c, const void *d_audio, int64_t f_begin, int64_t f_end);
int wave_tc16_prepare(phn_ctx *c);                                         // k_wave_tc16.cu
void wave_tc16_release(phn_ctx *c);
bool wave_tc16_applies(phn_ctx *c);
int launch_wave_tc16(phn_ctx *c, const void *d_audio, int64_t f_begin, int64_t f_end);
int mlp_tc_prepare(phn_ctx *c);                                            // fp16 weight images (k_mlp_tc.cu)
void mlp_tc_release(phn_ctx *c);
}  // namespace phn
