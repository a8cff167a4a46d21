/*
 * phnrec_b200.h — C ABI of the B200-native PhnRec recognition hot path.
 *
 * The reference (rampa069/PhnRec) has no FFI layer: its hot path sits behind four C++
 * class seams inside SpeechRec.  This header is the drop-in boundary
 * that replaces those seams; each entry point names the reference interface it
 * replaces (paths relative to the reference checkout).  Plain C types only: pointers,
 * sizes, integer status codes; caller-allocated outputs; no exceptions cross it.
 *
 * Batch convention ("ragged batch"): n_utt utterances are concatenated in one buffer;
 * off[u] .. off[u+1] delimits utterance u (off has n_utt+1 entries, off[0] = 0).
 *   byte_off  : offsets into the audio buffer, in BYTES
 *   frame_off : offsets into mel / posterior matrices, in FRAMES (rows)
 *   label_off : offsets into the label array, in LABELS
 *
 * There is NO CPU fallback: every compute entry point runs hand-written sm_100a CUDA
 * kernels and returns PHN_ERR_CUDA when no device / kernel image is available.
 */
#ifndef PHNREC_B200_H
#define PHNREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct phn_ctx phn_ctx;

/* One recognised segment; times are frame indices (x 100000 = HTK 100 ns units when
 * printed, phndec.cpp:230,292).  Replaces the DECODER_CALLBACK tuple
 * (unsigned msg, char* word, long_long start, long_long stop, float score) of
 * decoder.h:30 — `phn` indexes dicts/phonemes (phn_phoneme()). */
typedef struct {
    int32_t phn;
    int32_t start;
    int32_t end;
    float like;
} phn_label;

/* Status codes.  1..5 mirror NN_* (nn.h:35-42), 10..13 mirror EN_* (configz.h:29-33),
 * 20.. mirror DECERR_* (decoder.h:37-42); the rest are ours. */
enum {
    PHN_OK = 0,
    PHN_ERR_NN_FILE = 1,      /* NN_FILEERR / NN_READERR: weights nbin file missing or truncated */
    PHN_ERR_NN_FORMAT = 2,    /* NN_INVFORMAT: not a 3-layer .nbin */
    PHN_ERR_CFG_FILE = 10,    /* EN_FILEERR  */
    PHN_ERR_CFG_UNKVAR = 11,  /* EN_UNKVAR   */
    PHN_ERR_CFG_BADVAL = 12,  /* EN_BADVAL   */
    PHN_ERR_CFG_INVVAR = 13,  /* EN_INVVAR   */
    PHN_ERR_DEC_INPUT = 20,   /* DECERR_INPUTFILE-class: phoneme list / window files */
    PHN_ERR_UNSUPPORTED = 30, /* config selects something outside the hot path */
    PHN_ERR_ARG = 31,
    PHN_ERR_CAPACITY = 32,    /* caller buffer too small; required size is reported */
    PHN_ERR_NOMEM = 33,
    PHN_ERR_CUDA = 40         /* no device, no sm_100a image, or a CUDA runtime failure */
};

enum { PHN_WAVE_LIN16 = 0, PHN_WAVE_ALAW = 1 };        /* SpeechRec::wave_format, srec.h */
enum {
    PHN_MLP_EXACT_FP32 = 0, /* CUDA-core fp32, reference summation order: bit-identical posteriors */
    PHN_MLP_TC_F16 = 1      /* tcgen05/TMEM GEMMs, fp16 operands, fp32 accumulate */
};

typedef struct {
    int32_t sample_freq, wave_format, nbanks, vector_size, vector_step, fft_size;
    int32_t n_phonemes, n_states, n_outputs;     /* decoder uses the first n_phonemes*n_states outputs */
    int32_t band_inputs, merger_inputs, hidden;  /* .nbin header sizes (nn.cpp:464-531) */
    int32_t sent_mean_norm, time_pruning, mlp_mode, device;
    float wpenalty;
    int32_t n_params;   /* columns of the parameter matrix phn_mel returns / `-t par` saves: nbanks, or the PLP coefficients (params/kind = plp) */
} phn_info;

/* -- lifetime ---------------------------------------------------------------------- */
/* Replaces SpeechRec::Init (srec.cpp:235-707) for the offline phndec/LCRC path:
 * parses <cfg_dir>/config (configz.cpp:102-165 dialect, variable table srec.cpp:34-110),
 * loads weights/{band0,band1,merger}.nbin (nn.cpp:464-531), windows/band{0,1}.window
 * (traps.cpp:549-570), dicts/phonemes (phndec.cpp:305-350); builds the Hamming / mel
 * filterbank / FFT twiddle / DCT tables on the host with the reference's float
 * expressions (dspc.cpp:80-225, dspc.h:162-221) and uploads everything to `device`. */
int phn_create(const char *cfg_dir, int device, phn_ctx **out);
void phn_destroy(phn_ctx *ctx);
/* Message of the last failure on this context (or of the last failed phn_create when
 * ctx == NULL).  Text follows the reference's MERROR strings where one exists. */
const char *phn_last_error(const phn_ctx *ctx);
int phn_get_info(const phn_ctx *ctx, phn_info *info);
const char *phn_phoneme(const phn_ctx *ctx, int index); /* PhnDec::LoadPhnList, phndec.cpp:305 */
/* Config::GetString (configz.h:87-102) after the $C / $T substitution of srec.cpp:219-233;
 * NULL for a variable the table does not know. */
const char *phn_config_get(const phn_ctx *ctx, const char *section, const char *variable);

/* -- knobs ------------------------------------------------------------------------- */
int phn_set_penalty(phn_ctx *ctx, float wpenalty);   /* Decoder::SetWPenalty, decoder.h:70 / phnrec.cpp:212-221 */
int phn_set_wave_format(phn_ctx *ctx, int fmt);      /* SpeechRec::SetWaveFormat, phnrec.cpp:224-225 */
int phn_set_mlp_mode(phn_ctx *ctx, int mode);        /* ours: PHN_MLP_* */

/* srec.cpp:945 — frames of one utterance of `nbytes` bytes in the current wave format. */
int64_t phn_num_frames(const phn_ctx *ctx, int64_t nbytes);
/* Upper bound on labels for a batch with these frame offsets (sizes label arrays). */
int64_t phn_label_capacity(const phn_ctx *ctx, const int64_t *frame_off, int n_utt);

/* -- the three stages, HOST buffers (copies inside) --------------------------------- */
/* audio -> un-normalised log mel-bank energies, what `-t par` saves.
 * Replaces ConvertWaveformFormat + MelBanks::{AddWaveform,GetFeatures} (srec.cpp:709-791,
 * 942-971; melbanks.h:70-92).  frame_off [n_utt+1] is written; mel_out is
 * [frame_off[n_utt]][nbanks] (pass NULL to only get frame_off). */
int phn_mel(phn_ctx *ctx, const void *audio, const int64_t *byte_off, int n_utt, float *mel_out, int64_t *frame_off);

/* mel -> linear posteriors, what `-t post` saves.  Replaces SentenceBasedNormalization
 * + Traps::CalcFeaturesBunched (+ the 15-frame warm-up / tail driver) + NeuralNet::Forward x3
 * (srec.cpp:999-1070, traps.h:58-75, nn.h:50-59).  post_out is [frames][n_outputs]. */
int phn_posteriors(phn_ctx *ctx, const float *mel, const int64_t *frame_off, int n_utt, float *post_out);

/* linear posteriors -> labels: decSoftFunc=log (srec.cpp:1088-1097) then
 * Decoder::{Init,ProcessFrame,Done} (decoder.h:58-74, phndec.cpp:44-302), once per
 * penalty (the `-s post -p P` sweep).  penalties == NULL -> the context's penalty, n_pen = 1.
 * Output order: penalty-major, i.e. label_off has n_pen*n_utt+1 entries and utterance u
 * under penalty k is segment k*n_utt+u.  Returns PHN_ERR_CAPACITY (and label_off filled
 * with the needed counts) when label_cap is too small. */
int phn_decode(phn_ctx *ctx, const float *post, const int64_t *frame_off, int n_utt, const float *penalties, int n_pen,
               phn_label *labels, int64_t label_cap, int64_t *label_off);

/* audio -> labels with no host round trip in between; replaces SpeechRec::ProcessOffline
 * (srec.cpp:929-1111) for dfWaveform -> dfStrings.  frame_off_out may be NULL.
 * The audio is copied in groups of whole utterances on a second stream while K-wave of the
 * previous group runs (pass page-locked memory, phn_host_alloc_pinned, for the overlap). */
int phn_recognize(phn_ctx *ctx, const void *audio, const int64_t *byte_off, int n_utt, phn_label *labels,
                  int64_t label_cap, int64_t *label_off, int64_t *frame_off_out);

/* Asynchronous pair of phn_recognize for callers that keep the GPU busy (the CLI's list mode, servers): the reference
 * processes a list strictly one file after the other (SpeechRec::ProcessFileList, srec.cpp:1246-1291); here batch k+1 is
 * enqueued while batch k is still on the device - its audio goes up under batch k's nets, and batch k's decoder and label
 * read-back run under batch k+1's front end.  phn_recognize_async enqueues everything and returns; `audio` must stay
 * valid and unchanged until the matching phn_wait has returned (page-locked memory keeps the copy asynchronous).
 * phn_wait returns the OLDEST batch not yet waited for, exactly as phn_recognize would have (labels, label_off with
 * n_utt+1 entries, optional frame_off_out).  At most two batches may be in flight; results are independent of how calls
 * are interleaved.  On PHN_ERR_CAPACITY the batch stays queued (label_off holds the counts) and phn_wait may be repeated. */
int phn_recognize_async(phn_ctx *ctx, const void *audio, const int64_t *byte_off, int n_utt);
int phn_wait(phn_ctx *ctx, phn_label *labels, int64_t label_cap, int64_t *label_off, int64_t *frame_off_out);
int phn_pending(const phn_ctx *ctx);     /* batches enqueued by phn_recognize_async and not yet returned by phn_wait */

/* -- streaming (online) mode ---------------------------------------------------------
 * Replaces SpeechRec::ProcessOnline / ProcessLastBunch / ProcessTail (srec.cpp:793-927) with its live
 * normaliser Normalization::ProcessFrame (norm.cpp:216-234, configured by [onlinenorm] estim_interval /
 * mean_norm / var_norm, srec.cpp:594-601) - the path the reference drives from a sound card in 125 ms
 * blocks (RunLive, srec.cpp:1438-1490) - for MANY concurrent streams on one GPU.  A stream is the
 * reference's set of streaming objects: MelBanks' frame buffer (melbanks.cpp:151-204), the 31-frame FIFO
 * of Traps (traps.cpp:180-219), PhnDec with its 41-slot history (phndec.cpp:44-302) and the normaliser's
 * running estimate; their state stays in the context between pushes.
 *   phn_stream_open(ctx, n)   n streams, all in their initial state (replaces any earlier set).
 *   phn_stream_reset(ctx, s)  stream s back to its initial state, normaliser included.
 *   phn_stream_push(ctx, sids, n, audio, byte_off, last, labels, cap, label_off)
 *       one block of audio for each of the n streams sids[i] (each stream at most once per push): bytes
 *       [byte_off[i], byte_off[i+1]) of `audio` (host memory, the context's wave format; a lin16 block
 *       contributes floor(bytes/2) samples, like ConvertWaveformFormat called per block).  Blocks may have
 *       any length, also 0.  last[i] != 0 ends the utterance on that stream (ProcessTail + Decoder::Done):
 *       the stream then starts a new utterance with its next block; the normaliser's estimate is kept,
 *       as the reference's Normalization object outlives MelBanks::Reset / Traps::Reset.  `last` may be NULL.
 *       Returns the labels that became final during this push (TimePruning commits, phndec.cpp:191-234,
 *       and, for ending streams, the final traceback) - what the reference hands to the decoder callback
 *       (decoder.h:30-35) while it runs: labels of pushed stream i are labels[label_off[i] .. label_off[i+1]).
 *       The concatenation over a stream's pushes is exactly the label sequence of the reference's online
 *       path on the concatenated audio, whatever the block sizes, and however streams are interleaved.
 *       An utterance shorter than 15 frames is decoded over 15 frames, like the reference's (the tail of
 *       15 copies of the last frame also flushes 15 - T warm-up rows through the decoder).
 *   Needs posteriors/bunch_size to divide 15 (all shipped configs: 5); XML persistence of the estimates
 *   (onlinenorm/file) and scale_to_gvar are not supported (PHN_ERR_UNSUPPORTED). */
int phn_stream_open(phn_ctx *ctx, int n_streams);
int phn_stream_count(const phn_ctx *ctx);
int phn_stream_reset(phn_ctx *ctx, int sid);
int phn_stream_push(phn_ctx *ctx, const int *sids, int n, const void *audio, const int64_t *byte_off, const int *last,
                    phn_label *labels, int64_t label_cap, int64_t *label_off);

/* -- device-resident variants (inputs already in HBM; used by bench.py and servers) -- */
/* d_audio is a DEVICE pointer on the context's device; byte_off stays a host array.
 * Runs wave -> mean -> STC -> MLPs on the context's stream and Viterbi + traceback on the context's
 * decoder stream (so that the decoder of one call runs under the front end of the next), and
 * leaves mel / posteriors / labels in the context's device buffers.  Asynchronous:
 * call phn_sync() (joins both streams) or a phn_fetch_* before reading results. */
int phn_recognize_device(phn_ctx *ctx, const void *d_audio, const int64_t *byte_off, int n_utt);
/* Penalty sweep from posteriors already resident in the context (after phn_posteriors): decSoftFunc = log once, then the
 * decoder once per penalty (srec.cpp:1080-1104 with -p); asynchronous, results through phn_fetch_labels (penalty-major). */
int phn_decode_device(phn_ctx *ctx, const float *penalties, int n_pen);
int phn_sync(phn_ctx *ctx);
/* Copy results of the last *_device / host call back. Any output may be NULL.
 * phn_fetch_posteriors: in PHN_MLP_TC_F16 mode the audio -> labels calls (phn_recognize*) hand
 * ln p straight from the merger's epilogue to the decoder and never materialise the linear
 * posteriors; the fetch then returns PHN_ERR_ARG - use phn_posteriors() for the `-t post` data. */
int phn_fetch_labels(phn_ctx *ctx, phn_label *labels, int64_t label_cap, int64_t *label_off);
int phn_fetch_mel(phn_ctx *ctx, float *mel_out);
int phn_fetch_posteriors(phn_ctx *ctx, float *post_out);
/* ln p as the decoder of the last call consumed it (decSoftFunc = log, srec.cpp:1088-1097): [frames][3 * n_phonemes].
 * Works after every decoding call, including the fused tensor-core path that never materialises linear posteriors. */
int phn_fetch_logp(phn_ctx *ctx, float *logp_out);
/* Raw CUDA handles for callers that share the device (opaque integers/pointers). */
void *phn_stream(phn_ctx *ctx);          /* cudaStream_t */
void *phn_device_alloc(phn_ctx *ctx, int64_t nbytes);
void phn_device_free(phn_ctx *ctx, void *p);
void *phn_host_alloc_pinned(int64_t nbytes);
void phn_host_free_pinned(void *p);
int phn_memcpy_h2d(phn_ctx *ctx, void *dst, const void *src, int64_t nbytes); /* async on the context's stream */
int phn_memcpy_d2h(phn_ctx *ctx, void *dst, const void *src, int64_t nbytes); /* synchronous */
/* Deterministic synthetic audio generated ON the device (bench): utterance u
 * of `bytes_per_utt` bytes in the current wave format, seeded by seed and u. */
int phn_synth_audio_device(phn_ctx *ctx, void *d_audio, int64_t bytes_per_utt, int n_utt, uint64_t seed);

/* -- instrumentation ---------------------------------------------------------------- */
/* Per-kernel-family device time of the last call, measured with CUDA events on the
 * context's stream when profiling is enabled.  families: see PHN_K_*. */
enum { PHN_K_WAVE = 0, PHN_K_MEAN, PHN_K_STC, PHN_K_MLP, PHN_K_VIT, PHN_K_COUNT };
int phn_set_profiling(phn_ctx *ctx, int on);
int phn_last_timing(phn_ctx *ctx, float ms[PHN_K_COUNT], int64_t launches[PHN_K_COUNT]);

/* Kernel-development aid: which >= 0 arms a clock64() timeline of one tile of tensor-core net `which`
 * (0/1 band nets, 2 merger) for the following calls; which < 0 disarms and copies the 16 x 16 table out. */
int phn_debug_tc_timeline(phn_ctx *ctx, int which, long long *out);

/* Verification aid: logf as the device computes it (port of glibc 2.39 logf, the libm function behind the reference's
 * SoftLog srec.h:192-195 and sLn dspc.h:155-160) on the n consecutive float bit patterns starting at first_bits. */
int phn_debug_logf(phn_ctx *ctx, uint32_t first_bits, int64_t n, float *out);

/* N2: online normaliser arithmetic (Normalization::ProcessFrame, norm.cpp:216-234;
 * ChannelNormParams::{Accum,Update,Norm}, norm.cpp:92-148) on a [frames][nbanks] host
 * matrix, in place on the device: sums over the first `interval` frames; the frame that
 * completes the estimate (index interval-1) and every later one get `x -= mean` (mean_norm)
 * then `x *= invstd` (var_norm); earlier frames pass through (mean 0, inverse std 1).
 * var_norm without mean_norm -> PHN_ERR_ARG (the reference asserts, norm.cpp:150-155).
 * Pinned to the reference's own object by oracle/_ref/online_ref (tests/golden/ref_online_norm.npz). */
int phn_online_norm(phn_ctx *ctx, float *x, int64_t frames, int nbanks, int interval, int mean_norm, int var_norm);

/* ASCII model files -> the binary .nbin cache (host only, no GPU needed): NeuralNet::LoadAscii + SaveBinary
 * (nn.cpp:199-462, 533-592).  `weights`: "weigvec N" x2 then "biasvec N" x2; `norms` (may be NULL): "vec N" x2 =
 * input means and inverse standard deviations.  phn_create() does the same on its own when a net has no .nbin
 * (NeuralNet::Load, nn.cpp:594-621). */
int phn_convert_weights(const char *weights_path, const char *norms_path, const char *nbin_out);

/* Library / build identification. */
const char *phn_version(void);
int phn_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PHNREC_B200_H */
