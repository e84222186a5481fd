/*
 * ipfa_b200.h -- C ABI of the B200-native alignment hot path.
 *
 * Drop-in boundary for the numerics behind the reference's aligner calls
 *   aligner.get_segments(task)        /root/reference/src/iterative_utterance_alignment.py:216
 *                                     /root/reference/src/word_level_alignment.py:100
 *                                     /root/reference/src/search_on_speech.py:85
 * (speechbrain==0.5.11 CTCSegmentation -> ctc-segmentation==1.7.1
 *  `cython_fill_table` + backtrace + `determine_utterance_segments`,
 *  /root/reference/requirements.txt:13,87), for the accept/shrink/revert
 * decision of /root/reference/src/iterative_utterance_alignment.py:221-379,
 * and for the two standard-CTC comparators BASELINE.json's north_star names
 * (torch CTC loss, torchaudio forced align).
 *
 * Conventions
 *   - plain C, no torch types; every pointer in a *_device entry point is a
 *     DEVICE pointer owned by the caller, `stream` is a cudaStream_t passed as
 *     void*.  Nothing is allocated, freed, retained or synchronised inside a
 *     *_device call; workspace sizes come from the *_workspace_bytes helpers.
 *   - *_host entry points take HOST pointers and run the whole
 *     H2D -> kernels -> D2H sequence on an internal stream and an internal,
 *     grow-only device arena (this is what a cgo/ctypes binding that only has
 *     host buffers calls, and what bench.py's `e2e` times).
 *   - return value: 0 = IPFA_OK, otherwise an IPFA_ERR_* code; the text is in
 *     ipfa_status_string().  There is no CPU fallback: without a CUDA device
 *     every compute call returns IPFA_ERR_CUDA.
 *   - emissions `lp` are fp32 log-probabilities [N, Tmax, V] addressed through
 *     element strides (stride_n, stride_t); the vocabulary axis is contiguous.
 */
#ifndef IPFA_B200_H_
#define IPFA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPFA_OK 0
#define IPFA_ERR_INVALID_ARG 1   /* null pointer, negative size, label out of range */
#define IPFA_ERR_UNSUPPORTED 2   /* lattice wider than the widest kernel instance */
#define IPFA_ERR_WORKSPACE 3     /* workspace too small */
#define IPFA_ERR_CUDA 4          /* CUDA runtime error (ipfa_last_cuda_error()) */
#define IPFA_ERR_AUDIO_SHORTER_THAN_TEXT 5 /* ctcseg: N > T (reference raises AssertionError) */
#define IPFA_ERR_WINDOW 6        /* ipfa_ctcseg_host: the backtrace left the table window even at
                                    max_window_size (ctc-segmentation raises IndexError); the outputs
                                    hold the last attempt, status_out carries IPFA_WIN_WINDOW_TOO_SMALL */

/* per-window status bits written to the status_out arrays */
#define IPFA_WIN_OK 0
#define IPFA_WIN_INFEASIBLE 1    /* forced_align: T < L + repeats (torchaudio raises) */
#define IPFA_WIN_BAD_LABEL 2     /* a target id is < 0, >= V or equals blank (Viterbi) */
#define IPFA_WIN_TEXT_LONGER 4   /* ctcseg: N > T */
#define IPFA_WIN_WINDOW_TOO_SMALL 8 /* windowed ctcseg: the backtrace left the window (reference: IndexError,
                                      retry with twice the window) */

int ipfa_version(void);
const char *ipfa_status_string(int status);
const char *ipfa_last_cuda_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t ipfa_launch_count(void);
int ipfa_device_count(void);
/* Measurement aid (bench.py's roofline): with enable != 0 every launch of a lattice kernel (window
 * scorer, Viterbi fill, segmentation fill) is bracketed by CUDA events on the launching stream, except
 * while a graph is being captured; ipfa_profile_read_ms() waits for them and returns the longest bracket
 * since the last read (the dominant kernel's own duration, in ms; < 0 when none was taken).  Not
 * thread-safe; leave it off outside measurements. */
void ipfa_profile_kernels(int enable);
float ipfa_profile_read_ms(void);
/* The IPFA_* tuning switches (kernel-instance overrides used by tools/ and the tests) are read from
 * the environment once per process, not inside compute calls; this reads them again. */
void ipfa_tuning_reload(void);

/* ------------------------------------------------------------------------- *
 * Kernel (1): batched CTC alpha recursion over the blank-interleaved 2L+1
 * lattice.  Replaces torch.nn.functional.ctc_loss(..., reduction='none') as
 * the window scorer (the "acoustic CTC loss" of
 * /root/reference/README.md:3).
 *   lp        [N, Tmax, V] fp32 log-probs (strides in elements)
 *   targets   [N, Lmax] int32 (row stride tgt_stride), no blanks required
 *   in_len    [N] int32 frames per window, tgt_len [N] int32 labels per window
 *   nll_out   [N] fp32: -log p(target | window); +inf when infeasible
 *   workspace caller-owned scratch of ipfa_ctc_alpha_workspace_bytes() bytes; its first
 *             N + 4 int32 words end a call as: [0, N) arrival counters, [N] number
 *             of windows the linear-domain (fp64) instance handed to the log-domain
 *             instance, [N + 1] OR of the reasons (1 target names the blank, 2 emission
 *             ratio outside fp32, 4 reachable state too small, 8 state too large,
 *             16 / 32 scale step / neighbouring scales too far apart), [N + 2] / [N + 3]
 *             the same two words for the optional fp32 tier in front of it.  Diagnostic
 *             only: results are the same either way (csrc/ctc_alpha.cu).
 * Environment (read once per process): IPFA_ALPHA_LOG=1 scores with the log-domain instance alone;
 * IPFA_ALPHA_F32=1 puts the fp32 linear-domain tier in front (csrc/ctc_alpha_f32.cuh).
 * ------------------------------------------------------------------------- */
size_t ipfa_ctc_alpha_workspace_bytes(int N, int Tmax, int Lmax, int V);
int ipfa_ctc_alpha_device(const float *lp, int64_t stride_n, int64_t stride_t,
                          const int32_t *targets, int64_t tgt_stride,
                          const int32_t *in_len, const int32_t *tgt_len,
                          int N, int Tmax, int Lmax, int V, int blank,
                          float *nll_out, void *workspace, size_t workspace_bytes,
                          void *stream);
/* The same call for emissions whose VOCABULARY axis is not contiguous: element (n, t, v) lives at
 * lp[n * stride_n + t * stride_t + v * stride_v].  With a vocabulary-major producer ([N, V, T]:
 * stride_v = T, stride_t = 1) and a large vocabulary, the frames of each column a window needs are
 * contiguous runs, so the kernel reads the algorithmic T * (L + 1) * 4 bytes instead of streaming the dense
 * rows (BASELINE configs[3], V = 5000: 1.23 GB instead of 14.2 GB).  stride_v != 1 always takes the gather
 * panel; stride_v == 1 is ipfa_ctc_alpha_device. */
int ipfa_ctc_alpha_strided_device(const float *lp, int64_t stride_n, int64_t stride_t, int64_t stride_v,
                                  const int32_t *targets, int64_t tgt_stride,
                                  const int32_t *in_len, const int32_t *tgt_len,
                                  int N, int Tmax, int Lmax, int V, int blank,
                                  float *nll_out, void *workspace, size_t workspace_bytes,
                                  void *stream);
int ipfa_ctc_alpha_host(const float *lp, int64_t stride_n, int64_t stride_t,
                        const int32_t *targets, int64_t tgt_stride,
                        const int32_t *in_len, const int32_t *tgt_len,
                        int N, int Tmax, int Lmax, int V, int blank, float *nll_out);

/* ------------------------------------------------------------------------- *
 * Kernel (2a): CTC Viterbi forced alignment on the 2L+1 lattice with backpointers
 * as bit planes in HBM (3 words per state pair and 32-frame block: 1.5 bits per
 * state and frame), then a warp-parallel backtrace that also emits
 * per-frame scores and per-token spans/confidences.  Replaces
 * torchaudio.functional.forced_align + merge_tokens.
 *   paths_out   [N, Tmax] int32 token id per frame (-1 beyond in_len / on error)
 *   scores_out  [N, Tmax] fp32 lp[t, paths[t]]
 *   tok_start/tok_end [N, Lmax] int32 first frame / one-past-last frame per token
 *   tok_score   [N, Lmax] fp32 mean frame score of the token's span
 *   total_out   [N] fp32 Viterbi path log-probability (nullable)
 *   status_out  [N] int32 IPFA_WIN_* bits
 *   tok_* and total_out may be NULL.
 * ------------------------------------------------------------------------- */
size_t ipfa_ctc_viterbi_workspace_bytes(int N, int Tmax, int Lmax, int V);
int ipfa_ctc_viterbi_device(const float *lp, int64_t stride_n, int64_t stride_t,
                            const int32_t *targets, int64_t tgt_stride,
                            const int32_t *in_len, const int32_t *tgt_len,
                            int N, int Tmax, int Lmax, int V, int blank,
                            int32_t *paths_out, float *scores_out,
                            int32_t *tok_start, int32_t *tok_end, float *tok_score,
                            float *total_out, int32_t *status_out,
                            void *workspace, size_t workspace_bytes, void *stream);
/* ... and with a vocabulary stride, see ipfa_ctc_alpha_strided_device. */
int ipfa_ctc_viterbi_strided_device(const float *lp, int64_t stride_n, int64_t stride_t, int64_t stride_v,
                                    const int32_t *targets, int64_t tgt_stride,
                                    const int32_t *in_len, const int32_t *tgt_len,
                                    int N, int Tmax, int Lmax, int V, int blank,
                                    int32_t *paths_out, float *scores_out,
                                    int32_t *tok_start, int32_t *tok_end, float *tok_score,
                                    float *total_out, int32_t *status_out,
                                    void *workspace, size_t workspace_bytes, void *stream);
int ipfa_ctc_viterbi_host(const float *lp, int64_t stride_n, int64_t stride_t,
                          const int32_t *targets, int64_t tgt_stride,
                          const int32_t *in_len, const int32_t *tgt_len,
                          int N, int Tmax, int Lmax, int V, int blank,
                          int32_t *paths_out, float *scores_out,
                          int32_t *tok_start, int32_t *tok_end, float *tok_score,
                          float *total_out, int32_t *status_out);

/* ------------------------------------------------------------------------- *
 * Kernel (2b): CTC-segmentation (Kuerzinger) table fill + backtrace + utterance
 * scoring -- what `CTCSegmentation.get_segments` runs.  One fill per window
 * serves EVERY utterance-prefix of the window's text (the shrinking-transcript
 * iterations of /root/reference/src/iterative_utterance_alignment.py:203-385):
 * prefix k ends in column utt_begin[k]; its terminal frame is the first argmax
 * of that column.
 *   gt          [N, Cmax] int32 ground-truth token column per window:
 *               -1, blank, tokens..., blank, tokens..., blank (prepare_token_list)
 *   n_cols      [N] int32 columns used per window
 *   utt_begin   [N, Kmax+1] int32 utterance begin columns (prepare_token_list's
 *               utt_begin_indices), n_utts [N] int32
 *   index_duration  seconds per frame (config.index_duration)
 *   flags       bit0 blank_transition_cost_zero, bit1 preamble_transition_cost_zero,
 *               bit2 round-to-nearest frame index in scoring (default floor),
 *               bit3 align every prefix (else only the full text, prefix K)
 *   For window w and prefix k (1..K_w; slot k-1 along the prefix axis):
 *   seg_out       [N, Kmax, Kmax, 3] fp64 (start s, end s, score) of utterance u<k
 *   term_t_out    [N, Kmax] int32 terminal frame of prefix k
 *   timing_out    [N, Kmax, Cmax] int32 frame at which column c was entered (-1 unset)
 *                 (nullable; timings[c] = frame * index_duration)
 *   char_prob_out [N, Kmax, Tmax] fp32 per-frame path probability (nullable)
 *   state_out     [N, Kmax, Tmax] int32 per-frame column, -1 = stay ("epsilon"),
 *                 -2 = frame not on the path (nullable)
 *   status_out    [N] int32 IPFA_WIN_* bits (IPFA_WIN_TEXT_LONGER <-> AssertionError)
 *   With IPFA_SEG_ALL_PREFIXES clear only the full text is aligned: slot K_w - 1 is
 *   written, the other prefix slots are left untouched.  Slot k-1 fills seg[., k-1, 0..k-1].
 *   Tmax > 8000 frames (the reference's windowed table mode): ipfa_ctcseg_device returns
 *   IPFA_ERR_UNSUPPORTED (use ipfa_ctcseg_windowed_device); ipfa_ctcseg_host switches to the windowed
 *   kernels itself and doubles the window like the reference (status keeps IPFA_WIN_WINDOW_TOO_SMALL
 *   only if max_window_size = 100000 was reached).
 * ------------------------------------------------------------------------- */
#define IPFA_SEG_BLANK_COST_ZERO 1
#define IPFA_SEG_PREAMBLE_COST_ZERO 2
#define IPFA_SEG_ROUND_NEAREST 4
#define IPFA_SEG_ALL_PREFIXES 8
/* Windowed table mode only (T > window).  ctc-segmentation's source is not available to this build, so
 * the two spots of its sliding-window bookkeeping that could not be re-checked each have a switch;
 * clear = the package as recalled, set = the other reading:
 *   the largest per-column window step is int(mean_offset) + 1  /  ceil(mean_offset);
 *   cur_offset[s+1] = cur_offset[s] + offset runs in ascending order  /  as a shift (multi-column gt). */
#define IPFA_SEG_WINDOW_STEP_CEIL 16
#define IPFA_SEG_OFFSET_SHIFT 32
/* Launch policy of the full-table fill, results unaffected: spread the columns of every window over a
 * thread-block cluster of 2 / 4 SMs, boundary values handed over a chunk at a time through distributed
 * shared memory.  Opt-in: on the anchor sweep's windows it measured no faster than one SM. */
#define IPFA_SEG_SPREAD_2 64
#define IPFA_SEG_SPREAD_4 128

size_t ipfa_ctcseg_workspace_bytes(int N, int Tmax, int Cmax, int Kmax, int V);
int ipfa_ctcseg_device(const float *lp, int64_t stride_n, int64_t stride_t,
                       const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                       const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
                       int N, int Tmax, int Cmax, int Kmax, int V, int blank,
                       double index_duration, int score_len, int flags,
                       double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                       float *char_prob_out, int32_t *state_out, int32_t *status_out,
                       void *workspace, size_t workspace_bytes, void *stream);
int ipfa_ctcseg_host(const float *lp, int64_t stride_n, int64_t stride_t,
                     const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                     const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
                     int N, int Tmax, int Cmax, int Kmax, int V, int blank,
                     double index_duration, int score_len, int flags,
                     double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                     float *char_prob_out, int32_t *state_out, int32_t *status_out);

/* ------------------------------------------------------------------------- *
 * Kernel (3): on-device evaluation of the anchor loop's accept / shrink /
 * revert state machine (/root/reference/src/iterative_utterance_alignment.py:221-379)
 * over the K prefix alignments of each window produced by ipfa_ctcseg_device
 * with IPFA_SEG_ALL_PREFIXES, so only a few bytes per window return to host.
 *   seg          [N, Kmax, Kmax, 3] fp64 from ipfa_ctcseg_device
 *   n_utts       [N] int32
 *   text_len     [N, Kmax] int32 len(text) of each utterance (short-utterance rule)
 *   is_last      [N] int32 window is the last TSV row of its file (:263)
 *   threshold    anchor threshold (-2.0 default), short_len (30 default)
 *   decision_out [N, 4] int32:
 *       [0] accepted prefix length k (0 = everything discarded); the rows the reference keeps
 *           are the k segments of prefix k, the K - k dropped utterances go back to
 *           discarded_transcripts (last utterance first)
 *       [1] iterations the reference loop would have run
 *       [2] outcome code IPFA_SEL_*
 *       [3] index of the utterance (of prefix k) whose end is the new anchor;
 *           -1 = rewind to the window start (:277, :329), -2 = anchor unchanged (:263 with no
 *           segment above the threshold)
 *   anchor_out   [N] fp64 that utterance's end in seconds from the window start, rounded to
 *                0.01 like the `{end:.2f}` field the reference parses; NaN for -1 / -2
 * ------------------------------------------------------------------------- */
#define IPFA_SEL_ACCEPT_CURRENT 0
#define IPFA_SEL_KEEP_PREVIOUS 1
#define IPFA_SEL_DISCARD_ALL 2
#define IPFA_SEL_LAST_SEGMENT 3

int ipfa_anchor_select_device(const double *seg, const int32_t *n_utts,
                              const int32_t *text_len, const int32_t *is_last,
                              int N, int Kmax, double threshold, int short_len,
                              int32_t *decision_out, double *anchor_out, void *stream);

/* out[i] = float(f"{x[i]:.{decimals}f}"): the text round trip every start / end (2 decimals) and
 * score (4 decimals) goes through between `str(task)` and `float(segment[i])`
 * (/root/reference/src/iterative_utterance_alignment.py:218-230), as kernel (3) and the anchor sweep
 * apply it on the device: the exact binary value rounded half-to-even at that decimal, like printf. */
int ipfa_text_round_device(const double *x, int64_t n, int decimals, double *out, void *stream);

/* ------------------------------------------------------------------------- *
 * ipfa_ctcseg_device for windows that are SLICES of resident emissions: window w
 * starts at lp + win_off[w] (elements; device array) instead of lp + w * stride_n.
 * Everything else as ipfa_ctcseg_device.
 * ------------------------------------------------------------------------- */
int ipfa_ctcseg_windows_device(const float *lp, const int64_t *win_off, int64_t stride_t,
                               const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                               const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
                               int N, int Tmax, int Cmax, int Kmax, int V, int blank,
                               double index_duration, int score_len, int flags,
                               double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                               float *char_prob_out, int32_t *state_out, int32_t *status_out,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * Kernel (2b), windowed table mode -- audio longer than config.min_window_size frames
 * (8000).  ctc-segmentation keeps `window` rows per column and slides the window down the
 * audio column by column (offset from the previous column's arg-max), so the fill is
 * column-serial; the per-column offsets depend on the number of columns, so with
 * IPFA_SEG_ALL_PREFIXES every prefix gets its own fill.  Same arguments and outputs as
 * ipfa_ctcseg_device / ipfa_ctcseg_windows_device (win_off may be NULL -> w * stride_n) plus
 *   window       table rows (config.min_window_size, doubled by the caller after
 *                IPFA_WIN_WINDOW_TOO_SMALL up to config.max_window_size)
 *   term_t_out   terminal ROW inside the last column's window
 *   gt_cols      columns of the ground truth matrix: 1 for the tokenised text of the three entry
 *                points; G > 1 for ctc-segmentation's `classic` text converter (prepare_text), where
 *                gt is [N][Cmax][G], gt[c][s] = token spanning the last s+1 characters or -1, and
 *                state_out holds column | (s << 24) at the frames where a token was entered.
 *                window >= Tmax makes this the full-table algorithm for such tasks.
 * For T <= window and gt_cols == 1 the result equals ipfa_ctcseg_device's.
 * ------------------------------------------------------------------------- */
size_t ipfa_ctcseg_windowed_workspace_bytes(int N, int Tmax, int Cmax, int Kmax, int window, int gt_cols,
                                            int flags);
int ipfa_ctcseg_windowed_device(const float *lp, const int64_t *win_off, int64_t stride_n, int64_t stride_t,
                                const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                                const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
                                int N, int Tmax, int Cmax, int Kmax, int V, int blank,
                                double index_duration, int score_len, int flags, int window, int gt_cols,
                                double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                                float *char_prob_out, int32_t *state_out, int32_t *status_out,
                                void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------- *
 * Anchor sweep: the iterative anchor loop of many files, resident on the device
 * (BASELINE.json configs[4]).  Restates the per-row control flow of
 * /root/reference/src/iterative_utterance_alignment.py:67-402 over emissions that
 * were computed once per file; a window is the frame slice
 *   [int(clip_start * sample_rate) / frame_shift, + int(audio_samples / frame_shift))
 * of its file.  One lock-step iteration = window construction (:67-192), the
 * all-prefix CTC segmentation (:208-219), the accept/shrink/revert decision
 * (:221-379) and the anchor / pending-text update, for one window of every active
 * file; nothing returns to the host in between.
 *
 * Corpus (device arrays, read only).  File f owns rows [row_first[f], row_first[f+1])
 * and utterance slots utt_first[f] + i, i = 0..U_f (U_f utterances + one closing slot).
 * ------------------------------------------------------------------------- */
typedef struct ipfa_sweep_corpus {
    const float *lp;             /* [total frames, V] log-probs of all files, file after file */
    int64_t stride_t;            /* elements between frames (>= V) */
    int32_t V, blank, n_files, reserved;
    const int64_t *file_frame0;  /* [F] first frame of the file inside lp */
    const int32_t *file_frames;  /* [F] frames of the file */
    const int64_t *file_samples; /* [F] audio samples of the file (torchaudio.info num_frames) */
    const int32_t *row_first;    /* [F+1] */
    const int32_t *row_type;     /* [R] 0 = speech row, 1 = 'Non-Speech' row (:73) */
    const double *row_start;     /* [R] seconds (TSV 'Start' after fix_time_reference) */
    const double *row_end;       /* [R] seconds */
    const int32_t *row_utt_end;  /* [R] file-relative index one past the row's last utterance */
    const int32_t *utt_first;    /* [F+1] */
    const int32_t *utt_col;      /* [slots] position in the file's token stream of the blank in front
                                    of utterance i; slot U_f: the closing blank */
    const int32_t *utt_chars;    /* [slots] len(text) of utterance i (short-utterance rule, text length) */
    const int64_t *file_tok0;    /* [F] start of the file's token stream inside tokens */
    const int32_t *tokens;       /* per file: blank, tokens(utt 0), blank, tokens(utt 1), ..., blank
                                    (prepare_token_list without its leading -1) */
    const int32_t *file_ready;   /* optional [F] (NULL: every file's emissions are in lp): word f becomes non-zero
                                    once file f's emissions have arrived -- written by the caller's copy stream
                                    AFTER that file's upload, so that uploads overlap the sweep.  Honoured by
                                    ipfa_sweep_resident_device only: a CTA that draws file f waits for word f
                                    (upload the files in index order, zero the words before the launch). */
} ipfa_sweep_corpus;

/* Per-file loop state (device arrays [F], read and written by the sweep).  Initial values:
 * row = utt = exc = next_ns = n_windows = cells = frames = 0, anchor = follow_start = NaN (None),
 * prop = 0.0, status = IPFA_SWEEP_ACTIVE, recalc_row = -1, clip = (-1, -1). */
typedef struct ipfa_sweep_state {
    int32_t *row;          /* next TSV row of the file */
    int32_t *utt;          /* first utterance not accepted yet (the pending/discarded ones follow) */
    double *anchor;        /* new_segment_start in seconds, NaN = None (:37) */
    double *prop;          /* last text_to_audio_proportion (kept across rows like the reference) */
    int32_t *next_ns;      /* last next_row_is_non_speech (:113) */
    double *follow_start;  /* last following_row['Start'] (:112), NaN = none yet */
    int32_t *exc;          /* consecutive AssertionError count (:396) */
    int32_t *status;       /* IPFA_SWEEP_* */
    int32_t *need;         /* [F][3] (frames, columns, utterances) wanted when status == CAPACITY */
    int32_t *recalc_row;   /* row whose times the host already re-spread (:127-146), -1 = none */
    int32_t *n_windows;    /* windows aligned so far */
    int64_t *cells;        /* trellis cells (frames x columns) filled so far */
    int64_t *frames;       /* window frames aligned so far */
    int64_t *clip;         /* [F][2] (first sample, samples) of the last clip torchaudio.load accepted
                              (:149-156); -1 = none yet.  A clip that ends before it starts is refused
                              by torchaudio==0.11 ("num_frames must be -1 or greater than 0") and the
                              reference goes on with the previous clip's audio (:157-159). */
} ipfa_sweep_state;

typedef struct ipfa_sweep_params {
    double threshold;               /* -2.0 */
    double max_window_size;         /* 70.0 s  (:119) */
    double window_to_stop;          /* 500.0 s (:125) */
    double min_text_to_audio_prop;  /* 0.8     (:176) */
    double samples_to_frames_ratio; /* aligner.estimate_samples_to_frames_ratio() (:420) */
    double index_duration;          /* seconds per frame = samples_to_frames_ratio / fs */
    int32_t short_len;              /* 30 (:241) */
    int32_t max_exceptions;         /* 10 (:397) */
    int32_t sample_rate;            /* 16000 */
    int32_t frame_shift;            /* audio samples per emission frame (slicing of lp) */
    int32_t score_len;              /* 30 (scoring_length) */
    int32_t seg_flags;              /* IPFA_SEG_* (IPFA_SEG_ALL_PREFIXES is added) */
} ipfa_sweep_params;

#define IPFA_SWEEP_ACTIVE 0
#define IPFA_SWEEP_DONE 1            /* all rows consumed */
#define IPFA_SWEEP_STOP_WINDOW 2     /* clip_length >= window_to_stop (:125 break) */
#define IPFA_SWEEP_STOP_EXCEPTIONS 3 /* max_text_to_audio_prop_exec consecutive AssertionErrors (:397) */
#define IPFA_SWEEP_NEEDS_RECALC 4    /* :119-146 fix_text_to_time_proportion is host policy: re-spread the
                                        rows, upload them, set recalc_row = row, status = ACTIVE */
#define IPFA_SWEEP_CAPACITY 5        /* window exceeds (Tmax, Cmax, Kmax): see need[], grow and continue */
#define IPFA_SWEEP_NO_AUDIO 6        /* the file's FIRST clip was refused by torchaudio.load: the reference
                                        dies of a NameError at :162; host policy */

/* out_seg [slots][4] fp64: (clip_start, start, end, score) of every accepted utterance --
 * start/end rounded to 0.01 s and score to 1e-4 like the `str(task)` round trip (:219-230), the
 * short-utterance penalty already added (:241); absolute times are clip_start + start / end.
 * out_info [slots][2] int32: (ordinal, within its file, of the window that accepted the utterance;
 * file-relative TSV row that was being aligned -- the row whose Channel / Speaker_ID / Database the
 * result row carries, :258); the caller initialises it to -1.
 * One call = the file's pending window (re)built from the state, then n_steps iterations of
 * {all-prefix segmentation, tail: decision + accepted rows + state + next window}: 1 + 3 * n_steps
 * launches.  Nothing in a launch depends on the call count (`first_step` is informational), so a
 * CUDA graph captured around this call can be replayed; the status words say when every file is
 * finished. */
size_t ipfa_sweep_workspace_bytes(int n_files, int Tmax, int Cmax, int Kmax, int V);
int ipfa_sweep_step_device(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                           const ipfa_sweep_state *state, double *out_seg, int32_t *out_info,
                           int first_step, int n_steps, int Tmax, int Cmax, int Kmax,
                           void *workspace, size_t workspace_bytes, void *stream);

/* File-resident sweep: the same loop (same corpus / state / params / outputs, same rows, status words and
 * counters) in ONE persistent launch.  A CTA takes a file from a ticket counter -- files in index order, so
 * the caller lists the longest files first -- and runs that file's whole anchor loop on its SM: window
 * construction, table fill, backtrace of every prefix, decision, next window, until the file leaves
 * IPFA_SWEEP_ACTIVE; then it takes the next file.  No launch boundary and no other file sits on a file's
 * serial chain of windows (/root/reference/src/iterative_utterance_alignment.py:67-402 is that chain), and
 * the SMs are balanced by the queue.  Every file that is ACTIVE on entry has left ACTIVE when the launch
 * completes (DONE, or one of the host-policy states above: handle it, set ACTIVE, call again).
 * Covers dense emissions one bulk copy per 32-frame chunk can move (stride_t == V, V % 4 == 0, lp 16-byte
 * aligned), the reference's default table flags (IPFA_SEG_PREAMBLE_COST_ZERO set, IPFA_SEG_BLANK_COST_ZERO
 * clear), Cmax <= 4097, Tmax <= 8000; otherwise ipfa_sweep_resident_workspace_bytes returns 0 and the call
 * IPFA_ERR_UNSUPPORTED: use ipfa_sweep_step_device. */
size_t ipfa_sweep_resident_workspace_bytes(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                                           int Tmax, int Cmax, int Kmax);
int ipfa_sweep_resident_device(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                               const ipfa_sweep_state *state, double *out_seg, int32_t *out_info,
                               int Tmax, int Cmax, int Kmax, void *workspace, size_t workspace_bytes,
                               void *stream);

#ifdef __cplusplus
}
#endif
#endif /* IPFA_B200_H_ */
