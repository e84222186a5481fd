// anchor_sweep.cu -- the iterative anchor loop of MANY files, resident on the device
// (BASELINE.json configs[4]: "full anchor-loop sweep ... with on-device best-window
// selection").
//
// Restates the per-row control flow of
//   /root/reference/src/iterative_utterance_alignment.py:67-402
// for a corpus whose emissions were computed once per file and stay in HBM
// ("compute lpz for a file once, slice windows by
// frame offset instead of re-encoding each window").  Files are independent
// (:37 -- the anchor state is per file), so every lock-step iteration aligns ONE
// window of EVERY active file:
//
//   ctcseg fill/backtrace :208-219 all utterance prefixes of every window (ctcseg.cu)
//   sweep_tail_kernel     :221-379 accept / shrink / revert (anchor_select.cuh), then
//                         :231-260 accepted rows -> output slots; new anchor, pending
//                                  utterances, next row, then the file's NEXT window:
//                         :67-192  skip non-speech rows, pick the clip [anchor or row
//                                  start, row end), the text/audio proportion checks,
//                                  text-longer-than-audio (AssertionError :390-402); writes
//                                  the window's ground-truth column, utterance begins,
//                                  text lengths (build_window; sweep_build_kernel runs it
//                                  alone for the first iteration of a call)
//
// Nothing returns to the host between iterations; the host only polls the per-file
// status words every few iterations.  Two situations hand a file back to host POLICY
// (not host compute): the reference's `fix_text_to_time_proportion` re-spreading of the
// remaining rows (:119-146, pandas + VAD table) -> IPFA_SWEEP_NEEDS_RECALC, and a window
// larger than the launch capacity -> IPFA_SWEEP_CAPACITY (the caller grows the
// capacity and continues; the state is untouched).
//
// A window's transcript is always a contiguous range of the file's utterances: the
// utterances a window does not accept are exactly its trailing ones
// (`discarded_transcripts`, re-prepended at :94-96), so "pending + this row's
// utterances" = [utt, row_utt_end[row]).  Its ground-truth column is therefore a slice
// of the file's token stream `blank, tokens(u0), blank, tokens(u1), ..., blank` behind a
// leading -1 (prepare_token_list).
#include "anchor_select.cuh"

namespace ipfa {
extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;

int ctcseg_run(const float *lp, const int64_t *win_off, int64_t stride_n, int64_t stride_t,
               const int32_t *in_len, const int32_t *gt, int64_t gt_stride, const int32_t *n_cols,
               const int32_t *utt_begin, const int32_t *n_utts, int N, int Tmax, int Cmax, int Kmax, int V,
               int blank, double index_duration, int score_len, int flags, double *seg_out,
               int32_t *term_t_out, int32_t *timing_out, float *char_prob_out, int32_t *state_out,
               int32_t *status_out, void *workspace, size_t workspace_bytes, void *stream);

namespace {

inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// per-step window descriptors + results, carved from the caller's workspace
struct SweepWindows {
    int64_t *win_off;     // [F]
    int32_t *in_len;      // [F]
    int32_t *gt;          // [F][Cmax]
    int32_t *n_cols;      // [F]
    int32_t *utt_begin;   // [F][Kmax+1]
    int32_t *n_utts;      // [F]
    int32_t *text_len;    // [F][Kmax]
    int32_t *is_last;     // [F]
    double *clip_start;   // [F]
    double *seg;          // [F][Kmax][Kmax][3]
    int32_t *term_t;      // [F][Kmax]
    int32_t *win_status;  // [F]
    int32_t *decision;    // [F][4]
    double *anchor_rel;   // [F]
    void *seg_ws;
    size_t seg_ws_bytes;
};

size_t carve(SweepWindows *w, unsigned char *base, int F, int Tmax, int Cmax, int Kmax, int V) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *p = base ? base + off : nullptr;
        off += pad256(bytes);
        return p;
    };
    w->win_off = reinterpret_cast<int64_t *>(take((size_t)F * 8));
    w->clip_start = reinterpret_cast<double *>(take((size_t)F * 8));
    w->anchor_rel = reinterpret_cast<double *>(take((size_t)F * 8));
    w->seg = reinterpret_cast<double *>(take((size_t)F * Kmax * Kmax * 3 * 8));
    w->in_len = reinterpret_cast<int32_t *>(take((size_t)F * 4));
    w->gt = reinterpret_cast<int32_t *>(take((size_t)F * Cmax * 4));
    w->n_cols = reinterpret_cast<int32_t *>(take((size_t)F * 4));
    w->utt_begin = reinterpret_cast<int32_t *>(take((size_t)F * (Kmax + 1) * 4));
    w->n_utts = reinterpret_cast<int32_t *>(take((size_t)F * 4));
    w->text_len = reinterpret_cast<int32_t *>(take((size_t)F * Kmax * 4));
    w->is_last = reinterpret_cast<int32_t *>(take((size_t)F * 4));
    w->term_t = reinterpret_cast<int32_t *>(take((size_t)F * Kmax * 4));
    w->win_status = reinterpret_cast<int32_t *>(take((size_t)F * 4));
    w->decision = reinterpret_cast<int32_t *>(take((size_t)F * 4 * 4));
    w->seg_ws_bytes = ipfa_ctcseg_workspace_bytes(F, Tmax, Cmax, Kmax, V);
    w->seg_ws = take(w->seg_ws_bytes);
    return off;
}

// get_text_to_audio_proportion (/root/reference/src/utils/alignment_utils.py:84-106):
// text_length * 0.08 * 3 * sample_rate / audio_length, evaluated left to right in fp64
__device__ __forceinline__ double text_to_audio(long long text_length, int sample_rate, long long audio_length) {
    double x = __dmul_rn((double)text_length, 0.08);
    x = __dmul_rn(x, 3.0);
    x = __dmul_rn(x, (double)sample_rate);
    return __ddiv_rn(x, (double)audio_length);
}

constexpr int kBuildThreads = 128;

// :67-192 + :390-402 for file f: find the file's next window (or its terminal state) and write the
// window descriptor.  Called by all threads of a CTA; thread 0 walks the policy, all threads copy
// the ground-truth column.
__device__ __forceinline__ void build_window(const ipfa_sweep_corpus &c, const ipfa_sweep_params &p,
                                             const ipfa_sweep_state &s, const SweepWindows &w, int f, int Tmax,
                                             int Cmax, int Kmax) {
    __shared__ int sh_active, sh_u0, sh_K, sh_ncols;
    __syncthreads();
    if (threadIdx.x == 0) {
        int active = 0, u0 = 0, K = 0, n_cols = 0, T = 0, is_last = 0;
        long long f0 = 0;
        double clip_start = 0.0;
        int status = s.status[f];
        if (status == IPFA_SWEEP_ACTIVE) {
            const int row0 = c.row_first[f], n_rows = c.row_first[f + 1] - row0;
            const int slot0 = c.utt_first[f];
            int r = s.row[f];
            u0 = s.utt[f];
            double anchor = s.anchor[f], prop = s.prop[f], follow_start = s.follow_start[f];
            int next_ns = s.next_ns[f], exc = s.exc[f];
            long long clip_off = s.clip[2 * f], clip_len = s.clip[2 * f + 1];
            const int recalc_row = s.recalc_row[f];
            const int sr = p.sample_rate;
            while (true) {
                if (r >= n_rows) { status = IPFA_SWEEP_DONE; break; }
                const int R = row0 + r;
                if (c.row_type[R] == 1) {  // :73-77 non-speech row: the anchor jumps to its end
                    anchor = c.row_end[R];
                    ++r;
                    continue;
                }
                is_last = (r + 1 == n_rows);
                clip_start = isnan(anchor) ? c.row_start[R] : anchor;  // :82
                const double clip_end = c.row_end[R];
                const double clip_length = __dsub_rn(clip_end, clip_start);
                const int u1 = c.row_utt_end[R];
                K = u1 - u0;  // discarded (pending) utterances + this row's (:94-96)
                long long text_length = (K > 0) ? (K - 1) : 0;  // len(" ".join(transcript))
                for (int u = u0; u < u1; ++u) text_length += c.utt_chars[slot0 + u];
                const bool resumed = (r == recalc_row);  // host already re-spread this row (:127-146)
                if (!resumed) {
                    {   // :101-109 (ZeroDivisionError is swallowed there: the previous value stays)
                        const long long ns = (long long)__dmul_rn(clip_length, (double)sr);
                        if (ns != 0) prop = text_to_audio(text_length, sr, ns);
                    }
                    if (!is_last) {  // :111-113
                        next_ns = (c.row_type[R + 1] == 1);
                        follow_start = c.row_start[R + 1];
                    }
                    const bool speech_ending = prop > 10.0 && next_ns && !isnan(follow_start) &&
                                               fabs(__dsub_rn(follow_start, clip_start)) > 5.0;
                    const bool recalc = clip_length >= p.max_window_size || speech_ending;  // :119-123
                    if (clip_length >= p.window_to_stop) { status = IPFA_SWEEP_STOP_WINDOW; break; }  // :125
                    if (recalc) { status = IPFA_SWEEP_NEEDS_RECALC; break; }
                } else if (clip_length >= p.window_to_stop) {
                    status = IPFA_SWEEP_STOP_WINDOW;
                    break;
                }
                // :149-160 torchaudio.load(frame_offset=int(clip_start*sr), num_frames=int(clip_length*sr))
                // torchaudio 0.11 accepts num_frames == -1 (the rest of the file) or > 0 and refuses anything
                // else; the reference then keeps the audio of the file's previous clip (:157-159)
                long long offset = (long long)__dmul_rn(clip_start, (double)sr);
                long long audio_length = (long long)__dmul_rn(clip_length, (double)sr);
                if (offset >= 0 && (audio_length > 0 || audio_length == -1)) {
                    if (offset > c.file_samples[f]) offset = c.file_samples[f];
                    if (audio_length == -1 || audio_length > c.file_samples[f] - offset)
                        audio_length = c.file_samples[f] - offset;
                    clip_off = offset;
                    clip_len = audio_length;
                } else {
                    if (clip_off < 0) { status = IPFA_SWEEP_NO_AUDIO; break; }
                    offset = clip_off;
                    audio_length = clip_len;
                }
                if (audio_length > 0) prop = text_to_audio(text_length, sr, audio_length);  // :163
                if (!is_last) {  // :167-192
                    if (audio_length <= 0) { ++r; continue; }                        // :170-173
                    if (prop < p.min_text_to_audio_prop) { anchor = clip_start; ++r; continue; }  // :176-182
                    if (prop > 10.0 && next_ns && !isnan(follow_start) &&
                        fabs(__dsub_rn(follow_start, clip_start)) > 5.0) {
                        // find_a_valid_text_to_audio_proportion (alignment_utils.py:174-196): drop
                        // trailing utterances until the text is shorter than the frame count
                        const long long max_chars = (long long)__ddiv_rn((double)audio_length, p.samples_to_frames_ratio);
                        long long len_k = text_length;
                        int k = K;
                        bool found = false;
                        for (; k >= 1; --k) {
                            if (len_k < max_chars) { found = true; break; }
                            len_k -= c.utt_chars[slot0 + u0 + k - 1] + (k > 1 ? 1 : 0);
                        }
                        if (found) K = k;
                    }
                }
                if (K <= 0) { status = IPFA_SWEEP_NEEDS_RECALC; break; }  // empty transcript: host policy
                const int ratio = p.frame_shift;
                f0 = offset / ratio;
                T = (int)(audio_length / ratio);
                if (f0 + T > c.file_frames[f]) T = (int)max(0LL, (long long)c.file_frames[f] - f0);
                n_cols = c.utt_col[slot0 + u0 + K] - c.utt_col[slot0 + u0] + 2;
                if (n_cols > T) {  // AssertionError("Audio is shorter than text!") :390-402
                    ++exc;
                    if (exc >= p.max_exceptions) { status = IPFA_SWEEP_STOP_EXCEPTIONS; break; }
                    ++r;
                    continue;
                }
                if (T > Tmax || n_cols > Cmax || K > Kmax) {
                    status = IPFA_SWEEP_CAPACITY;
                    s.need[f * 3 + 0] = T; s.need[f * 3 + 1] = n_cols; s.need[f * 3 + 2] = K;
                    break;
                }
                active = 1;
                break;
            }
            s.row[f] = r;
            s.anchor[f] = anchor;
            s.prop[f] = prop;
            s.follow_start[f] = follow_start;
            s.next_ns[f] = next_ns;
            s.exc[f] = exc;
            s.clip[2 * f] = clip_off;
            s.clip[2 * f + 1] = clip_len;
            s.status[f] = status;
        }
        sh_active = active; sh_u0 = u0; sh_K = K; sh_ncols = n_cols;
        w.in_len[f] = active ? T : 0;
        w.n_cols[f] = active ? n_cols : 0;
        w.n_utts[f] = active ? K : 0;
        w.is_last[f] = is_last;
        w.clip_start[f] = clip_start;
        w.win_off[f] = active ? (c.file_frame0[f] + f0) * c.stride_t : 0;
    }
    __syncthreads();
    if (!sh_active) return;
    const int u0 = sh_u0, K = sh_K, n_cols = sh_ncols;
    const int slot0 = c.utt_first[f] + u0;
    const int col0 = c.utt_col[slot0];
    const int32_t *tok = c.tokens + c.file_tok0[f] + col0;
    int32_t *gt = w.gt + (int64_t)f * Cmax;
    for (int i = threadIdx.x; i < n_cols; i += kBuildThreads) gt[i] = (i == 0) ? -1 : tok[i - 1];
    int32_t *ub = w.utt_begin + (int64_t)f * (Kmax + 1);
    for (int k = threadIdx.x; k <= Kmax; k += kBuildThreads)
        ub[k] = 1 + c.utt_col[slot0 + min(k, K)] - col0;
    int32_t *tl = w.text_len + (int64_t)f * Kmax;
    for (int k = threadIdx.x; k < Kmax; k += kBuildThreads) tl[k] = (k < K) ? c.utt_chars[slot0 + k] : 0;
}

__global__ void __launch_bounds__(kBuildThreads)
sweep_build_kernel(const ipfa_sweep_corpus c, const ipfa_sweep_params p, const ipfa_sweep_state s,
                   const SweepWindows w, int Tmax, int Cmax, int Kmax) {
    build_window(c, p, s, w, blockIdx.x, Tmax, Cmax, Kmax);
}

// The tail of an iteration, one CTA per file: the accept / shrink / revert decision (:221-379,
// anchor_select_one), the accepted rows and the loop state (:231-260, :388), then the file's NEXT
// window (build_window) -- one launch instead of three between backtrace and the next fill.
__global__ void __launch_bounds__(kBuildThreads)
sweep_tail_kernel(const ipfa_sweep_corpus c, const ipfa_sweep_params p, const ipfa_sweep_state s,
                  const SweepWindows w, int Tmax, int Cmax, int Kmax, double *__restrict__ out_seg,
                  int32_t *__restrict__ out_info) {
    const int f = blockIdx.x;
    const int K = w.n_utts[f];
    if (K > 0 && threadIdx.x == 0) {  // the file had a window this iteration
        const double *seg_w = w.seg + (int64_t)f * Kmax * Kmax * 3;
        const AnchorDecision d = anchor_select_one(seg_w, w.text_len + (int64_t)f * Kmax, min(K, Kmax), Kmax,
                                                   w.is_last[f] != 0, p.threshold, p.short_len);
        const int k = d.accepted;
        const int u0 = s.utt[f];
        const int64_t slot0 = c.utt_first[f] + u0;
        const double clip_start = w.clip_start[f];
        const double penalty = __dmul_rn(2.0, p.threshold);
        const double *seg = seg_w + (int64_t)max(k - 1, 0) * Kmax * 3;
        const int row = s.row[f];
        for (int u = 0; u < k; ++u) {
            double score = round_decimals(seg[u * 3 + 2], 1.0e4);
            if (c.utt_chars[slot0 + u] < p.short_len) score = __dadd_rn(score, penalty);  // :241
            double *o = out_seg + (slot0 + u) * 4;
            o[0] = clip_start;
            o[1] = round_decimals(seg[u * 3 + 0], 100.0);
            o[2] = round_decimals(seg[u * 3 + 1], 100.0);
            o[3] = score;
            out_info[(slot0 + u) * 2] = s.n_windows[f];  // ordinal of this window in its file
            out_info[(slot0 + u) * 2 + 1] = row;
        }
        if (d.anchor_u == -1) s.anchor[f] = clip_start;                              // :277, :329
        else if (d.anchor_u >= 0) s.anchor[f] = __dadd_rn(clip_start, d.anchor);     // :249
        s.utt[f] = u0 + k;
        s.exc[f] = 0;  // :388
        s.row[f] = row + 1;
        s.n_windows[f] += 1;
        s.cells[f] += (int64_t)w.in_len[f] * w.n_cols[f];
        s.frames[f] += w.in_len[f];
        w.decision[f * 4 + 0] = d.accepted; w.decision[f * 4 + 1] = d.n_iter;
        w.decision[f * 4 + 2] = d.outcome;  w.decision[f * 4 + 3] = d.anchor_u;
    }
    build_window(c, p, s, w, f, Tmax, Cmax, Kmax);
}

}  // namespace
}  // namespace ipfa

using namespace ipfa;

extern "C" size_t ipfa_sweep_workspace_bytes(int n_files, int Tmax, int Cmax, int Kmax, int V) {
    if (n_files <= 0 || Tmax <= 0 || Cmax <= 0 || Kmax <= 0) return 256;
    SweepWindows w;
    return carve(&w, nullptr, n_files, Tmax, Cmax, Kmax, V) + 256;
}

extern "C" int ipfa_sweep_step_device(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                                      const ipfa_sweep_state *state, double *out_seg, int32_t *out_info,
                                      int first_step, int n_steps, int Tmax, int Cmax, int Kmax,
                                      void *workspace, size_t workspace_bytes, void *stream) {
    (void)first_step;
    if (!corpus || !params || !state || !out_seg || !out_info || !workspace || n_steps < 0 || Tmax <= 0 ||
        Cmax <= 1 || Kmax <= 0)
        return IPFA_ERR_INVALID_ARG;
    const ipfa_sweep_corpus &c = *corpus;
    const ipfa_sweep_state &s = *state;
    const int F = c.n_files;
    if (F == 0 || n_steps == 0) return IPFA_OK;
    if (F < 0 || !c.lp || c.V <= 0 || c.blank < 0 || c.blank >= c.V || c.stride_t < c.V || !c.file_frame0 ||
        !c.file_frames || !c.file_samples || !c.row_first || !c.row_type || !c.row_start || !c.row_end ||
        !c.row_utt_end || !c.utt_first || !c.utt_col || !c.utt_chars || !c.file_tok0 || !c.tokens ||
        !s.row || !s.utt || !s.anchor || !s.prop || !s.next_ns || !s.follow_start || !s.exc || !s.status ||
        !s.need || !s.recalc_row || !s.n_windows || !s.cells || !s.frames || !s.clip || params->sample_rate <= 0 ||
        params->frame_shift <= 0 || !(params->samples_to_frames_ratio > 0.0) || !(params->index_duration > 0.0) || params->score_len <= 0)
        return IPFA_ERR_INVALID_ARG;
    if (Tmax > 8000) return IPFA_ERR_UNSUPPORTED;  // windowed table mode
    if (workspace_bytes < ipfa_sweep_workspace_bytes(F, Tmax, Cmax, Kmax, c.V)) return IPFA_ERR_WORKSPACE;
    SweepWindows w;
    carve(&w, static_cast<unsigned char *>(workspace), F, Tmax, Cmax, Kmax, c.V);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const ipfa_sweep_params p = *params;
    // the window of every file for the first iteration (idempotent: the tail of the previous call
    // already built it from the same state)
    NvtxRange range("ipfa.sweep_step (build + n x {segmentation, tail})");
    sweep_build_kernel<<<F, kBuildThreads, 0, st>>>(c, p, s, w, Tmax, Cmax, Kmax);
    ++g_launch_count;
    for (int i = 0; i < n_steps; ++i) {
        int rc = ctcseg_run(c.lp, w.win_off, 0, c.stride_t, w.in_len, w.gt, Cmax, w.n_cols, w.utt_begin, w.n_utts,
                            F, Tmax, Cmax, Kmax, c.V, c.blank, p.index_duration, p.score_len,
                            p.seg_flags | IPFA_SEG_ALL_PREFIXES, w.seg, w.term_t, nullptr, nullptr, nullptr,
                            w.win_status, w.seg_ws, w.seg_ws_bytes, stream);
        if (rc) return rc;
        NvtxRange range_tail("ipfa.sweep.tail (select + rows + next window)");
        sweep_tail_kernel<<<F, kBuildThreads, 0, st>>>(c, p, s, w, Tmax, Cmax, Kmax, out_seg, out_info);
        ++g_launch_count;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
