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
#include <stdio.h>
#include <stdlib.h>

#include "anchor_select.cuh"
#include "ctcseg_walk.cuh"

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

// CTA barrier that is safe when a warp arrives in pieces.  Thread 0 walks long single-thread policy paths
// (with loops the compiler cannot bound) in these kernels.  ptxas does not always reconverge the warp behind
// such a region, and __syncthreads() is the ALIGNED barrier: in the file-resident kernel lanes 1..31 of warp 0
// reached it without lane 0, the hardware counted the warp as arrived, and lane 0 ran one barrier behind
// everybody else from then on until the CTA deadlocked (an explicit __syncwarp() in front was elided).  The
// non-aligned barrier.sync (this intrinsic) makes the compiler converge the warp first whenever it may be split.
__device__ __forceinline__ void cta_sync() { __barrier_sync(0); }

// :67-192 + :390-402 for file f: find the file's next window (or its terminal state) and write the
// window descriptor.  Called by all threads of a CTA; thread 0 walks the policy, all threads copy
// the ground-truth column.
// `slot`: where the descriptor goes (the lock-step kernels keep one per file: slot == f; the file-resident
// kernel one per CTA).
__device__ __forceinline__ void build_window(const ipfa_sweep_corpus &c, const ipfa_sweep_params &p,
                                             const ipfa_sweep_state &s, const SweepWindows &w, int f, int slot,
                                             int Tmax, int Cmax, int Kmax) {
    __shared__ int sh_active, sh_u0, sh_K, sh_ncols;
    cta_sync();
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
        w.in_len[slot] = active ? T : 0;
        w.n_cols[slot] = active ? n_cols : 0;
        w.n_utts[slot] = active ? K : 0;
        w.is_last[slot] = is_last;
        w.clip_start[slot] = clip_start;
        w.win_off[slot] = active ? (c.file_frame0[f] + f0) * c.stride_t : 0;
    }
    cta_sync();
    if (!sh_active) return;
    const int u0 = sh_u0, K = sh_K, n_cols = sh_ncols;
    const int slot0 = c.utt_first[f] + u0;
    const int col0 = c.utt_col[slot0];
    const int32_t *tok = c.tokens + c.file_tok0[f] + col0;
    int32_t *gt = w.gt + (int64_t)slot * Cmax;
    for (int i = threadIdx.x; i < n_cols; i += blockDim.x) gt[i] = (i == 0) ? -1 : tok[i - 1];
    int32_t *ub = w.utt_begin + (int64_t)slot * (Kmax + 1);
    for (int k = threadIdx.x; k <= Kmax; k += blockDim.x)
        ub[k] = 1 + c.utt_col[slot0 + min(k, K)] - col0;
    int32_t *tl = w.text_len + (int64_t)slot * Kmax;
    for (int k = threadIdx.x; k < Kmax; k += blockDim.x) tl[k] = (k < K) ? c.utt_chars[slot0 + k] : 0;
}

__global__ void __launch_bounds__(kBuildThreads)
sweep_build_kernel(const ipfa_sweep_corpus c, const ipfa_sweep_params p, const ipfa_sweep_state s,
                   const SweepWindows w, int Tmax, int Cmax, int Kmax) {
    build_window(c, p, s, w, blockIdx.x, blockIdx.x, Tmax, Cmax, Kmax);
}

// The decision of one window and what follows from it (:221-379, :231-260, :388), one thread: accepted rows
// -> output slots, new anchor, pending utterances, next row.  seg_w: [Kmax][Kmax][3] segments of every prefix.
__device__ __forceinline__ void apply_decision(const ipfa_sweep_corpus &c, const ipfa_sweep_params &p,
                                               const ipfa_sweep_state &s, const SweepWindows &w, int f, int slot,
                                               int K, int Kmax, const double *seg_w, double *out_seg,
                                               int32_t *out_info) {
    const AnchorDecision d = anchor_select_one(seg_w, w.text_len + (int64_t)slot * Kmax, min(K, Kmax), Kmax,
                                               w.is_last[slot] != 0, p.threshold, p.short_len);
    const int k = d.accepted;
    const int u0 = s.utt[f];
    const int64_t slot0 = c.utt_first[f] + u0;
    const double clip_start = w.clip_start[slot];
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
    s.cells[f] += (int64_t)w.in_len[slot] * w.n_cols[slot];
    s.frames[f] += w.in_len[slot];
    w.decision[slot * 4 + 0] = d.accepted; w.decision[slot * 4 + 1] = d.n_iter;
    w.decision[slot * 4 + 2] = d.outcome;  w.decision[slot * 4 + 3] = d.anchor_u;
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
    if (K > 0 && threadIdx.x == 0)  // the file had a window this iteration
        apply_decision(c, p, s, w, f, f, K, Kmax, w.seg + (int64_t)f * Kmax * Kmax * 3, out_seg, out_info);
    build_window(c, p, s, w, f, f, Tmax, Cmax, Kmax);
}


// ===========================================================================
// File-resident sweep: ONE persistent launch, a CTA owns a FILE (one CTA per SM; two for corpora of many
// short files, see resident_plan).
//
// The lock-step kernels above pay, per iteration, three dependent launches and the longest window of
// the group; a file's anchor loop is a serial chain of windows (the next window starts where the last
// one was anchored), so the sweep's duration is (longest chain) x (latency of an iteration), and on top
// of it whatever the lock step adds.  Here a CTA takes a file from a ticket counter (files longest
// first) and runs the file's whole loop by itself -- build window -> table fill -> backtrace of every
// prefix -> decision -> next window -- without leaving the SM: no launch boundary and no other file on
// the chain, the window's ground truth, utterance boundaries and per-prefix segments stay in shared
// memory, and a CTA that finishes its file takes the next one, so the SMs are balanced by a work queue
// instead of by groups of streams.  The arithmetic is the lock-step path's: the same frame recursion and
// transition test as ctcseg_fill_kernel (1, 2 or 4 columns per thread -- chosen per window --, dense panel, the
// reference's default flags), the same walk and scoring (ctcseg_walk.cuh), the same policy functions (build_window,
// apply_decision) -- rows, state and counters are identical (tests/test_gpu_sweep.py).
//
// A window uses the first ceil((columns - 1) / (32 KC)) warps of the CTA for its fill (named barrier 1 over
// exactly those warps); emission chunks of 32 frames arrive as 16-byte cp.async pieces on a ring of three stages;
// the window's backpointer words stay in shared memory when they fit; every prefix is walked by its own warp and
// the utterance scores are (prefix, utterance) tasks over all warps.  With ipfa_sweep_corpus.file_ready a CTA waits
// for its file's emissions to arrive, so the upload of the corpus overlaps the sweep.
struct ResidentParams {
    ipfa_sweep_corpus c;
    ipfa_sweep_params p;
    ipfa_sweep_state s;
    SweepWindows w;          // descriptors, one slot per CTA
    int Tmax, Cmax, Kmax, pitch, bits_bytes;
    int kc;                  // 0: columns per thread chosen per window; 1 / 2 / 4: fixed (IPFA_SWEEP_KC, tuning)
    uint32_t *bp;            // [slots][words_per_slot] backpointer words (windows too large for shared memory)
    int64_t words_per_slot;
    int32_t *timing;         // [slots][Kmax][Cmax]
    float *cprob;            // [slots][Kmax][Tmax]
    int *ticket;
    double *out_seg;
    int32_t *out_info;
    int phases;              // IPFA_SWEEP_PHASES=1 (tuning): CTA 0 prints the cycles it spent in each phase
};

constexpr int kResChunk = 32;  // frames per emission chunk

// Barrier `id` over the first `threads` threads of the CTA, once per frame of the fill: the ALIGNED form (one
// BAR.SYNC, no divergence check; the non-aligned intrinsic costs a slow path per frame once a warp has ever
// split -- measured 38 -> 50 ms on the 100 h sweep).  Its callers are warp-uniform code; resident_fill converges
// every warp explicitly (__syncwarp) at the top of each emission chunk.
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Table fill of one window by the first NTA threads of the CTA (all of them call this).
template <int KC, int PITCH>
__device__ __forceinline__ void resident_fill(const float *lp_win, int T, int NC, int V, int pitch_rt, int blank,
                                              const int32_t *gt_s, const int32_t *ub_s, int K, float *ring,
                                              float *xline, int line_len, int NTA,
                                              uint32_t *bp, int32_t *colarg_s, int tid) {
    constexpr int SPW = 32 / KC;
    constexpr float kNegInf = -__builtin_huge_valf();
    constexpr float kProbMax = -1000000000.0f;  // cdef float prob_max = -1000000000
    const int pitch = PITCH ? PITCH : pitch_rt;
    int col[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = 1 + tid * KC + k;
        col[k] = (c < NC) ? gt_s[c] : blank;
    }
    bool track = false;
    {
        const int c0 = 1 + tid * KC;
        for (int u = 1; u <= K; ++u) {
            const int c = ub_s[u];
            track |= (c >= c0 && c < c0 + KC);
        }
    }
    const bool warp_tracks = __any_sync(0xffffffffu, track);
    const int nchunks = (T + kResChunk - 1) / kResChunk;
    // Emission chunks (32 frames x V floats, contiguous) come in as 16-byte cp.async pieces issued by ALL the
    // window's threads, three stages deep.  (Not the elected-thread bulk copy + mbarrier of the other kernels: with
    // thread 0 apt to run apart from its warp in this kernel -- see cta_sync() -- nobody here polls for something
    // another lane of the same warp has yet to start; cp.async.wait_group blocks in hardware.  Same speed: a
    // chunk is 4 KB.)
    const int v4 = V >> 2;
    auto issue = [&](int chunk) {
        if (chunk < nchunks) {
            const int rows = min(kResChunk, T - chunk * kResChunk);
            float *dst = ring + (size_t)(chunk % 3) * kResChunk * pitch;
            const float *src = lp_win + (int64_t)chunk * kResChunk * V;
            const int pieces = rows * v4;
            for (int q = tid; q < pieces; q += NTA) cp_async_16(dst + q * 4, src + q * 4);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);
    float val[KC], cmax[KC];
    int carg[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) { val[k] = kProbMax; cmax[k] = kNegInf; carg[k] = -1; }
    uint32_t word = 0;
    int shift = 0;
    uint32_t *bp_ptr = bp + tid;
    int t = 0;
    // one frame (the arithmetic of ctcseg_fill_kernel's FAST instance); returns this thread's KC decision bits
    auto frame = [&](const float *row, const float *rd, float *wr) -> uint32_t {
        const float eb = row[blank];
        float ec[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) ec[k] = row[col[k]];
        float left_[KC], up_[KC], stayp_[KC];
        const int t_now = t;
        const float prev = rd[tid];
#pragma unroll
        for (int k = KC - 1; k >= 0; --k) {
            const float left = (k == 0) ? prev : val[k - 1];  // table[t-1, c-1]
            const float up = val[k];                          // table[t-1, c]
            const float sw = __fadd_rn(left, ec[k]);
            const float stay_p = fmaxf(eb, ec[k]);
            const float st = __fadd_rn(up, stay_p);
            left_[k] = left; up_[k] = up; stayp_[k] = stay_p;
            val[k] = fmaxf(sw, st);
        }
        ++t;
        wr[tid + 1] = val[KC - 1];
        named_bar_sync(1, NTA);
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            // the reference backtrace's transition test, on the same fp32 values
            const float v = val[k];
            const float d_sw = fabsf(__fsub_rn(ec[k], __fsub_rn(v, left_[k])));
            const float d_st = fabsf(__fsub_rn(stayp_[k], __fsub_rn(v, up_[k])));
            bits |= (d_st > d_sw) ? (1u << k) : 0u;
        }
        if (warp_tracks) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (cmax[k] < val[k]) { cmax[k] = val[k]; carg[k] = t_now; }
            }
        }
        return bits;
    };
    auto push_bits = [&](uint32_t bits) {
        word |= bits << shift;
        shift += KC;
        if (shift == 32) {
            *bp_ptr = word;
            bp_ptr += NTA;
            word = 0;
            shift = 0;
        }
    };
    float *line0 = xline, *line1 = xline + line_len;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        __syncwarp();              // (the per-frame barrier below is the aligned one)
        cp_async_wait<1>();        // this thread's pieces of chunk `chunk` have landed ...
        named_bar_sync(1, NTA);    // ... everybody's have, and everybody is done with the stage refilled next
        issue(chunk + 2);
        const float *panel = ring + (size_t)(chunk % 3) * kResChunk * pitch;
        const int rows = min(kResChunk, T - chunk * kResChunk);
        int r = 0;
        if (chunk == 0) {
            // t = 0: every column c >= 1 is max(switch = prob_max, stay = prob_max)
            if (warp_tracks) {
#pragma unroll
                for (int k = 0; k < KC; ++k) { cmax[k] = kProbMax; carg[k] = 0; }
            }
            push_bits(0);
            t = 1;
            line0[tid + 1] = kProbMax;
            named_bar_sync(1, NTA);
            r = 1;
        }
        const float *row = panel + r * pitch;
        // frame t reads line[(t-1)&1], writes line[t&1]; chunks are 32 frames so r has t's parity
        while (r < rows) {
            if (shift == 0 && !(r & 1) && r + SPW <= rows) {
                uint32_t acc = 0;
#pragma unroll
                for (int f = 0; f < SPW; ++f)
                    acc |= frame(row + f * pitch, (f & 1) ? line0 : line1, (f & 1) ? line1 : line0) << (f * KC);
                *bp_ptr = acc;
                bp_ptr += NTA;
                row += SPW * pitch;
                r += SPW;
            } else {
                push_bits((r & 1) ? frame(row, line0, line1) : frame(row, line1, line0));
                row += pitch;
                ++r;
            }
        }
    }
    if (shift != 0) *bp_ptr = word;
    if (track) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = 1 + tid * KC + k;
            for (int u = 1; u <= K; ++u)
                if (ub_s[u] == c) colarg_s[u] = carg[k];
        }
    }
    cp_async_wait<0>();
}

struct ResidentSmem {
    size_t ring, xline, gt, walk, seg, ub, colarg, bits, total;
};
constexpr int kResWalkWords = seg_walk_smem_words<4>();  // per-warp staging words of the walks: the largest instance
static_assert(seg_walk_smem_words<1>() <= kResWalkWords && seg_walk_smem_words<2>() <= kResWalkWords, "walk staging");
__host__ __device__ inline ResidentSmem resident_smem(int pitch, int threads, int Cmax, int Kmax, int bits_bytes) {
    ResidentSmem m;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    m.ring = take((size_t)3 * kResChunk * pitch * 4);
    m.seg = take((size_t)Kmax * Kmax * 3 * 8);
    m.xline = take((size_t)2 * (threads + 1) * 4);
    m.gt = take((size_t)Cmax * 4);
    m.walk = take((size_t)(threads / 32) * kResWalkWords * 4);
    m.ub = take((size_t)(Kmax + 1) * 4);
    m.colarg = take((size_t)(Kmax + 1) * 4);
    m.bits = take((size_t)bits_bytes);  // backpointer words of the window in flight, when they fit
    m.total = off;
    return m;
}

// Table fill and the walks of every prefix of ONE window with KC columns per thread (the CTA's first
// ceil(columns / (32 KC)) warps fill; one warp per prefix walks).
struct ResidentWindow {
    const float *lp_win;
    int T, NC, K, slot;
    float *ring, *xline;
    int line_len;
    int32_t *gt_s, *ub_s, *colarg_s;
    uint32_t *bits_s, *bp_global, *walk_s;
    double *seg_s;
    bool round_nearest;
    long long *t_fill_end;  // diagnostic (thread 0 of CTA 0), nullable
};
template <int KC, int PITCH>
__device__ __forceinline__ void resident_align(const ResidentParams &P, const ResidentWindow &W) {
    const ipfa_sweep_corpus &c = P.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int T = W.T, NC = W.NC, K = W.K, Kmax = P.Kmax, Cmax = P.Cmax, Tmax = P.Tmax;
    const int nact = (NC - 1 + 32 * KC - 1) / (32 * KC);  // warps that own a column
    // the window's backpointer words: shared memory when they fit (the walks then have no L2 latency to hide)
    uint32_t *bp = ((int64_t)((T + 32 / KC - 1) / (32 / KC)) * 32 * nact * 4 <= P.bits_bytes) ? W.bits_s : W.bp_global;
    if (warp < nact)
        resident_fill<KC, PITCH>(W.lp_win, T, NC, c.V, P.pitch, c.blank, W.gt_s, W.ub_s, K, W.ring, W.xline,
                                 W.line_len, 32 * nact, bp, W.colarg_s, tid);
    cta_sync();
    if (W.t_fill_end) *W.t_fill_end = clock64();
    // every prefix of the window: one warp each (ctc_segmentation() backtrace; the utterances are scored after)
    for (int kslot = warp; kslot < K; kslot += nwarps) {
        int32_t *timing = P.timing + ((int64_t)W.slot * Kmax + kslot) * Cmax;
        float *cprob = P.cprob + ((int64_t)W.slot * Kmax + kslot) * Tmax;
        double *seg = W.seg_s + (int64_t)kslot * Kmax * 3;
        for (int t = lane; t < T; t += 32) cprob[t] = 0.0f;
        for (int cc = lane; cc < NC; cc += 32) timing[cc] = -1;
        const int c_end = W.ub_s[kslot + 1];
        const bool feasible = T > 0 && c_end >= 1 && c_end < NC && c_end + 1 <= T;
        const int t_term = feasible ? W.colarg_s[kslot + 1] : -1;
        if (!feasible || t_term < 0) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
            continue;
        }
        __syncwarp();
        SegWalkArgs a;
        a.lp = W.lp_win; a.stride_t = c.stride_t; a.T = T; a.Cmax = Cmax; a.blank = c.blank; a.NT = 32 * nact;
        a.score_len = P.p.score_len; a.round_nearest = W.round_nearest; a.index_duration = P.p.index_duration;
        a.bp_w = bp; a.ub = W.ub_s; a.gt_s = W.gt_s; a.timing = timing; a.cprob = cprob; a.state = nullptr;
        a.seg = seg; a.raw = W.walk_s; a.skip_scoring = true;
        seg_walk_prefix<KC, true>(a, kslot, t_term, c_end, lane);
        if (lane == 0) W.colarg_s[kslot + 1] |= 0x40000000;  // walked: its utterances are scored below
    }
}

template <int PITCH, int MAXT>
__global__ void __launch_bounds__(MAXT) sweep_resident_kernel(const ResidentParams P) {
    extern __shared__ __align__(128) unsigned char rs_smem[];
    __shared__ int sh_file;
    const ipfa_sweep_corpus &c = P.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int slot = blockIdx.x;
    const int Tmax = P.Tmax, Cmax = P.Cmax, Kmax = P.Kmax;
    const ResidentSmem m = resident_smem(PITCH ? PITCH : P.pitch, blockDim.x, Cmax, Kmax, P.bits_bytes);
    float *ring = reinterpret_cast<float *>(rs_smem + m.ring);
    double *seg_s = reinterpret_cast<double *>(rs_smem + m.seg);
    float *xline = reinterpret_cast<float *>(rs_smem + m.xline);
    int32_t *gt_s = reinterpret_cast<int32_t *>(rs_smem + m.gt);
    uint32_t *walk_s = reinterpret_cast<uint32_t *>(rs_smem + m.walk) + (size_t)warp * kResWalkWords;
    int32_t *ub_s = reinterpret_cast<int32_t *>(rs_smem + m.ub);
    int32_t *colarg_s = reinterpret_cast<int32_t *>(rs_smem + m.colarg);
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(rs_smem + m.bits);
    const int line_len = blockDim.x + 1;
    if (tid == 0) {
        xline[0] = 0.0f;          // z(t) = table[t, 0] = 0 (preamble_transition_cost_zero)
        xline[line_len] = 0.0f;
    }
    uint32_t *bp_global = P.bp + (int64_t)slot * P.words_per_slot;
    const bool round_nearest = (P.p.seg_flags & IPFA_SEG_ROUND_NEAREST) != 0;
    const int slice_cap = P.bits_bytes / 4 / nwarps;
    // diagnostic phase counters of thread 0 of CTA 0 (IPFA_SWEEP_PHASES): in shared memory, so that they cost the
    // loops below no registers -- ph[0..4]: window, staging, fill, walk + scores, decision; [5] walks, [6] scores,
    // [7] time stamp, [8] fill end, [9] walks end, [10] windows
    __shared__ long long ph[11];
    const bool timed = P.phases && tid == 0 && slot == 0;
#define IPFA_PHASE(i) do { if (timed) { const long long now_ = clock64(); ph[i] += now_ - ph[7]; ph[7] = now_; } } while (0)
    if (timed) for (int i = 0; i < 11; ++i) ph[i] = 0;
    while (true) {
        cta_sync();
        if (tid == 0) sh_file = atomicAdd(P.ticket, 1);
        cta_sync();
        const int f = sh_file;
        if (f >= c.n_files) {
            if (timed)
                printf("ipfa resident CTA 0: %d windows; cycles per window: next window %lld, staging %lld, fill %lld, "
                       "backtrace+scoring %lld (walks %lld, scoring %lld), decision %lld\n", (int)ph[10],
                       ph[0] / max(ph[10], 1LL), ph[1] / max(ph[10], 1LL), ph[2] / max(ph[10], 1LL),
                       ph[3] / max(ph[10], 1LL), ph[5] / max(ph[10], 1LL), ph[6] / max(ph[10], 1LL),
                       ph[4] / max(ph[10], 1LL));
            break;
        }
        if (c.file_ready) {
            // the file's emissions may still be on their way (uploads overlapping the sweep): one thread polls the
            // word the copy stream writes behind the file's upload, the rest of the CTA waits at the barrier
            if (tid == 0) {
                const volatile int32_t *ready = c.file_ready + f;
                while (*ready == 0) __nanosleep(500);
                __threadfence();
            }
            cta_sync();
        }
        if (timed) ph[7] = clock64();
        build_window(c, P.p, P.s, P.w, f, slot, Tmax, Cmax, Kmax);
        while (true) {
            cta_sync();
            IPFA_PHASE(0);
            const int K = P.w.n_utts[slot];
            if (K <= 0) break;  // the file is finished or waits for host policy (status word)
            const int T = P.w.in_len[slot], NC = P.w.n_cols[slot];
            const float *lp_win = c.lp + P.w.win_off[slot];
            {
                const int32_t *gt = P.w.gt + (int64_t)slot * Cmax;
                for (int i = tid; i < NC; i += blockDim.x) {
                    int g = gt[i];
                    if (g < 0 || g >= c.V) g = c.blank;
                    gt_s[i] = g;
                }
                const int32_t *ub = P.w.utt_begin + (int64_t)slot * (Kmax + 1);
                for (int u = tid; u <= K; u += blockDim.x) { ub_s[u] = ub[u]; colarg_s[u] = -1; }
            }
            cta_sync();
            IPFA_PHASE(1);
            if (timed) ph[10] += 1;
            // Columns per thread, chosen per window: a file's windows are a serial chain, and a window with few
            // warps is latency bound (154 cycles per frame with four warps against 100 with eleven), so narrow
            // windows take one column per thread; wide ones are issue bound and take four.
            {
                ResidentWindow W;
                W.lp_win = lp_win; W.T = T; W.NC = NC; W.K = K; W.slot = slot; W.ring = ring; W.xline = xline;
                W.line_len = line_len; W.gt_s = gt_s; W.ub_s = ub_s; W.colarg_s = colarg_s; W.bits_s = bits_s;
                W.bp_global = bp_global; W.walk_s = walk_s; W.seg_s = seg_s; W.round_nearest = round_nearest;
                W.t_fill_end = timed ? &ph[8] : nullptr;
                const int kcw = P.kc ? P.kc : (NC - 1 <= 512 ? 1 : NC - 1 <= 1024 ? 2 : 4);
                if (kcw == 1) resident_align<1, PITCH>(P, W);
                else if (kcw == 2) resident_align<2, PITCH>(P, W);
                else resident_align<4, PITCH>(P, W);
            }
            if (timed) { ph[2] += ph[8] - ph[7]; ph[7] = ph[8]; }
            cta_sync();
            if (timed) { ph[9] = clock64(); ph[5] += ph[9] - ph[7]; }
            // determine_utterance_segments: utterance u of prefix k is its own task, spread over ALL the warps
            // (the longest prefix alone would score its k utterances one after the other)
            for (int task = warp; task < K * (K + 1) / 2; task += nwarps) {
                int k = 0;
                while ((k + 1) * (k + 2) / 2 <= task) ++k;  // prefix slot: tasks k(k+1)/2 .. + k
                const int u = task - k * (k + 1) / 2;
                if (!(colarg_s[k + 1] & 0x40000000) || colarg_s[k + 1] < 0) continue;  // infeasible prefix: NaN rows are in
                score_one_segment(ub_s, P.timing + ((int64_t)slot * Kmax + k) * Cmax,
                                  P.cprob + ((int64_t)slot * Kmax + k) * Tmax, u, T, Cmax, P.p.index_duration,
                                  P.p.score_len, round_nearest, lane, seg_s + (int64_t)k * Kmax * 3,
                                  seg_scratch<4>(walk_s),
                                  // (the walks are over: the backpointer area is this warp's to stage char_probs in)
                                  reinterpret_cast<float *>(bits_s) + (size_t)warp * slice_cap, slice_cap);
            }
            cta_sync();
            if (timed) ph[6] += clock64() - ph[9];
            IPFA_PHASE(3);
            if (tid == 0) apply_decision(c, P.p, P.s, P.w, f, slot, K, Kmax, seg_s, P.out_seg, P.out_info);
            IPFA_PHASE(4);
            build_window(c, P.p, P.s, P.w, f, slot, Tmax, Cmax, Kmax);
        }
    }
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

namespace {
struct ResidentPlan {
    int kc, threads, pitch, slots, bits_bytes, ctas;
    bool fixed_pitch;
    size_t smem;
};
// (one CTA per SM: the sizes of resident_plan's default case)
bool resident_plan_one(const ipfa_sweep_corpus &c, const ipfa_sweep_params &, int Tmax, int Cmax, int Kmax,
                       ResidentPlan *pl, int sms) {
    const int64_t want = ((int64_t)Tmax + 32) * (Cmax + 128) / 8;
    const int64_t room = 160 * 1024;
    pl->bits_bytes = (int)((((want < room ? want : room) + 15) / 16) * 16);
    if (pl->bits_bytes < 116 * 1024) pl->bits_bytes = 116 * 1024;
    pl->smem = resident_smem(pl->pitch, pl->threads, Cmax, Kmax, pl->bits_bytes).total;
    if (pl->smem > 225 * 1024) return false;
    pl->slots = c.n_files < sms ? c.n_files : sms;
    return true;
}

// The file-resident kernel covers the anchor loop's own configuration: dense rows that one bulk copy per
// chunk can move (stride_t == V, V a multiple of 4), the reference's default table flags, windows of at most
// 4096 columns; anything else runs on the lock-step path (ipfa_sweep_step_device).
bool resident_plan(const ipfa_sweep_corpus &c, const ipfa_sweep_params &p, int Tmax, int Cmax, int Kmax,
                   ResidentPlan *pl) {
    const int table_flags = p.seg_flags & (IPFA_SEG_BLANK_COST_ZERO | IPFA_SEG_PREAMBLE_COST_ZERO);
    if (table_flags != IPFA_SEG_PREAMBLE_COST_ZERO) return false;
    if (c.V % 4 != 0 || c.V > 256 || c.stride_t != c.V || (reinterpret_cast<uintptr_t>(c.lp) & 15) != 0) return false;
    if (Tmax > 8000 || Cmax - 1 > 4096) return false;
    // Columns per thread are chosen per WINDOW (resident_kernel): 1 up to 512 columns, 2 up to 1024, 4 beyond --
    // so 16 warps hold any window of up to 2048 columns (the 128-register instances), 32 warps up to 4096.
    // IPFA_SWEEP_KC=1|2|4 (tuning) fixes the choice for every window.
    pl->kc = 0;
    int warps = (Cmax - 1 <= 512) ? (Cmax - 1 + 31) / 32 : (Cmax - 1 <= 2048 ? 16 : 32);
    if (const char *e = tuning("IPFA_SWEEP_KC")) {
        const int v = atoi(e);
        if ((v == 1 || v == 2 || v == 4) && Cmax - 1 <= v * 1024) {
            pl->kc = v;
            warps = (Cmax - 1 + 32 * v - 1) / (32 * v);
        }
    }
    pl->threads = 32 * (warps < 4 ? 4 : warps);
    pl->pitch = c.V;
    pl->fixed_pitch = (c.V == 32);
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        return false;
    // Files in flight per SM.  One: a file's chain is latency bound and two chains on one SM each run slower, so
    // as long as the longest file bounds the sweep every file gets an SM of its own.  Two, when there are at least
    // three files per SM (a corpus of many short files, where the queue and not one chain sets the duration): a
    // lone window issues one instruction every four cycles per sub-partition, and a second file fills the gaps.
    // (Every SM then holds two CTAs -- the grid is exactly 2 x SMs -- so the block scheduler has nothing to decide.)
    // IPFA_SWEEP_CTAS=1|2 (tuning) overrides.
    pl->ctas = (c.n_files >= 3 * sms && pl->threads <= 512) ? 2 : 1;
    if (const char *e = tuning("IPFA_SWEEP_CTAS")) {
        const int v = atoi(e);
        if (v == 1 || (v == 2 && pl->threads <= 512)) pl->ctas = v;
    }
    {
        // shared memory for the backpointer words: the launch capacity's worth, at most 160 KB with one CTA per SM
        // (and never less than 116 KB: more than half of an SM's shared memory keeps every CTA on an SM of its
        // own), 64 KB with two.
        const int64_t want = ((int64_t)Tmax + 32) * (Cmax + 128) / 8;  // one bit per cell, padded
        const int64_t room = (pl->ctas == 2 ? 64 : 160) * 1024;
        pl->bits_bytes = (int)((((want < room ? want : room) + 15) / 16) * 16);
        if (pl->ctas == 1 && pl->bits_bytes < 116 * 1024) pl->bits_bytes = 116 * 1024;
    }
    const ResidentSmem m = resident_smem(pl->pitch, pl->threads, Cmax, Kmax, pl->bits_bytes);
    pl->smem = m.total;
    if (pl->ctas == 2 && pl->smem > 110 * 1024) {  // two do not fit: back to one
        pl->ctas = 1;
        return resident_plan_one(c, p, Tmax, Cmax, Kmax, pl, sms);
    }
    if (pl->smem > 225 * 1024) return false;
    pl->slots = c.n_files < sms * pl->ctas ? c.n_files : sms * pl->ctas;
    return true;
}

struct ResidentCarve {
    SweepWindows w;
    uint32_t *bp;
    int64_t words_per_slot;
    int32_t *timing;
    float *cprob;
    int *ticket;
};
size_t carve_resident(ResidentCarve *r, unsigned char *base, const ResidentPlan &pl, int Tmax, int Cmax, int Kmax) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *q = base ? base + off : nullptr;
        off += pad256(bytes);
        return q;
    };
    const int S = pl.slots;
    SweepWindows &w = r->w;
    w.win_off = reinterpret_cast<int64_t *>(take((size_t)S * 8));
    w.clip_start = reinterpret_cast<double *>(take((size_t)S * 8));
    w.anchor_rel = nullptr;
    w.seg = nullptr;
    w.in_len = reinterpret_cast<int32_t *>(take((size_t)S * 4));
    w.gt = reinterpret_cast<int32_t *>(take((size_t)S * Cmax * 4));
    w.n_cols = reinterpret_cast<int32_t *>(take((size_t)S * 4));
    w.utt_begin = reinterpret_cast<int32_t *>(take((size_t)S * (Kmax + 1) * 4));
    w.n_utts = reinterpret_cast<int32_t *>(take((size_t)S * 4));
    w.text_len = reinterpret_cast<int32_t *>(take((size_t)S * Kmax * 4));
    w.is_last = reinterpret_cast<int32_t *>(take((size_t)S * 4));
    w.term_t = nullptr;
    w.win_status = nullptr;
    w.decision = reinterpret_cast<int32_t *>(take((size_t)S * 4 * 4));
    w.seg_ws = nullptr;
    w.seg_ws_bytes = 0;
    // one bit per cell; a window is padded to whole words in time (<= 32 frames) and whole warps in columns (<= 128)
    r->words_per_slot = ((int64_t)Tmax + 32) * (Cmax + 128) / 32 + 32;
    r->bp = reinterpret_cast<uint32_t *>(take((size_t)S * (size_t)r->words_per_slot * 4));
    r->timing = reinterpret_cast<int32_t *>(take((size_t)S * Kmax * (size_t)Cmax * 4));
    r->cprob = reinterpret_cast<float *>(take((size_t)S * Kmax * (size_t)Tmax * 4));
    r->ticket = reinterpret_cast<int *>(take(256));
    return off;
}

template <int PITCH, int MAXT>
cudaError_t launch_resident(const ResidentParams &P, const ResidentPlan &pl, cudaStream_t st) {
    auto kern = sweep_resident_kernel<PITCH, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    kern<<<pl.slots, pl.threads, pl.smem, st>>>(P);
    return cudaGetLastError();
}
}  // namespace

extern "C" size_t ipfa_sweep_resident_workspace_bytes(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                                                      int Tmax, int Cmax, int Kmax) {
    ResidentPlan pl;
    if (!corpus || !params || Tmax <= 0 || Cmax <= 1 || Kmax <= 0 || corpus->n_files <= 0 ||
        !resident_plan(*corpus, *params, Tmax, Cmax, Kmax, &pl))
        return 0;
    ResidentCarve r;
    return carve_resident(&r, nullptr, pl, Tmax, Cmax, Kmax) + 256;
}

extern "C" int ipfa_sweep_resident_device(const ipfa_sweep_corpus *corpus, const ipfa_sweep_params *params,
                                          const ipfa_sweep_state *state, double *out_seg, int32_t *out_info,
                                          int Tmax, int Cmax, int Kmax, void *workspace, size_t workspace_bytes,
                                          void *stream) {
    if (!corpus || !params || !state || !out_seg || !out_info || !workspace || Tmax <= 0 || Cmax <= 1 || Kmax <= 0)
        return IPFA_ERR_INVALID_ARG;
    const ipfa_sweep_corpus &c = *corpus;
    const ipfa_sweep_state &s = *state;
    if (c.n_files == 0) return IPFA_OK;
    if (c.n_files < 0 || !c.lp || c.V <= 0 || c.blank < 0 || c.blank >= c.V || c.stride_t < c.V || !c.file_frame0 ||
        !c.file_frames || !c.file_samples || !c.row_first || !c.row_type || !c.row_start || !c.row_end ||
        !c.row_utt_end || !c.utt_first || !c.utt_col || !c.utt_chars || !c.file_tok0 || !c.tokens ||
        !s.row || !s.utt || !s.anchor || !s.prop || !s.next_ns || !s.follow_start || !s.exc || !s.status ||
        !s.need || !s.recalc_row || !s.n_windows || !s.cells || !s.frames || !s.clip || params->sample_rate <= 0 ||
        params->frame_shift <= 0 || !(params->samples_to_frames_ratio > 0.0) || !(params->index_duration > 0.0) ||
        params->score_len <= 0)
        return IPFA_ERR_INVALID_ARG;
    ResidentPlan pl;
    if (!resident_plan(c, *params, Tmax, Cmax, Kmax, &pl)) return IPFA_ERR_UNSUPPORTED;
    ResidentCarve r;
    if (workspace_bytes < carve_resident(&r, nullptr, pl, Tmax, Cmax, Kmax) + 256) return IPFA_ERR_WORKSPACE;
    carve_resident(&r, static_cast<unsigned char *>(workspace), pl, Tmax, Cmax, Kmax);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ResidentParams P;
    P.c = c; P.p = *params; P.s = s; P.w = r.w;
    P.Tmax = Tmax; P.Cmax = Cmax; P.Kmax = Kmax; P.pitch = pl.pitch; P.bits_bytes = pl.bits_bytes; P.kc = pl.kc;
    P.bp = r.bp; P.words_per_slot = r.words_per_slot; P.timing = r.timing; P.cprob = r.cprob; P.ticket = r.ticket;
    P.out_seg = out_seg; P.out_info = out_info;
    P.phases = tuning("IPFA_SWEEP_PHASES") != nullptr;
    NvtxRange range("ipfa.sweep_resident (one persistent launch: files x {window, fill, backtrace, decision})");
    cudaError_t e = cudaMemsetAsync(r.ticket, 0, 4, st);
    if (e == cudaSuccess) {
        const int prof_slot = profile_begin(st);
        const bool regs64 = pl.threads > 512 || pl.ctas == 2;  // (two CTAs of 512 threads need the 64-register instances)
        if (pl.fixed_pitch && !regs64) e = launch_resident<32, 512>(P, pl, st);
        else if (pl.fixed_pitch) e = launch_resident<32, 1024>(P, pl, st);
        else if (!regs64) e = launch_resident<0, 512>(P, pl, st);
        else e = launch_resident<0, 1024>(P, pl, st);
        profile_end(prof_slot, st);
        ++g_launch_count;
    }
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
