// anchor_select.cu -- kernel (3): the anchor loop's accept / shrink / revert
// decision, evaluated on the device over the K prefix alignments of a window.
//
// Restates /root/reference/src/iterative_utterance_alignment.py:221-379
// (SURVEY.md section 8(a) row A10).  The reference re-runs the aligner once per
// iteration with the last utterance dropped; ipfa_ctcseg_device already holds
// the segments of every prefix, so the loop below only reads them.  Scores are
// compared the way the reference sees them: rounded to 4 decimals
// (`{score:3.4f}` in CTCSegmentationTask.__str__, parsed back with float()),
// plus 2*threshold for utterances shorter than `short_len` characters (:241).
//
// decision_out[w] = { accepted prefix length (0 = everything discarded),
//                     iterations the reference loop would have run,
//                     IPFA_SEL_* outcome,
//                     index of the utterance whose end is the new anchor
//                       (-1: rewind to the window start, -2: anchor unchanged) }
// anchor_out[w]   = that utterance's end (seconds from the window start,
//                   rounded to 0.01 like the `{end:.2f}` field), NaN if none.
#include "ipfa_common.cuh"

namespace ipfa {
extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;

__device__ __forceinline__ double round_decimals(double x, double scale) {
    if (!(fabs(x) < 1.0e15)) return x;
    return __ddiv_rn(rint(__dmul_rn(x, scale)), scale);
}

__global__ void anchor_select_kernel(const double *__restrict__ seg, const int32_t *__restrict__ n_utts,
                                     const int32_t *__restrict__ text_len,
                                     const int32_t *__restrict__ is_last, int N, int Kmax,
                                     double threshold, int short_len, int32_t *__restrict__ decision_out,
                                     double *__restrict__ anchor_out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int K = max(0, min(n_utts[w], Kmax));
    const double *seg_w = seg + (int64_t)w * Kmax * Kmax * 3;
    const int32_t *tl = text_len + (int64_t)w * Kmax;
    const double penalty = __dmul_rn(2.0, threshold);
    auto score_of = [&](int k, int u) -> double {  // utterance u of the k-utterance prefix
        double s = round_decimals(seg_w[((int64_t)(k - 1) * Kmax + u) * 3 + 2], 1.0e4);
        if (tl[u] < short_len) s = __dadd_rn(s, penalty);
        return s;
    };

    int accepted = 0, n_iter = 0, outcome = IPFA_SEL_DISCARD_ALL;
    int anchor_k = 0, anchor_u = -2;  // tracked new_segment_start: (prefix, utterance)
    bool bad = true, have_prev = false;
    int prev_k = 0;
    int k = K;
    while (k >= 1) {
        ++n_iter;
        double score = 0.0;
        for (int u = 0; u < k; ++u) {  // :221-260
            score = score_of(k, u);
            if (score < threshold) {
                bad = true;
            } else {
                bad = false;
                anchor_k = k;
                anchor_u = u;
            }
        }
        if (is_last[w]) {  // :263
            accepted = k;
            outcome = IPFA_SEL_LAST_SEGMENT;
            break;
        }
        if (bad && !have_prev) {  // :269
            if (k == 1) {         // :272 nothing left to drop
                accepted = 0; outcome = IPFA_SEL_DISCARD_ALL; anchor_u = -1; anchor_k = 0;
                break;
            }
            --k;                  // :281
            continue;
        }
        if (have_prev) {  // :291
            const double prev_score = score_of(prev_k, prev_k - 2);  // previous_segmentation[-2]
            if (score > -1.0 && !(prev_score == score)) {             // :298
                accepted = k; outcome = IPFA_SEL_ACCEPT_CURRENT; anchor_k = k; anchor_u = k - 1;
                break;
            }
            if (prev_score >= score) {                                // :306
                accepted = prev_k; outcome = IPFA_SEL_KEEP_PREVIOUS; anchor_k = prev_k; anchor_u = prev_k - 1;
                break;
            }
            if (k == 1) {                                             // :319
                if (!bad) {
                    accepted = 1; outcome = IPFA_SEL_ACCEPT_CURRENT; anchor_k = 1; anchor_u = 0;
                } else {
                    accepted = 0; outcome = IPFA_SEL_DISCARD_ALL; anchor_k = 0; anchor_u = -1;
                }
                break;
            }
            if (bad) {                                                // :340
                accepted = prev_k; outcome = IPFA_SEL_KEEP_PREVIOUS; anchor_k = prev_k; anchor_u = prev_k - 1;
                break;
            }
            prev_k = k;                                               // :348
            --k;
            continue;
        }
        // first repetition, alignment not bad (:357)
        if (score > -1.0 || k == 1) {  // :360, :367
            accepted = k; outcome = IPFA_SEL_ACCEPT_CURRENT; anchor_k = k; anchor_u = k - 1;
            break;
        }
        have_prev = true;              // :372
        prev_k = k;
        --k;
    }
    decision_out[w * 4 + 0] = accepted;
    decision_out[w * 4 + 1] = n_iter;
    decision_out[w * 4 + 2] = outcome;
    decision_out[w * 4 + 3] = anchor_u;
    double a = __longlong_as_double(0x7ff8000000000000LL);
    if (anchor_u >= 0 && anchor_k >= 1)
        a = round_decimals(seg_w[((int64_t)(anchor_k - 1) * Kmax + anchor_u) * 3 + 1], 100.0);
    anchor_out[w] = a;
}
}  // namespace ipfa

using namespace ipfa;

extern "C" int ipfa_anchor_select_device(const double *seg, const int32_t *n_utts, const int32_t *text_len,
                                         const int32_t *is_last, int N, int Kmax, double threshold,
                                         int short_len, int32_t *decision_out, double *anchor_out,
                                         void *stream) {
    if (N == 0) return IPFA_OK;
    if (!seg || !n_utts || !text_len || !is_last || !decision_out || !anchor_out || N < 0 || Kmax <= 0)
        return IPFA_ERR_INVALID_ARG;
    const int threads = 128;
    anchor_select_kernel<<<(N + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        seg, n_utts, text_len, is_last, N, Kmax, threshold, short_len, decision_out, anchor_out);
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
