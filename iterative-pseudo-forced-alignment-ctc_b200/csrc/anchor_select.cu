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
#include "anchor_select.cuh"

namespace ipfa {
extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;

__global__ void anchor_select_kernel(const double *__restrict__ seg, const int32_t *__restrict__ n_utts,
                                     const int32_t *__restrict__ text_len,
                                     const int32_t *__restrict__ is_last, int N, int Kmax,
                                     double threshold, int short_len, int32_t *__restrict__ decision_out,
                                     double *__restrict__ anchor_out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int K = max(0, min(n_utts[w], Kmax));
    const AnchorDecision d = anchor_select_one(seg + (int64_t)w * Kmax * Kmax * 3, text_len + (int64_t)w * Kmax, K,
                                               Kmax, is_last[w] != 0, threshold, short_len);
    decision_out[w * 4 + 0] = d.accepted;
    decision_out[w * 4 + 1] = d.n_iter;
    decision_out[w * 4 + 2] = d.outcome;
    decision_out[w * 4 + 3] = d.anchor_u;
    anchor_out[w] = d.anchor;
}

__global__ void text_round_kernel(const double *__restrict__ x, int64_t n, double scale, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = round_decimals(x[i], scale);
}
}  // namespace ipfa

using namespace ipfa;

extern "C" int ipfa_text_round_device(const double *x, int64_t n, int decimals, double *out, void *stream) {
    if (n == 0) return IPFA_OK;
    if (!x || !out || n < 0 || decimals < 0 || decimals > 9) return IPFA_ERR_INVALID_ARG;
    double scale = 1.0;
    for (int i = 0; i < decimals; ++i) scale *= 10.0;
    const int threads = 256;
    text_round_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, scale, out);
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

extern "C" int ipfa_anchor_select_device(const double *seg, const int32_t *n_utts, const int32_t *text_len,
                                         const int32_t *is_last, int N, int Kmax, double threshold,
                                         int short_len, int32_t *decision_out, double *anchor_out,
                                         void *stream) {
    if (N == 0) return IPFA_OK;
    if (!seg || !n_utts || !text_len || !is_last || !decision_out || !anchor_out || N < 0 || Kmax <= 0)
        return IPFA_ERR_INVALID_ARG;
    NvtxRange range("ipfa.anchor_select");
    const int threads = 128;
    anchor_select_kernel<<<(N + threads - 1) / threads, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        seg, n_utts, text_len, is_last, N, Kmax, threshold, short_len, decision_out, anchor_out);
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
