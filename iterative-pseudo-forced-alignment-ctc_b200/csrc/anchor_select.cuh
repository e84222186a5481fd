// anchor_select.cuh -- the accept / shrink / revert decision of one window as a device function
// (shared by anchor_select_kernel and the anchor sweep's per-iteration tail kernel).
// Restates /root/reference/src/iterative_utterance_alignment.py:221-379; see anchor_select.cu.
#pragma once
#include "ipfa_common.cuh"

namespace ipfa {

// float(f"{x:.Nf}") with scale = 10^N -- the text round trip of `str(task)` (:218-230).  printf rounds
// the EXACT binary value (ties to even), so the product x * scale may not be rounded before the
// nearest integer is taken: p + e is the exact product (e from the fused multiply-add), the integer is
// chosen from p's fraction and, when that is exactly one half, from the sign of e.  The quotient of two
// exactly representable numbers, correctly rounded, is the double strtod returns for the digits.
__device__ __forceinline__ double round_decimals(double x, double scale) {
    if (!(fabs(x) < 1.0e11)) return x;
    const double p = __dmul_rn(x, scale);
    const double e = __fma_rn(x, scale, -p);
    const double f = floor(p);
    const double h = __dsub_rn(__dsub_rn(p, f), 0.5);
    double r;
    if (h > 0.0 || (h == 0.0 && e > 0.0)) r = f + 1.0;
    else if (h < 0.0 || e < 0.0) r = f;
    else r = (fmod(f, 2.0) == 0.0) ? f : f + 1.0;
    return copysign(fabs(__ddiv_rn(r, scale)), x);  // "-0.0000" stays -0.0
}

struct AnchorDecision {
    int accepted, n_iter, outcome, anchor_u;
    double anchor;  // end of the anchor utterance, seconds from the window start, rounded to 0.01; NaN if none
};

// seg_w: [Kmax][Kmax][3] segments of every prefix of one window; tl: [Kmax] text lengths.
__device__ __forceinline__ AnchorDecision anchor_select_one(const double *seg_w,
                                                            const int32_t *tl, int K, int Kmax,
                                                            bool is_last_window, double threshold,
                                                            int short_len) {
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
            if (!(score >= threshold)) {  // (a NaN score -- a prefix that was not aligned -- is never an anchor)
                bad = true;
            } else {
                bad = false;
                anchor_k = k;
                anchor_u = u;
            }
        }
        if (is_last_window) {  // :263
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
    AnchorDecision d;
    d.accepted = accepted; d.n_iter = n_iter; d.outcome = outcome; d.anchor_u = anchor_u;
    d.anchor = __longlong_as_double(0x7ff8000000000000LL);
    if (anchor_u >= 0 && anchor_k >= 1)
        d.anchor = round_decimals(seg_w[((int64_t)(anchor_k - 1) * Kmax + anchor_u) * 3 + 1], 100.0);
    return d;
}

}  // namespace ipfa
