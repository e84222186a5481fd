/*
 * oracle/ctc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32) of the two standard-CTC comparators that
 * BASELINE.json's north_star names as "the reference CPU path (torch CTC loss /
 * forced align)".  The reference repo itself holds no CTC arithmetic (it
 * delegates to pip packages), so the algorithms restated
 * here are the published ones of the installed libraries:
 *
 *   oracle_ctc_alpha    follows  ATen/native/LossCTC.cpp::ctc_loss_cpu_template
 *                       (torch 2.11; fp32 log-sum-exp alpha over the 2L+1
 *                       blank-interleaved lattice)
 *   oracle_ctc_viterbi  follows  torchaudio/csrc/forced_align/cpu/compute.cpp::
 *                       forced_align_impl (torchaudio 2.11; strict-greater
 *                       comparisons, ties fall to "stay"; SURVEY.md section 8(a) row A8,
 *                       /opt/prime-rl/.venv/lib/python3.12/site-packages/torchaudio/
 *                       functional/_alignment.py:11-73 is its Python surface)
 *
 * Pinning: tests/golden/ctc_golden.npz holds outputs of the installed
 * torch.nn.functional.ctc_loss / torchaudio.functional.forced_align on seeded
 * inputs (generator script tests/golden/make_golden.py); tests/test_oracle_ctc.py
 * checks this file against them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * into this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline int target_prime(const int32_t *tg, int s, int blank) {
    return (s & 1) ? tg[s >> 1] : blank;
}

/* One window.  lp: [T, V] row-major with row stride `stride_t` (elements).
 * Returns the negative log likelihood (+inf when the target is infeasible). */
float oracle_ctc_alpha_one(const float *lp, int64_t stride_t, int T, int V,
                           const int32_t *tg, int L, int blank, float *scratch) {
    (void)V;
    const int S = 2 * L + 1;
    const float NEG = -INFINITY;
    float *prev = scratch, *cur = scratch + S;
    if (T <= 0) return (L == 0) ? 0.0f : INFINITY;
    for (int s = 0; s < S; ++s) prev[s] = NEG;
    prev[0] = lp[blank];
    if (L > 0) prev[1] = lp[tg[0]];
    for (int t = 1; t < T; ++t) {
        const float *row = lp + (int64_t)t * stride_t;
        for (int s = 0; s < S; ++s) {
            const int cp = target_prime(tg, s, blank);
            float la1 = prev[s], la2 = NEG, la3 = NEG, m = la1;
            if (s > 0) {
                la2 = prev[s - 1];
                if (la2 > m) m = la2;
            }
            if (s > 1 && target_prime(tg, s - 2, blank) != cp) {
                la3 = prev[s - 2];
                if (la3 > m) m = la3;
            }
            if (m == NEG) m = 0.0f;
            cur[s] = logf(expf(la1 - m) + expf(la2 - m) + expf(la3 - m)) + m + row[cp];
        }
        float *tmp = prev; prev = cur; cur = tmp;
    }
    if (L == 0) return -prev[0];
    float l1 = prev[S - 1], l2 = prev[S - 2];
    float m = l1 > l2 ? l1 : l2;
    if (m == NEG) m = 0.0f;
    return -(logf(expf(l1 - m) + expf(l2 - m)) + m);
}

/* Batch: lp [N, Tmax, V] via (stride_n, stride_t); targets [N, Lmax] row stride tgt_stride. */
void oracle_ctc_alpha_batch(const float *lp, int64_t stride_n, int64_t stride_t,
                            const int32_t *targets, int64_t tgt_stride,
                            const int32_t *in_len, const int32_t *tgt_len,
                            int N, int V, int blank, float *nll_out) {
#pragma omp parallel
    {
        float *scratch = NULL;
        int cap = 0;
#pragma omp for schedule(dynamic, 1)
        for (int n = 0; n < N; ++n) {
            int L = tgt_len[n], S = 2 * L + 1;
            if (2 * S > cap) {
                free(scratch);
                cap = 2 * S;
                scratch = (float *)malloc(sizeof(float) * (size_t)cap);
            }
            nll_out[n] = oracle_ctc_alpha_one(lp + n * stride_n, stride_t, in_len[n], V,
                                              targets + n * tgt_stride, L, blank, scratch);
        }
        free(scratch);
    }
}

/* CTC Viterbi forced alignment, one window.
 * paths_out[T] receives the token id per frame, scores_out[T] (nullable) the
 * emission log-prob of that token.  Returns 0, or 1 when T < L + repeats
 * (torchaudio raises there). */
int oracle_ctc_viterbi_one(const float *lp, int64_t stride_t, int T, int V,
                           const int32_t *tg, int L, int blank,
                           int32_t *paths_out, float *scores_out) {
    (void)V;
    const int S = 2 * L + 1;
    const float NEG = -INFINITY;
    int R = 0;
    for (int i = 1; i < L; ++i) R += (tg[i] == tg[i - 1]);
    if (T < L + R || T <= 0) return 1;
    float *alphas = (float *)malloc(sizeof(float) * 2 * (size_t)S);
    int8_t *bp = (int8_t *)malloc((size_t)T * (size_t)S);
    memset(bp, -1, (size_t)T * (size_t)S);
    for (int s = 0; s < 2 * S; ++s) alphas[s] = NEG;
    int start = (T - (L + R) > 0) ? 0 : 1;
    int end = (S == 1) ? 1 : 2;
    for (int i = start; i < end; ++i) alphas[i] = lp[target_prime(tg, i, blank)];
    for (int t = 1; t < T; ++t) {
        const float *row = lp + (int64_t)t * stride_t;
        if (T - t <= L + R) {
            if ((start % 2 == 1) && tg[start / 2] != tg[start / 2 + 1]) start += 1;
            start += 1;
        }
        if (t <= L + R) {
            if (end % 2 == 0 && end < 2 * L && tg[end / 2 - 1] != tg[end / 2]) end += 1;
            end += 1;
        }
        int startloop = start;
        float *cur = alphas + (t % 2) * S, *prev = alphas + ((t - 1) % 2) * S;
        for (int j = 0; j < S; ++j) cur[j] = NEG;
        if (start == 0) {
            cur[0] = prev[0] + row[blank];
            bp[(size_t)t * S] = 0;
            startloop += 1;
        }
        for (int i = startloop; i < end; ++i) {
            float x0 = prev[i], x1 = prev[i - 1], x2 = NEG;
            if (i % 2 != 0 && i != 1 && tg[i / 2] != tg[i / 2 - 1]) x2 = prev[i - 2];
            float result;
            if (x2 > x1 && x2 > x0) { result = x2; bp[(size_t)t * S + i] = 2; }
            else if (x1 > x0 && x1 > x2) { result = x1; bp[(size_t)t * S + i] = 1; }
            else { result = x0; bp[(size_t)t * S + i] = 0; }
            cur[i] = result + row[target_prime(tg, i, blank)];
        }
    }
    const float *last = alphas + ((T - 1) % 2) * S;
    int ltr = (S == 1) ? 0 : (last[S - 1] > last[S - 2] ? S - 1 : S - 2);
    for (int t = T - 1; t > -1; --t) {
        int lbl = target_prime(tg, ltr, blank);
        paths_out[t] = lbl;
        if (scores_out) scores_out[t] = lp[(int64_t)t * stride_t + lbl];
        ltr -= bp[(size_t)t * S + ltr];
    }
    free(alphas);
    free(bp);
    return 0;
}

void oracle_ctc_viterbi_batch(const float *lp, int64_t stride_n, int64_t stride_t,
                              const int32_t *targets, int64_t tgt_stride,
                              const int32_t *in_len, const int32_t *tgt_len,
                              int N, int Tmax, int V, int blank,
                              int32_t *paths_out, float *scores_out, int32_t *status_out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int n = 0; n < N; ++n) {
        status_out[n] = oracle_ctc_viterbi_one(
            lp + n * stride_n, stride_t, in_len[n], V, targets + n * tgt_stride, tgt_len[n],
            blank, paths_out + (int64_t)n * Tmax, scores_out ? scores_out + (int64_t)n * Tmax : NULL);
    }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
