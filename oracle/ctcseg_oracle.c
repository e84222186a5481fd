/*
 * oracle/ctcseg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32) of the only compiled routine on the
 * reference's alignment path: `cython_fill_table` of the third-party package
 * ctc-segmentation==1.7.1 (/root/reference/requirements.txt:13), reached from
 *   /root/reference/src/iterative_utterance_alignment.py:216
 *   /root/reference/src/word_level_alignment.py:100
 *   /root/reference/src/search_on_speech.py:85
 * through speechbrain==0.5.11 `CTCSegmentation.get_segments`
 * (/root/reference/requirements.txt:87).
 *
 * PARITY UNPINNED: neither package is vendored under /root/reference nor
 * installed in this image and there is no network, and the reference's own
 * tests (src/test/test_ctc_segmentation.py) only print.  This file restates
 * the published algorithm of ctc_segmentation_dyn.pyx (Kuerzinger et al.,
 * "CTC-Segmentation of Large Corpora for German End-to-end Speech
 * Recognition", 2020) as laid out in SURVEY.md section 8(a) row A4; every place
 * that could not be re-checked against the package source is tagged [verify].
 *
 * Semantics (row A4):
 *   table[0,0] = 0; column 0 (ground-truth symbol -1) costs nothing to stay in
 *   when flag preamble_transition_cost_zero (bit 1) is set;
 *   switch(t,c) = table[t-1, c-1] + lpz[t, g_c]          (-1e9 when t==0 or c==0)
 *   stay(t,c)   = table[t-1, c]   + max(lpz[t,blank], lpz[t,g_c])   (-1e9 when t==0)
 *   table[t,c]  = max(switch, stay); per-column FIRST argmax over t
 *   (strict '<' update).  Windowed variant when T > table rows: per-column
 *   sliding offset.  Multi-column ground truth (classic text converter):
 *   switch takes the max over the s-th previous column for each token
 *   spanning s+1 characters.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call
 * into this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PROB_MAX (-1000000000.0f) /* cdef float prob_max = -1000000000 */

static inline float fmax2(float a, float b) { return a > b ? a : b; }

/*
 * table: [W, N] row-major, pre-filled by the caller with -1e10 (config.max_prob)
 * lpz:   [T, V] row-major (row stride V)
 * gt:    [N, G] int64 row-major (-1 = no token)
 * offsets: [N] out
 * flags: bit0 blank_transition_cost_zero, bit1 preamble_transition_cost_zero,
 *        bit4 (16) the largest window step is ceil(mean_offset) instead of int(mean_offset) + 1,
 *        bit5 (32) the per-candidate offsets are SHIFTED instead of cascaded in ascending order
 *        (the two [verify] spots of the windowed mode; defaults = the package as recalled)
 * argmax_out (nullable): [N] first-argmax row of every column
 * returns t (argmax row of the last column); c is N-1.
 */
int oracle_ctcseg_fill(float *table, int W, int N, const float *lpz, int T, int V,
                       const int64_t *gt, int G, int64_t *offsets, int blank, int flags,
                       int32_t *argmax_out) {
    const int blank_cost_zero = flags & 1;
    const int preamble_cost_zero = flags & 2;
    int offset = 0, offset_sum = 0;
    int last_argmax = -1;
    float last_max = 0.0f;
    int *cur_offset = (int *)malloc(sizeof(int) * (size_t)G);
    for (int s = 0; s < G; ++s) cur_offset[s] = -1; /* np.zeros(G) - 1 */ /* [verify] */
    /* mean offset between two window positions */
    const float mean_offset = (float)(T - W) / (float)N;
    /* lower_offset = int(mean_offset); higher_offset = lower_offset + 1   [verify] (flag 16: ceil) */
    const int higher_offset = (flags & 16) ? (int)ceilf(mean_offset) : (int)mean_offset + 1;

    table[0] = 0.0f;
    for (int c = 0; c < N; ++c) {
        if (c > 0) {
            int lim = (T - W) - offset_sum;
            int hi = higher_offset < lim ? higher_offset : lim;
            int lo = last_argmax - W / 2;
            if (lo < 0) lo = 0;
            offset = lo < hi ? lo : hi;
            /* for s in range(G - 1): cur_offset[s + 1] = cur_offset[s] + offset   [verify] -- in
             * ascending order every entry builds on the one just written; flag 32 shifts instead
             * (entry s = the window movement over the last s+1 columns). */
            if (flags & 32) {
                for (int s = G - 2; s >= 0; --s) cur_offset[s + 1] = cur_offset[s] + offset;
            } else {
                for (int s = 0; s < G - 1; ++s) cur_offset[s + 1] = cur_offset[s] + offset;
            }
            cur_offset[0] = offset;
            offset_sum += offset;
        }
        offsets[c] = offset_sum;
        last_argmax = -1;
        last_max = 0.0f;
        for (int t = (c == 0 ? 1 : 0); t < W; ++t) {
            const float *row = lpz + (int64_t)(t + offset_sum) * V;
            float switch_prob = PROB_MAX, max_lpz_prob = PROB_MAX;
            for (int s = 0; s < G; ++s) {
                int64_t g = gt[(int64_t)c * G + s];
                if (g != -1) {
                    float p;
                    /* reading the window-shifted row of column c-(s+1) */
                    int tp = t - 1 + cur_offset[s];
                    if (tp >= W || tp < 0 || c - (s + 1) < 0 || t - 1 < 0) {
                        p = PROB_MAX;
                    } else {
                        p = table[(int64_t)tp * N + (c - (s + 1))] + row[g];
                    }
                    switch_prob = fmax2(switch_prob, p);
                    max_lpz_prob = fmax2(max_lpz_prob, row[g]);
                }
            }
            float stay_prob;
            if (t - 1 < 0) {
                stay_prob = PROB_MAX;
            } else if (c == 0) {
                stay_prob = preamble_cost_zero ? 0.0f
                                               : table[(int64_t)(t - 1) * N] + row[blank]; /* [verify] */
            } else if (blank_cost_zero) {
                stay_prob = table[(int64_t)(t - 1) * N + c]; /* [verify] SURVEY A4 */
            } else {
                stay_prob = table[(int64_t)(t - 1) * N + c] + fmax2(row[blank], max_lpz_prob);
            }
            float v = fmax2(switch_prob, stay_prob);
            table[(int64_t)t * N + c] = v;
            if (last_argmax == -1 || last_max < v) {
                last_max = v;
                last_argmax = t;
            }
        }
        if (argmax_out) argmax_out[c] = last_argmax;
    }
    free(cur_offset);
    return last_argmax;
}
