// ctcseg_walk.cuh -- backtrace of the 1-bit backpointers of one (window, prefix) and the scoring of its
// utterances (ctc_segmentation() backtrace + determine_utterance_segments),
// as device functions of ONE WARP: shared by ctcseg_backtrace_kernel (ctcseg.cu, one launch over all
// windows and prefixes) and the file-resident anchor sweep (anchor_sweep.cu, where the warps of the CTA
// that filled the table walk its prefixes right away).
#pragma once
#include "ipfa_common.cuh"

namespace ipfa {

// --- numpy-order float64 sums ------------------------------------------------
// np.ndarray.mean on a contiguous float64 vector = pairwise sum (8 accumulators
// for n <= 128, recursive halving above) / n.  char_probs is float64 holding
// fp32 values; mirroring the order keeps the score bit-identical.
__device__ __forceinline__ double np_sum_upto128(const float *__restrict__ a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, (double)a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = (double)a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], (double)a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, (double)a[i]);
    return res;
}
static __device__ __noinline__ double np_pairwise_sum_rec(const float *a, int n) {
    if (n <= 128) return np_sum_upto128(a, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(np_pairwise_sum_rec(a, n2), np_pairwise_sum_rec(a + n2, n - n2));
}
// the scoring windows are score_len (30) frames: the common case stays inline, register-only
__device__ __forceinline__ double np_pairwise_sum(const float *a, int n) {
    return (n <= 128) ? np_sum_upto128(a, n) : np_pairwise_sum_rec(a, n);
}

// determine_utterance_segments for utterance u of one alignment, one warp:
// start/end from the column timings, score = min over the windowed means of char_probs (numpy summation
// order, fp64).
// `scratch` (nullable): 40 doubles of this warp's shared memory.  With it, the sliding windows of full length
// share their partial sums: numpy's pairwise sum of n <= 128 values keeps eight accumulators, r[j] = a[j] +
// a[j+8] + a[j+16] + ... (blocks of eight, in order), combines them as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) and adds
// the n % 8 tail values one by one.  r[j] of the window starting at t is S[t+j] with S[x] = a[x] + a[x+8] + ...
// (the same additions in the same order), so 32 neighbouring windows need 39 values of S instead of 32 x 8
// accumulators: identical bits, about half the loads and additions.
// `slice` (nullable, `slice_cap` floats of this warp's shared memory): the utterance's stretch of char_probs is
// copied there first, so that the window sums read shared memory instead of paying a cache round trip per step.
__device__ __forceinline__ void score_one_segment(const int32_t *ub, const int32_t *timing, const float *cprob_g, int u,
                                                  int T, int Cmax, double dur, int n, bool round_nearest, int lane,
                                                  double *seg, double *scratch, float *slice = nullptr,
                                                  int slice_cap = 0) {
    const float *cprob = cprob_g;
    auto tm = [&](int cc) -> double {
        if (cc < 0 || cc >= Cmax) return 0.0;
        const int f = timing[cc];
        return f < 0 ? 0.0 : __dmul_rn((double)f, dur);
    };
    const int b = ub[u], e = ub[u + 1];
    const double mid_b = __ddiv_rn(__dadd_rn(tm(b), tm(b - 1)), 2.0);
    const double start = fmax(__dadd_rn(tm(b + 1), -0.5), mid_b);
    const double mid_e = __ddiv_rn(__dadd_rn(tm(e), tm(e - 1)), 2.0);
    const double end = fmin(__dadd_rn(tm(e - 1), 0.5), mid_e);
    const double qs = __ddiv_rn(start, dur), qe = __ddiv_rn(end, dur);
    const long long s_t = (long long)(round_nearest ? rint(qs) : floor(qs));
    const long long e_t = (long long)(round_nearest ? rint(qe) : floor(qe));
    double score;
    if (e_t <= s_t) {
        score = -10000000000.0;
    } else if (e_t - s_t <= n) {
        const int lo = (int)max(0LL, min(s_t, (long long)T));
        const int hi = (int)max(0LL, min(e_t, (long long)T));
        const int cnt = hi - lo;
        score = (cnt > 0) ? __ddiv_rn(np_pairwise_sum(cprob + lo, cnt), (double)cnt)
                          : __longlong_as_double(0x7ff8000000000000LL);
    } else {
        // min over t of mean(char_probs[t : t + n]).  x -> x / n is monotone, so for the windows of
        // full length the minimum of the means is the minimum of the sums divided once (the fp64
        // division is a long instruction sequence); windows clipped by the end of the audio keep
        // their own division.
        double best = 0.0, best_sum = __longlong_as_double(0x7ff0000000000000LL);
        int limit = T;  // no window of this utterance reads char_probs at or beyond it
        {
            const int lo_all = (int)max(0LL, min(s_t, (long long)T)), hi_all = (int)max(0LL, min(e_t, (long long)T));
            limit = hi_all;
            if (slice != nullptr && hi_all - lo_all <= slice_cap) {
                __syncwarp();
                for (int i = lo_all + lane; i < hi_all; i += 32) slice[i - lo_all] = cprob_g[i];
                __syncwarp();
                cprob = slice - lo_all;  // (every index below lies in [lo_all, hi_all))
            }
        }
        auto window = [&](long long t) {  // the generic path: one window, numpy's order
            const int lo = (int)max(0LL, min(t, (long long)T));
            const int hi = (int)max(0LL, min(t + n, (long long)T));
            const int cnt = hi - lo;
            if (cnt <= 0) return;
            const double sum = np_pairwise_sum(cprob + lo, cnt);
            if (cnt == n) best_sum = fmin(best_sum, sum);
            else best = fmin(best, __ddiv_rn(sum, (double)cnt));
        };
        if (scratch != nullptr && n >= 8 && n <= 128) {
            const int nb = n / 8;  // blocks of eight
            auto partial = [&](long long x) -> double {  // S[x]; 0 where the full windows never look
                if (x < 0 || x + 8 * (nb - 1) >= limit) return 0.0;
                double r = (double)cprob[x];
                for (int m = 1; m < nb; ++m) r = __dadd_rn(r, (double)cprob[x + 8 * m]);
                return r;
            };
            for (long long base = s_t; base < e_t - n; base += 32) {
                __syncwarp();
                scratch[lane] = partial(base + lane);
                if (lane < 7) scratch[32 + lane] = partial(base + 32 + lane);
                __syncwarp();
                const long long t = base + lane;
                if (t < e_t - n) {
                    if (t >= 0 && t + n <= T) {
                        const double *r = scratch + lane;
                        double sum = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                        for (int i = 8 * nb; i < n; ++i) sum = __dadd_rn(sum, (double)cprob[t + i]);
                        best_sum = fmin(best_sum, sum);
                    } else {
                        window(t);
                    }
                }
            }
        } else {
            for (long long t = s_t + lane; t < e_t - n; t += 32) window(t);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            best = fmin(best, __shfl_xor_sync(0xffffffffu, best, off));
            best_sum = fmin(best_sum, __shfl_xor_sync(0xffffffffu, best_sum, off));
        }
        if (best_sum < __longlong_as_double(0x7ff0000000000000LL))
            best = fmin(best, __ddiv_rn(best_sum, (double)n));
        score = best;
    }
    if (lane == 0) {
        seg[u * 3] = start;
        seg[u * 3 + 1] = end;
        seg[u * 3 + 2] = score;
    }
}

// ... for utterances 0..n_utt-1 of one alignment.
__device__ __forceinline__ void score_segments(const int32_t *ub, const int32_t *timing, const float *cprob,
                                               int n_utt, int T, int Cmax, double dur, int n, bool round_nearest,
                                               int lane, double *seg, double *scratch = nullptr) {
    for (int u = 0; u < n_utt; ++u)
        score_one_segment(ub, timing, cprob, u, T, Cmax, dur, n, round_nearest, lane, seg, scratch);
}

// Bits 0, KC, 2KC, ... of x packed into the low 32 / KC bits.
template <int KC>
__device__ __forceinline__ uint32_t compress_stride(uint32_t x) {
    if constexpr (KC == 1) {
        return x;
    } else if constexpr (KC == 2) {
        x &= 0x55555555u;
        x = (x | (x >> 1)) & 0x33333333u;
        x = (x | (x >> 2)) & 0x0f0f0f0fu;
        x = (x | (x >> 4)) & 0x00ff00ffu;
        return (x | (x >> 8)) & 0x0000ffffu;
    } else if constexpr (KC == 4) {
        x &= 0x11111111u;
        x = (x | (x >> 3)) & 0x03030303u;
        x = (x | (x >> 6)) & 0x000f000fu;
        return (x | (x >> 12)) & 0x000000ffu;
    } else {
        x &= 0x01010101u;
        x = (x | (x >> 7)) & 0x00030003u;
        return (x | (x >> 14)) & 0x0000000fu;
    }
}

// What the walk of one (window, prefix) needs.  `gt_s`: the window's ground-truth column with the symbols
// outside [0, V) already replaced by the blank (shared memory); `raw`: this warp's staging area,
// seg_walk_smem_words<KC>() 32-bit words of shared memory.
struct SegWalkArgs {
    const float *lp;         // first frame of the window
    int64_t stride_t;
    int T, Cmax, blank, NT, score_len;
    bool round_nearest;
    double index_duration;
    const uint32_t *bp_w;    // the window's backpointer words [frame / SPW][NT]
    const int32_t *ub;       // [K + 1] utterance-boundary columns
    const int32_t *gt_s;
    int32_t *timing;         // [Cmax], initialised to -1
    float *cprob;            // [>= T], initialised to 0
    int32_t *state;          // nullable
    double *seg;             // [kslot + 1][3]
    uint32_t *raw;
    bool skip_scoring = false;  // the caller scores the utterances itself (file-resident sweep: a warp per (prefix, utterance))
};

template <int KC>
__host__ __device__ constexpr int seg_walk_smem_words() {
    constexpr int NCW2 = (31 / KC + 2) + 32 / KC + 1;
    return NCW2 * KC * 2;
}

// 40 doubles inside a warp's staging words (8-byte aligned whatever the caller's layout), for score_one_segment
template <int KC>
__device__ __forceinline__ double *seg_scratch(uint32_t *raw) {
    static_assert(seg_walk_smem_words<KC>() >= 82, "staging area holds the partial sums");
    return reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(raw) + 7) & ~(uintptr_t)7);
}

// COHERENT: the backpointers were written by THIS kernel (file-resident sweep): plain loads instead of
// the read-only data path.
template <int KC, bool COHERENT>
__device__ __forceinline__ void seg_walk_prefix(const SegWalkArgs &a, int kslot, int t_term, int c_end,
                                                int lane) {
    constexpr int SPW = 32 / KC;        // frames per word
    constexpr int NW = KC;              // word-rows per 32-frame block
    constexpr int NCW = 31 / KC + 2;    // thread-columns reachable inside one block
    // staged one block ahead, before the walk of the current block is known: the start column
    // can drop by up to 32 more columns (one switch per frame) -> 32 / KC + 1 more thread-columns
    constexpr int NCW2 = NCW + 32 / KC + 1;
    constexpr int NWORDS2 = NCW2 * NW;
    constexpr int LOG2KC = (KC == 1) ? 0 : (KC == 2) ? 1 : (KC == 4) ? 2 : 3;
    const int T = a.T;
    const float *lp = a.lp;
    const int32_t *gt_s = a.gt_s;
    int32_t *timing = a.timing;
    float *cprob = a.cprob;
    int32_t *state = a.state;
    const uint32_t *bp_w = a.bp_w;
    const int NT = a.NT;
    uint32_t *raw = a.raw;
    // (one column per thread: a staged word already IS a column's 32 frames -- col32 is raw, nothing to transpose)
    uint32_t *col32 = (KC == 1) ? raw : raw + NWORDS2;
    // ---- walk the 1-bit backpointers from (t_term, c_end) to (0, 0) ----------
    // Per 32-frame block the words the walk can reach sit in shared memory.  The walk does not
    // step frame by frame: inside a word the bits of one column are the frames at which that
    // column was ENTERED, so "the next switch at or below frame t" is one mask + find-leading-one;
    // the serial chain is one LDS per column change (about C steps per window instead of T).
    // The words of block b-1 are requested before block b is walked (a superset wide enough
    // for wherever the walk ends up), so no global-memory latency sits on the chain either.
    // Column 0 / thread-column -1 is staged as zeros: its bits read 0 (stay).
    constexpr int NQ = (NWORDS2 + 31) / 32;
    uint32_t v[NQ];
    auto request = [&](int blk, int i_base) {  // words of block blk, thread-columns i_base - [0, NCW2)
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = lane + 32 * u;
            const int row = q / NCW2, crel = q - row * NCW2;
            const int col = i_base - crel;
            const int wrow = blk * NW + row;
            v[u] = 0;
            if (q < NWORDS2 && col >= 0 && wrow * SPW < T) {
                const uint32_t *src = bp_w + (int64_t)wrow * NT + col;
                v[u] = COHERENT ? *src : __ldg(src);
            }
        }
    };
    // per-frame outputs of a block are stored one block late: their emission gathers are issued
    // before the next block's walk and consumed after it
    int pend_t = -1, pend_c = 0, pend_sw = 0;
    float pend_eb = 0.0f, pend_ec = 0.0f;
    int lc = c_end - 1;  // lattice column (table column - 1); -1 is table column 0
    int i_base = lc >> LOG2KC;  // arithmetic shift: -1 -> thread-column -1
    request(t_term >> 5, i_base);
    for (int blk = t_term >> 5; blk >= 0; --blk) {
        const int t_hi = min(t_term, blk * 32 + 31);
        const int t_lo = blk * 32;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = lane + 32 * u;
            if (q < NWORDS2) raw[q] = v[u];
        }
        __syncwarp();
        const int i_cur = i_base;
        i_base = lc >> LOG2KC;
        if (blk > 0) request(blk - 1, i_base);  // in flight during this block's walk
        if (pend_t >= 0) {
            const float *row = lp + (int64_t)pend_t * a.stride_t;
            pend_eb = row[a.blank];
            pend_ec = row[gt_s[pend_c]];
        }
        // per-column words of the whole block: col32[d] bit f = "column top - d was entered at frame
        // t_lo + f" (the KC-strided bits of the NW word-rows compressed and concatenated)
        const int top = i_cur * KC + (KC - 1);  // lattice column of col32[0]
        if constexpr (KC != 1) {
            for (int d = lane; d < NCW2 * KC; d += 32) {
                const int crel = d >> LOG2KC, k = (KC - 1) - (d & (KC - 1));
                uint32_t bits = 0;
#pragma unroll
                for (int row = 0; row < NW; ++row)
                    bits |= compress_stride<KC>(raw[row * NCW2 + crel] >> k) << (row * SPW);
                col32[d] = bits;
            }
            __syncwarp();
        }
        // switch frames of this block as a 32-bit mask (bit = frame - t_lo): the serial chain is
        // one LDS + mask + find-leading-one per column change
        uint32_t S = 0;
        const int c_hi = lc + 1;  // table column at frame t_hi
        {
            // frame 0 of the table is never visited: the reference's loop ends at (0, 0)
            const uint32_t keep = (blk == 0) ? ~1u : 0xffffffffu;
            int d = top - lc;
            uint32_t below = ((2u << (t_hi - t_lo)) - 1u) & keep;  // frames <= the current one
            while (true) {
                const uint32_t m = col32[d] & below;
                if (m == 0) break;                 // stays down to the first frame of the block
                const int f = 31 - __clz(m);       // the column was entered at this frame
                S |= 1u << f;
                ++d;
                below &= (1u << f) - 1u;
                if (below == 0) break;
            }
            lc = top - d;
        }
        // the previous block's outputs (their gathers had this block's walk to arrive)
        if (pend_t >= 0) {
            const float p = (pend_c == 0) ? pend_eb : (pend_sw ? pend_ec : fmaxf(pend_eb, pend_ec));
            cprob[pend_t] = p;
            if (pend_sw && pend_c > 0) timing[pend_c] = pend_t;
            if (state) state[pend_t] = pend_sw ? pend_c : -1;
        }
        // lane = frame: column at frame t = c_hi - (switches at later frames of the block)
        const int t = t_lo + lane;
        pend_t = -1;
        if (t <= t_hi && t >= 1) {
            pend_t = t;
            pend_c = c_hi - __popc((S >> lane) >> 1);
            pend_sw = (S >> lane) & 1u;
        }
    }
    if (pend_t >= 0) {
        const float *row = lp + (int64_t)pend_t * a.stride_t;
        pend_eb = row[a.blank];
        pend_ec = row[gt_s[pend_c]];
        const float p = (pend_c == 0) ? pend_eb : (pend_sw ? pend_ec : fmaxf(pend_eb, pend_ec));
        cprob[pend_t] = p;
        if (pend_sw && pend_c > 0) timing[pend_c] = pend_t;
        if (state) state[pend_t] = pend_sw ? pend_c : -1;
    }
    __syncwarp();
    __threadfence_block();

    // ---- determine_utterance_segments ---------------------------------------
    if (a.skip_scoring) return;
    // (the staging words of the walk are free now: scratch for the shared partial sums)
    score_segments(a.ub, timing, cprob, kslot + 1, T, a.Cmax, a.index_duration, a.score_len, a.round_nearest, lane,
                   a.seg, seg_scratch<KC>(a.raw));
}

}  // namespace ipfa
