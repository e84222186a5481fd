// ctc_viterbi.cu -- kernel (2a): CTC Viterbi forced alignment on the 2L+1
// lattice with backpointer bit planes in HBM (1.5 bits per state and frame) + warp-parallel
// backtrace with per-frame scores and per-token spans / confidences.
//
// Replaces torchaudio.functional.forced_align (torchaudio/csrc/forced_align/cpu/
// compute.cpp::forced_align_impl; Python surface functional/_alignment.py:11-73)
// and merge_tokens (_alignment.py:94-127); SURVEY.md section 8(a) row A8.  Tie rule
// mirrored exactly: strict-greater comparisons, ties fall to "stay", final state
// prefers S-2 on a tie.  torchaudio's start/end band is not restated: states it
// masks are exactly the ones that cannot lie on a complete path, so the natural
// -inf propagation yields identical backpointers along every complete path.
//
// Fill: same pair-per-thread register layout as ctc_alpha.cu (max-plus instead
// of log-sum-exp; fp32 single-rounding adds so scores are bit-identical to the
// CPU).  Backpointers are bit PLANES: per state and 32-frame block one word whose bit f says
// "entered from a lower state at frame f" (blank: from the label below; label: from its blank,
// and a second plane for the skip transition) -- 3 words per state pair and block, stored
// lane-contiguous, written once, coalesced.
// Backtrace: one warp per window hops from move to move (find-leading-one on the planes of the
// block, staged in shared memory one block ahead) instead of stepping frame by frame.
#include "emission_pipe.cuh"
#include "lattice_shapes.cuh"

namespace ipfa {

extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;
bool use_dense_panel(int V, int Lmax);

struct ViterbiParams {
    const float *lp;
    int64_t stride_n, stride_t, stride_v;
    const int32_t *targets;
    int64_t tgt_stride;
    const int32_t *in_len;
    const int32_t *tgt_len;
    int N, Tmax, V, blank;
    int pitch, tc, u_cap, l_cap;
    size_t group_smem;
    uint32_t *bp;            // [N][words_per_window]
    int64_t words_per_window;
    int32_t *final_state;    // [N]
    float *total_out;        // [N] nullable
    int32_t *status_out;     // [N]
    // length buckets: this launch walks windows order[0 .. *count) (null: all N windows in order)
    const int32_t *order;
    const int32_t *count;
};

// PITCH > 0: compile-time panel pitch (dense rows of <= 32 symbols): the unrolled frames of a
// backpointer word address the panel with immediate offsets.
template <int P, int WARPS, bool DENSE, int PITCH>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32 * WARPS)
ctc_viterbi_fill_kernel(const ViterbiParams prm) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    constexpr int NT = 32 * WARPS;
    constexpr float NEG = -__builtin_huge_valf();
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int group = (WARPS == 1) ? (threadIdx.x >> 5) : 0;
    const int tid = (WARPS == 1) ? (threadIdx.x & 31) : threadIdx.x;
    int w = blockIdx.x * GROUPS + group;
    if (w >= (prm.count ? *prm.count : prm.N)) return;
    if (prm.order) w = prm.order[w];

    const int pitch = PITCH ? PITCH : prm.pitch;
    unsigned char *gsm = smem_raw + (size_t)group * prm.group_smem;
    float *ring = reinterpret_cast<float *>(gsm);
    float *xline = ring + (size_t)kStages * prm.tc * pitch;      // [2][NT + 1] neighbour exchange
    float *fin = xline + 2 * (NT + 1);                            // [2]
    int *cnt = reinterpret_cast<int *>(fin + 2);                // [2] repeats, bad labels
    int *cols = cnt + 2;                                        // [u_cap]

    const int T = prm.in_len[w];
    const int L = max(0, min(prm.tgt_len[w], prm.l_cap));
    const int32_t *tg = prm.targets + (int64_t)w * prm.tgt_stride;
    const int blank = prm.blank;

    if (tid < 2) {
        cnt[tid] = 0;
        if constexpr (WARPS > 1) xline[tid * (NT + 1)] = NEG;  // left neighbour of thread 0
    }
    group_sync<WARPS>();

    int col[P];
    bool skip[P];
    int n_rep = 0, n_bad = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;
        // states past the end of the target run on garbage: nothing real reads them
        const bool lab_ok = j < L;
        int lab = lab_ok ? tg[j] : blank;
        if (lab_ok && (lab < 0 || lab >= prm.V || lab == blank)) ++n_bad;
        if (lab < 0 || lab >= prm.V) lab = blank;
        const int prev = (j >= 1 && lab_ok) ? tg[j - 1] : -1;
        skip[p] = lab_ok && j >= 1 && prev != lab;
        if (lab_ok && j >= 1 && prev == lab) ++n_rep;
        col[p] = DENSE ? lab : (lab_ok ? j + 1 : 0);
    }
    if (n_rep) atomicAdd(&cnt[0], n_rep);
    if (n_bad) atomicAdd(&cnt[1], n_bad);
    int colb = DENSE ? blank : 0;
    int U = L + 1;  // panel columns
    if constexpr (!DENSE) {
        for (int j = tid; j <= L; j += NT) {
            int c = (j == 0) ? blank : tg[j - 1];
            if (c < 0 || c >= prm.V) c = blank;
            cols[j] = c;
        }
        group_sync<WARPS>();
        // ascending, unique column list; the emission ring (idle until the prologue) is scratch
        int *scratch = reinterpret_cast<int *>(ring);
        U = sort_unique_columns<WARPS>(cols, L + 1, scratch, tid);
        const int *pos = scratch + 2 * (L + 1);
#pragma unroll
        for (int p = 0; p < P; ++p) col[p] = pos[col[p]];
        colb = pos[0];
    }
    group_sync<WARPS>();
    const int R = cnt[0];
    int status = cnt[1] ? IPFA_WIN_BAD_LABEL : IPFA_WIN_OK;
    if (T <= 0 || T < L + R) status |= IPFA_WIN_INFEASIBLE;
    if (status & IPFA_WIN_INFEASIBLE) {
        if (tid == 0) {
            prm.status_out[w] = status;
            prm.final_state[w] = -1;
            if (prm.total_out) prm.total_out[w] = NEG;
        }
        return;
    }

    EmissionPipe<WARPS, DENSE> pipe;
    pipe.init(ring, cols, prm.lp + (int64_t)w * prm.stride_n, prm.stride_t, T, U, prm.V, pitch, prm.tc,
              reinterpret_cast<uint64_t *>(gsm + prm.group_smem - 32), tid, 0, false, prm.stride_v);
    pipe.prologue(tid);

    float ab[P], al[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { ab[p] = NEG; al[p] = NEG; }
    // Backpointers as bit planes, one 32-bit word per state and 32-frame block: bit f of
    //   mvb[p]  "the blank state of pair p was ENTERED from the label below (s-1) at frame f"
    //   mvl[p]  "the label state of pair p was entered from its blank (s-1) at frame f"
    //   by2[p]  "the label state of pair p was entered by the skip transition (s-2) at frame f"
    // so the backtrace can hop from move to move with mask + find-leading-one instead of
    // stepping frames, and a decision costs one predicated OR.  3P words per thread and block
    // (1.5 bits per state and frame), stored plane by plane: every warp store is 128 contiguous bytes.
    // (ptxas 12.9 segfaults on some of the 8-pairs-per-thread instances unless their label plane also
    // carries the skip moves; the backtrace ORs the two planes, so both encodings read the same.)
    constexpr bool kMergedPlanes = (P == 8);
    uint32_t mvb[P], mvl[P], by2[P];
#pragma unroll
    for (int p = 0; p < P; ++p) { mvb[p] = 0; mvl[p] = 0; by2[p] = 0; }
    uint32_t bit = 1;  // 1 << (t & 31)
    uint32_t *bp_ptr = prm.bp + (int64_t)w * prm.words_per_window + tid;
    auto flush = [&]() {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            bp_ptr[(3 * p + 0) * NT] = mvb[p];
            bp_ptr[(3 * p + 1) * NT] = mvl[p];
            bp_ptr[(3 * p + 2) * NT] = by2[p];
            mvb[p] = 0; mvl[p] = 0; by2[p] = 0;
        }
        bp_ptr += 3 * P * NT;
    };
    auto advance = [&](const int frames) {  // `frames` frames done; a group never straddles a block
        bit <<= frames;
        if (bit == 0) { flush(); bit = 1; }
    };

    // One frame, `off` floats past the column cursors, `sh` frames past the one `bit` stands for
    // (a compile-time constant after inlining, so the unrolled group shares one block test).
    const float *pb = nullptr;
    const float *pl[P];
    auto frame = [&](const int off, const float *rd, float *wr, const int sh) {
        const uint32_t fbit = bit << sh;
        const float eb = pb[off];
        float el[P];
#pragma unroll
        for (int p = 0; p < P; ++p) el[p] = pl[p][off];
        float prev;
        if constexpr (WARPS > 1) {
            prev = rd[tid];
        } else {
            prev = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
            if (tid == 0) prev = NEG;
        }
#pragma unroll
        for (int p = P - 1; p >= 0; --p) {
            const float lm1 = (p == 0) ? prev : al[p - 1];
            // label state: x0 stay, x1 from blank, x2 skip.  torchaudio's rule
            //   if (x2 > x1 && x2 > x0) 2; else if (x1 > x0 && x1 > x2) 1; else 0
            // has mutually exclusive branches (x2 > x1 vs x1 > x2), so both predicates are
            // evaluated independently.
            const float x0 = al[p], x1 = ab[p], x2 = skip[p] ? lm1 : NEG;
            const bool take2 = (x2 > x1) && (x2 > x0);
            const bool take1 = (x1 > x0) && (x1 > x2);
            const float res = take2 ? x2 : (take1 ? x1 : x0);
            // blank state: x0 stay, x1 from previous label
            const bool takeb = lm1 > ab[p];
            const float nb = (takeb ? lm1 : ab[p]) + eb;
            al[p] = res + el[p];
            ab[p] = nb;
                    mvb[p] |= takeb ? fbit : 0u;
            mvl[p] |= (take1 || (kMergedPlanes && take2)) ? fbit : 0u;
            by2[p] |= take2 ? fbit : 0u;
        }
        if constexpr (WARPS > 1) {
            wr[tid + 1] = al[P - 1];
            __syncthreads();
        }
    };
    auto bump = [&](const int frames) {
        pb += frames * pitch;
#pragma unroll
        for (int p = 0; p < P; ++p) pl[p] += frames * pitch;
    };

    float *line0 = xline, *line1 = xline + NT + 1;
    for (int chunk = 0; chunk < pipe.nchunks; ++chunk) {
        const float *panel = pipe.acquire(chunk, tid);
        const int t0 = chunk * pipe.tc;
        const int rows = min(pipe.tc, T - t0);
        int r = 0;
        if (chunk == 0) {
            if (tid == 0) {
                // torchaudio: start = (T - (L+R) > 0) ? 0 : 1 -- the leading blank is off every
                // complete path when T == L+R; keeping it changes nothing.
                ab[0] = panel[colb];
                if (L > 0) al[0] = panel[col[0]];
            }
            advance(1);  // frame 0 has no incoming transition
            if constexpr (WARPS > 1) {
                line0[tid + 1] = al[P - 1];
                __syncthreads();
            }
            r = 1;
        }
        // frame t reads line[(t-1)&1], writes line[t&1]; tc is even so r has t's parity
        pb = panel + r * pitch + colb;
#pragma unroll
        for (int p = 0; p < P; ++p) pl[p] = panel + r * pitch + col[p];
        // single frames up to a multiple of 4 (t0 is a multiple of tc, tc of 4 whenever tc > 2): from
        // there groups of 4 frames sit inside one 32-frame block and share one advance()
        // (the 8-pairs-per-thread instances keep single frames: ptxas 12.9 segfaults on their
        // unrolled form)
        const int align4 = (P == 8 || (pipe.tc & 3)) ? rows : min(rows, (r + 3) & ~3);
        for (; r < align4; ++r) {
            if (r & 1) frame(0, line0, line1, 0); else frame(0, line1, line0, 0);
            advance(1);
            bump(1);
        }
        if constexpr (P < 8) {
            for (; r + 3 < rows; r += 4) {
                frame(0, line1, line0, 0);
                frame(pitch, line0, line1, 1);
                frame(2 * pitch, line1, line0, 2);
                frame(3 * pitch, line0, line1, 3);
                advance(4);
                bump(4);
            }
        }
        for (; r < rows; ++r) {
            if (r & 1) frame(0, line0, line1, 0); else frame(0, line1, line0, 0);
            advance(1);
            bump(1);
        }
    }
    if (bit != 1) flush();

#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;
        if (j == L) fin[0] = ab[p];
        if (j == L - 1) fin[1] = al[p];
    }
    group_sync<WARPS>();
    if (tid == 0) {
        const int S = 2 * L + 1;
        int ltr = 0;
        float best = fin[0];
        if (S > 1) {
            if (fin[0] > fin[1]) { ltr = S - 1; best = fin[0]; }
            else { ltr = S - 2; best = fin[1]; }
        }
        prm.final_state[w] = ltr;
        prm.status_out[w] = status;
        if (prm.total_out) prm.total_out[w] = best;
    }
}

struct BacktraceParams {
    const float *lp;
    int64_t stride_n, stride_t, stride_v;
    const int32_t *targets;
    int64_t tgt_stride;
    const int32_t *in_len;
    const int32_t *tgt_len;
    int N, Tmax, Lmax, V, blank, NT;
    const uint32_t *bp;
    int64_t words_per_window;
    const int32_t *final_state;
    const int32_t *order;    // length bucket of this launch (see ViterbiParams)
    const int32_t *count;
    int32_t *paths_out;
    float *scores_out;
    int32_t *tok_start, *tok_end;
    float *tok_score;
};

// One warp per window.  The backpointers are bit planes (one word per state and 32-frame block,
// bit f = "entered from a lower state at frame f"), so the walk hops from move to move:
// "the next move at or below frame t" is word & below-mask + find-leading-one -- about 2L+1 steps
// per window instead of T.  Per block the planes of the thread-columns the walk can reach are
// staged in shared memory; the words of block b-1 are requested before block b is walked (a
// superset wide enough for wherever the walk ends up), and the per-frame score gather of a block
// is issued before the next block's walk and stored after it, so no memory latency sits on the
// chain.  Lane = frame for the outputs: state at frame t = state at the block's top minus the
// moves above t (two popcounts).  Token spans come lane-parallel from neighbouring frames' states.
template <int P>
__global__ void __launch_bounds__(128) ctc_viterbi_backtrace_kernel(const BacktraceParams prm) {
    constexpr int SPT = 2 * P;                 // states per thread-column
    constexpr int NPL = 3 * P;                 // planes per thread-column
    constexpr int NTH = 64 / SPT + 2;          // thread-columns reachable inside one block (<= 64 states)
    constexpr int NTH2 = NTH + 64 / SPT + 1;   // ... plus wherever the block below may start
    constexpr int NWORDS = NTH2 * NPL;
    constexpr int LOG2SPT = (P == 1) ? 1 : (P == 2) ? 2 : (P == 4) ? 3 : 4;
    extern __shared__ __align__(16) unsigned char bt_smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    int w = blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= (prm.count ? *prm.count : prm.N)) return;
    if (prm.order) w = prm.order[w];
    uint32_t *raw = reinterpret_cast<uint32_t *>(bt_smem) + (size_t)wib * (NWORDS + prm.Lmax);
    int32_t *tg_s = reinterpret_cast<int32_t *>(raw + NWORDS);
    const int T = min(prm.in_len[w], prm.Tmax);
    const int L = max(0, min(prm.tgt_len[w], prm.Lmax));
    const int32_t *tg = prm.targets + (int64_t)w * prm.tgt_stride;
    int32_t *paths = prm.paths_out + (int64_t)w * prm.Tmax;
    float *scores = prm.scores_out ? prm.scores_out + (int64_t)w * prm.Tmax : nullptr;
    const float *lp = prm.lp + (int64_t)w * prm.stride_n;
    int s = prm.final_state[w];

    // tail beyond in_len, or the whole row when the window was rejected
    const int t_valid = (s < 0) ? 0 : max(T, 0);
    for (int t = t_valid + lane; t < prm.Tmax; t += 32) {
        paths[t] = -1;
        if (scores) scores[t] = 0.0f;
    }
    const bool want_tok = prm.tok_start != nullptr;
    int32_t *tok_s = want_tok ? prm.tok_start + (int64_t)w * prm.Lmax : nullptr;
    int32_t *tok_e = want_tok ? prm.tok_end + (int64_t)w * prm.Lmax : nullptr;
    float *tok_p = (want_tok && prm.tok_score) ? prm.tok_score + (int64_t)w * prm.Lmax : nullptr;
    if (want_tok) {
        for (int l = lane; l < prm.Lmax; l += 32) {
            tok_s[l] = -1;
            tok_e[l] = -1;
            if (tok_p) tok_p[l] = 0.0f;
        }
    }
    for (int l = lane; l < L; l += 32) tg_s[l] = tg[l];
    __syncwarp();
    if (s < 0 || T <= 0) return;

    const uint32_t *bp_w = prm.bp + (int64_t)w * prm.words_per_window;
    const int NT = prm.NT;
    constexpr int NQ = (NWORDS + 31) / 32;
    uint32_t v[NQ];
    auto request = [&](int blk, int c_base) {  // planes of thread-columns c_base - [0, NTH2) of block blk
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = lane + 32 * u;
            const int plane = q / NTH2, crel = q - plane * NTH2;
            const int col = c_base - crel;
            v[u] = 0;
            if (q < NWORDS && col >= 0) v[u] = __ldg(bp_w + ((int64_t)blk * NPL + plane) * NT + col);
        }
    };
    int s_above = -1;  // state of frame t_hi + 1 (none above the last frame)
    // outputs of the previous (higher) block, stored one iteration late
    int pend_t = -1, pend_lab = 0;
    float pend_sc = 0.0f;
    int c_base = s >> LOG2SPT;
    request((T - 1) >> 5, c_base);

    for (int blk = (T - 1) >> 5; blk >= 0; --blk) {
        const int t_hi = min(T - 1, blk * 32 + 31);
        const int t_lo = blk * 32;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NQ; ++u) {
            const int q = lane + 32 * u;
            if (q < NWORDS) raw[q] = v[u];
        }
        __syncwarp();
        const int c_cur = c_base;
        c_base = s >> LOG2SPT;
        if (blk > 0) request(blk - 1, c_base);  // in flight during this block's walk
        if (pend_t >= 0 && scores) pend_sc = lp[(int64_t)pend_t * prm.stride_t + (int64_t)pend_lab * prm.stride_v];
        // moves of this block as two 32-bit masks (bit = frame - t_lo)
        uint32_t S_mv = 0, S_b2 = 0;
        const int s_top = s;
        {
            // frame 0 has no incoming transition
            uint32_t below = ((2u << (t_hi - t_lo)) - 1u) & ((blk == 0) ? ~1u : 0xffffffffu);
            while (below) {
                const int col = s >> LOG2SPT, k = s & (SPT - 1);
                const uint32_t *wp = raw + (3 * (k >> 1) + (k & 1)) * NTH2 + (c_cur - col);
                const uint32_t b2 = (k & 1) ? wp[NTH2] : 0u;  // label: the by-two plane follows
                const uint32_t m = (wp[0] | b2) & below;
                if (m == 0) break;                  // stays down to the first frame of the block
                const int f = 31 - __clz(m);        // the state was entered at this frame
                const uint32_t two = (b2 >> f) & 1u;
                S_mv |= 1u << f;
                S_b2 |= two << f;
                s = max(s - 1 - (int)two, 0);
                below &= (1u << f) - 1u;
            }
        }
        if (pend_t >= 0) {
            paths[pend_t] = pend_lab;
            if (scores) scores[pend_t] = pend_sc;
        }
        const int s_below = s;  // state of frame t_lo - 1 (blk > 0)
        const int t = t_lo + lane;
        const int my_state = s_top - __popc((S_mv >> lane) >> 1) - __popc((S_b2 >> lane) >> 1);
        // neighbours' states for the token spans
        int st_up = __shfl_down_sync(0xffffffffu, my_state, 1);
        int st_dn = __shfl_up_sync(0xffffffffu, my_state, 1);
        if (t == t_hi) st_up = s_above;
        if (lane == 0) st_dn = (t_lo == 0) ? -1 : s_below;
        pend_t = -1;
        if (t <= t_hi) {
            const int lab = (my_state & 1) ? tg_s[my_state >> 1] : prm.blank;
            pend_t = t;
            pend_lab = lab;
            if (want_tok && (my_state & 1)) {
                if (st_dn != my_state) tok_s[my_state >> 1] = t;
                if (st_up != my_state) tok_e[my_state >> 1] = t + 1;
            }
        }
        s_above = __shfl_sync(0xffffffffu, my_state, 0);  // state of frame t_lo
    }
    if (pend_t >= 0) {
        paths[pend_t] = pend_lab;
        if (scores) scores[pend_t] = lp[(int64_t)pend_t * prm.stride_t + (int64_t)pend_lab * prm.stride_v];
    }
    if (want_tok && tok_p) {
        __syncwarp();
        for (int l = lane; l < L; l += 32) {
            const int a = tok_s[l], b = tok_e[l];
            if (a >= 0 && b > a) {
                const int lab = tg_s[l];
                float acc = 0.0f;
                for (int t = a; t < b; ++t) acc += lp[(int64_t)t * prm.stride_t + (int64_t)lab * prm.stride_v];
                tok_p[l] = acc / (float)(b - a);
            }
        }
    }
}

// ---------------------------------------------------------------------------
static int64_t viterbi_words_per_window(int Tmax, LatticeShape s) {
    // per 32-frame block and thread: 2 "moved" planes + 1 "by two" plane per state pair
    return (int64_t)((Tmax + 31) / 32) * 3 * s.PER * 32 * s.WARPS;
}

template <int P, int WARPS, bool DENSE, int PITCH>
static int launch_fill_p(ViterbiParams prm, int Lmax, cudaStream_t stream) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    const int U = PITCH ? PITCH : (DENSE ? prm.V : (Lmax + 1));
    const size_t budget = (WARPS == 1) ? (16 * 1024) : (160 * 1024);
    PipeGeometry g = pipe_geometry(U, budget, !DENSE && prm.stride_v > prm.stride_t);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = DENSE ? 0 : ((Lmax + 1 + 3) & ~3);
    prm.l_cap = Lmax;
    size_t group_smem = g.ring_bytes + (2 * (32 * WARPS + 1) + 2) * sizeof(float) + 2 * sizeof(int) +
                        (size_t)prm.u_cap * sizeof(int) + 40;
    group_smem = (group_smem + 15) & ~(size_t)15;
    prm.group_smem = group_smem;
    const size_t smem = group_smem * GROUPS;
    auto kern = ctc_viterbi_fill_kernel<P, WARPS, DENSE, PITCH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    const int threads = (WARPS == 1) ? 128 : 32 * WARPS;
    const int blocks = (prm.N + GROUPS - 1) / GROUPS;
    const int prof_slot = profile_begin(stream);
    kern<<<blocks, threads, smem, stream>>>(prm);
    profile_end(prof_slot, stream);
    ++g_launch_count;
    e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

template <int P, int WARPS, bool DENSE>
static int launch_fill(const ViterbiParams &prm, int Lmax, cudaStream_t stream) {
    if constexpr (DENSE) {
        if (prm.V <= 32) return launch_fill_p<P, WARPS, true, 32>(prm, Lmax, stream);
    }
    return launch_fill_p<P, WARPS, DENSE, 0>(prm, Lmax, stream);
}

template <bool DENSE>
static int dispatch_fill(const ViterbiParams &prm, int Lmax, LatticeShape s, cudaStream_t stream) {
#define IPFA_X(P_, W_) \
    if (s.PER == P_ && s.WARPS == W_) return launch_fill<P_, W_, DENSE>(prm, Lmax, stream);
    IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    return IPFA_ERR_UNSUPPORTED;
}

template <int P>
static int launch_backtrace_p(const BacktraceParams &prm, cudaStream_t stream) {
    constexpr int NWORDS = (64 / (2 * P) + 2 + 64 / (2 * P) + 1) * (3 * P);
    const size_t smem = (size_t)4 * (NWORDS + prm.Lmax) * sizeof(uint32_t);
    auto kern = ctc_viterbi_backtrace_kernel<P>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    }
    kern<<<(prm.N + 3) / 4, 128, smem, stream>>>(prm);
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

static int launch_backtrace(const BacktraceParams &prm, int P, cudaStream_t stream) {
    switch (P) {
        case 1: return launch_backtrace_p<1>(prm, stream);
        case 2: return launch_backtrace_p<2>(prm, stream);
        case 4: return launch_backtrace_p<4>(prm, stream);
        case 8: return launch_backtrace_p<8>(prm, stream);
        default: return IPFA_ERR_UNSUPPORTED;
    }
}


}  // namespace ipfa

using namespace ipfa;

static inline size_t vit_pad256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t ipfa_ctc_viterbi_workspace_bytes(int N, int Tmax, int Lmax, int V) {
    (void)V;
    LatticeShape s, s6;
    if (N <= 0 || Tmax < 0 || Lmax < 0 || !pick_lattice_shape(Lmax + 1, N, &s, "IPFA_VITERBI_SHAPE", use_dense_panel(V, Lmax) ? 3 : 6)) return 256;
    size_t bp = (size_t)N * (size_t)viterbi_words_per_window(Tmax, s) * sizeof(uint32_t);
    // (a strided call, stride_v != 1, always takes the gather panel's shape: size for the wider of the two)
    if (pick_lattice_shape(Lmax + 1, N, &s6, "IPFA_VITERBI_SHAPE", 6)) {
        const size_t bp6 = (size_t)N * (size_t)viterbi_words_per_window(Tmax, s6) * sizeof(uint32_t);
        if (bp6 > bp) bp = bp6;
    }
    // planes, final states, bucket lists [2][N] + counters
    return vit_pad256(bp) + vit_pad256((size_t)N * 4) + vit_pad256((size_t)N * 8) + 256 + 256;
}

extern "C" int ipfa_ctc_viterbi_device(const float *lp, int64_t stride_n, int64_t stride_t,
                                       const int32_t *targets, int64_t tgt_stride,
                                       const int32_t *in_len, const int32_t *tgt_len, int N, int Tmax,
                                       int Lmax, int V, int blank, int32_t *paths_out, float *scores_out,
                                       int32_t *tok_start, int32_t *tok_end, float *tok_score,
                                       float *total_out, int32_t *status_out, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    return ipfa_ctc_viterbi_strided_device(lp, stride_n, stride_t, 1, targets, tgt_stride, in_len, tgt_len, N, Tmax,
                                           Lmax, V, blank, paths_out, scores_out, tok_start, tok_end, tok_score,
                                           total_out, status_out, workspace, workspace_bytes, stream);
}

extern "C" int ipfa_ctc_viterbi_strided_device(const float *lp, int64_t stride_n, int64_t stride_t, int64_t stride_v,
                                               const int32_t *targets, int64_t tgt_stride,
                                               const int32_t *in_len, const int32_t *tgt_len, int N, int Tmax,
                                               int Lmax, int V, int blank, int32_t *paths_out, float *scores_out,
                                               int32_t *tok_start, int32_t *tok_end, float *tok_score,
                                               float *total_out, int32_t *status_out, void *workspace,
                                               size_t workspace_bytes, void *stream) {
    if (N == 0) return IPFA_OK;
    if (stride_v < 1) return IPFA_ERR_INVALID_ARG;
    if (!lp || !in_len || !tgt_len || !paths_out || !status_out || N < 0 || Tmax < 0 || V <= 0 ||
        Lmax < 0 || blank < 0 || blank >= V || (Lmax > 0 && !targets) || !workspace ||
        ((tok_start == nullptr) != (tok_end == nullptr)))
        return IPFA_ERR_INVALID_ARG;
    LatticeShape s;
    // (rows that are not contiguous over the vocabulary -- stride_v != 1 -- always go through the gather panel)
    const bool dense = stride_v == 1 && use_dense_panel(V, Lmax);
    if (!pick_lattice_shape(Lmax + 1, N, &s, "IPFA_VITERBI_SHAPE", dense ? 3 : 6)) return IPFA_ERR_UNSUPPORTED;
    if (workspace_bytes < ipfa_ctc_viterbi_workspace_bytes(N, Tmax, Lmax, V)) return IPFA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t wpw = viterbi_words_per_window(Tmax, s);
    uint32_t *bp = static_cast<uint32_t *>(workspace);
    const size_t bp_bytes = ((size_t)N * (size_t)wpw * sizeof(uint32_t) + 255) & ~(size_t)255;
    int32_t *final_state = reinterpret_cast<int32_t *>(static_cast<unsigned char *>(workspace) + bp_bytes);

    int32_t *order = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(final_state) +
                                                 vit_pad256((size_t)N * 4));
    int32_t *count = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(order) +
                                                 vit_pad256((size_t)N * 8));

    ViterbiParams fp{};
    fp.lp = lp; fp.stride_n = stride_n; fp.stride_t = stride_t; fp.stride_v = stride_v;
    fp.targets = targets; fp.tgt_stride = tgt_stride; fp.in_len = in_len; fp.tgt_len = tgt_len;
    fp.N = N; fp.Tmax = Tmax; fp.V = V; fp.blank = blank;
    fp.bp = bp; fp.words_per_window = wpw; fp.final_state = final_state;
    fp.total_out = total_out; fp.status_out = status_out;
    BacktraceParams bt{};
    bt.lp = lp; bt.stride_n = stride_n; bt.stride_t = stride_t; bt.stride_v = stride_v;
    bt.targets = targets; bt.tgt_stride = tgt_stride; bt.in_len = in_len; bt.tgt_len = tgt_len;
    bt.N = N; bt.Tmax = Tmax; bt.Lmax = Lmax; bt.V = V; bt.blank = blank;
    bt.bp = bp; bt.words_per_window = wpw; bt.final_state = final_state;
    bt.paths_out = paths_out; bt.scores_out = scores_out;
    bt.tok_start = tok_start; bt.tok_end = tok_end; bt.tok_score = tok_score;

    // two length buckets when the batch is large and a half-width instance exists
    const int big_units = 32 * s.WARPS * s.PER, small_units = big_units / 2;
    LatticeShape s_small;
    const bool bucketed = N >= kBucketMinWindows && small_units >= 32 && !tuning("IPFA_NO_BUCKETS") &&
                          pick_lattice_shape(small_units, N, &s_small, "IPFA_VITERBI_SMALL_SHAPE", dense ? 3 : 6) &&
                          32 * s_small.WARPS * s_small.PER == small_units;
    if (!bucketed) {
        int rc;
        {
            NvtxRange range("ipfa.viterbi.fill");
            rc = dense ? dispatch_fill<true>(fp, Lmax, s, st) : dispatch_fill<false>(fp, Lmax, s, st);
        }
        if (rc) return rc;
        bt.NT = 32 * s.WARPS;
        NvtxRange range("ipfa.viterbi.backtrace");
        return launch_backtrace(bt, s.PER, st);
    }
    NvtxRange range_b("ipfa.viterbi.length_buckets(fill+backtrace x2)");
    cudaError_t e = cudaMemsetAsync(count, 0, 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    length_bucket_kernel<<<(N + 255) / 256, 256, 0, st>>>(tgt_len, N, small_units, order, count);
    ++g_launch_count;
    for (int cls = 0; cls < 2; ++cls) {
        const LatticeShape sh = cls == 0 ? s_small : s;
        fp.order = order + (int64_t)cls * N; fp.count = count + cls;
        bt.order = fp.order; bt.count = fp.count;
        int rc = dense ? dispatch_fill<true>(fp, Lmax, sh, st) : dispatch_fill<false>(fp, Lmax, sh, st);
        if (rc) return rc;
        bt.NT = 32 * sh.WARPS;
        rc = launch_backtrace(bt, sh.PER, st);
        if (rc) return rc;
    }
    return IPFA_OK;
}
