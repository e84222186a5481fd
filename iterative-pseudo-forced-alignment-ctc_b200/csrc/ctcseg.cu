// ctcseg.cu -- kernel (2b): CTC-segmentation table fill, backtrace and
// utterance scoring: the numerics behind `aligner.get_segments(task)`
//   /root/reference/src/iterative_utterance_alignment.py:216
//   /root/reference/src/word_level_alignment.py:100
//   /root/reference/src/search_on_speech.py:85
// i.e. ctc-segmentation==1.7.1 (/root/reference/requirements.txt:13; not vendored):
//   cython_fill_table            -> ctcseg_fill_kernel       
//   ctc_segmentation() backtrace -> ctcseg_backtrace_kernel  (A5)
//   determine_utterance_segments -> same kernel, epilogue    (A6)
//
// Differences in STRUCTURE (results are identical):
//  * the table is filled time-major (row t needs only row t-1) and never
//    materialised: thread i keeps KC consecutive columns of the current row in
//    registers; the one cross-thread value per frame (column c-1 of the previous
//    thread) moves by __shfl_up / one shared word.
//  * the reference's backtrace re-derives each transition from fp32 table
//    differences ("|stay - est_stay| > |switch - est_switch|").  Every operand
//    of that test is in registers when the cell is filled, so the test is
//    evaluated THERE, with the same fp32 operations, and its outcome is stored
//    as a 1-bit backpointer (bp[w][t / SPW][thread], 32/KC frames per word).
//    The backtrace then only follows bits -- bit-identical decisions, 32x less
//    HBM than the fp32 table.
//  * one fill serves all utterance-prefixes of the text (the shrinking
//    transcript iterations of iterative_utterance_alignment.py:203-385): the
//    per-column first-argmax over t is tracked in registers, each prefix k is
//    backtraced from (argmax_t table[:, utt_begin[k]], utt_begin[k]).
// Full-table mode only (T <= min_window_size = 8000 frames); the windowed
// variant is SURVEY.md section 8(f) rank 2.
#include "ctcseg_walk.cuh"
#include "emission_pipe.cuh"
#include "lattice_shapes.cuh"

namespace ipfa {

extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;
bool use_dense_panel(int V, int Lmax);

constexpr float kProbMax = -1000000000.0f;  // cdef float prob_max = -1000000000

struct SegFillParams {
    const float *lp;
    const int64_t *win_off;  // optional per-window element offset into lp (replaces w * stride_n)
    int64_t stride_n, stride_t;
    const int32_t *in_len;
    const int32_t *gt;
    int64_t gt_stride;
    const int32_t *n_cols;
    const int32_t *utt_begin;  // [N][Kmax+1]
    const int32_t *n_utts;
    int N, Tmax, Cmax, Kmax, V, blank, flags;
    int pitch, tc, u_cap;
    size_t group_smem;
    uint32_t *bp;
    int64_t words_per_window;
    int32_t *colarg;  // [N][Cmax] first argmax_t of the utterance-boundary columns (-1 elsewhere)
};

// Lattice column lc of a window = table column lc + 1.  Column 0 (ground truth -1, the free
// preamble) is not a lattice unit: its value z(t) is 0 for every t (preamble_transition_cost_zero)
// or a running sum of blank emissions, and only enters as the left neighbour of column 1.
// PITCH > 0: compile-time panel pitch (dense rows of <= 32 symbols), so the unrolled frames of a
// backpointer word address the panel with immediate offsets.
// FAST: the reference's default flags (blank_transition_cost_zero = False,
// preamble_transition_cost_zero = True) compiled in, so neither costs an instruction per cell.
// CL > 1: a window's columns are spread over a thread-block CLUSTER of CL CTAs (CTA `rank` owns the
// columns of threads rank*NT ...), for launches that leave SMs idle: the fill is issue bound and a
// window otherwise lives on ONE SM, so CL SMs issue its frames CL times as fast.  The one
// value per frame that crosses a CTA boundary (last column of rank r -> first column of rank r+1) is
// handed over a CHUNK at a time: the producer collects the 32 boundary values of an emission chunk in
// its own shared memory and, at the end of the chunk, sends them into a 4-slot ring in the consumer's
// shared memory as asynchronous remote stores (st.async, distributed shared memory) that complete an
// mbarrier over there -- no fence, nothing to wait for on the producer's side; the consumer waits on
// that mbarrier once per chunk and then reads the values locally, frame by frame.  The CTAs of a window
// therefore run one chunk (plus the hand-over latency) behind each other, like pipeline stages, and
// nothing is exchanged per frame (round 1's per-frame remote ring was slower than no split at all).
// Inside a frame the value the right neighbour waits for is published first and the emissions are
// fetched a frame ahead (see frame()), which is what the few warps of such a CTA are bound by.
// Back-pressure: the consumer reports the chunks it has finished; the producer never runs more than
// two chunks ahead of that, so a ring slot is not rewritten while its last value is still needed.
constexpr int kSpreadSlots = 4;   // ring of chunk-sized boundary buffers (a chunk = 32 frames)
constexpr int kSpreadBytes = kSpreadSlots * 32 * 4 + kSpreadSlots * 8 + 32 * 4 + 16;

template <int KC, int WARPS, bool DENSE, int PITCH, bool FAST, int CL = 1>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32 * WARPS)
ctcseg_fill_kernel(const SegFillParams prm) {
    static_assert(CL == 1 || WARPS > 1, "clusters split multi-warp windows");
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    constexpr int NT = 32 * WARPS;
    constexpr int SPW = 32 / KC;  // frames per backpointer word
    constexpr float kNegInf = -__builtin_huge_valf();
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int group = (WARPS == 1) ? (threadIdx.x >> 5) : 0;
    const int ltid = (WARPS == 1) ? (threadIdx.x & 31) : threadIdx.x;   // thread inside the CTA / group
    const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
    const int w = (CL > 1) ? (int)(blockIdx.x / CL) : (int)(blockIdx.x * GROUPS + group);
    if (w >= prm.N) return;  // (whole clusters leave together)
    const int tid = rank * NT + ltid;                                    // thread inside the window
    const int pitch = PITCH ? PITCH : prm.pitch;

    unsigned char *gsm = smem_raw + (size_t)group * prm.group_smem;
    float *ring = reinterpret_cast<float *>(gsm);
    float *xline = ring + (size_t)kStages * prm.tc * pitch;     // [2][NT + 1] neighbour exchange
    int *cols = reinterpret_cast<int *>(xline + (CL > 1 ? 3 : 2) * (NT + 1));    // [u_cap]

    const int T = min(prm.in_len[w], prm.Tmax);
    const int NC = max(0, min(prm.n_cols[w], prm.Cmax));
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const int blank = prm.blank;
    const bool blank_cost_zero = FAST ? false : (prm.flags & IPFA_SEG_BLANK_COST_ZERO) != 0;
    const bool preamble_cost_zero = FAST ? true : (prm.flags & IPFA_SEG_PREAMBLE_COST_ZERO) != 0;
    int32_t *colarg_w = prm.colarg + (int64_t)w * prm.Cmax;

    for (int c = tid; c < prm.Cmax; c += CL * NT) colarg_w[c] = -1;
    if (T <= 0 || NC <= 1) return;

    int col[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = 1 + tid * KC + k;
        int g = (c < NC) ? gt[c] : blank;
        if (g < 0 || g >= prm.V) g = blank;
        col[k] = DENSE ? g : ((c < NC) ? c : 0);
    }
    int colb = DENSE ? blank : 0;
    int U = NC;  // panel columns
    if constexpr (!DENSE) {
        for (int j = ltid; j < NC; j += NT) {
            int g = (j == 0) ? blank : gt[j];
            if (g < 0 || g >= prm.V) g = blank;
            cols[j] = g;
        }
        group_sync<WARPS>();
        // ascending, unique column list; the emission ring (idle until the prologue) is scratch
        int *scratch = reinterpret_cast<int *>(ring);
        U = sort_unique_columns<WARPS>(cols, NC, scratch, ltid);
        const int *pos = scratch + 2 * NC;
#pragma unroll
        for (int k = 0; k < KC; ++k) col[k] = pos[col[k]];
        colb = pos[0];
    }
    // only the utterance-boundary columns need their first argmax over t; a warp that owns
    // none of them skips the tracking code altogether
    bool track = false;
    {
        const int K = max(0, min(prm.n_utts[w], prm.Kmax));
        const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
        const int c0 = 1 + tid * KC;
        for (int u = 1; u <= K; ++u) {
            const int c = ub[u];
            track |= (c >= c0 && c < c0 + KC);
        }
    }
    const bool warp_tracks = __any_sync(0xffffffffu, track);
    if constexpr (WARPS > 1) {
        if (ltid < (CL > 1 ? 3 : 2)) xline[ltid * (NT + 1)] = 0.0f;  // z(t) = table[t, 0]; table[0, 0] = 0
    }
    // cluster hand-over (CL > 1): [kSpreadSlots][32] incoming boundary values, their flags, the outgoing
    // staging line, the right neighbour's progress word
    float *sp_in = reinterpret_cast<float *>(gsm + prm.group_smem - 32 - kSpreadBytes);
    uint64_t *sp_bar = reinterpret_cast<uint64_t *>(sp_in + kSpreadSlots * 32);   // one mbarrier per slot
    float *sp_out = reinterpret_cast<float *>(sp_bar + kSpreadSlots);
    volatile uint32_t *sp_progress = reinterpret_cast<volatile uint32_t *>(sp_out + 32);
    uint32_t sp_in_next = 0, sp_bar_next = 0, sp_progress_prev = 0;  // remote (shared::cluster) addresses
    if constexpr (CL > 1) {
        if (ltid == 0) {
            *sp_progress = 0;
            for (int i = 0; i < kSpreadSlots; ++i) mbar_init(&sp_bar[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            // every slot is armed for one chunk of 32 values: the producer's asynchronous stores complete it
            for (int i = 0; i < kSpreadSlots; ++i) mbar_expect_tx(&sp_bar[i], 128);
        }
        if (ltid == NT - 1) sp_out[0] = kProbMax;  // table[0, c] = prob_max for every column c >= 1
        if (rank + 1 < CL) {
            sp_in_next = cluster_map(sp_in, rank + 1);
            sp_bar_next = cluster_map(sp_bar, rank + 1);
        }
        if (rank > 0) sp_progress_prev = cluster_map(const_cast<uint32_t *>(sp_progress), rank - 1);
        cluster_sync_all();  // every ring is initialised before any neighbour stores into it
    }
    group_sync<WARPS>();

    EmissionPipe<WARPS, DENSE> pipe;
    pipe.init(ring, cols, prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n), prm.stride_t, T, U,
              prm.V, pitch, prm.tc,
              reinterpret_cast<uint64_t *>(gsm + prm.group_smem - 32), ltid);
    pipe.prologue(ltid);

    float val[KC], cmax[KC];
    int carg[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) { val[k] = kProbMax; cmax[k] = kNegInf; carg[k] = -1; }
    float z = 0.0f;  // column 0 at the previous frame
    uint32_t word = 0;
    int shift = 0;
    uint32_t *bp_ptr = prm.bp + (int64_t)w * prm.words_per_window + tid;
    int t = 0;

    // one frame; returns this thread's KC decision bits
    int li = 0;  // CL > 1: exchange line of the frame being written (three lines, frame t -> line t % 3)
    // CL > 1: the emissions of a frame are fetched one frame ahead (eb_n / ec_n), so that no shared-memory
    // load latency sits between the barrier and the values the neighbours wait for
    float eb_n = 0.0f, ec_n[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) ec_n[k] = 0.0f;
    auto prefetch = [&](const float *row) {
        eb_n = row[colb];
#pragma unroll
        for (int k = 0; k < KC; ++k) ec_n[k] = row[col[k]];
    };
    auto frame = [&](const float *row, const float *rd, float *wr) -> uint32_t {
        float eb, ec[KC];
        if constexpr (CL > 1) {
            eb = eb_n;
#pragma unroll
            for (int k = 0; k < KC; ++k) ec[k] = ec_n[k];
            prefetch(row + pitch);  // (past the chunk's last row: stale rows of the ring, re-fetched at the chunk start)
        } else {
            eb = row[colb];
#pragma unroll
            for (int k = 0; k < KC; ++k) ec[k] = row[col[k]];
        }
        float left_[KC], up_[KC], stayp_[KC];
        const int t_now = t;
        if constexpr (CL > 1) {
            // Publish first.  The value the right neighbour waits for -- this thread's LAST column -- depends
            // only on this thread's own values of the previous frame, so it is computed, stored and the CTA
            // barrier is passed BEFORE the left neighbour's value (one shared-memory load later) is needed for
            // the first column: with the few warps a CTA of a spread window has, the frame time is this
            // dependent chain, not the issue rate.  The loads after the barrier read the line the previous
            // frame wrote while faster neighbours may already write the next one: three lines.
            static_assert(KC >= 2, "spread windows keep at least two columns per thread");
            float *lw = xline + li * (NT + 1);
            const float *lr = xline + (li == 0 ? 2 : li - 1) * (NT + 1);
            li = (li == 2) ? 0 : li + 1;
            {
                constexpr int k = KC - 1;
                const float left = val[k - 1], up = val[k];
                const float stay_p = fmaxf(eb, ec[k]);
                left_[k] = left; up_[k] = up; stayp_[k] = stay_p;
                val[k] = fmaxf(__fadd_rn(left, ec[k]), __fadd_rn(up, stay_p));
            }
            lw[ltid + 1] = val[KC - 1];
            if (rank + 1 < CL && ltid == NT - 1) sp_out[t & 31] = val[KC - 1];
            __syncthreads();
            float prev = lr[ltid];
            // table[t-1] of the left neighbour's last column: delivered with the chunk of frame t-1
            if (rank > 0 && ltid == 0) prev = sp_in[(((t - 1) >> 5) & (kSpreadSlots - 1)) * 32 + ((t - 1) & 31)];
#pragma unroll
            for (int k = KC - 2; k >= 0; --k) {
                const float left = (k == 0) ? prev : val[k - 1];
                const float up = val[k];
                const float stay_p = fmaxf(eb, ec[k]);
                left_[k] = left; up_[k] = up; stayp_[k] = stay_p;
                val[k] = fmaxf(__fadd_rn(left, ec[k]), __fadd_rn(up, stay_p));
            }
            ++t;
        } else {
        float prev;
        if constexpr (WARPS > 1) {
            prev = rd[ltid];
        } else {
            prev = __shfl_up_sync(0xffffffffu, val[KC - 1], 1);
            if (ltid == 0) prev = z;
        }
        // Only table[t-1] -> table[t] is on the chain that the next frame (and the neighbour thread,
        // through the exchange line and the CTA barrier) waits for; the transition test and the
        // arg-max bookkeeping of this frame are computed AFTER the barrier, in the shadow of the next
        // frame's shared-memory loads.
#pragma unroll
        for (int k = KC - 1; k >= 0; --k) {
            const float left = (k == 0) ? prev : val[k - 1];  // table[t-1, c-1]
            const float up = val[k];                          // table[t-1, c]
            const float sw = __fadd_rn(left, ec[k]);
            const float stay_p = fmaxf(eb, ec[k]);
            const float st = blank_cost_zero ? up : __fadd_rn(up, stay_p);
            left_[k] = left; up_[k] = up; stayp_[k] = stay_p;
            val[k] = fmaxf(sw, st);
        }
        ++t;
        if (!preamble_cost_zero) z = fmaxf(kProbMax, __fadd_rn(z, eb));
        if constexpr (WARPS > 1) {
            wr[ltid + 1] = val[KC - 1];
            if (!preamble_cost_zero && ltid == 0) wr[0] = z;
            __syncthreads();
        }
        }
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
            bp_ptr += CL * NT;
            word = 0;
            shift = 0;
        }
    };

    float *line0 = xline, *line1 = xline + NT + 1;
    for (int chunk = 0; chunk < pipe.nchunks; ++chunk) {
        const float *panel = pipe.acquire(chunk, ltid);
        const int rows = min(pipe.tc, T - chunk * pipe.tc);
        if constexpr (CL > 1) {
            // frames 32c+1 .. 32c+32 read the neighbour's values of frames 32c .. 32c+31: chunk c's buffer
            // (the first frame of the chunk reads the last value of chunk c-1, delivered before)
            if (rank > 0 && ltid == 0)
                mbar_wait(&sp_bar[chunk & (kSpreadSlots - 1)], (uint32_t)(chunk / kSpreadSlots) & 1u);
        }
        int r = 0;
        if (chunk == 0) {
            // t = 0: every column c >= 1 is max(switch = prob_max, stay = prob_max)
            if (warp_tracks) {
#pragma unroll
                for (int k = 0; k < KC; ++k) { cmax[k] = kProbMax; carg[k] = 0; }
            }
            push_bits(0);
            t = 1;
            if constexpr (WARPS > 1) {
                line0[ltid + 1] = kProbMax;  // (CL > 1: line 0 = the line of frame 0)
                __syncthreads();
                li = 1;
            }
            r = 1;
        }
        const float *row = panel + r * pitch;
        if constexpr (CL > 1) prefetch(row);
        // frame t reads line[(t-1)&1], writes line[t&1]; tc is even so r has t's parity
        while (r < rows) {
            if (shift == 0 && !(r & 1) && r + SPW <= rows) {
                // a whole backpointer word: SPW frames with compile-time shifts and line parity
                uint32_t acc = 0;
#pragma unroll
                for (int f = 0; f < SPW; ++f) {
                    acc |= frame(row + f * pitch, (f & 1) ? line0 : line1, (f & 1) ? line1 : line0) << (f * KC);
                }
                *bp_ptr = acc;
                bp_ptr += CL * NT;
                row += SPW * pitch;
                r += SPW;
            } else {
                push_bits((r & 1) ? frame(row, line0, line1) : frame(row, line1, line0));
                row += pitch;
                ++r;
            }
        }
        if constexpr (CL > 1) {
            // (every frame ended with a CTA barrier: sp_out holds the whole chunk.)  The 32 values go to the
            // neighbour as asynchronous remote stores that complete the slot's mbarrier there: no fence,
            // nothing for this CTA to wait for.
            if (rank + 1 < CL && ltid < 32) {
                if (ltid == 0) {  // never more than two chunks ahead of what the neighbour has finished
                    while ((int)((uint32_t)chunk - *sp_progress) > 2) {}
                }
                __syncwarp();
                const uint32_t slot = (uint32_t)(chunk & (kSpreadSlots - 1));
                st_async_u32(sp_in_next + (slot * 32 + ltid) * 4, __float_as_uint(sp_out[ltid]), sp_bar_next + slot * 8);
            }
            if (rank > 0 && ltid == 0) {  // done with the slot: arm it for chunk + kSpreadSlots, then say so
                mbar_expect_tx(&sp_bar[chunk & (kSpreadSlots - 1)], 128);
                st_cluster_u32(sp_progress_prev, (uint32_t)(chunk + 1));
            }
        }
    }
    if (shift != 0) *bp_ptr = word;
    if constexpr (CL > 1) cluster_sync_all();  // no CTA leaves while a neighbour may still store into it
    if (warp_tracks) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int c = 1 + tid * KC + k;
            if (c < NC) colarg_w[c] = carg[k];
        }
    }
}

struct SegBackParams {
    const float *lp;
    const int64_t *win_off;
    int64_t stride_n, stride_t;
    const int32_t *in_len;
    const int32_t *gt;
    int64_t gt_stride;
    const int32_t *n_cols;
    const int32_t *utt_begin;  // [N][Kmax+1]
    const int32_t *n_utts;
    int N, Tmax, Cmax, Kmax, V, blank, flags, NT, score_len;
    double index_duration;
    const uint32_t *bp;
    int64_t words_per_window;
    const int32_t *colarg;
    double *seg_out;       // [N][Kmax][Kmax][3]
    int32_t *term_t_out;   // [N][Kmax]
    int32_t *timing;       // [N][Kmax][Cmax] (output or scratch)
    float *char_prob;      // [N][Kmax][Tmax] (output or scratch)
    int32_t *state_out;    // [N][Kmax][Tmax] nullable
    int32_t *status_out;   // [N]
};

// One warp per (window, prefix).
template <int KC>
__global__ void __launch_bounds__(128) ctcseg_backtrace_kernel(const SegBackParams prm) {
    extern __shared__ __align__(16) unsigned char bt_smem[];  // per warp: seg_walk_smem_words<KC>() + gt[Cmax]
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (gw >= prm.N * prm.Kmax) return;
    const int w = gw / prm.Kmax;
    const int kslot = gw - w * prm.Kmax;  // prefix length k = kslot + 1
    const int K = max(0, min(prm.n_utts[w], prm.Kmax));
    const int T = min(prm.in_len[w], prm.Tmax);
    const int NC = max(0, min(prm.n_cols[w], prm.Cmax));
    const bool all_prefixes = prm.flags & IPFA_SEG_ALL_PREFIXES;

    if (kslot == 0 && lane == 0) prm.status_out[w] = (NC > T) ? IPFA_WIN_TEXT_LONGER : IPFA_WIN_OK;
    if (kslot < 0 || kslot >= K || (!all_prefixes && kslot != K - 1)) return;

    const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const float *lp = prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n);
    const int64_t slot = (int64_t)w * prm.Kmax + kslot;
    int32_t *timing = prm.timing + slot * prm.Cmax;
    float *cprob = prm.char_prob + slot * prm.Tmax;
    int32_t *state = prm.state_out ? prm.state_out + slot * prm.Tmax : nullptr;
    double *seg = prm.seg_out + slot * prm.Kmax * 3;

    const int c_end = ub[kslot + 1];  // terminal column of this prefix
    for (int t = lane; t < prm.Tmax; t += 32) {
        cprob[t] = 0.0f;
        if (state) state[t] = -2;
    }
    for (int c = lane; c < prm.Cmax; c += 32) timing[c] = -1;
    const bool feasible = T > 0 && c_end >= 1 && c_end < NC && c_end + 1 <= T;
    int t_term = feasible ? prm.colarg[(int64_t)w * prm.Cmax + c_end] : -1;
    if (lane == 0) prm.term_t_out[slot] = t_term;
    if (!feasible || t_term < 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
        return;
    }
    __syncwarp();

    // the walk and the scoring are one warp's job (ctcseg_walk.cuh, shared with the file-resident sweep)
    uint32_t *raw = reinterpret_cast<uint32_t *>(bt_smem) +
                    (size_t)(threadIdx.x >> 5) * (seg_walk_smem_words<KC>() + prm.Cmax);
    int32_t *gt_s = reinterpret_cast<int32_t *>(raw + seg_walk_smem_words<KC>());
    for (int cc = lane; cc <= c_end; cc += 32) {
        int g = gt[cc];
        if (g < 0 || g >= prm.V) g = prm.blank;
        gt_s[cc] = g;
    }
    SegWalkArgs a;
    a.lp = lp; a.stride_t = prm.stride_t; a.T = T; a.Cmax = prm.Cmax; a.blank = prm.blank; a.NT = prm.NT;
    a.score_len = prm.score_len; a.round_nearest = (prm.flags & IPFA_SEG_ROUND_NEAREST) != 0;
    a.index_duration = prm.index_duration; a.bp_w = prm.bp + (int64_t)w * prm.words_per_window;
    a.ub = ub; a.gt_s = gt_s; a.timing = timing; a.cprob = cprob; a.state = state; a.seg = seg; a.raw = raw;
    seg_walk_prefix<KC, false>(a, kslot, t_term, c_end, lane);
}

// ===========================================================================
// Windowed table mode: audio longer than
// `min_window_size` frames (8000 = 160 s).  ctc-segmentation then keeps only W rows per
// column and slides that window down the audio column by column:
//     offset_c = min(max(0, argmax_{c-1} - W/2), min(ceil((T-W)/N), (T-W) - sum of offsets))
// where argmax_{c-1} is the first arg-max over ALL W rows of the previous column -- so a
// column cannot start before the previous one is complete, and inside a column
//     x_t = max(switch_t, x_{t-1} + e_t)
// is a chain of W fp32-rounded adds.  The fill is therefore column-serial with a serial
// scan per column (phase B below); everything around the chain -- emission gathers, switch
// candidates, the tolerance test of the reference's backtrace (evaluated at fill time and
// stored as one bit per cell, like the full-table kernel), first-argmax -- is data parallel
// over the rows of a tile.  One CTA per (window, prefix): the per-column offsets depend on
// the number of columns, so prefixes do NOT share a fill in this mode.  This is the rare
// path; parity with the reference's arithmetic, not speed, is the point.
struct SegWinParams {
    const float *lp;
    const int64_t *win_off;
    int64_t stride_n, stride_t;
    const int32_t *in_len;
    const int32_t *gt;
    int64_t gt_stride;
    const int32_t *n_cols;
    const int32_t *utt_begin;  // [N][Kmax+1]
    const int32_t *n_utts;
    int N, Tmax, Cmax, Kmax, V, blank, flags, window, score_len;
    double index_duration;
    float *cols2;        // [problems][2][window] previous / current table column
    uint32_t *bits;      // [problems][Cmax][words] 1 bit per (column, row): reverse transition = switch
    int64_t words;       // 32-bit words per column = ceil(window / 32)
    int32_t *offsets;    // [problems][Cmax] window offset of every column
    int32_t *term;       // [problems] first argmax row of the last column
    double *seg_out;     // [N][Kmax][Kmax][3]
    int32_t *term_t_out; // [N][Kmax]
    int32_t *timing;     // [N][Kmax][Cmax]
    float *char_prob;    // [N][Kmax][Tmax]
    int32_t *state_out;  // nullable
    int32_t *status_out; // [N]
};

constexpr int kWinTile = 2048;
constexpr int kWinThreads = 256;

__global__ void __launch_bounds__(kWinThreads) ctcseg_windowed_fill_kernel(const SegWinParams prm) {
    const int prob = blockIdx.x;
    const int PP = (prm.flags & IPFA_SEG_ALL_PREFIXES) ? prm.Kmax : 1;  // problems per window
    const int w = prob / PP;
    const int kslot = (PP > 1) ? prob - w * PP : max(0, min(prm.n_utts[w], prm.Kmax)) - 1;
    const int K = max(0, min(prm.n_utts[w], prm.Kmax));
    const bool all_prefixes = prm.flags & IPFA_SEG_ALL_PREFIXES;
    // The window's status word is initialised HERE, by the fill of its first problem: the backtrace
    // launch that follows only ORs IPFA_WIN_WINDOW_TOO_SMALL into it (a store there could overwrite the
    // bit another prefix of the same window had already set).
    if ((PP == 1 || kslot == 0) && threadIdx.x == 0)
        prm.status_out[w] = (max(0, min(prm.n_cols[w], prm.Cmax)) > min(prm.in_len[w], prm.Tmax)) ? IPFA_WIN_TEXT_LONGER
                                                                                                   : IPFA_WIN_OK;
    if (kslot < 0 || kslot >= K || (!all_prefixes && kslot != K - 1)) return;
    const int T = min(prm.in_len[w], prm.Tmax);
    const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
    const int NC = min(ub[kslot + 1] + 1, min(prm.n_cols[w], prm.Cmax));  // columns of this prefix
    if (T <= 0 || NC <= 1 || NC > T) return;
    const int W = min(prm.window, T);
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const float *lp = prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n);
    float *colA = prm.cols2 + (int64_t)prob * 2 * prm.window;
    float *colB = colA + prm.window;
    uint32_t *bits = prm.bits + (int64_t)prob * prm.Cmax * prm.words;
    int32_t *offs = prm.offsets + (int64_t)prob * prm.Cmax;
    const bool blank_cost_zero = prm.flags & IPFA_SEG_BLANK_COST_ZERO;
    const bool preamble_cost_zero = prm.flags & IPFA_SEG_PREAMBLE_COST_ZERO;
    const int tid = threadIdx.x;

    __shared__ float s_sw[kWinTile], s_e[kWinTile], s_eg[kWinTile], s_x[kWinTile + 1];
    __shared__ int s_off, s_offsum, s_arg;
    __shared__ float s_max, s_carry;
    __shared__ float s_rv[kWinThreads / 32];
    __shared__ int s_ri[kWinThreads / 32];

    const float mean_offset = (float)(T - W) / (float)NC;
    const int higher_offset = (prm.flags & IPFA_SEG_WINDOW_STEP_CEIL) ? (int)ceilf(mean_offset) : (int)mean_offset + 1;
    if (tid == 0) { s_off = 0; s_offsum = 0; s_arg = -1; s_max = 0.0f; }
    __syncthreads();

    float *prev = colA, *cur = colB;
    for (int c = 0; c < NC; ++c) {
        if (tid == 0) {
            int offset = s_off;
            if (c > 0) {
                const int lim = (T - W) - s_offsum;
                const int hi = min(higher_offset, lim);
                const int lo = max(s_arg - W / 2, 0);
                offset = min(lo, hi);
                s_off = offset;
                s_offsum += offset;
            }
            offs[c] = s_offsum;
            s_arg = -1;
            s_max = 0.0f;
        }
        __syncthreads();
        const int offset = s_off, offset_sum = s_offsum;
        int g = gt[c];
        if (c > 0 && (g < 0 || g >= prm.V)) g = prm.blank;
        uint32_t *bits_c = bits + (int64_t)c * prm.words;
        for (int t0 = 0; t0 < W; t0 += kWinTile) {
            const int rows = min(kWinTile, W - t0);
            // phase A: switch candidates and stay emissions of the tile (loads of 4 rows in flight)
#pragma unroll 4
            for (int r = tid; r < rows; r += kWinThreads) {
                const int t = t0 + r;
                const float *row = lp + (int64_t)(t + offset_sum) * prm.stride_t;
                const float eb = row[prm.blank];
                float eg = kProbMax, sw = kProbMax, e;
                if (c > 0) {
                    eg = row[g];
                    const int tp = t - 1 + offset;
                    if (!(tp >= W || tp < 0 || t - 1 < 0)) sw = prev[tp] + eg;
                    e = blank_cost_zero ? 0.0f : fmaxf(eb, eg);
                } else {
                    e = eb;  // column 0: running blank sum when the preamble is not free
                }
                s_sw[r] = sw; s_e[r] = e; s_eg[r] = eg;
            }
            __syncthreads();
            // phase B: the serial chain x_t = max(switch_t, x_{t-1} + e_t).  Only FADD -> FMNMX sits on
            // it: operands are fetched 8 rows at a time, the arg-max is left to phase C.
            if (tid == 0) {
                float x = (t0 == 0) ? 0.0f : s_carry;
                s_x[0] = x;  // x_{t0-1}
                int r = 0;
                if (t0 == 0) {
                    x = (c == 0) ? 0.0f : kProbMax;  // table[0,0] = 0; t - 1 < 0: both candidates are prob_max
                    s_x[1] = x;
                    r = 1;
                }
                if (c == 0 && preamble_cost_zero) {
                    for (; r < rows; ++r) { x = fmaxf(s_sw[r], 0.0f); s_x[r + 1] = x; }
                } else {
                    constexpr int UB = 8;
                    for (; r + UB <= rows; r += UB) {
                        float a[UB], e[UB];
#pragma unroll
                        for (int j = 0; j < UB; ++j) { a[j] = s_sw[r + j]; e[j] = s_e[r + j]; }
#pragma unroll
                        for (int j = 0; j < UB; ++j) { x = fmaxf(a[j], __fadd_rn(x, e[j])); a[j] = x; }
#pragma unroll
                        for (int j = 0; j < UB; ++j) s_x[r + 1 + j] = a[j];
                    }
                    for (; r < rows; ++r) { x = fmaxf(s_sw[r], __fadd_rn(x, s_e[r])); s_x[r + 1] = x; }
                }
                s_carry = x;
            }
            __syncthreads();
            // phase C: the column goes to the workspace; the tolerance test of the reference's
            // backtrace, |stay_prob - est_stay| > |switch_prob - est_switch| -> switch, as one bit;
            // first arg-max of the tile (strict '<' keeps the earliest row)
            float bv = -__builtin_huge_valf();
            int bi = 0x7fffffff;
#pragma unroll 2
            for (int r = tid; r < ((rows + 31) & ~31); r += kWinThreads) {
                const int t = t0 + r;
                bool sw_bit = false;
                if (r < rows) {
                    const float x = s_x[r + 1];
                    cur[t] = x;
                    if (!(c == 0 && t == 0) && bv < x) { bv = x; bi = t; }  // column 0's loop starts at t = 1
                    const int tp = t - 1 + offset;
                    if (c > 0 && t > 0 && tp >= 0 && tp < W) {
                        const float eg = s_eg[r];
                        const float stay_prob = fmaxf(lp[(int64_t)(t + offset_sum) * prm.stride_t + prm.blank], eg);
                        const float est_switch = x - prev[tp];
                        const float est_stay = x - s_x[r];
                        sw_bit = fabsf(stay_prob - est_stay) > fabsf(eg - est_switch);
                    }
                }
                const uint32_t word = __ballot_sync(0xffffffffu, sw_bit);
                if ((tid & 31) == 0 && r < rows) bits_c[t >> 5] = word;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if ((tid & 31) == 0) { s_rv[tid >> 5] = bv; s_ri[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int q = 1; q < kWinThreads / 32; ++q)
                    if (s_rv[q] > bv || (s_rv[q] == bv && s_ri[q] < bi)) { bv = s_rv[q]; bi = s_ri[q]; }
                if (bi != 0x7fffffff && (s_arg == -1 || s_max < bv)) { s_max = bv; s_arg = bi; }
            }
            __syncthreads();
        }
        float *tmp = prev; prev = cur; cur = tmp;
    }
    if (tid == 0) prm.term[prob] = s_arg;
}

// One warp per (window, prefix): follow the bits from (argmax of the last column, last column)
// to (0, 0) in window coordinates, then score the utterances.
__global__ void __launch_bounds__(128) ctcseg_windowed_backtrace_kernel(const SegWinParams prm) {
    const int lane = threadIdx.x & 31;
    const int prob = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int PP = (prm.flags & IPFA_SEG_ALL_PREFIXES) ? prm.Kmax : 1;  // problems per window
    if (prob >= prm.N * PP) return;
    const int w = prob / PP;
    const int kslot = (PP > 1) ? prob - w * PP : max(0, min(prm.n_utts[w], prm.Kmax)) - 1;
    const int K = max(0, min(prm.n_utts[w], prm.Kmax));
    const int T = min(prm.in_len[w], prm.Tmax);
    const int NCw = max(0, min(prm.n_cols[w], prm.Cmax));
    const bool all_prefixes = prm.flags & IPFA_SEG_ALL_PREFIXES;
    // (status_out[w] was initialised by the fill kernel; this launch only ORs into it)
    if (kslot < 0 || kslot >= K || (!all_prefixes && kslot != K - 1)) return;
    const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const float *lp = prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n);
    const int64_t slot = (int64_t)w * prm.Kmax + kslot;
    int32_t *timing = prm.timing + slot * prm.Cmax;
    float *cprob = prm.char_prob + slot * prm.Tmax;
    int32_t *state = prm.state_out ? prm.state_out + slot * prm.Tmax : nullptr;
    double *seg = prm.seg_out + slot * prm.Kmax * 3;
    const int NC = min(ub[kslot + 1] + 1, NCw);
    for (int t = lane; t < prm.Tmax; t += 32) {
        cprob[t] = 0.0f;
        if (state) state[t] = -2;
    }
    for (int c = lane; c < prm.Cmax; c += 32) timing[c] = -1;
    const bool feasible = T > 0 && NC > 1 && NC <= T;
    const int t_term = feasible ? prm.term[prob] : -1;
    if (lane == 0) prm.term_t_out[slot] = t_term;  // window row; absolute frame = row + offset of the column
    if (!feasible || t_term < 0) {
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
        return;
    }
    __syncwarp();
    const int W = min(prm.window, T);
    const uint32_t *bits = prm.bits + (int64_t)prob * prm.Cmax * prm.words;
    const int32_t *offs = prm.offsets + (int64_t)prob * prm.Cmax;
    bool too_small = false;
    if (lane == 0) {
        int t = t_term, c = NC - 1;
        while (t != 0 || c != 0) {
            const int oc = offs[c];
            const int offset = (c > 0) ? oc - offs[c - 1] : 0;
            // the reference reads table[t - 1 + offset, c - 1] here: beyond the window -> IndexError
            // (it doubles the window and starts over); row 0 of a later column is unreachable
            if ((c > 0 && t - 1 + offset >= W) || (t == 0 && c > 0)) { too_small = true; break; }
            const float *row = lp + (int64_t)(t + oc) * prm.stride_t;
            const float eb = row[prm.blank];
            if (c == 0) {
                cprob[oc + t] = eb;
                if (state) state[oc + t] = -1;
                --t;
                continue;
            }
            int g = gt[c];
            if (g < 0 || g >= prm.V) g = prm.blank;
            const float eg = row[g];
            const bool sw = (bits[(int64_t)c * prm.words + (t >> 5)] >> (t & 31)) & 1u;
            if (sw) {
                timing[c] = oc + t;
                cprob[oc + t] = eg;
                if (state) state[oc + t] = c;
                --c;
                t -= 1 - offset;
            } else {
                cprob[oc + t] = fmaxf(eb, eg);
                if (state) state[oc + t] = -1;
                --t;
            }
        }
    }
    too_small = __shfl_sync(0xffffffffu, too_small, 0);
    __syncwarp();
    __threadfence_block();
    if (too_small) {
        if (lane == 0) atomicOr(prm.status_out + w, IPFA_WIN_WINDOW_TOO_SMALL);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
        return;
    }
    score_segments(ub, timing, cprob, kslot + 1, T, prm.Cmax, prm.index_duration, prm.score_len,
                   (prm.flags & IPFA_SEG_ROUND_NEAREST) != 0, lane, seg);
}

// ---------------------------------------------------------------------------
// Multi-column ground truth: ctc-segmentation's `classic` text
// converter gives every character position up to G candidate tokens, the s-th spanning the last
// s+1 characters, so a cell has up to G switch transitions (from columns c-1-s).  Same
// column-serial scheme as the windowed kernels above (it IS the general form of the algorithm:
// window = T gives the full table); G+1 table columns are kept, and the reference's backtrace
// test -- argmin_s |switch_prob_s - est_switch_s| against |stay_prob - est_stay| -- is stored as one
// byte per cell: 0 stay, 1 + min_s switch, 255 "a read beyond the window" (IndexError there).
constexpr int kMaxGtCols = 32;

struct SegMultiParams {
    SegWinParams w;      // gt is [N][Cmax][G] here (gt_stride elements per window)
    int G;
    float *ring;         // [problems][G + 1][window] last G + 1 table columns
    uint8_t *codes;      // [problems][Cmax][window]
};

__global__ void __launch_bounds__(kWinThreads) ctcseg_multi_fill_kernel(const SegMultiParams mp) {
    const SegWinParams &prm = mp.w;
    const int G = mp.G;
    const int prob = blockIdx.x;
    const int PP = (prm.flags & IPFA_SEG_ALL_PREFIXES) ? prm.Kmax : 1;  // problems per window
    const int w = prob / PP;
    const int kslot = (PP > 1) ? prob - w * PP : max(0, min(prm.n_utts[w], prm.Kmax)) - 1;
    const int K = max(0, min(prm.n_utts[w], prm.Kmax));
    const bool all_prefixes = prm.flags & IPFA_SEG_ALL_PREFIXES;
    // The window's status word is initialised HERE, by the fill of its first problem: the backtrace
    // launch that follows only ORs IPFA_WIN_WINDOW_TOO_SMALL into it (a store there could overwrite the
    // bit another prefix of the same window had already set).
    if ((PP == 1 || kslot == 0) && threadIdx.x == 0)
        prm.status_out[w] = (max(0, min(prm.n_cols[w], prm.Cmax)) > min(prm.in_len[w], prm.Tmax)) ? IPFA_WIN_TEXT_LONGER
                                                                                                   : IPFA_WIN_OK;
    if (kslot < 0 || kslot >= K || (!all_prefixes && kslot != K - 1)) return;
    const int T = min(prm.in_len[w], prm.Tmax);
    const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
    const int NC = min(ub[kslot + 1] + 1, min(prm.n_cols[w], prm.Cmax));
    if (T <= 0 || NC <= 1 || NC > T) return;
    const int W = min(prm.window, T);
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const float *lp = prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n);
    float *ring = mp.ring + (int64_t)prob * (G + 1) * prm.window;
    uint8_t *codes = mp.codes + (int64_t)prob * prm.Cmax * prm.window;
    int32_t *offs = prm.offsets + (int64_t)prob * prm.Cmax;
    const bool blank_cost_zero = prm.flags & IPFA_SEG_BLANK_COST_ZERO;
    const bool preamble_cost_zero = prm.flags & IPFA_SEG_PREAMBLE_COST_ZERO;
    const int tid = threadIdx.x;

    __shared__ float s_sw[kWinTile], s_e[kWinTile], s_mx[kWinTile], s_x[kWinTile + 1];
    __shared__ int s_off, s_offsum, s_arg, s_curoff[kMaxGtCols], s_g[kMaxGtCols];
    __shared__ float s_max, s_carry;

    const float mean_offset = (float)(T - W) / (float)NC;
    const int higher_offset = (prm.flags & IPFA_SEG_WINDOW_STEP_CEIL) ? (int)ceilf(mean_offset) : (int)mean_offset + 1;
    if (tid == 0) {
        s_off = 0; s_offsum = 0; s_arg = -1; s_max = 0.0f;
        for (int s = 0; s < G; ++s) s_curoff[s] = -1;  // np.zeros(G) - 1
    }
    __syncthreads();

    for (int c = 0; c < NC; ++c) {
        if (tid == 0) {
            if (c > 0) {
                const int lim = (T - W) - s_offsum;
                const int hi = min(higher_offset, lim);
                const int lo = max(s_arg - W / 2, 0);
                const int offset = min(lo, hi);
                if (prm.flags & IPFA_SEG_OFFSET_SHIFT) {
                    for (int s = G - 2; s >= 0; --s) s_curoff[s + 1] = s_curoff[s] + offset;
                } else {  // ascending: every entry builds on the one just written
                    for (int s = 0; s < G - 1; ++s) s_curoff[s + 1] = s_curoff[s] + offset;
                }
                s_curoff[0] = offset;
                s_off = offset;
                s_offsum += offset;
            }
            offs[c] = s_offsum;
            s_arg = -1;
            s_max = 0.0f;
            for (int s = 0; s < G; ++s) {
                int g = gt[(int64_t)c * G + s];
                if (g >= prm.V) g = -1;
                s_g[s] = g;
            }
        }
        __syncthreads();
        const int offset_sum = s_offsum;
        float *cur = ring + (int64_t)(c % (G + 1)) * prm.window;
        uint8_t *codes_c = codes + (int64_t)c * prm.window;
        for (int t0 = 0; t0 < W; t0 += kWinTile) {
            const int rows = min(kWinTile, W - t0);
            for (int r = tid; r < rows; r += kWinThreads) {
                const int t = t0 + r;
                const float *row = lp + (int64_t)(t + offset_sum) * prm.stride_t;
                const float eb = row[prm.blank];
                float sw = kProbMax, mx = kProbMax;
                for (int s = 0; s < G; ++s) {
                    const int g = s_g[s];
                    if (g < 0) continue;
                    const float e = row[g];
                    const int tp = t - 1 + s_curoff[s];
                    float p = kProbMax;
                    if (!(tp >= W || tp < 0 || t - 1 < 0 || c - (s + 1) < 0))
                        p = ring[(int64_t)((c - 1 - s) % (G + 1)) * prm.window + tp] + e;
                    sw = fmaxf(sw, p);
                    mx = fmaxf(mx, e);
                }
                s_sw[r] = sw;
                s_mx[r] = mx;
                s_e[r] = (c == 0) ? eb : (blank_cost_zero ? 0.0f : fmaxf(eb, mx));
            }
            __syncthreads();
            if (tid == 0) {
                float x = (t0 == 0) ? 0.0f : s_carry;
                float best = s_max;
                int arg = s_arg;
                int r = 0;
                if (t0 == 0) {
                    if (c == 0) {
                        x = 0.0f;
                    } else {
                        x = kProbMax;
                        arg = 0; best = x;
                    }
                    s_x[1] = x;
                    r = 1;
                }
                for (; r < rows; ++r) {
                    const float stay = (c == 0 && preamble_cost_zero) ? 0.0f : x + s_e[r];
                    x = fmaxf(s_sw[r], stay);
                    s_x[r + 1] = x;
                    if (arg == -1 || best < x) { best = x; arg = t0 + r; }
                }
                s_x[0] = (t0 == 0) ? 0.0f : s_carry;
                s_carry = x;
                s_max = best;
                s_arg = arg;
            }
            __syncthreads();
            for (int r = tid; r < rows; r += kWinThreads) {
                const int t = t0 + r;
                const float x = s_x[r + 1];
                cur[t] = x;
                uint8_t code = 0;
                if (c > 0 && t > 0) {
                    const float *row = lp + (int64_t)(t + offset_sum) * prm.stride_t;
                    float min_delta = __int_as_float(0x7f800000);
                    int min_s = -1;
                    float max_lpz = -10000000000.0f;  // config.max_prob
                    bool oob = false;
                    for (int s = 0; s < G; ++s) {
                        const int g = s_g[s];
                        if (g < 0 || c - (s + 1) < 0) continue;  // (a token reaching before column 0 cannot exist)
                        const int tp = t - 1 + s_curoff[s];
                        if (tp >= W) { oob = true; break; }
                        const float switch_prob = row[g];
                        const float est = x - ring[(int64_t)((c - 1 - s) % (G + 1)) * prm.window + tp];
                        const float d = fabsf(switch_prob - est);
                        if (d < min_delta) { min_delta = d; min_s = s; }
                        max_lpz = fmaxf(max_lpz, switch_prob);
                    }
                    if (oob) {
                        code = 255;
                    } else {
                        const float stay_prob = fmaxf(row[prm.blank], max_lpz);
                        const float est_stay = x - s_x[r];
                        if (fabsf(stay_prob - est_stay) > min_delta) code = (uint8_t)(1 + min_s);
                    }
                }
                codes_c[t] = code;
            }
            __syncthreads();
        }
    }
    if (tid == 0) prm.term[prob] = s_arg;
}

__global__ void __launch_bounds__(128) ctcseg_multi_backtrace_kernel(const SegMultiParams mp) {
    const SegWinParams &prm = mp.w;
    const int G = mp.G;
    const int lane = threadIdx.x & 31;
    const int prob = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int PP = (prm.flags & IPFA_SEG_ALL_PREFIXES) ? prm.Kmax : 1;  // problems per window
    if (prob >= prm.N * PP) return;
    const int w = prob / PP;
    const int kslot = (PP > 1) ? prob - w * PP : max(0, min(prm.n_utts[w], prm.Kmax)) - 1;
    const int K = max(0, min(prm.n_utts[w], prm.Kmax));
    const int T = min(prm.in_len[w], prm.Tmax);
    const int NCw = max(0, min(prm.n_cols[w], prm.Cmax));
    const bool all_prefixes = prm.flags & IPFA_SEG_ALL_PREFIXES;
    // (status_out[w] was initialised by the fill kernel; this launch only ORs into it)
    if (kslot < 0 || kslot >= K || (!all_prefixes && kslot != K - 1)) return;
    const int32_t *ub = prm.utt_begin + (int64_t)w * (prm.Kmax + 1);
    const int32_t *gt = prm.gt + (int64_t)w * prm.gt_stride;
    const float *lp = prm.lp + (prm.win_off ? prm.win_off[w] : (int64_t)w * prm.stride_n);
    const int64_t slot = (int64_t)w * prm.Kmax + kslot;
    int32_t *timing = prm.timing + slot * prm.Cmax;
    float *cprob = prm.char_prob + slot * prm.Tmax;
    int32_t *state = prm.state_out ? prm.state_out + slot * prm.Tmax : nullptr;
    double *seg = prm.seg_out + slot * prm.Kmax * 3;
    const int NC = min(ub[kslot + 1] + 1, NCw);
    for (int t = lane; t < prm.Tmax; t += 32) {
        cprob[t] = 0.0f;
        if (state) state[t] = -2;
    }
    for (int c = lane; c < prm.Cmax; c += 32) timing[c] = -1;
    const bool feasible = T > 0 && NC > 1 && NC <= T;
    const int t_term = feasible ? prm.term[prob] : -1;
    if (lane == 0) prm.term_t_out[slot] = t_term;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (!feasible || t_term < 0) {
        for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
        return;
    }
    __syncwarp();
    const uint8_t *codes = mp.codes + (int64_t)prob * prm.Cmax * prm.window;
    const int32_t *offs = prm.offsets + (int64_t)prob * prm.Cmax;
    bool too_small = false;
    if (lane == 0) {
        int t = t_term, c = NC - 1;
        while (t != 0 || c != 0) {
            if (t == 0 || c < 0) { too_small = true; break; }  // row 0 of a later column is unreachable
            const int oc = offs[c];
            const float *row = lp + (int64_t)(t + oc) * prm.stride_t;
            const float eb = row[prm.blank];
            if (c == 0) {
                cprob[oc + t] = eb;
                if (state) state[oc + t] = -1;
                --t;
                continue;
            }
            const uint8_t code = codes[(int64_t)c * prm.window + t];
            if (code == 255) { too_small = true; break; }
            // max_lpz_prob and the `offset` the reference's loop leaves behind (that of the LAST
            // candidate it looked at, not of the chosen one)
            float max_lpz = -10000000000.0f;
            int offset = 0;
            for (int s = 0; s < G; ++s) {
                int g = gt[(int64_t)c * G + s];
                if (g < 0 || g >= prm.V || c - (s + 1) < 0) continue;
                max_lpz = fmaxf(max_lpz, row[g]);
                offset = oc - offs[c - 1 - s];
            }
            if (code > 0) {
                const int min_s = code - 1;
                for (int s = 0; s <= min_s; ++s) timing[c - s] = oc + t;
                cprob[oc + t] = max_lpz;
                if (state) state[oc + t] = c | (min_s << 24);
                c -= 1 + min_s;
                t -= 1 - offset;
            } else {
                cprob[oc + t] = fmaxf(eb, max_lpz);
                if (state) state[oc + t] = -1;
                --t;
            }
        }
    }
    too_small = __shfl_sync(0xffffffffu, too_small, 0);
    __syncwarp();
    __threadfence_block();
    if (too_small) {
        if (lane == 0) atomicOr(prm.status_out + w, IPFA_WIN_WINDOW_TOO_SMALL);
        for (int u = lane; u <= kslot; u += 32) { seg[u * 3] = nan; seg[u * 3 + 1] = nan; seg[u * 3 + 2] = nan; }
        return;
    }
    score_segments(ub, timing, cprob, kslot + 1, T, prm.Cmax, prm.index_duration, prm.score_len,
                   (prm.flags & IPFA_SEG_ROUND_NEAREST) != 0, lane, seg);
}

// ---------------------------------------------------------------------------
using SegShape = LatticeShape;
static bool pick_seg_shape(int cols, int n_windows, int V, SegShape *s) {
    return pick_lattice_shape(cols, n_windows, s, "IPFA_SEG_SHAPE", use_dense_panel(V, cols) ? 3 : 6, true);
}
static int64_t seg_words_per_window(int Tmax, SegShape s) {
    const int spw = 32 / s.PER;
    return (int64_t)((Tmax + spw - 1) / spw) * 32 * s.WARPS;
}
static inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

template <int KC, int WARPS, bool DENSE, int PITCH, bool FAST>
static int launch_seg_fill_p(SegFillParams prm, cudaStream_t stream) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    const int U = PITCH ? PITCH : (DENSE ? prm.V : prm.Cmax);
    const size_t budget = (WARPS == 1) ? (16 * 1024) : (160 * 1024);
    PipeGeometry g = pipe_geometry(U, budget);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = DENSE ? 0 : ((prm.Cmax + 3) & ~3);
    size_t group_smem = g.ring_bytes + 2 * (32 * WARPS + 1) * sizeof(float) + (size_t)prm.u_cap * sizeof(int) + 40;
    group_smem = (group_smem + 15) & ~(size_t)15;
    prm.group_smem = group_smem;
    const size_t smem = group_smem * GROUPS;
    auto kern = ctcseg_fill_kernel<KC, WARPS, DENSE, PITCH, FAST>;
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

// Cluster launch (CL CTAs per window, KC columns per thread, WARPS warps per CTA): dense 32-symbol
// panel with the reference's default flags only -- the configuration of the anchor loop.
template <int KC, int WARPS, int CL>
static int launch_seg_fill_cluster(SegFillParams prm, cudaStream_t stream) {
    PipeGeometry g = pipe_geometry(32, 160 * 1024);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = 0;
    size_t group_smem = g.ring_bytes + 3 * (32 * WARPS + 1) * sizeof(float) + kSpreadBytes + 40;
    group_smem = (group_smem + 15) & ~(size_t)15;
    prm.group_smem = group_smem;
    auto kern = ctcseg_fill_kernel<KC, WARPS, true, 32, true, CL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(prm.N * CL));
    cfg.blockDim = dim3(32 * WARPS);
    cfg.dynamicSmemBytes = group_smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, prm);
    ++g_launch_count;
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

// Spread the columns of every window over CL SMs (cluster launch).  OPT-IN (IPFA_SEG_SPREAD_2 / _4 in
// `flags`): measured on the anchor sweep's windows it does not pay -- 88 us per fill of a 1700-frame,
// 700-column window on one SM, 94 us over two, 89 us over four (profiles/r02_fill_spread.txt) -- because
// with a handful of warps per CTA the frame time is the dependent chain LDS(neighbour) -> FADD -> FMNMX ->
// STS -> BAR (~100 cycles), which is as long as the 132 issue cycles the unsplit window needs.
// Instances: dense 32-symbol panel, default flags, 2 columns per thread (the anchor loop's configuration).
static bool try_seg_fill_cluster(const SegFillParams &prm, SegShape s, cudaStream_t stream, int *rc) {
    const bool fast = (prm.flags & (IPFA_SEG_BLANK_COST_ZERO | IPFA_SEG_PREAMBLE_COST_ZERO)) ==
                      IPFA_SEG_PREAMBLE_COST_ZERO;
    if (!fast || prm.V > 32 || s.PER != 2) return false;
    int cl = (prm.flags & IPFA_SEG_SPREAD_4) ? 4 : (prm.flags & IPFA_SEG_SPREAD_2) ? 2 : 1;
    if (cl == 4 && (s.WARPS % 4 != 0 || s.WARPS < 8)) cl = 2;
    if (cl == 2 && (s.WARPS % 2 != 0 || s.WARPS < 4)) cl = 1;
    if (cl == 1) return false;
#define IPFA_C(CL_, W_) if (cl == CL_ && s.WARPS == CL_ * W_) { *rc = launch_seg_fill_cluster<2, W_, CL_>(prm, stream); return true; }
    IPFA_C(2, 2) IPFA_C(2, 3) IPFA_C(2, 4) IPFA_C(2, 5) IPFA_C(2, 6) IPFA_C(2, 8) IPFA_C(2, 10) IPFA_C(2, 12)
    IPFA_C(4, 2) IPFA_C(4, 3) IPFA_C(4, 4) IPFA_C(4, 5) IPFA_C(4, 6)
#undef IPFA_C
    return false;
}

template <int KC, int WARPS, bool DENSE>
static int launch_seg_fill(const SegFillParams &prm, cudaStream_t stream) {
    const bool fast = (prm.flags & (IPFA_SEG_BLANK_COST_ZERO | IPFA_SEG_PREAMBLE_COST_ZERO)) ==
                      IPFA_SEG_PREAMBLE_COST_ZERO;
    if constexpr (DENSE) {
        if (prm.V <= 32 && fast) return launch_seg_fill_p<KC, WARPS, true, 32, true>(prm, stream);
    }
    return launch_seg_fill_p<KC, WARPS, DENSE, 0, false>(prm, stream);
}

template <bool DENSE>
static int dispatch_seg_fill(const SegFillParams &prm, SegShape s, cudaStream_t stream) {
#define IPFA_X(K_, W_) \
    if (s.PER == K_ && s.WARPS == W_) return launch_seg_fill<K_, W_, DENSE>(prm, stream);
    IPFA_FOR_EACH_SHAPE(IPFA_X)
    IPFA_FOR_EACH_EXTRA_SHAPE(IPFA_X)
#undef IPFA_X
    return IPFA_ERR_UNSUPPORTED;
}

}  // namespace ipfa

using namespace ipfa;

extern "C" size_t ipfa_ctcseg_workspace_bytes(int N, int Tmax, int Cmax, int Kmax, int V) {
    (void)V;
    SegShape s;
    if (N <= 0 || Tmax < 0 || Cmax <= 0 || Kmax <= 0 || !pick_seg_shape(Cmax, N, V, &s)) return 256;
    size_t b = pad256((size_t)N * (size_t)seg_words_per_window(Tmax, s) * 4);
    b += pad256((size_t)N * Cmax * 4);                   // colarg
    b += pad256((size_t)N * Kmax * (size_t)Cmax * 4);    // timing scratch
    b += pad256((size_t)N * Kmax * (size_t)Tmax * 4);    // char_prob scratch
    return b + 256;
}

namespace ipfa {
// Shared by ipfa_ctcseg_device (windows at w * stride_n) and ipfa_ctcseg_windows_device /
// the anchor sweep (windows at win_off[w], slices of corpus-resident emissions).
int ctcseg_run(const float *lp, const int64_t *win_off, int64_t stride_n, int64_t stride_t,
               const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
               const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
               int N, int Tmax, int Cmax, int Kmax, int V, int blank,
               double index_duration, int score_len, int flags, double *seg_out,
               int32_t *term_t_out, int32_t *timing_out, float *char_prob_out,
               int32_t *state_out, int32_t *status_out, void *workspace,
               size_t workspace_bytes, void *stream) {
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !gt || !n_cols || !utt_begin || !n_utts || !seg_out || !term_t_out ||
        !status_out || !workspace || N < 0 || Tmax <= 0 || Cmax <= 0 || Kmax <= 0 || V <= 0 || blank < 0 ||
        blank >= V || score_len <= 0 || !(index_duration > 0.0))
        return IPFA_ERR_INVALID_ARG;
    if (Tmax > 8000) return IPFA_ERR_UNSUPPORTED;  // windowed table mode: ipfa_ctcseg_windowed_device
    SegShape s;
    if (!pick_seg_shape(Cmax, N, V, &s)) return IPFA_ERR_UNSUPPORTED;
    if (workspace_bytes < ipfa_ctcseg_workspace_bytes(N, Tmax, Cmax, Kmax, V)) return IPFA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    const int64_t wpw = seg_words_per_window(Tmax, s);
    uint32_t *bp = reinterpret_cast<uint32_t *>(ws);
    ws += pad256((size_t)N * (size_t)wpw * 4);
    int32_t *colarg = reinterpret_cast<int32_t *>(ws);
    ws += pad256((size_t)N * Cmax * 4);
    int32_t *timing_scratch = reinterpret_cast<int32_t *>(ws);
    ws += pad256((size_t)N * Kmax * (size_t)Cmax * 4);
    float *cprob_scratch = reinterpret_cast<float *>(ws);

    SegFillParams fp{};
    fp.lp = lp; fp.win_off = win_off; fp.stride_n = stride_n; fp.stride_t = stride_t; fp.in_len = in_len;
    fp.gt = gt; fp.gt_stride = gt_stride; fp.n_cols = n_cols; fp.utt_begin = utt_begin; fp.n_utts = n_utts;
    fp.Kmax = Kmax;
    fp.N = N; fp.Tmax = Tmax; fp.Cmax = Cmax; fp.V = V; fp.blank = blank; fp.flags = flags;
    fp.bp = bp; fp.words_per_window = wpw; fp.colarg = colarg;
    int rc = IPFA_OK;
    {
        NvtxRange range("ipfa.ctcseg.fill");
        if (!(use_dense_panel(V, Cmax) && try_seg_fill_cluster(fp, s, st, &rc)))
            rc = use_dense_panel(V, Cmax) ? dispatch_seg_fill<true>(fp, s, st) : dispatch_seg_fill<false>(fp, s, st);
    }
    if (rc) return rc;
    NvtxRange range_bt("ipfa.ctcseg.backtrace+segments");

    SegBackParams bk{};
    bk.lp = lp; bk.win_off = win_off; bk.stride_n = stride_n; bk.stride_t = stride_t; bk.in_len = in_len;
    bk.gt = gt; bk.gt_stride = gt_stride; bk.n_cols = n_cols; bk.utt_begin = utt_begin; bk.n_utts = n_utts;
    bk.N = N; bk.Tmax = Tmax; bk.Cmax = Cmax; bk.Kmax = Kmax; bk.V = V; bk.blank = blank; bk.flags = flags;
    bk.NT = 32 * s.WARPS; bk.score_len = score_len; bk.index_duration = index_duration;
    bk.bp = bp; bk.words_per_window = wpw; bk.colarg = colarg;
    bk.seg_out = seg_out; bk.term_t_out = term_t_out;
    bk.timing = timing_out ? timing_out : timing_scratch;
    bk.char_prob = char_prob_out ? char_prob_out : cprob_scratch;
    bk.state_out = state_out; bk.status_out = status_out;
    const int warps = N * Kmax;
    const int blocks = (warps + 3) / 4;
    {
        const int kc = s.PER;
        const size_t bt_smem = (size_t)4 * (2 * (31 / kc + 2 + 32 / kc + 1) * kc + Cmax) * sizeof(uint32_t);
        cudaError_t ea = cudaSuccess;
#define IPFA_BT(K_)                                                                                     \
    if (kc == K_) {                                                                                     \
        if (bt_smem > 48 * 1024)                                                                        \
            ea = cudaFuncSetAttribute(ctcseg_backtrace_kernel<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)bt_smem);                                                    \
        if (ea == cudaSuccess) ctcseg_backtrace_kernel<K_><<<blocks, 128, bt_smem, st>>>(bk);           \
    }
        IPFA_BT(1) IPFA_BT(2) IPFA_BT(4) IPFA_BT(8)
#undef IPFA_BT
        if (ea != cudaSuccess) { g_last_cuda_error = ea; return IPFA_ERR_CUDA; }
    }
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
}  // namespace ipfa

extern "C" int ipfa_ctcseg_device(const float *lp, int64_t stride_n, int64_t stride_t,
                                  const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                                  const int32_t *n_cols, const int32_t *utt_begin, const int32_t *n_utts,
                                  int N, int Tmax, int Cmax, int Kmax, int V, int blank,
                                  double index_duration, int score_len, int flags, double *seg_out,
                                  int32_t *term_t_out, int32_t *timing_out, float *char_prob_out,
                                  int32_t *state_out, int32_t *status_out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    return ipfa::ctcseg_run(lp, nullptr, stride_n, stride_t, in_len, gt, gt_stride, n_cols, utt_begin, n_utts, N,
                            Tmax, Cmax, Kmax, V, blank, index_duration, score_len, flags, seg_out, term_t_out,
                            timing_out, char_prob_out, state_out, status_out, workspace, workspace_bytes, stream);
}

extern "C" int ipfa_ctcseg_windows_device(const float *lp, const int64_t *win_off, int64_t stride_t,
                                          const int32_t *in_len, const int32_t *gt, int64_t gt_stride,
                                          const int32_t *n_cols, const int32_t *utt_begin,
                                          const int32_t *n_utts, int N, int Tmax, int Cmax, int Kmax, int V,
                                          int blank, double index_duration, int score_len, int flags,
                                          double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                                          float *char_prob_out, int32_t *state_out, int32_t *status_out,
                                          void *workspace, size_t workspace_bytes, void *stream) {
    if (N > 0 && !win_off) return IPFA_ERR_INVALID_ARG;
    return ipfa::ctcseg_run(lp, win_off, 0, stride_t, in_len, gt, gt_stride, n_cols, utt_begin, n_utts, N, Tmax,
                            Cmax, Kmax, V, blank, index_duration, score_len, flags, seg_out, term_t_out,
                            timing_out, char_prob_out, state_out, status_out, workspace, workspace_bytes, stream);
}

// ---------------------------------------------------------------------------
// windowed table mode
static size_t win_words(int window) { return (size_t)((window + 31) / 32); }

extern "C" size_t ipfa_ctcseg_windowed_workspace_bytes(int N, int Tmax, int Cmax, int Kmax, int window,
                                                       int gt_cols, int flags) {
    if (N <= 0 || Tmax <= 0 || Cmax <= 0 || Kmax <= 0 || window <= 0 || gt_cols <= 0) return 256;
    const size_t P = (size_t)N * ((flags & IPFA_SEG_ALL_PREFIXES) ? Kmax : 1);  // one fill per aligned prefix
    const size_t W = (size_t)(window < Tmax ? window : Tmax);
    size_t b = pad256(P * ((size_t)gt_cols + 1) * W * 4);  // the last gt_cols + 1 table columns
    b += (gt_cols == 1) ? pad256(P * (size_t)Cmax * win_words((int)W) * 4)   // 1-bit transitions
                        : pad256(P * (size_t)Cmax * W);                      // 1-byte transitions
    b += pad256(P * (size_t)Cmax * 4);                  // offsets
    b += pad256(P * 4);                                 // terminal rows
    b += pad256((size_t)N * Kmax * (size_t)Cmax * 4);   // timing scratch (indexed by output slot)
    b += pad256((size_t)N * Kmax * (size_t)Tmax * 4);   // char_prob scratch
    return b + 256;
}

extern "C" int ipfa_ctcseg_windowed_device(const float *lp, const int64_t *win_off, int64_t stride_n,
                                           int64_t stride_t, const int32_t *in_len, const int32_t *gt,
                                           int64_t gt_stride, const int32_t *n_cols, const int32_t *utt_begin,
                                           const int32_t *n_utts, int N, int Tmax, int Cmax, int Kmax, int V,
                                           int blank, double index_duration, int score_len, int flags,
                                           int window, int gt_cols, double *seg_out, int32_t *term_t_out,
                                           int32_t *timing_out, float *char_prob_out, int32_t *state_out,
                                           int32_t *status_out, void *workspace, size_t workspace_bytes,
                                           void *stream) {
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !gt || !n_cols || !utt_begin || !n_utts || !seg_out || !term_t_out ||
        !status_out || !workspace || N < 0 || Tmax <= 0 || Cmax <= 0 || Kmax <= 0 || V <= 0 || blank < 0 ||
        blank >= V || score_len <= 0 || !(index_duration > 0.0) || window <= 0 || gt_cols <= 0)
        return IPFA_ERR_INVALID_ARG;
    if (gt_cols > kMaxGtCols) return IPFA_ERR_UNSUPPORTED;
    if (workspace_bytes < ipfa_ctcseg_windowed_workspace_bytes(N, Tmax, Cmax, Kmax, window, gt_cols, flags))
        return IPFA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t P = (size_t)N * ((flags & IPFA_SEG_ALL_PREFIXES) ? Kmax : 1);
    const int W = window < Tmax ? window : Tmax;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    SegWinParams p{};
    p.lp = lp; p.win_off = win_off; p.stride_n = stride_n; p.stride_t = stride_t; p.in_len = in_len;
    p.gt = gt; p.gt_stride = gt_stride; p.n_cols = n_cols; p.utt_begin = utt_begin; p.n_utts = n_utts;
    p.N = N; p.Tmax = Tmax; p.Cmax = Cmax; p.Kmax = Kmax; p.V = V; p.blank = blank; p.flags = flags;
    p.window = W; p.score_len = score_len; p.index_duration = index_duration;
    p.words = (int64_t)win_words(W);
    p.cols2 = reinterpret_cast<float *>(ws);            ws += pad256(P * ((size_t)gt_cols + 1) * (size_t)W * 4);
    p.bits = reinterpret_cast<uint32_t *>(ws);
    ws += (gt_cols == 1) ? pad256(P * (size_t)Cmax * win_words(W) * 4) : pad256(P * (size_t)Cmax * (size_t)W);
    p.offsets = reinterpret_cast<int32_t *>(ws);        ws += pad256(P * (size_t)Cmax * 4);
    p.term = reinterpret_cast<int32_t *>(ws);           ws += pad256(P * 4);
    int32_t *timing_scratch = reinterpret_cast<int32_t *>(ws); ws += pad256((size_t)N * Kmax * (size_t)Cmax * 4);
    float *cprob_scratch = reinterpret_cast<float *>(ws);
    p.seg_out = seg_out; p.term_t_out = term_t_out;
    p.timing = timing_out ? timing_out : timing_scratch;
    p.char_prob = char_prob_out ? char_prob_out : cprob_scratch;
    p.state_out = state_out; p.status_out = status_out;
    if (gt_cols == 1) {
        ctcseg_windowed_fill_kernel<<<(unsigned)P, kWinThreads, 0, st>>>(p);
        ++g_launch_count;
        ctcseg_windowed_backtrace_kernel<<<(unsigned)((P + 3) / 4), 128, 0, st>>>(p);
        ++g_launch_count;
    } else {
        SegMultiParams mp{};
        mp.w = p;
        mp.G = gt_cols;
        mp.ring = p.cols2;
        mp.codes = reinterpret_cast<uint8_t *>(p.bits);
        ctcseg_multi_fill_kernel<<<(unsigned)P, kWinThreads, 0, st>>>(mp);
        ++g_launch_count;
        ctcseg_multi_backtrace_kernel<<<(unsigned)((P + 3) / 4), 128, 0, st>>>(mp);
        ++g_launch_count;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
