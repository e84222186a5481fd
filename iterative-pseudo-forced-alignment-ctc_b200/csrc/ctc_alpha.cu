// ctc_alpha.cu -- kernel (1): batched CTC alpha recursion over the
// blank-interleaved 2L+1 lattice (window scoring, "acoustic CTC loss").
//
// Replaces torch.nn.functional.ctc_loss(lp, targets, in_len, tgt_len, blank,
// reduction='none') (ATen LossCTC.cpp::ctc_loss_cpu_template; SURVEY.md section 8(a)
// row A9), which BASELINE.json's north_star names as the CPU comparator.
//
// Layout: a window is owned by a group of WARPS warps.  The lattice is cut into
// (blank, label) state PAIRS; thread i of the group keeps P consecutive pairs
// -- states 2(iP+p), 2(iP+p)+1 -- in registers, in the log2 domain.  Per frame:
//   blank_p <- lse(blank_p, label_{p-1})                  + e[blank]
//   label_p <- lse(label_p, blank_p, skip_p?label_{p-1})  + e[label_p]
// so the only cross-thread dependency is ONE value (the previous thread's last
// label state): one __shfl_up per frame inside a warp, one shared-memory word +
// CTA barrier per frame across warps.  Emission columns arrive through the
// cp.async ring of emission_pipe.cuh, several frames ahead of the recursion.
// The T-serial chain is MUFU/latency bound (2 ex2 + 1 lg2 per label state,
// 1 + 1 per blank state), not HBM bound; see DESIGN.md.  The same template has a second,
// linear-domain instance (LIN, below: fp64 probabilities, per-thread power-of-two scales, an
// exactness guard) that runs first on dense panels of <= 256 pairs and hands the windows it
// cannot vouch for to this log-domain instance through a redo list.
#include "emission_pipe.cuh"
#include "lattice_shapes.cuh"

#include <type_traits>

namespace ipfa {

struct AlphaParams {
    const float *lp;
    int64_t stride_n, stride_t, stride_v;
    const int32_t *targets;
    int64_t tgt_stride;
    const int32_t *in_len;
    const int32_t *tgt_len;
    const int32_t *order;  // length bucket of this launch: windows order[0 .. *count) (null: all N)
    const int32_t *count;
    int N, V, blank;
    int halves;            // groups per window: 2 = meet-in-the-middle walk, 1 = forward only
    int pitch, tc;         // pipe geometry
    int u_cap;             // capacity of the per-group column list (gather mode)
    int l_cap;             // Lmax the launch was sized for
    size_t group_smem;     // bytes of shared memory per group
    float *nll_out;
    float *join_vec;       // [N][2 halves][2][l_cap + 1] state vectors at the cut (workspace)
    int *join_count;       // [N] arrivals at the cut, zeroed before the launch (workspace)
    int32_t *redo;         // linear-domain instance: windows handed to the log-domain instance
    int *redo_count;
};

// ---- linear-domain instance (LIN) -----------------------------------------------------------
// The log-domain recursion pays 4 MUFU operations (8 cycles each per warp) per state pair and
// frame.  The LIN instance keeps the states as PROBABILITIES in fp64 -- the B200 FP64 pipe issues
// a warp DADD / DMUL every 2 cycles (tools/microbench_fp64.cu) -- divided by the blank emission
// of every frame walked so far (a factor common to the whole lattice, accumulated as a sum of
// logs on the side), so that a pair costs 2 DADD + 1 DMUL + 1 select and no MUFU:
//   blank_p <- blank_p + label_{p-1}
//   label_p <- (label_p + (skip_p ? blank_p' : blank_p)) * r_p,   r = exp(lp[label] - lp[blank])
// Range: every thread carries its own power-of-two scale 2^E for its 2P states, re-chosen every
// second emission chunk (<= 64 frames) so that its largest state (or the value handed over by its
// left neighbour, if larger) sits at 2^kLinTarget; the value
// received from the left neighbour is rescaled by 2^(E - E_left), constant between two
// re-scalings.  Exactness guard: at every re-scaling each state the lattice can have reached
// (by the graph alone: frame index >= minimal arrival time) must hold a normal number
// >= 2^kLinTinyExp, nothing may exceed 2^kLinHugeExp, every emission ratio must be a normal
// fp32, and neighbouring scales must be within 2^1000 of each other.  A window that breaks any
// of these (or whose total is zero: infeasible targets) is appended to the redo list and scored
// by the log-domain instance right after -- the LIN instance never writes a result it cannot
// vouch for.
constexpr int kLinTarget = 100;
constexpr int kLinTinyExp = -700;
constexpr int kLinHugeExp = 1000;
constexpr int kLinEmpty = -(1 << 28);
constexpr int kLinRescaleChunks = 2;  // chunks (of <= 32 frames) between two re-scalings

__device__ __forceinline__ double lin_pow2(int d) {  // 2^d, d clamped to the normal range
    d = max(-1022, min(1023, d));
    return __hiloint2double((1023 + d) << 20, 0);
}
__device__ __forceinline__ int lin_exponent(double v) {  // floor(log2 v) of a positive normal v
    return (__double2hiint(v) >> 20) - 1023;
}

// PITCH > 0: compile-time panel pitch (dense rows of <= 32 symbols), so the unrolled frames
// address the panel with immediate offsets from one pointer per column.
//
// Two groups per window.  The recursion is T-serial, so a window is cut at its middle frame
// m = (T-1)/2 and walked from both ends at once (twice the resident warp-chains, half the
// chain length): group 0 runs alpha over frames 0..m; group 1 runs the SAME recursion over
// frames T-1..m+1 with the target reversed, which is beta with the emission of its own frame
// included (the blank-interleaved lattice and its skip rule are symmetric under reversal).
// Each group leaves its state vector in the workspace; the one that finishes second joins
// them:  p = sum_s alpha_m(s) * sum_{s' in succ(s)} b_{m+1}(s').
constexpr int kBidirMinFrames = 16;  // shorter windows are walked by group 0 alone

// LIN: bits of the emission ratio exp(x - x_blank) as an fp32, and their packing into the high
// word of the fp64 with the same value (mantissa rounded to 20 bits).  `ratio_range(bits)` is
// <= kRatioRangeMax exactly when the ratio is a normal fp32 <= 1e38 (zero, denormals, inf and NaN
// all map above it).
__device__ __forceinline__ uint32_t ratio_raw(const float x, const float xb) {
    return __float_as_uint(ex2_approx((x - xb) * kLog2e));
}
__device__ __forceinline__ uint32_t ratio_range(const uint32_t bits) { return bits - 0x00800000u; }
constexpr uint32_t kRatioRangeMax = 0x7e967699u - 0x00800000u;
__device__ __forceinline__ float ratio_pack(const uint32_t bits) {
    return __uint_as_float(((bits + 4u) >> 3) + 0x38000000u);
}

template <int P, int WARPS, bool DENSE, int PITCH, bool LIN = false>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32 * WARPS)
ctc_alpha_kernel(const AlphaParams prm) {
    static_assert(!LIN || (WARPS <= 2 && DENSE), "the linear-domain instance: one or two warps per half window, dense panel");
    using State = std::conditional_t<LIN, double, float>;
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    constexpr int NT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int group = (WARPS == 1) ? (threadIdx.x >> 5) : 0;
    const int tid = (WARPS == 1) ? (threadIdx.x & 31) : threadIdx.x;
    const int g2 = blockIdx.x * GROUPS + group;  // (window, half)
    // WARPS==1: whole warp leaves; WARPS>1: whole CTA leaves
    if (g2 >= prm.halves * (prm.count ? *prm.count : prm.N)) return;
    int w = (prm.halves == 2) ? g2 >> 1 : g2;
    const int half = (prm.halves == 2) ? g2 & 1 : 0;
    if (prm.order) w = prm.order[w];

    const int pitch = PITCH ? PITCH : prm.pitch;
    unsigned char *gsm = smem_raw + (size_t)group * prm.group_smem;
    float *ring = reinterpret_cast<float *>(gsm);
    float *xline = ring + (size_t)kStages * prm.tc * pitch;  // [2][NT + 1] neighbour exchange (WARPS > 1)
    float *fin = xline + 2 * (NT + 1);                        // [2 + WARPS]
    int *cols = reinterpret_cast<int *>(fin + 2 + WARPS);     // [u_cap] (gather mode)
    // LIN with two warps: fp64 exchange lines [2][NT + 1], scale exchange [NT + 1], scratch [NT]
    double *dline = reinterpret_cast<double *>(
        (reinterpret_cast<uintptr_t>(cols + prm.u_cap) + 7) & ~(uintptr_t)7);
    int *eline = reinterpret_cast<int *>(dline + 2 * (NT + 1));
    double *dscratch = reinterpret_cast<double *>(eline + NT + 2);

    const int T_all = prm.in_len[w];
    const int L = max(0, min(prm.tgt_len[w], prm.l_cap));
    const int32_t *tg = prm.targets + (int64_t)w * prm.tgt_stride;
    const int blank = prm.blank;

    if (T_all <= 0) {
        if (tid == 0 && half == 0) prm.nll_out[w] = (L == 0) ? 0.0f : __int_as_float(0x7f800000);
        return;
    }
    const bool bidir = prm.halves == 2 && T_all >= kBidirMinFrames;
    if (!bidir && half == 1) return;
    const int m = bidir ? (T_all - 1) >> 1 : T_all - 1;  // last frame of the forward half
    const bool rev = half == 1;
    const int t_lo = rev ? m + 1 : 0;
    const int T = rev ? T_all - 1 - m : m + 1;           // frames this group walks

    // per-thread lattice constants; the reverse half sees the target back to front
    auto target = [&](int j) { return rev ? tg[L - 1 - j] : tg[j]; };
    int col[P];      // panel column of label_p
    bool skip[P];    // s-2 transition allowed into label_p
    bool bad = false;
    bool flag = false;   // LIN: this window goes to the redo list
    int why = 0;         // ... and why (diagnostic bits, OR-ed into redo_count[1])
    int rep_excl = 0;    // LIN: repeated labels (target[k] == target[k-1]) before this thread's pairs
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;  // target index of this pair's label
        // States past the end of the target (j >= L) are left to run on garbage: they only
        // feed states further right, never a real one, and every value stays finite.
        const bool lab_ok = j < L;
        int lab = lab_ok ? target(j) : blank;
        if (lab < 0 || lab >= prm.V) { bad = true; lab = blank; }
        // (the blank column of the LIN panel holds raw logs, not ratios: a target that names
        // the blank symbol is left to the log-domain instance)
        if (LIN && lab_ok && lab == blank) { flag = true; why |= 1; }
        const int prev = (j >= 1 && lab_ok) ? target(j - 1) : -1;
        skip[p] = lab_ok && j >= 1 && prev != lab;
        if (LIN) rep_excl += (lab_ok && j >= 1 && prev == lab) ? 1 : 0;
        col[p] = DENSE ? lab : (lab_ok ? j + 1 : 0);
    }
    if constexpr (LIN) {  // exclusive prefix over the lanes (two warps: + the first warp's total)
        int incl = rep_excl;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if ((tid & 31) >= off) incl += v;
        }
        rep_excl = incl - rep_excl;
        if constexpr (WARPS > 1) {
            if (tid == 31) fin[0] = __int_as_float(incl);
            __syncthreads();
            if (tid >= 32) rep_excl += __float_as_int(fin[0]);
            __syncthreads();
        }
    }
    int colb = DENSE ? blank : 0;
    int U = L + 1;  // panel columns
    if constexpr (!DENSE) {
        for (int j = tid; j <= L; j += NT) {
            int c = (j == 0) ? blank : target(j - 1);
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
    if constexpr (WARPS > 1) {
        if (tid < 2) xline[tid * (NT + 1)] = kNegBig;  // left neighbour of thread 0: log(0)
        if (LIN && tid < 2) dline[tid * (NT + 1)] = 0.0;
        if (LIN && tid == 0) eline[0] = 0;
    }
    group_sync<WARPS>();

    EmissionPipe<WARPS, DENSE> pipe;
    pipe.init(ring, cols, prm.lp + (int64_t)w * prm.stride_n, prm.stride_t, T, U, prm.V, pitch,
              prm.tc, reinterpret_cast<uint64_t *>(gsm + prm.group_smem - 32), tid, t_lo, rev, prm.stride_v);
    pipe.prologue(tid);

    // LIN: first frame (walk index) at which the lattice can have reached blank_p / label_p by the
    // graph alone -- pair index + repeated labels so far (a repeat needs a blank in between);
    // kNever for the states past the target.
    constexpr int kNever = 0x7fffffff;
    int needb[P], needl[P];
    {
        int cnt = rep_excl;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j = tid * P + p;
            const int isrep = (j < L && j >= 1 && !skip[p]) ? 1 : 0;
            cnt += isrep;
            needb[p] = (j <= L) ? j + cnt - isrep : kNever;
            needl[p] = (j < L) ? j + cnt : kNever;
        }
    }
    int skipm[P];  // LIN: skip[p] as an all-ones / zero mask the compiler keeps in a register
#pragma unroll
    for (int p = 0; p < P; ++p) {
        skipm[p] = skip[p] ? -1 : 0;
        if constexpr (LIN) asm volatile("" : "+r"(skipm[p]));
    }
    State ab[P], al[P];  // blank / label alphas (log2 domain; LIN: scaled probabilities)
#pragma unroll
    for (int p = 0; p < P; ++p) { ab[p] = LIN ? State(0) : State(kNegBig); al[p] = ab[p]; }
    // LIN: stored = true * 2^E / prod(blank emissions so far); rs = 2^(E - E of the left lane)
    int E = 0;
    double rs = (tid == 0) ? 0.0 : 1.0;
    float sb = 0.0f;        // sum of the blank log-emissions since the last re-scaling (natural log)
    double sb_total = 0.0;  // ... and before it

    // One frame of the recursion, `off` floats past the column cursors.  `rd`/`wr`: this frame's
    // read / write lines of the cross-warp exchange (WARPS > 1); inside a warp the neighbour
    // comes by shuffle.
    const float *pb = nullptr;   // cursor on the blank column
    const float *pl[P];          // cursors on the label columns
    auto frame = [&](const int off, const float *rd, float *wr) {
        if constexpr (LIN) {
            if constexpr (PITCH != 32) sb += pb[off];  // (PITCH == 32: summed by the conversion)
            double r[P];  // the panel holds the high word of the fp64 ratio (20 mantissa bits)
#pragma unroll
            for (int p = 0; p < P; ++p) r[p] = __hiloint2double(__float_as_int(pl[p][off]), 0);
            // lane 0 has no left neighbour: its rs is 0
            double prev;
            if constexpr (WARPS > 1) prev = dline[(rd - xline) + tid] * rs;
            else prev = __shfl_up_sync(0xffffffffu, al[P - 1], 1) * rs;
#pragma unroll
            for (int p = P - 1; p >= 0; --p) {
                const double lm1 = (p == 0) ? prev : al[p - 1];
                const double nb = ab[p] + lm1;
                // x = skip ? nb : ab, as one bit-select per word on a mask held in a register
                const int m = skipm[p];
                const double x = __hiloint2double((__double2hiint(nb) & m) | (__double2hiint(ab[p]) & ~m),
                                                  (__double2loint(nb) & m) | (__double2loint(ab[p]) & ~m));
                al[p] = (al[p] + x) * r[p];
                ab[p] = nb;
            }
            if constexpr (WARPS > 1) {
                dline[(wr - xline) + tid + 1] = al[P - 1];
                __syncthreads();
            }
        } else {
            const float eb = pb[off];
            float el[P];
#pragma unroll
            for (int p = 0; p < P; ++p) el[p] = pl[p][off];
            float prev;
            if constexpr (WARPS > 1) {
                prev = rd[tid];
            } else {
                prev = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
                if (tid == 0) prev = kNegBig;
            }
#pragma unroll
            for (int p = P - 1; p >= 0; --p) {
                const float lm1 = (p == 0) ? prev : al[p - 1];
                // blank_p <- lse(blank_p, label_{p-1});  label_p <- lse(label_p, blank_p [, label_{p-1}])
                // and lse(blank_p, label_{p-1}) is shared between the two when the skip is allowed.
                const float nb = lse2_2(ab[p], lm1);
                const float x = skip[p] ? nb : ab[p];
                al[p] = lse2_2(al[p], x) + el[p];
                ab[p] = nb + eb;
            }
            if constexpr (WARPS > 1) {
                wr[tid + 1] = al[P - 1];
                __syncthreads();
            }
        }
    };
    // LIN: re-scaling + exactness guard, between two frames; `tcur` = walk index of the last frame
    // done.  Pair j's label is first alive at frame j + (repeats up to j), its blank one frame after
    // the previous label.
    // (two warps: `rdl` = offset of the exchange line the next frame reads, i.e. the one the last
    // frame wrote; it is rewritten in the new scales)
    auto renorm = [&](const int tcur, const int rdl) {
        if constexpr (LIN) {
            sb_total += (double)sb;
            sb = 0.0f;
            // The states are non-negative, so their order is the order of their high words as
            // integers (a NaN or inf sorts above every finite value and trips the `huge` test; a
            // negative value -- arithmetic on a flagged window -- sorts below `tiny`).
            constexpr int tiny_hi = (1023 + kLinTinyExp) << 20, huge_hi = (1023 + kLinHugeExp) << 20;
            int hmax = 0;
            bool ok = true;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                // the states past the target never feed a real one: their high word is cleared
                // (what is left is a denormal, i.e. nothing) so that they stay out of the maximum
                int hb = __double2hiint(ab[p]), hl = __double2hiint(al[p]);
                if (needb[p] == kNever) { hb = 0; ab[p] = __hiloint2double(0, __double2loint(ab[p])); }
                if (needl[p] == kNever) { hl = 0; al[p] = __hiloint2double(0, __double2loint(al[p])); }
                if (tcur >= needb[p]) ok = ok && (hb >= tiny_hi);
                if (tcur >= needl[p]) ok = ok && (hl >= tiny_hi);
                hmax = max(hmax, max(hb, hl));
            }
            if (!ok) why |= 4;
            if (!(hmax < huge_hi)) why |= 8;
            ok = ok && (hmax < huge_hi);
            // true exponents of this thread's largest state and of the value its left neighbour hands over
            const int A = (hmax >= (1 << 20)) ? (hmax >> 20) - 1023 - E : kLinEmpty;
            double b;
            int El;
            if constexpr (WARPS > 1) {
                eline[tid + 1] = E;
                __syncthreads();
                b = dline[rdl + tid];
                El = eline[tid];
            } else {
                b = __shfl_up_sync(0xffffffffu, al[P - 1], 1);
                El = __shfl_up_sync(0xffffffffu, E, 1);
            }
            // (the line holds what the last frame wrote: a left neighbour past the target has not
            // been cleared there)
            const int B = (tid > 0 && tid * P - 1 < L && __double2hiint(b) >= (1 << 20)) ? lin_exponent(b) - El
                                                                                       : kLinEmpty;
            const int X = max(A, B);
            const bool empty = X <= kLinEmpty / 2;
            int Enew = kLinTarget - X;
            // the lanes right of the frontier take the scale of the frontier lane (the lanes that
            // hold something are a prefix of the group)
            int Ead;
            if constexpr (WARPS > 1) {
                int *etmp = reinterpret_cast<int *>(dscratch);
                etmp[tid] = Enew;
                const int n_holding = __syncthreads_count(!empty);
                Ead = etmp[max(n_holding - 1, 0)];
            } else {
                const unsigned ne = __ballot_sync(0xffffffffu, !empty);
                Ead = __shfl_sync(0xffffffffu, Enew, ne ? 31 - __clz(ne) : 0);
            }
            if (empty) Enew = Ead;
            const int d = Enew - E;
            if (!empty && (d > 1000 || d < -1000)) { ok = false; why |= 16; }
            const double f = lin_pow2(max(-1000, min(1000, d)));
#pragma unroll
            for (int p = 0; p < P; ++p) { ab[p] *= f; al[p] *= f; }
            E = Enew;
            int d2;
            if constexpr (WARPS > 1) {
                __syncthreads();  // everybody has read the old scales and the old line
                eline[tid + 1] = E;
                dline[rdl + tid + 1] = al[P - 1];
                __syncthreads();
                d2 = E - eline[tid];
            } else {
                d2 = E - __shfl_up_sync(0xffffffffu, E, 1);
            }
            if (tid > 0 && !empty && (d2 > 1000 || d2 < -1000)) { ok = false; why |= 32; }
            rs = (tid == 0) ? 0.0 : lin_pow2(max(-1000, min(1000, d2)));
            flag = flag || !ok;
        }
    };
    float *line0 = xline, *line1 = xline + NT + 1;
    // The j-th frame a chunk processes reads line[(j-1)&1] and writes line[j&1] (tc is even, so
    // j has the parity of the frame's position in the walk).  DIR = +1: rows 0, 1, ...;
    // DIR = -1 (reverse half): rows rows-1, rows-2, ...; the 4 unrolled frames sit at immediate
    // offsets from the cursors either way.
    // LIN: an emission becomes the RATIO exp(x - x_blank), stored as the high word of its fp64
    // (rounded to 20 mantissa bits); the blank column keeps its raw logs, which the conversion
    // also sums on the side (PITCH == 32; frame() does it for the generic pitch).  A ratio that is
    // not a normal fp32 <= 1e38 (zero, denormal, inf, NaN) sends the window to the redo list.
    const int lane = tid & 31, wrp = tid >> 5;
    const bool keep = lane == colb;  // PITCH == 32: one lane per column; this lane owns the blank column
    // PITCH == 32: largest |x - x_blank| this lane has converted; the ratio is a normal fp32 exactly
    // when it is <= kMaxLogRatio (one FMNMX per emission; a NaN slips through here and is caught as
    // a state above 2^kLinHugeExp, its packed ratio being 2^128)
    float dmax = 0.0f;
    constexpr float kMaxLogRatio = 87.0f;
    // whole chunk at once (8 rows in flight: the loads of a batch precede its stores)
    auto prescale_chunk = [&](float *panel, const int rows) {
        if constexpr (LIN && PITCH == 32) {
            // (two warps: rows wrp, wrp + 2, ...; each warp sums the blank logs of its own rows)
            float *cell = panel + lane;
            const float *bcell = panel + colb;
            for (int r0 = wrp; r0 < rows; r0 += 8 * WARPS) {
                float xv[8], bv[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int rr = min(r0 + k * WARPS, rows - 1);
                    xv[k] = cell[rr * 32];
                    bv[k] = bcell[rr * 32];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const bool in = r0 + k * WARPS < rows;
                    if (in) sb += bv[k];
                    const float d = xv[k] - bv[k];
                    if (in) dmax = fmaxf(dmax, fabsf(d));
                    const uint32_t raw = __float_as_uint(ex2_approx(d * kLog2e));
                    if (!keep && in) cell[(r0 + k * WARPS) * 32] = ratio_pack(raw);
                }
            }
        } else if constexpr (LIN) {
            float4 *p4 = reinterpret_cast<float4 *>(panel);
            const int n4 = (rows * pitch) >> 2;
            for (int q = tid; q < n4; q += NT) {
                const int e0 = q << 2, row = e0 / pitch, c0 = e0 - row * pitch;
                const float xb = panel[row * pitch + colb];
                float4 v = p4[q];
                const uint32_t b0 = ratio_raw(v.x, xb), b1 = ratio_raw(v.y, xb);
                const uint32_t b2 = ratio_raw(v.z, xb), b3 = ratio_raw(v.w, xb);
                if (c0 != colb) { v.x = ratio_pack(b0); flag = flag || (ratio_range(b0) > kRatioRangeMax && c0 < prm.V); }
                if (c0 + 1 != colb) { v.y = ratio_pack(b1); flag = flag || (ratio_range(b1) > kRatioRangeMax && c0 + 1 < prm.V); }
                if (c0 + 2 != colb) { v.z = ratio_pack(b2); flag = flag || (ratio_range(b2) > kRatioRangeMax && c0 + 2 < prm.V); }
                if (c0 + 3 != colb) { v.w = ratio_pack(b3); flag = flag || (ratio_range(b3) > kRatioRangeMax && c0 + 3 < prm.V); }
                p4[q] = v;
            }
        }
    };
    // PRE (LIN, PITCH == 32): while walking a chunk, every group of 4 frames also converts 4 rows
    // of the NEXT chunk (cursors qx / qb) -- independent work that fills the recursion's stalls,
    // instead of a latency-bound conversion pass that all the warps of a sub-partition hit together.
    float *qx = nullptr;
    const float *qb = nullptr;
    auto run_rows = [&](auto mode, int j, const int rows) {  // mode: +-1 walk, +-2 walk and convert
        constexpr int DIR = decltype(mode)::value > 0 ? 1 : -1;
        constexpr bool PRE = decltype(mode)::value == 2 || decltype(mode)::value == -2;
        const int step = DIR * pitch;
        auto bump = [&](const int frames) {
            pb += frames * step;
#pragma unroll
            for (int p = 0; p < P; ++p) pl[p] += frames * step;
        };
        if ((j & 1) && j < rows) { frame(0, line0, line1); bump(1); ++j; }
        for (; j + 3 < rows; j += 4) {
            // the 4 rows of the next chunk that go with these 4 frames: 4 / WARPS per warp (the
            // cursors of warp w start w rows in)
            constexpr int RPW = PRE ? 4 / WARPS : 1;
            uint32_t raw[RPW];
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    const float b = qb[i * WARPS * step];
                    sb += b;
                    const float d = qx[i * WARPS * step] - b;
                    dmax = fmaxf(dmax, fabsf(d));
                    raw[i] = __float_as_uint(ex2_approx(d * kLog2e));
                }
            }
            frame(0, line1, line0);
            frame(step, line0, line1);
            frame(2 * step, line1, line0);
            frame(3 * step, line0, line1);
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < RPW; ++i) {
                    if (!keep) qx[i * WARPS * step] = ratio_pack(raw[i]);
                }
                qx += 4 * step;
                qb += 4 * step;
            }
            bump(4);
        }
        for (; j + 1 < rows; j += 2) {
            frame(0, line1, line0);
            frame(step, line0, line1);
            bump(2);
        }
        if (j < rows) frame(0, line1, line0);
    };
    // LIN with bulk-copied chunks of PITCH == 32: chunk c+1 is converted while chunk c is walked
    const bool lin_overlap = LIN && PITCH == 32 && pipe.bulk;

    int last_par = 0;  // parity of the exchange line the last frame wrote
    for (int chunk = 0; chunk < pipe.nchunks; ++chunk) {
        float *panel;
        int rows;
        bool pre = false;  // LIN: this walk converts the next chunk
        if constexpr (!LIN) {
            panel = const_cast<float *>(pipe.acquire(chunk, tid));
            rows = pipe.chunk_rows(chunk);
        } else if (lin_overlap) {
            rows = pipe.chunk_rows(chunk);
            if (chunk == 0) {
                panel = const_cast<float *>(pipe.acquire(0, tid));
                prescale_chunk(panel, rows);
            } else {
                // chunk `chunk` is converted already; the stage of chunk-1 is free: refill it
                panel = pipe.stage_ptr(chunk);
                fence_proxy_async();
                group_sync<WARPS>();
                pipe.issue(chunk + kStages - 1, tid);
            }
            if (chunk + 1 < pipe.nchunks) {
                pipe.wait_landed(chunk + 1);
                const int rows_next = pipe.chunk_rows(chunk + 1);
                pre = chunk > 0 && rows_next == rows && (rows & 3) == 0;
                if (!pre) prescale_chunk(pipe.stage_ptr(chunk + 1), rows_next);
            }
            group_sync<WARPS>();
        } else {
            panel = const_cast<float *>(pipe.acquire(chunk, tid));
            rows = pipe.chunk_rows(chunk);
        }
        if constexpr (LIN) {
            if (!lin_overlap) {
                prescale_chunk(panel, rows);
                group_sync<WARPS>();
            }
        } else {
            // in-place: natural log -> log2, clamp log(0) to the finite stand-in
            float4 *p4 = reinterpret_cast<float4 *>(panel);
            const int n4 = (rows * pitch) >> 2;
            for (int q = tid; q < n4; q += NT) {
                float4 v = p4[q];
                v.x = fmaxf(v.x * kLog2e, kNegBig); v.y = fmaxf(v.y * kLog2e, kNegBig);
                v.z = fmaxf(v.z * kLog2e, kNegBig); v.w = fmaxf(v.w * kLog2e, kNegBig);
                p4[q] = v;
            }
            // (an odd pitch -- vocabulary-major gather -- can leave up to three elements behind the last float4)
            for (int q = 4 * n4 + tid; q < rows * pitch; q += NT) panel[q] = fmaxf(panel[q] * kLog2e, kNegBig);
            group_sync<WARPS>();
        }
        const int first_row = rev ? rows - 1 : 0;  // the chunk's first frame in walking order
        int j = 0;
        if (chunk == 0) {  // first frame of the walk: only states 0 and 1 are alive
            if constexpr (LIN) {
                if (tid == 0) {
                    ab[0] = 1.0;
                    if (L > 0) al[0] = __hiloint2double(__float_as_int(panel[first_row * pitch + col[0]]), 0);
                }
                if constexpr (PITCH != 32) sb = panel[first_row * pitch + colb];
            } else if (tid == 0) {
                ab[0] = panel[first_row * pitch + colb];
                if (L > 0) al[0] = panel[first_row * pitch + col[0]];
            }
            if constexpr (WARPS > 1) {
                if constexpr (LIN) dline[tid + 1] = al[P - 1];
                else line0[tid + 1] = al[P - 1];
                __syncthreads();
            }
            j = 1;
        }
        // re-scaling + guard: after the first frame, then every kLinRescaleChunks-th chunk
        if (chunk % kLinRescaleChunks == 0) renorm(chunk * pipe.tc + j - 1, ((j - 1) & 1) ? NT + 1 : 0);
        last_par = (rows - 1) & 1;
        const int row = rev ? rows - 1 - j : j;
        pb = panel + row * pitch + colb;
#pragma unroll
        for (int p = 0; p < P; ++p) pl[p] = panel + row * pitch + col[p];
        if constexpr (LIN && PITCH == 32) {
            if (pre) {  // j == 0 here: the 4-frame groups cover the whole chunk
                float *nxt = pipe.stage_ptr(chunk + 1) + (rev ? row - wrp : row + wrp) * pitch;
                qx = nxt + lane;
                qb = nxt + colb;
                if (rev) run_rows(std::integral_constant<int, -2>{}, j, rows);
                else run_rows(std::integral_constant<int, 2>{}, j, rows);
                group_sync<WARPS>();
            }
        }
        if (!pre) {
            if (rev) run_rows(std::integral_constant<int, -1>{}, j, rows);
            else run_rows(std::integral_constant<int, 1>{}, j, rows);
        }
    }

    // state -> log2 of its true value (LIN: undo the scale and the blank normalisation)
    bool flag_any = false;
    double lin_bias = 0.0;
    if constexpr (LIN) {
        if (!(dmax <= kMaxLogRatio) && !keep && lane < prm.V) { flag = true; why |= 2; }
        renorm(T - 1, last_par ? NT + 1 : 0);
        if constexpr (WARPS > 1) {
            flag_any = __syncthreads_or(flag) != 0;
            if constexpr (PITCH == 32) {  // each warp summed the blank logs of its own rows
                dscratch[tid] = sb_total;
                __syncthreads();
                sb_total = 0.0;
#pragma unroll
                for (int q = 0; q < WARPS; ++q) sb_total += dscratch[lane + 32 * q];
            }
        } else {
            flag_any = __any_sync(0xffffffffu, flag);
        }
        if (flag) atomicOr(prm.redo_count + 1, why ? why : 64);
        lin_bias = sb_total * 1.4426950408889634 - (double)E;
    }
    auto log2_of = [&](const State v) -> float {
        if constexpr (LIN) {
            if (!(v >= 2.2250738585072014e-308)) return kNegBig;
            const int hi = __double2hiint(v);
            const double mant = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(v));
            return (float)((double)((hi >> 20) - 1023) + lin_bias + (double)lg2_approx((float)mant));
        } else {
            return v;
        }
    };
    // LIN: hand the window to the log-domain instance instead of writing a result
    auto redo_window = [&]() {
        prm.redo[atomicAdd(prm.redo_count, 1)] = w;
        prm.join_count[w] = 0;
    };
    if (!bidir) {
        // final states 2L (blank of pair L) and 2L-1 (label of pair L-1)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j = tid * P + p;
            if (j == L) fin[0] = log2_of(ab[p]);
            if (j == L - 1) fin[1] = log2_of(al[p]);
        }
        if (L == 0 && tid == 0) fin[1] = kNegBig;
        group_sync<WARPS>();
        if (tid == 0) {
            const float v = lse2_2(fin[0], fin[1]);
            float nll = -v * kLn2;
            if (v < kNegThreshold || bad) nll = __int_as_float(0x7f800000);
            if (LIN && !bad && (flag_any || !(v >= kNegThreshold))) redo_window();
            else prm.nll_out[w] = nll;
        }
        return;
    }

    // ---- join the two halves ------------------------------------------------
    // vec[w][half][0][j] = blank of pair j, vec[w][half][1][j] = label of pair j, j = 0..l_cap,
    // each in its own half's pair numbering.
    const int64_t vstride = prm.l_cap + 1;
    float *vec_w = prm.join_vec + (int64_t)w * 4 * vstride;
    float *mine = vec_w + (int64_t)half * 2 * vstride;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;
        if (j <= L) {
            mine[j] = log2_of(ab[p]);
            mine[vstride + j] = (j < L) ? log2_of(al[p]) : kNegBig;
        }
    }
    __threadfence();
    group_sync<WARPS>();
    // (LIN: bit 8 of the arrival counter carries the first half's redo flag to the joiner)
    if (tid == 0) fin[0] = __int_as_float(atomicAdd(prm.join_count + w, flag_any ? 257 : 1));
    group_sync<WARPS>();
    if ((__float_as_int(fin[0]) & 255) == 0) return;  // the other half is still walking; it will join
    flag_any = flag_any || (__float_as_int(fin[0]) >> 8) != 0;
    __threadfence();
    // In forward numbering: A = alpha_m, B = b_{m+1}.  Reverse pair i holds the blank of forward
    // pair L - i and the label of forward pair L - 1 - i.
    const float *fwd = vec_w, *bwd = vec_w + 2 * vstride;
    bool staged = false;
    if constexpr (LIN) {
        // both halves' vectors (4 x (l_cap + 1) floats <= 4 KB) come through the emission ring,
        // idle by now: independent loads, one L2 round trip instead of one per dependent step
        // (unless a tuning override shrank the ring below that)
        const int nvec = 4 * (int)vstride;
        staged = nvec <= kStages * prm.tc * pitch;
        if (staged) {
            float *stage = ring;
            for (int i = tid; i < nvec; i += NT) stage[i] = __ldcg(vec_w + i);
            group_sync<WARPS>();
            fwd = stage;
            bwd = stage + 2 * vstride;
        }
    }
    auto ld = [&](const float *q) { return staged ? *q : __ldcg(q); };
    auto A_b = [&](int j) { return ld(fwd + j); };
    auto A_l = [&](int j) { return ld(fwd + vstride + j); };
    auto B_b = [&](int j) { return ld(bwd + (L - j)); };
    auto B_l = [&](int j) { return ld(bwd + vstride + (L - 1 - j)); };
    float acc = kNegBig;
    for (int j = tid; j <= L; j += NT) {
        // blank j -> {blank j, label j}
        float succ = B_b(j);
        if (j < L) succ = lse2_2(succ, B_l(j));
        acc = lse2_2(acc, A_b(j) + succ);
        if (j < L) {  // label j -> {label j, blank j+1, label j+1 if it differs}
            float s2 = lse2_2(B_l(j), B_b(j + 1));
            if (j + 1 < L && tg[j + 1] != tg[j]) s2 = lse2_2(s2, B_l(j + 1));
            acc = lse2_2(acc, A_l(j) + s2);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = lse2_2(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if constexpr (WARPS > 1) {
        if ((tid & 31) == 0) fin[2 + (tid >> 5)] = acc;
        __syncthreads();
        if (tid == 0) {
            for (int q = 1; q < WARPS; ++q) acc = lse2_2(acc, fin[2 + q]);
        }
    }
    if (tid == 0) {
        float nll = -acc * kLn2;
        if (acc < kNegThreshold || bad) nll = __int_as_float(0x7f800000);
        if (LIN && !bad && (flag_any || !(acc >= kNegThreshold))) redo_window();
        else prm.nll_out[w] = nll;
    }
}

}  // namespace ipfa

#include "ctc_alpha_f32.cuh"

namespace ipfa {

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;

template <int P, int WARPS, bool DENSE, int PITCH, bool LIN = false>
static int launch_alpha_p(AlphaParams prm, int Lmax, cudaStream_t stream) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    const int U = PITCH ? PITCH : (DENSE ? prm.V : (Lmax + 1));
    // shared-memory budget per group for the emission ring
    const size_t budget = (WARPS == 1) ? (16 * 1024) : (160 * 1024);
    PipeGeometry g = pipe_geometry(U, budget, !DENSE && prm.stride_v > prm.stride_t);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = DENSE ? 0 : ((Lmax + 1 + 3) & ~3);
    prm.l_cap = Lmax;
    size_t group_smem = g.ring_bytes + (2 * (32 * WARPS + 1) + 2 + WARPS) * sizeof(float) + (size_t)prm.u_cap * sizeof(int) + 40;
    if (LIN && WARPS > 1)  // fp64 exchange lines, scale line, scratch (+ alignment)
        group_smem += 2 * (32 * WARPS + 1) * sizeof(double) + (32 * WARPS + 2) * sizeof(int) + 32 * WARPS * sizeof(double) + 16;
    group_smem = (group_smem + 15) & ~(size_t)15;
    prm.group_smem = group_smem;
    const size_t smem = group_smem * GROUPS;
    auto kern = ctc_alpha_kernel<P, WARPS, DENSE, PITCH, LIN>;
    // (one attribute call per instance and device while the request does not grow: it costs a
    // microsecond of host time per launch, which shows next to a 60 us kernel)
    static size_t smem_set[16] = {0};
    int dev_id = 0;
    cudaError_t e = cudaGetDevice(&dev_id);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    if (dev_id < 0 || dev_id >= 16 || smem > smem_set[dev_id]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
        if (dev_id >= 0 && dev_id < 16) smem_set[dev_id] = smem;
    }
    const int threads = (WARPS == 1) ? 128 : 32 * WARPS;
    const int blocks = (prm.halves * prm.N + GROUPS - 1) / GROUPS;
    const int prof_slot = profile_begin(stream);
    kern<<<blocks, threads, smem, stream>>>(prm);
    profile_end(prof_slot, stream);
    ++g_launch_count;
    e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

template <int P, int WARPS, bool DENSE>
static int launch_alpha(const AlphaParams &prm, int Lmax, cudaStream_t stream) {
    // (ptxas 12.9 segfaults on the widest instance with the compile-time pitch; it keeps the
    // runtime pitch -- 8192 state pairs over a 32-symbol vocabulary is not a real workload)
    if constexpr (DENSE && !(P == 8 && WARPS == 32)) {
        if (prm.V <= 32) return launch_alpha_p<P, WARPS, true, 32>(prm, Lmax, stream);
    }
    return launch_alpha_p<P, WARPS, DENSE, 0>(prm, Lmax, stream);
}

template <bool DENSE>
static int dispatch_alpha(const AlphaParams &prm, int Lmax, LatticeShape s, cudaStream_t stream) {
#define IPFA_X(P_, W_) \
    if (s.PER == P_ && s.WARPS == W_) return launch_alpha<P_, W_, DENSE>(prm, Lmax, stream);
    IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    return IPFA_ERR_UNSUPPORTED;
}

// linear-domain instance: W (1 or 2) warps per half window, P pairs per lane
static int launch_alpha_lin(const AlphaParams &prm, int Lmax, int P, int W, cudaStream_t stream) {
#define IPFA_LIN(P_, W_)                                                                      \
    if (P == P_ && W == W_) {                                                                 \
        if (prm.V <= 32) return launch_alpha_p<P_, W_, true, 32, true>(prm, Lmax, stream);    \
        return launch_alpha_p<P_, W_, true, 0, true>(prm, Lmax, stream);                      \
    }
    IPFA_LIN(1, 1) IPFA_LIN(2, 1) IPFA_LIN(4, 1) IPFA_LIN(8, 1)
    IPFA_LIN(1, 2) IPFA_LIN(2, 2) IPFA_LIN(4, 2)
#undef IPFA_LIN
    return IPFA_ERR_UNSUPPORTED;
}
// fp32 tier (ctc_alpha_f32.cuh): one warp per half window, P pairs per lane, V <= 32
template <int P>
static int launch_alpha_f32_p(AlphaParams prm, int Lmax, cudaStream_t stream) {
    PipeGeometry g = pipe_geometry(32, 16 * 1024);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = 0;
    prm.l_cap = Lmax;
    prm.group_smem = (g.ring_bytes + 16 + 32 + 15) & ~(size_t)15;
    const size_t smem = prm.group_smem * 4;
    auto kern = ctc_alpha_f32_kernel<P>;
    static size_t smem_set[16] = {0};
    int dev_id = 0;
    cudaError_t e = cudaGetDevice(&dev_id);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    if (dev_id < 0 || dev_id >= 16 || smem > smem_set[dev_id]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
        if (dev_id >= 0 && dev_id < 16) smem_set[dev_id] = smem;
    }
    const int blocks = (prm.halves * prm.N + 3) / 4;
    const int prof_slot = profile_begin(stream);
    kern<<<blocks, 128, smem, stream>>>(prm);
    profile_end(prof_slot, stream);
    ++g_launch_count;
    e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}
static int launch_alpha_f32(const AlphaParams &prm, int Lmax, int P, cudaStream_t stream) {
    switch (P) {
        case 1: return launch_alpha_f32_p<1>(prm, Lmax, stream);
        case 2: return launch_alpha_f32_p<2>(prm, Lmax, stream);
        case 4: return launch_alpha_f32_p<4>(prm, Lmax, stream);
        case 8: return launch_alpha_f32_p<8>(prm, Lmax, stream);
    }
    return IPFA_ERR_UNSUPPORTED;
}
constexpr int kLinMaxPairs = 256;
// pairs per lane / warps per half window for `units` state pairs: one warp per half window.
// The two-warp instances (fp64 exchange line + CTA barrier per frame, as in the log-domain
// instance) double the resident chains but measured slower on BASELINE configs[1] (113.6 us
// against 106.4 us, profiles/r01_alpha_lin_parity_timing.txt); they stay reachable through
// IPFA_ALPHA_LIN_SHAPE=P,W (tuning).
static void lin_shape(int units, long long chains, int *P, int *W) {
    *W = 1;
    *P = units <= 32 ? 1 : units <= 64 ? 2 : units <= 128 ? 4 : 8;
    (void)chains;
    if (const char *e = tuning("IPFA_ALPHA_LIN_SHAPE")) {
        int p = 0, w = 0;
        if (sscanf(e, "%d,%d", &p, &w) == 2 && (w == 1 || w == 2) && (p == 1 || p == 2 || p == 4 || p == 8) &&
            !(p == 8 && w == 2) && 32 * p * w >= units) { *P = p; *W = w; }
    }
}

bool use_dense_panel(int V, int Lmax) { return V <= 64 || V <= 2 * (Lmax + 1); }

}  // namespace ipfa

using namespace ipfa;

static inline size_t alpha_pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// arrival counters [N] + the two halves' state vectors at the cut [N][2][2][Lmax + 1]
extern "C" size_t ipfa_ctc_alpha_workspace_bytes(int N, int, int Lmax, int) {
    const size_t n = (size_t)(N > 0 ? N : 1), l1 = (size_t)(Lmax > 0 ? Lmax : 0) + 1;
    // arrival counters (+ the two tiers' redo counters and reason bits right behind them: one memset),
    // join vectors, length-bucket lists [2][N] + their counters, the two redo lists [N]
    return alpha_pad256((n + 4) * sizeof(int32_t)) + alpha_pad256(n * 4 * l1 * sizeof(float)) +
           alpha_pad256(n * 2 * sizeof(int32_t)) + 256 + 2 * alpha_pad256(n * sizeof(int32_t)) + 256;
}

extern "C" int ipfa_ctc_alpha_device(const float *lp, int64_t stride_n, int64_t stride_t,
                                     const int32_t *targets, int64_t tgt_stride,
                                     const int32_t *in_len, const int32_t *tgt_len, int N, int Tmax,
                                     int Lmax, int V, int blank, float *nll_out, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    return ipfa_ctc_alpha_strided_device(lp, stride_n, stride_t, 1, targets, tgt_stride, in_len, tgt_len, N, Tmax,
                                         Lmax, V, blank, nll_out, workspace, workspace_bytes, stream);
}

extern "C" int ipfa_ctc_alpha_strided_device(const float *lp, int64_t stride_n, int64_t stride_t, int64_t stride_v,
                                             const int32_t *targets, int64_t tgt_stride,
                                             const int32_t *in_len, const int32_t *tgt_len, int N, int Tmax,
                                             int Lmax, int V, int blank, float *nll_out, void *workspace,
                                             size_t workspace_bytes, void *stream) {
    (void)Tmax;
    NvtxRange range("ipfa.ctc_alpha (window scorer: tiers + redo lists)");
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !tgt_len || !nll_out || N < 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V ||
        (Lmax > 0 && !targets) || stride_v < 1)
        return IPFA_ERR_INVALID_ARG;
    if (!workspace) return IPFA_ERR_INVALID_ARG;
    if (workspace_bytes < ipfa_ctc_alpha_workspace_bytes(N, Tmax, Lmax, V)) return IPFA_ERR_WORKSPACE;
    // Dense panels: two groups per window (the halves of the meet-in-the-middle walk), and the
    // MUFU pipe wants ~6 resident warp-chains per SM sub-partition more than it wants units per
    // thread.  The gather panel (large vocabularies) is bound by its loads, not by the chain:
    // one group per window.
    // (rows that are not contiguous over the vocabulary -- stride_v != 1 -- always go through the gather panel)
    const bool dense = stride_v == 1 && use_dense_panel(V, Lmax);
    const int halves = dense ? 2 : 1;
    LatticeShape s;
    if (!pick_lattice_shape(Lmax + 1, halves * N, &s, "IPFA_ALPHA_SHAPE", 6)) return IPFA_ERR_UNSUPPORTED;
    AlphaParams prm{};
    prm.halves = halves;
    prm.lp = lp; prm.stride_n = stride_n; prm.stride_t = stride_t; prm.stride_v = stride_v;
    prm.targets = targets; prm.tgt_stride = tgt_stride;
    prm.in_len = in_len; prm.tgt_len = tgt_len; prm.order = nullptr; prm.count = nullptr;
    prm.N = N; prm.V = V; prm.blank = blank; prm.nll_out = nll_out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    prm.join_count = static_cast<int *>(workspace);
    prm.join_vec = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) +
                                             alpha_pad256(((size_t)N + 4) * sizeof(int32_t)));
    cudaError_t e = cudaMemsetAsync(prm.join_count, 0, ((size_t)N + 4) * sizeof(int32_t), st);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    // Dense panels whose lattice fits one warp: the linear-domain instance scores the windows and
    // lists the ones it cannot vouch for; the log-domain instance below then runs over that list.
    const bool lin = dense && Lmax + 1 <= kLinMaxPairs && !tuning("IPFA_ALPHA_LOG");
    const size_t l1 = (size_t)Lmax + 1;
    int32_t *order = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(prm.join_vec) +
                                                 alpha_pad256((size_t)N * 4 * l1 * sizeof(float)));
    int32_t *count = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(order) +
                                                 alpha_pad256((size_t)N * 2 * sizeof(int32_t)));
    int32_t *redo = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(count) + 256);
    int32_t *redo_f32 = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(redo) +
                                                    alpha_pad256((size_t)N * sizeof(int32_t)));
    if (lin) {
        int P = 0, W = 0;
        lin_shape(Lmax + 1, (long long)halves * N, &P, &W);
        // Tiers: fp32 linear-domain (V <= 32) -> fp64 linear-domain over the windows the first could
        // not vouch for -> log-domain over the windows the second could not vouch for.  The lists
        // and their lengths stay on the device; a tier's groups beyond its list's length leave at once.
        // (the fp32 tier is OFF unless IPFA_ALPHA_F32=1: on BASELINE configs[1] the states of one lane
        // lie 100-150 binades apart, fp32 cannot hold that next to the growth between two re-scalings,
        // 58 % of the windows are handed over and the step gets slower -- ctc_alpha_f32.cuh)
        const bool f32 = V <= 32 && W == 1 && !tuning("IPFA_ALPHA_LIN_SHAPE") &&
                         tuning("IPFA_ALPHA_F32") && tuning("IPFA_ALPHA_F32")[0] == '1';
        const bool buckets = N >= kBucketMinWindows && W == 1 && P >= 2 && !tuning("IPFA_NO_BUCKETS");
        if (buckets) {
            e = cudaMemsetAsync(count, 0, 2 * sizeof(int32_t), st);
            if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
            length_bucket_kernel<<<(N + 255) / 256, 256, 0, st>>>(tgt_len, N, 32 * P / 2, order, count);
            ++g_launch_count;
        }
        // first tier over every window (in two length buckets when the batch is large)
        prm.redo = f32 ? redo_f32 : redo;
        prm.redo_count = prm.join_count + N + (f32 ? 2 : 0);
        for (int cls = 0; cls < (buckets ? 2 : 1); ++cls) {
            if (buckets) { prm.order = order + (int64_t)cls * N; prm.count = count + cls; }
            const int Pc = (buckets && cls == 0) ? P / 2 : P;
            const int rc = f32 ? launch_alpha_f32(prm, Lmax, Pc, st) : launch_alpha_lin(prm, Lmax, Pc, W, st);
            if (rc) return rc;
        }
        if (f32) {  // second tier: the fp64 instance over the first tier's list
            prm.order = redo_f32;
            prm.count = prm.join_count + N + 2;
            prm.redo = redo;
            prm.redo_count = prm.join_count + N;
            const int rc = launch_alpha_lin(prm, Lmax, P, 1, st);
            if (rc) return rc;
        }
        prm.order = redo;
        prm.count = prm.join_count + N;
        prm.redo = nullptr;
        prm.redo_count = nullptr;
        if (dense) return dispatch_alpha<true>(prm, Lmax, s, st);
    }
    // two length buckets when the batch is large and a half-width instance exists (decided on the
    // device, see length_bucket_kernel)
    const int big_units = 32 * s.WARPS * s.PER, small_units = big_units / 2;
    LatticeShape s_small;
    const bool bucketed = N >= kBucketMinWindows && small_units >= 32 && !tuning("IPFA_NO_BUCKETS") &&
                          pick_lattice_shape(small_units, halves * N, &s_small, "IPFA_ALPHA_SMALL_SHAPE", 6) &&
                          32 * s_small.WARPS * s_small.PER == small_units;
    if (!bucketed) {
        if (dense) return dispatch_alpha<true>(prm, Lmax, s, st);
        return dispatch_alpha<false>(prm, Lmax, s, st);
    }
    e = cudaMemsetAsync(count, 0, 2 * sizeof(int32_t), st);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    length_bucket_kernel<<<(N + 255) / 256, 256, 0, st>>>(tgt_len, N, small_units, order, count);
    ++g_launch_count;
    for (int cls = 0; cls < 2; ++cls) {
        prm.order = order + (int64_t)cls * N;
        prm.count = count + cls;
        const LatticeShape sh = cls == 0 ? s_small : s;
        const int rc = dense ? dispatch_alpha<true>(prm, Lmax, sh, st) : dispatch_alpha<false>(prm, Lmax, sh, st);
        if (rc) return rc;
    }
    return IPFA_OK;
}
