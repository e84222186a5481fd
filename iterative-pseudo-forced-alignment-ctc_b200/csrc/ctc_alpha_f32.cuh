// ctc_alpha_f32.cuh -- an optional tier in FRONT of the window scorer's fp64 linear-domain instance:
// the same recursion (ctc_alpha.cu, LIN) with fp32 states.  Included by ctc_alpha.cu (same AlphaParams,
// same join vectors, same redo-list protocol).  OFF by default (IPFA_ALPHA_F32=1 turns it on): it is
// exact where it answers, but on BASELINE configs[1] it answers for 42 % of the windows only -- see
// "Range" below and DESIGN.md 5.1.2 -- so the step is slower with it than without.
//
// Why a tier in front of the fp64 one: on B200 a warp DADD / DMUL occupies the issue port of an SM
// sub-partition for two cycles and an FADD / FMUL for one (tools/microbench_fp64.cu;
// tools/microbench_f32x2.cu: the packed f32x2 forms also take two, so packing buys nothing), and the
// fp64 states cost 8 LOP3 + 4 MOV per frame for the skip select and the zero low words.  In fp32
// a pair is FADD, FSEL, FADD, FMUL: 23 issue slots per frame of 4 pairs instead of 47.
//
//   blank_p <- blank_p + label_{p-1}
//   label_p <- (label_p + (skip_p ? blank_p' : blank_p)) * r_p,   r = exp(lp[label] - lp[blank])
//
// What fp32 changes is the RANGE, not the algorithm: 254 binades instead of 2046.  The per-thread
// power-of-two scale is re-chosen every kF32Rescale = 8 frames (the lane's largest state back to
// 2^kF32Target), and the exactness guard (once per chunk) is the LIN guard with fp32 bounds: every state the lattice
// can have reached must hold at least 2^kF32TinyExp, nothing may exceed 2^kF32HugeExp (inf / NaN sort
// above it), every emission ratio must lie within e^+-kF32MaxLogRatio, neighbouring lanes' scales
// within 2^kF32MaxStep.  A window that breaks any of these is appended to the redo list and the
// fp64 tier scores it right after (which in turn hands what IT cannot vouch for to the log-domain
// instance): peaked emissions whose states grow by more than ~2^126 within 8 frames, -inf
// emissions, infeasible targets, targets that name the blank.  This tier never writes a result it
// cannot vouch for.
//
// Range, measured (tools/exp_alpha_tiers.py and a float64 simulation of the lattice): on random
// emissions with T = 1000, L = 100 the reachable states of ONE lane (4 pairs) lie 100-150 binades apart
// -- the trailing states of the lattice keep the weight of paths that stayed behind for hundreds of
// frames -- and a lane's maximum grows by up to ~90 binades within 8 frames.  254 binades do not hold
// both, whatever the target; the fp64 tier's 2046 do.
//
// Instance: one warp per half window (meet-in-the-middle walk as in ctc_alpha.cu), P pairs per lane,
// dense panel of pitch 32 (V <= 32), four half windows per 128-thread CTA, no CTA barrier.
#pragma once

namespace ipfa {

// A lane's largest state is put at 2^kF32Target every kF32Rescale frames; until the next re-scaling it
// may grow by 2^(127 - kF32Target), and the lane's other reachable states may lie
// 2^(kF32Target - kF32TinyExp) below it.  Both need room: a state the lattice has just reached gains
// paths combinatorially fast (C(t, j) of them, ~2^6 per frame with the emission ratios on random
// emissions), and the states of one lane drift apart by the random walks of their labels' ratios
// (~2^60 after 500 frames).  Re-scaling is the cheap part (every 8 frames); the guard proper -- every
// reachable state >= 2^kF32TinyExp, nothing >= 2^kF32HugeExp -- runs once per emission chunk: an
// overflow in between is sticky (inf / NaN survive the recursion and the re-scaling).
constexpr int kF32Target = 50;
constexpr int kF32TinyExp = -120;
constexpr int kF32HugeExp = 120;
constexpr int kF32MaxStep = 126;
#ifndef IPFA_F32_RESCALE
#define IPFA_F32_RESCALE 8
#endif
constexpr int kF32Rescale = IPFA_F32_RESCALE;  // frames between two re-scalings (divides the chunk length)
constexpr float kF32MaxLogRatio = 60.0f;  // |lp[label] - lp[blank]| beyond this goes to the fp64 tier
constexpr int kF32Empty = -(1 << 28);

__device__ __forceinline__ float f32_pow2(int d) {  // 2^d, d within the normal range
    return __int_as_float((127 + d) << 23);
}

template <int P>
__global__ void __launch_bounds__(128)
ctc_alpha_f32_kernel(const AlphaParams prm) {
    constexpr int PITCH = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int group = threadIdx.x >> 5;
    const int tid = threadIdx.x & 31;
    const int g2 = blockIdx.x * 4 + group;  // (window, half)
    if (g2 >= prm.halves * (prm.count ? *prm.count : prm.N)) return;
    int w = (prm.halves == 2) ? g2 >> 1 : g2;
    const int half = (prm.halves == 2) ? g2 & 1 : 0;
    if (prm.order) w = prm.order[w];

    unsigned char *gsm = smem_raw + (size_t)group * prm.group_smem;
    float *ring = reinterpret_cast<float *>(gsm);
    float *fin = ring + (size_t)kStages * prm.tc * PITCH;  // [2]

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
    const int m = bidir ? (T_all - 1) >> 1 : T_all - 1;
    const bool rev = half == 1;
    const int t_lo = rev ? m + 1 : 0;
    const int T = rev ? T_all - 1 - m : m + 1;

    auto target = [&](int j) { return rev ? tg[L - 1 - j] : tg[j]; };
    int col[P];
    bool skip[P];
    bool bad = false, flag = false;
    int why = 0, rep_excl = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;
        const bool lab_ok = j < L;
        int lab = lab_ok ? target(j) : blank;
        if (lab < 0 || lab >= prm.V) { bad = true; lab = blank; }
        if (lab_ok && lab == blank) { flag = true; why |= 1; }  // the blank column holds raw logs
        const int prev = (j >= 1 && lab_ok) ? target(j - 1) : -1;
        skip[p] = lab_ok && j >= 1 && prev != lab;
        rep_excl += (lab_ok && j >= 1 && prev == lab) ? 1 : 0;
        col[p] = lab;
    }
    {   // exclusive prefix of the repeat counts over the lanes
        int incl = rep_excl;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (tid >= off) incl += v;
        }
        rep_excl = incl - rep_excl;
    }
    const int colb = blank;

    EmissionPipe<1, true> pipe;
    pipe.init(ring, nullptr, prm.lp + (int64_t)w * prm.stride_n, prm.stride_t, T, L + 1, prm.V, PITCH, prm.tc,
              reinterpret_cast<uint64_t *>(gsm + prm.group_smem - 32), tid, t_lo, rev);
    pipe.prologue(tid);

    // first frame (walk index) at which the graph alone lets the lattice reach blank_p / label_p
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
    float ab[P], al[P];  // stored = true * 2^E / prod(blank emissions so far)
#pragma unroll
    for (int p = 0; p < P; ++p) { ab[p] = 0.0f; al[p] = 0.0f; }
    int E = 0;
    float rs = (tid == 0) ? 0.0f : 1.0f;  // 2^(E - E of the left lane); lane 0 has no left neighbour
    float sb = 0.0f;                      // blank log-emissions since the last re-scaling (natural log)
    double sb_total = 0.0;

    const float *pl[P];
    auto frame = [&](const int off) {
        float r[P];
#pragma unroll
        for (int p = 0; p < P; ++p) r[p] = pl[p][off];
        const float prev = __shfl_up_sync(0xffffffffu, al[P - 1], 1) * rs;
#pragma unroll
        for (int p = P - 1; p >= 0; --p) {
            const float lm1 = (p == 0) ? prev : al[p - 1];
            const float nb = ab[p] + lm1;
            const float x = skip[p] ? nb : ab[p];
            al[p] = (al[p] + x) * r[p];
            ab[p] = nb;
        }
    };
    // Re-scaling between two frames: the lane's largest state (or the value its left neighbour hands
    // over, if larger) goes to 2^kF32Target.  Exponent fields are compared as they are (biased).
    auto rescale = [&]() {
        float mx = fmaxf(ab[0], al[0]);
#pragma unroll
        for (int p = 1; p < P; ++p) mx = fmaxf(mx, fmaxf(ab[p], al[p]));
        const int mb = __float_as_int(mx);
        const int bb = __float_as_int(__shfl_up_sync(0xffffffffu, al[P - 1], 1));
        const int El = __shfl_up_sync(0xffffffffu, E, 1);
        const int A = (mb >= (1 << 23)) ? (mb >> 23) - E : kF32Empty;
        const int B = (tid > 0 && bb >= (1 << 23)) ? (bb >> 23) - El : kF32Empty;
        const int X = max(A, B);
        const bool empty = X <= kF32Empty / 2;
        int Enew = (kF32Target + 127) - X;
        // lanes right of the frontier take the frontier lane's scale (the holders are a prefix)
        const unsigned ne = __ballot_sync(0xffffffffu, !empty);
        const int Ead = __shfl_sync(0xffffffffu, Enew, ne ? 31 - __clz(ne) : 0);
        if (empty) Enew = Ead;
        const int d = Enew - E;
        const float f = f32_pow2(max(-kF32MaxStep, min(kF32MaxStep, d)));
#pragma unroll
        for (int p = 0; p < P; ++p) { ab[p] *= f; al[p] *= f; }
        E = Enew;
        const int d2 = E - __shfl_up_sync(0xffffffffu, E, 1);
        rs = (tid == 0) ? 0.0f : f32_pow2(max(-kF32MaxStep, min(kF32MaxStep, d2)));
        if (!empty && (abs(d) > kF32MaxStep || (tid > 0 && abs(d2) > kF32MaxStep))) { flag = true; why |= 16; }
    };
    // Exactness guard, right after a re-scaling; tcur = walk index of the last frame done.
    // Non-negative floats order like their bit patterns (inf / NaN above every finite value; a negative
    // value -- arithmetic on a flagged window -- below `tiny`).
    auto guard = [&](const int tcur) {
        sb_total += (double)sb;
        sb = 0.0f;
        constexpr int tiny_bits = (127 + kF32TinyExp) << 23, huge_bits = (127 + kF32HugeExp) << 23;
        bool ok = true;
        int hmax = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int hb = __float_as_int(ab[p]), hl = __float_as_int(al[p]);
            if (tcur >= needb[p]) ok = ok && (hb >= tiny_bits);
            if (tcur >= needl[p]) ok = ok && (hl >= tiny_bits);
            hmax = max(hmax, max(hb, hl));
        }
        if (!ok) why |= 4;
        if (!(hmax < huge_bits)) { ok = false; why |= 8; }
        flag = flag || !ok;
    };

    // An emission becomes the RATIO exp(x - x_blank) (plain fp32, in place), one lane per column.  The
    // blank column becomes 0 (its lane adds -inf to the exponent): nothing reads it as a ratio except
    // the states past the end of the target, whose column cursor points there -- they stay exactly 0
    // and out of the maxima.  The blank logs themselves are summed on the side.
    const float lane_bias = (tid == colb) ? __int_as_float(0xff800000) : 0.0f;
    float dmax = 0.0f;
    auto prescale_chunk = [&](float *panel, const int rows) {
        float *cell = panel + tid;
        const float *bcell = panel + colb;
        for (int r0 = 0; r0 < rows; r0 += 8) {
            float xv[8], bv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int rr = min(r0 + k, rows - 1);
                xv[k] = cell[rr * 32];
                bv[k] = bcell[rr * 32];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool in = r0 + k < rows;
                if (in) sb += bv[k];
                const float d = xv[k] - bv[k];
                if (in) dmax = fmaxf(dmax, fabsf(d));
                const float ratio = ex2_approx(__fmaf_rn(d, kLog2e, lane_bias));
                if (in) cell[(r0 + k) * 32] = ratio;
            }
        }
    };
    // PRE: while walking a chunk, every group of 4 frames also converts 4 rows of the NEXT chunk --
    // independent work in the shadow of the recursion's dependent instructions.
    float *qx = nullptr;
    const float *qb = nullptr;
    auto run_rows = [&](auto mode, int j, const int rows) {  // mode: +-1 walk, +-2 walk and convert
        constexpr int DIR = decltype(mode)::value > 0 ? 1 : -1;
        constexpr bool PRE = decltype(mode)::value == 2 || decltype(mode)::value == -2;
        constexpr int step = DIR * PITCH;
        auto bump = [&](const int frames) {
#pragma unroll
            for (int p = 0; p < P; ++p) pl[p] += frames * step;
        };
        for (; j + 3 < rows; j += 4) {
            float ratio[4];
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float b = qb[i * step];
                    sb += b;
                    const float d = qx[i * step] - b;
                    dmax = fmaxf(dmax, fabsf(d));
                    ratio[i] = ex2_approx(__fmaf_rn(d, kLog2e, lane_bias));
                }
            }
            frame(0);
            frame(step);
            frame(2 * step);
            frame(3 * step);
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) qx[i * step] = ratio[i];
                qx += 4 * step;
                qb += 4 * step;
            }
            bump(4);
        }
        for (; j < rows; ++j) { frame(0); bump(1); }
    };
    const bool overlap = pipe.bulk;

    for (int chunk = 0; chunk < pipe.nchunks; ++chunk) {
        float *panel;
        const int rows = pipe.chunk_rows(chunk);
        bool pre = false;
        if (overlap) {
            if (chunk == 0) {
                panel = const_cast<float *>(pipe.acquire(0, tid));
                prescale_chunk(panel, rows);
            } else {
                // chunk `chunk` is converted already; the stage of chunk-1 is free: refill it
                panel = pipe.stage_ptr(chunk);
                fence_proxy_async();
                __syncwarp();
                pipe.issue(chunk + kStages - 1, tid);
            }
            if (chunk + 1 < pipe.nchunks) {
                pipe.wait_landed(chunk + 1);
                const int rows_next = pipe.chunk_rows(chunk + 1);
                pre = chunk > 0 && rows_next == rows && (rows % kF32Rescale) == 0;
                if (!pre) prescale_chunk(pipe.stage_ptr(chunk + 1), rows_next);
            }
            __syncwarp();
        } else {
            panel = const_cast<float *>(pipe.acquire(chunk, tid));
            prescale_chunk(panel, rows);
            __syncwarp();
        }
        const int first_row = rev ? rows - 1 : 0;
        int j = 0;
        if (chunk == 0) {  // first frame of the walk: only states 0 and 1 are alive
            if (tid == 0) {
                ab[0] = 1.0f;
                if (L > 0) al[0] = panel[first_row * PITCH + col[0]];
            }
            j = 1;
            rescale();
        }
        // the chunk in pieces of kF32Rescale frames, a re-scaling in front of every piece but the
        // first of the walk, the guard in front of every chunk
        for (int piece = 0; piece < rows; piece += kF32Rescale) {
            const int end = min(rows, piece + kF32Rescale);
            if (piece > 0 || chunk > 0) rescale();
            if (piece == 0 && chunk > 0) guard(chunk * pipe.tc - 1);
            const int start = max(j, piece);
            const int row = rev ? rows - 1 - start : start;
#pragma unroll
            for (int p = 0; p < P; ++p) pl[p] = panel + row * PITCH + col[p];
            if (pre) {  // whole pieces of 16 frames: the 4-frame groups cover them
                float *nxt = pipe.stage_ptr(chunk + 1) + row * PITCH;
                qx = nxt + tid;
                qb = nxt + colb;
                if (rev) run_rows(std::integral_constant<int, -2>{}, start, end);
                else run_rows(std::integral_constant<int, 2>{}, start, end);
            } else {
                if (rev) run_rows(std::integral_constant<int, -1>{}, start, end);
                else run_rows(std::integral_constant<int, 1>{}, start, end);
            }
        }
        if (pre) __syncwarp();
    }

    if (!(dmax <= kF32MaxLogRatio) && tid < prm.V) { flag = true; why |= 2; }
    rescale();
    guard(T - 1);
    const bool flag_own = __any_sync(0xffffffffu, flag);
    if (flag) atomicOr(prm.redo_count + 1, why ? why : 64);
    const double lin_bias = sb_total * 1.4426950408889634 - (double)E;
    auto log2_of = [&](const float v) -> float {
        if (!(v >= 1.17549435e-38f)) return kNegBig;
        const int bits = __float_as_int(v);
        const float mant = __int_as_float((bits & 0x007fffff) | 0x3f800000);
        return (float)((double)((bits >> 23) - 127) + lin_bias + (double)lg2_approx(mant));
    };
    auto redo_window = [&]() {
        prm.redo[atomicAdd(prm.redo_count, 1)] = w;
        prm.join_count[w] = 0;
    };
    if (!bidir) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int j = tid * P + p;
            if (j == L) fin[0] = log2_of(ab[p]);
            if (j == L - 1) fin[1] = log2_of(al[p]);
        }
        if (L == 0 && tid == 0) fin[1] = kNegBig;
        __syncwarp();
        if (tid == 0) {
            const float v = lse2_2(fin[0], fin[1]);
            float nll = -v * kLn2;
            if (v < kNegThreshold || bad) nll = __int_as_float(0x7f800000);
            if (!bad && (flag_own || !(v >= kNegThreshold))) redo_window();
            else prm.nll_out[w] = nll;
        }
        return;
    }

    // ---- join the two halves (as in ctc_alpha_kernel) -------------------------------------------
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
    __syncwarp();
    int arrivals = 0;
    if (tid == 0) arrivals = atomicAdd(prm.join_count + w, flag_own ? 257 : 1);
    arrivals = __shfl_sync(0xffffffffu, arrivals, 0);
    if ((arrivals & 255) == 0) return;  // the other half is still walking; it will join
    const bool flag_any = flag_own || (arrivals >> 8) != 0;
    __threadfence();
    const float *fwd = vec_w, *bwd = vec_w + 2 * vstride;
    const int nvec = 4 * (int)vstride;
    const bool staged = nvec <= kStages * prm.tc * PITCH;
    if (staged) {  // both halves' vectors through the idle emission ring: one L2 round trip
        float *stage = ring;
        for (int i = tid; i < nvec; i += 32) stage[i] = __ldcg(vec_w + i);
        __syncwarp();
        fwd = stage;
        bwd = stage + 2 * vstride;
    }
    auto ld = [&](const float *q) { return staged ? *q : __ldcg(q); };
    auto A_b = [&](int j) { return ld(fwd + j); };
    auto A_l = [&](int j) { return ld(fwd + vstride + j); };
    auto B_b = [&](int j) { return ld(bwd + (L - j)); };
    auto B_l = [&](int j) { return ld(bwd + vstride + (L - 1 - j)); };
    float acc = kNegBig;
    for (int j = tid; j <= L; j += 32) {
        float succ = B_b(j);
        if (j < L) succ = lse2_2(succ, B_l(j));
        acc = lse2_2(acc, A_b(j) + succ);
        if (j < L) {
            float s2 = lse2_2(B_l(j), B_b(j + 1));
            if (j + 1 < L && tg[j + 1] != tg[j]) s2 = lse2_2(s2, B_l(j + 1));
            acc = lse2_2(acc, A_l(j) + s2);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = lse2_2(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (tid == 0) {
        float nll = -acc * kLn2;
        if (acc < kNegThreshold || bad) nll = __int_as_float(0x7f800000);
        if (!bad && (flag_any || !(acc >= kNegThreshold))) redo_window();
        else prm.nll_out[w] = nll;
    }
}

}  // namespace ipfa
