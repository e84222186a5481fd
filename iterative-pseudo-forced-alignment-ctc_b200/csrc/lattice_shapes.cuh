// lattice_shapes.cuh -- (states-per-thread, warps-per-window) instances shared by
// the three lattice kernels, and the heuristic that picks one.
//
// A window's lattice is spread over 32*WARPS threads holding `PER` units each
// (a unit = one (blank,label) pair for the ctc lattices, one column for ctcseg).
// The recursion is T-serial, so throughput comes from resident warps: when the
// batch is small relative to the GPU (few windows per SM) the lattice is spread
// over MORE warps with fewer units per thread; when the batch is large the
// widest-per-thread shape that fits is used (fewest barriers and shuffles).
#pragma once
#include <stdio.h>
#include <stdlib.h>

namespace ipfa {

struct LatticeShape {
    int PER, WARPS;
};

// X(PER, WARPS)
#define IPFA_FOR_EACH_SHAPE(X)                                                        \
    X(1, 1) X(1, 2) X(1, 4) X(1, 8) X(1, 16)                                          \
    X(2, 1) X(2, 2) X(2, 4) X(2, 8) X(2, 16)                                          \
    X(4, 1) X(4, 2) X(4, 4) X(4, 8) X(4, 16)                                          \
    X(8, 1) X(8, 2) X(8, 4) X(8, 8) X(8, 16) X(8, 32)

// Extra (non power-of-two) widths, instantiated by the CTC-segmentation fill only: the anchor
// sweep sizes one launch for the widest window in flight, and 2x padding costs ~30% there.
#define IPFA_FOR_EACH_EXTRA_SHAPE(X)                                                  \
    X(2, 6) X(2, 10) X(2, 12) X(2, 20) X(2, 24)                                       \
    X(4, 6) X(4, 10) X(4, 12) X(4, 20) X(4, 24)

inline bool shape_exists(int per, int warps, bool extra = false) {
#define IPFA_X(P_, W_) if (per == P_ && warps == W_) return true;
    IPFA_FOR_EACH_SHAPE(IPFA_X)
    if (extra) { IPFA_FOR_EACH_EXTRA_SHAPE(IPFA_X) }
#undef IPFA_X
    return false;
}

// units: lattice units the widest window needs; n_windows: batch size;
// target_warps: resident warps wanted on the whole GPU before widening PER.
inline bool pick_lattice_shape(int units, int n_windows, LatticeShape *out, const char *env_name,
                               int warps_per_smsp = 3, bool extra = false) {
    if (const char *e = tuning(env_name)) {  // tuning override "PER,WARPS"
        int p = 0, w = 0;
        if (sscanf(e, "%d,%d", &p, &w) == 2 && shape_exists(p, w, extra) && 32 * w * p >= units) {
            out->PER = p; out->WARPS = w;
            return true;
        }
    }
    // Candidates: every instance wide enough and at most 2x more padded than the tightest one.
    // Among those that put >= warps_per_smsp warps on every SM sub-partition (measured: 3 is
    // enough for the dense-panel kernels, the gather panel wants 6 to cover its load latency)
    // take the least padded (ties: more units per thread = fewer barriers/shuffles per unit);
    // when the batch is too small for that, take the one with the most warps.
    static const LatticeShape base_shapes[] = {
#define IPFA_X(P_, W_) {P_, W_},
        IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    };
    static const LatticeShape extra_shapes[] = {
#define IPFA_X(P_, W_) {P_, W_},
        IPFA_FOR_EACH_EXTRA_SHAPE(IPFA_X)
#undef IPFA_X
    };
    LatticeShape all[sizeof(base_shapes) / sizeof(base_shapes[0]) + sizeof(extra_shapes) / sizeof(extra_shapes[0])];
    int n_all = 0;
    for (const auto &c : base_shapes) all[n_all++] = c;
    if (extra)
        for (const auto &c : extra_shapes) all[n_all++] = c;
    long long min_pad = -1;
    for (int i = 0; i < n_all; ++i) {
        const long long pad = 32LL * all[i].WARPS * all[i].PER;
        if (pad >= units && (min_pad < 0 || pad < min_pad)) min_pad = pad;
    }
    if (min_pad < 0) return false;
    const long long want_warps = 148LL * 4 * warps_per_smsp;
    // Batch too small to fill the sub-partitions (latency regime; measured on the anchor sweep,
    // 21-201 windows in flight): time follows the padded width, 4 units per thread cost ~5%
    // and 8 units ~20% over 1-2 units.
    auto small_batch_cost = [](const LatticeShape &c) {
        const double pad = 32.0 * c.WARPS * c.PER;
        return pad * (c.PER >= 8 ? 1.2 : c.PER >= 4 ? 1.05 : 1.0);
    };
    LatticeShape best{0, 0};
    bool best_ok = false;
    long long best_pad = 0;
    for (int i = 0; i < n_all; ++i) {
        const LatticeShape c = all[i];
        const long long pad = 32LL * c.WARPS * c.PER;
        if (pad < units || pad > 2 * min_pad) continue;
        const bool ok = (long long)n_windows * c.WARPS >= want_warps;
        bool take;
        if (best.PER == 0) take = true;
        else if (ok != best_ok) take = ok;
        else if (ok) take = pad < best_pad || (pad == best_pad && c.PER > best.PER);
        else take = small_batch_cost(c) < small_batch_cost(best) ||
                    (small_batch_cost(c) == small_batch_cost(best) && c.WARPS > best.WARPS);
        if (take) { best = c; best_ok = ok; best_pad = pad; }
    }
    *out = best;
    return true;
}

}  // namespace ipfa
