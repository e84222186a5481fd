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

inline bool shape_exists(int per, int warps) {
#define IPFA_X(P_, W_) if (per == P_ && warps == W_) return true;
    IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    return false;
}

// units: lattice units the widest window needs; n_windows: batch size;
// target_warps: resident warps wanted on the whole GPU before widening PER.
inline bool pick_lattice_shape(int units, int n_windows, LatticeShape *out, const char *env_name,
                               int warps_per_smsp = 3) {
    if (const char *e = getenv(env_name)) {  // tuning override "PER,WARPS"
        int p = 0, w = 0;
        if (sscanf(e, "%d,%d", &p, &w) == 2 && shape_exists(p, w) && 32 * w * p >= units) {
            out->PER = p; out->WARPS = w;
            return true;
        }
    }
    // Candidates: every instance wide enough and at most 2x more padded than the tightest one.
    // Among those that put >= warps_per_smsp warps on every SM sub-partition (measured: 3 is
    // enough for the dense-panel kernels, the gather panel wants 6 to cover its load latency)
    // take the least padded (ties: more units per thread = fewer barriers/shuffles per unit);
    // when the batch is too small for that, take the one with the most warps.
    static const LatticeShape all[] = {
#define IPFA_X(P_, W_) {P_, W_},
        IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    };
    long long min_pad = -1;
    for (const auto &c : all) {
        const long long pad = 32LL * c.WARPS * c.PER;
        if (pad >= units && (min_pad < 0 || pad < min_pad)) min_pad = pad;
    }
    if (min_pad < 0) return false;
    const long long want_warps = 148LL * 4 * warps_per_smsp;
    LatticeShape best{0, 0};
    bool best_ok = false;
    long long best_pad = 0;
    for (const auto &c : all) {
        const long long pad = 32LL * c.WARPS * c.PER;
        if (pad < units || pad > 2 * min_pad) continue;
        const bool ok = (long long)n_windows * c.WARPS >= want_warps;
        bool take;
        if (best.PER == 0) take = true;
        else if (ok != best_ok) take = ok;
        else if (ok) take = pad < best_pad || (pad == best_pad && c.PER > best.PER);
        else take = c.WARPS > best.WARPS || (c.WARPS == best.WARPS && pad < best_pad);
        if (take) { best = c; best_ok = ok; best_pad = pad; }
    }
    *out = best;
    return true;
}

}  // namespace ipfa
