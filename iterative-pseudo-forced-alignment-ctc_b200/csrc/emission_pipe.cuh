// emission_pipe.cuh -- time-chunked staging of one window's emission columns
// into shared memory (cp.async ring), shared by the alpha / Viterbi / ctcseg
// lattice kernels.
//
// A "group" is the set of threads that owns one window: a single warp
// (WARPS == 1, several windows per CTA, no CTA barrier anywhere) or the whole
// CTA (WARPS > 1).  Two panel layouts:
//   DENSE  : the panel row is the full vocabulary row lp[t, 0:V] (V small).
//            When the rows of a chunk are contiguous in global AND shared memory
//            (stride_t == V == pitch, 16-byte aligned) the whole chunk moves with ONE
//            bulk async copy (TMA engine, cp.async.bulk -> SASS UBLKCP) issued by one
//            thread and completed on an mbarrier; otherwise 16-byte / 4-byte cp.async.
//   gather : the panel row holds only the window's own columns,
//            panel[t][j] = lp[t, cols[j]] with cols[0] = blank; 4-byte cp.async.
#pragma once
#include <stdlib.h>

#include "ipfa_common.cuh"

namespace ipfa {

constexpr int kStages = 3;

template <int WARPS>
__device__ __forceinline__ void group_sync() {
    if constexpr (WARPS == 1) {
        __syncwarp();
    } else {
        __syncthreads();
    }
}


// Gather mode: sort + dedup a window's column list so that the lanes of one gather
// instruction walk ascending addresses (a warp touches a handful of 128-byte lines of the
// vocabulary row instead of 32 -- the L1 wavefront count is what bounds the V = 5000 gather).
// In : cols[0..U) raw column ids (>= 0).  scratch: 3*U ints (the emission ring, not yet in use).
// Out: cols[0..U') ascending and unique; scratch[2U + j] = panel position of raw column j.
// Returns U'.  All threads of the group call it; ends with a group barrier.
template <int WARPS>
__device__ __forceinline__ int sort_unique_columns(int *cols, int U, int *scratch, int tid) {
    constexpr int NT = 32 * WARPS;
    int *raw = scratch, *first = scratch + U, *pos = scratch + 2 * U;
    for (int j = tid; j < U; j += NT) raw[j] = cols[j];
    group_sync<WARPS>();
    for (int j = tid; j < U; j += NT) {
        const int v = raw[j];
        int f = 1;
        for (int k = 0; k < j; ++k)
            if (raw[k] == v) { f = 0; break; }
        first[j] = f;
    }
    group_sync<WARPS>();
    int n_unique = 0;
    for (int k = 0; k < U; ++k) n_unique += first[k];
    for (int j = tid; j < U; j += NT) {
        const int v = raw[j];
        int r = 0;
        for (int k = 0; k < U; ++k) r += (first[k] != 0) & (raw[k] < v);
        pos[j] = r;
        if (first[j]) cols[r] = v;
    }
    group_sync<WARPS>();
    return n_unique;
}

template <int WARPS, bool DENSE>
struct EmissionPipe {
    static constexpr int NT = 32 * WARPS;
    float *ring;          // [kStages][tc][pitch]
    const int *cols;      // smem column list (gather mode)
    const float *base;    // &lp[w, 0, 0]
    int64_t stride_t;
    int64_t stride_v;     // gather mode: elements between two vocabulary entries of one frame (1: rows contiguous)
    int T, U, V, pitch, tc, nchunks;
    int t_lo;             // first frame of the range this pipe walks
    bool rev;             // chunks (and the caller's rows) run from the last frame down
    bool vec16, bulk;
    uint64_t *bars;       // [kStages] mbarriers (bulk mode)

    // All threads of the group call init(); it ends with a group barrier.
    __device__ __forceinline__ void init(float *ring_, const int *cols_, const float *base_,
                                         int64_t stride_t_, int T_, int U_, int V_, int pitch_, int tc_,
                                         uint64_t *bars_, int tid, int t_lo_ = 0, bool rev_ = false,
                                         int64_t stride_v_ = 1) {
        // T_ frames starting at t_lo_; rev_: logical chunk 0 holds the LAST tc frames of the range
        ring = ring_; cols = cols_; base = base_; stride_t = stride_t_; stride_v = stride_v_;
        T = T_; U = U_; V = V_; pitch = pitch_; tc = tc_; bars = bars_; t_lo = t_lo_; rev = rev_;
        nchunks = (T + tc - 1) / tc;
        vec16 = DENSE && (V % 4 == 0) && (stride_t % 4 == 0) &&
                ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
        bulk = vec16 && stride_t == V && pitch == V;
        if (bulk) {
            if (tid == 0) {
#pragma unroll
                for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        group_sync<WARPS>();
    }

    __device__ __forceinline__ int chunk_rows(int chunk) const { return min(tc, T - chunk * tc); }
    // first (lowest) frame of a chunk; smem row r of the chunk holds frame chunk_t0 + r
    __device__ __forceinline__ int chunk_t0(int chunk) const {
        return rev ? t_lo + T - chunk * tc - chunk_rows(chunk) : t_lo + chunk * tc;
    }

    __device__ __forceinline__ float *stage_ptr(int chunk) const {
        return ring + (size_t)(chunk % kStages) * tc * pitch;
    }

    // All threads of the group call this; always followed by a commit.
    __device__ __forceinline__ void issue(int chunk, int tid) {
        if (chunk < nchunks) {
            float *dst = stage_ptr(chunk);
            const int t0 = chunk_t0(chunk);
            const int rows = chunk_rows(chunk);
            if constexpr (DENSE) {
                if (bulk) {
                    if (tid == 0) {
                        const uint32_t bytes = (uint32_t)rows * (uint32_t)V * 4u;
                        uint64_t *bar = &bars[chunk % kStages];
                        mbar_expect_tx(bar, bytes);
                        bulk_copy_g2s(dst, base + (int64_t)t0 * stride_t, bytes, bar);
                    }
                } else if (vec16) {
                    const int v4 = V >> 2;
                    const int pieces = rows * v4;
                    for (int q = tid; q < pieces; q += NT) {
                        const int r = q / v4, c4 = q - r * v4;
                        cp_async_16(dst + r * pitch + c4 * 4, base + (int64_t)(t0 + r) * stride_t + c4 * 4);
                    }
                } else {
                    const int total = rows * V;
                    for (int q = tid; q < total; q += NT) {
                        const int r = q / V, c = q - r * V;
                        cp_async_4(dst + r * pitch + c, base + (int64_t)(t0 + r) * stride_t + c);
                    }
                }
            } else if (stride_v > stride_t) {
                // vocabulary-major emissions ([V, T] per window): the frames of one column are contiguous,
                // so the lanes of a warp walk TIME -- one 128-byte run per (column, chunk) -- and the
                // panel (frame-major, odd pitch: conflict-free) is filled column by column
                const int lane = tid & 31;
                for (int j = tid >> 5; j < U; j += WARPS) {
                    const float *src = base + (int64_t)cols[j] * stride_v + (int64_t)t0 * stride_t;
                    for (int r = lane; r < rows; r += 32) cp_async_4(dst + r * pitch + j, src + (int64_t)r * stride_t);
                }
            } else {
                for (int r = 0; r < rows; ++r) {
                    const float *src = base + (int64_t)(t0 + r) * stride_t;
                    float *d = dst + r * pitch;
                    for (int j = tid; j < U; j += NT) cp_async_4(d + j, src + (int64_t)cols[j] * stride_v);
                }
            }
        }
        cp_async_commit();
    }

    __device__ __forceinline__ void prologue(int tid) {
#pragma unroll
        for (int c = 0; c < kStages - 1; ++c) issue(c, tid);
    }

    // bulk mode: block until chunk `chunk` has landed (may be called again for the same chunk)
    __device__ __forceinline__ void wait_landed(int chunk) {
        mbar_wait(&bars[chunk % kStages], (uint32_t)(chunk / kStages) & 1u);
    }

    // Make chunk `chunk` visible to the group and refill the stage freed by chunk-1.
    __device__ __forceinline__ const float *acquire(int chunk, int tid) {
        if (bulk) {
            mbar_wait(&bars[chunk % kStages], (uint32_t)(chunk / kStages) & 1u);
            // this thread's generic-proxy accesses to the stage about to be refilled (reads, and
            // the in-place prescale of the alpha kernel) are ordered before the async-proxy write
            fence_proxy_async();
        } else {
            cp_async_wait<kStages - 2>();
        }
        group_sync<WARPS>();
        issue(chunk + kStages - 1, tid);
        return stage_ptr(chunk);
    }
};

// Host-side sizing shared by the launchers.
struct PipeGeometry {
    int pitch;      // floats per panel row
    int tc;         // frames per chunk
    size_t ring_bytes;
};

inline PipeGeometry pipe_geometry(int U_panel, size_t budget_bytes, bool odd_pitch = false) {
    PipeGeometry g;
    g.pitch = odd_pitch ? (U_panel | 1) : ((U_panel + 3) & ~3);  // odd: column-wise fills hit distinct banks
    size_t per_frame = (size_t)kStages * g.pitch * sizeof(float);
    int tc = (int)(budget_bytes / per_frame);
    if (tc > 32) tc = 32;
    if (const char *e = tuning("IPFA_PIPE_TC")) {  // tuning override: frames per chunk
        const int v = atoi(e);
        if (v >= 2 && v < tc) tc = v;
    }
    tc &= ~1;  // even: frame parity == row parity inside a chunk (double-buffered exchange lines)
    if (tc < 2) tc = 2;
    g.tc = tc;
    g.ring_bytes = per_frame * tc;
    return g;
}

}  // namespace ipfa
