// host_api.cu -- library state, status strings and the *_host entry points
// (host buffers in, host buffers out: H2D -> kernels -> D2H on an internal
// stream with a grow-only device arena).  See include/ipfa_b200.h.
#include <mutex>
#include <stdio.h>

#include "ipfa_common.cuh"

namespace ipfa {
cudaError_t g_last_cuda_error = cudaSuccess;
uint64_t g_launch_count = 0;

namespace {
constexpr int kMaxChunks = 32;

struct Arena {
    std::mutex mu;
    cudaStream_t stream = nullptr;    // copy-in stream (H2D)
    cudaStream_t compute = nullptr;   // kernels + D2H
    cudaEvent_t ready[kMaxChunks] = {};
    unsigned char *base = nullptr;
    size_t cap = 0, used = 0;

    int ensure(size_t bytes) {
        if (!stream) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) { g_last_cuda_error = e; stream = nullptr; return IPFA_ERR_CUDA; }
            e = cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking);
            if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
            for (int i = 0; i < kMaxChunks; ++i) {
                e = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming);
                if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
            }
        }
        if (bytes > cap) {
            if (base) { cudaStreamSynchronize(stream); cudaStreamSynchronize(compute); cudaFree(base); base = nullptr; cap = 0; }
            size_t want = bytes + (bytes >> 3) + (1 << 20);
            cudaError_t e = cudaMalloc(&base, want);
            if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
            cap = want;
        }
        used = 0;
        return IPFA_OK;
    }
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
};
Arena g_arena;
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

#define IPFA_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) { g_last_cuda_error = e__; return IPFA_ERR_CUDA; } \
    } while (0)

// copy a [N, rows, V] fp32 host block (row pitch V) with batch stride stride_n
int upload_lp(float *dst, const float *src, int64_t stride_n, int N, int64_t per_window, cudaStream_t st) {
    if (stride_n == per_window || N == 1) {
        IPFA_CUDA(cudaMemcpyAsync(dst, src, (size_t)N * per_window * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        IPFA_CUDA(cudaMemcpy2DAsync(dst, per_window * sizeof(float), src, stride_n * sizeof(float),
                                    per_window * sizeof(float), N, cudaMemcpyHostToDevice, st));
    }
    return IPFA_OK;
}

// Windows per chunk of the host pipeline: H2D of chunk i+1 overlaps the kernels of chunk i.
// ~16 MB of emissions per chunk: copies run at 53.8 of the 55.4 GB/s this box reaches with one 131 MB
// copy (tools/exp_h2d.py) and only the last chunk's kernels are exposed; 32 MB chunks copy at 54.7 GB/s
// but expose twice the kernel tail -- measured 2.63 ms against 2.61 ms per 131 MB step.
int windows_per_chunk(int N, int64_t per_window_floats) {
    const int64_t bytes = per_window_floats * 4;
    int64_t c = bytes > 0 ? (16LL << 20) / bytes : N;
    if (c < 1) c = 1;
    if (c * kMaxChunks < N) c = (N + kMaxChunks - 1) / kMaxChunks;
    if (c > N) c = N;
    return (int)c;
}
}  // namespace
}  // namespace ipfa

using namespace ipfa;

extern "C" int ipfa_version(void) { return 100; }

extern "C" const char *ipfa_status_string(int status) {
    switch (status) {
        case IPFA_OK: return "ok";
        case IPFA_ERR_INVALID_ARG: return "invalid argument";
        case IPFA_ERR_UNSUPPORTED: return "lattice wider than the widest kernel instance";
        case IPFA_ERR_WORKSPACE: return "workspace too small";
        case IPFA_ERR_CUDA: return "CUDA runtime error";
        case IPFA_ERR_AUDIO_SHORTER_THAN_TEXT: return "Audio is shorter than text!";
        default: return "unknown status";
    }
}

extern "C" const char *ipfa_last_cuda_error(void) { return cudaGetErrorString(g_last_cuda_error); }
extern "C" uint64_t ipfa_launch_count(void) { return g_launch_count; }
extern "C" int ipfa_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int ipfa_ctc_alpha_host(const float *lp, int64_t stride_n, int64_t stride_t,
                                   const int32_t *targets, int64_t tgt_stride, const int32_t *in_len,
                                   const int32_t *tgt_len, int N, int Tmax, int Lmax, int V, int blank,
                                   float *nll_out) {
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !tgt_len || !nll_out || N < 0 || Tmax < 0 || stride_t != V ||
        (Lmax > 0 && !targets))
        return IPFA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g_arena.mu);
    const int64_t per_window = (int64_t)Tmax * V;
    const size_t ws = ipfa_ctc_alpha_workspace_bytes(N, Tmax, Lmax, V);
    size_t need = pad256((size_t)N * per_window * 4) + pad256((size_t)N * (Lmax > 0 ? Lmax : 1) * 4) +
                  3 * pad256((size_t)N * 4) + pad256(ws) + 1024;
    int rc = g_arena.ensure(need);
    if (rc) return rc;
    cudaStream_t st = g_arena.stream;
    float *d_lp = g_arena.take<float>((size_t)N * per_window);
    int32_t *d_tg = g_arena.take<int32_t>((size_t)N * (Lmax > 0 ? Lmax : 1));
    int32_t *d_il = g_arena.take<int32_t>(N);
    int32_t *d_tl = g_arena.take<int32_t>(N);
    float *d_out = g_arena.take<float>(N);
    void *d_ws = g_arena.take<unsigned char>(ws);
    if (Lmax > 0) {
        if (tgt_stride == Lmax)
            IPFA_CUDA(cudaMemcpyAsync(d_tg, targets, (size_t)N * Lmax * 4, cudaMemcpyHostToDevice, st));
        else
            IPFA_CUDA(cudaMemcpy2DAsync(d_tg, (size_t)Lmax * 4, targets, (size_t)tgt_stride * 4, (size_t)Lmax * 4,
                                        N, cudaMemcpyHostToDevice, st));
    }
    IPFA_CUDA(cudaMemcpyAsync(d_il, in_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_tl, tgt_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    const int C = windows_per_chunk(N, per_window);
    cudaStream_t cs = g_arena.compute;
    // every copy is queued before the first kernel launch, so the copy engine never waits for the host
    int n_chunks = 0;
    for (int w0 = 0; w0 < N; w0 += C, ++n_chunks) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        rc = upload_lp(d_lp + (int64_t)w0 * per_window, lp + (int64_t)w0 * stride_n, stride_n, n, per_window, st);
        if (rc) return rc;
        IPFA_CUDA(cudaEventRecord(g_arena.ready[n_chunks], st));
    }
    for (int w0 = 0, ci = 0; w0 < N; w0 += C, ++ci) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        IPFA_CUDA(cudaStreamWaitEvent(cs, g_arena.ready[ci], 0));
        rc = ipfa_ctc_alpha_device(d_lp + (int64_t)w0 * per_window, per_window, V, d_tg + (int64_t)w0 * Lmax, Lmax,
                                   d_il + w0, d_tl + w0, n, Tmax, Lmax, V, blank, d_out + w0, d_ws, ws, cs);
        if (rc) return rc;
    }
    IPFA_CUDA(cudaMemcpyAsync(nll_out, d_out, (size_t)N * 4, cudaMemcpyDeviceToHost, cs));
    IPFA_CUDA(cudaStreamSynchronize(cs));
    return IPFA_OK;
}

extern "C" int ipfa_ctc_viterbi_host(const float *lp, int64_t stride_n, int64_t stride_t,
                                     const int32_t *targets, int64_t tgt_stride, const int32_t *in_len,
                                     const int32_t *tgt_len, int N, int Tmax, int Lmax, int V, int blank,
                                     int32_t *paths_out, float *scores_out, int32_t *tok_start,
                                     int32_t *tok_end, float *tok_score, float *total_out,
                                     int32_t *status_out) {
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !tgt_len || !paths_out || !status_out || N < 0 || Tmax < 0 || stride_t != V ||
        (Lmax > 0 && !targets) || ((tok_start == nullptr) != (tok_end == nullptr)))
        return IPFA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g_arena.mu);
    const int64_t per_window = (int64_t)Tmax * V;
    size_t ws = ipfa_ctc_viterbi_workspace_bytes(N, Tmax, Lmax, V);
    {   // chunks may pick a lattice shape with a different backpointer pitch
        const int c = windows_per_chunk(N, per_window);
        const size_t ws_c = ipfa_ctc_viterbi_workspace_bytes(c, Tmax, Lmax, V);
        const int tail = N % c;
        const size_t ws_t = tail ? ipfa_ctc_viterbi_workspace_bytes(tail, Tmax, Lmax, V) : 0;
        if (ws_c > ws) ws = ws_c;
        if (ws_t > ws) ws = ws_t;
    }
    const size_t lcap = (size_t)(Lmax > 0 ? Lmax : 1);
    size_t need = pad256((size_t)N * per_window * 4) + pad256((size_t)N * lcap * 4) * 4 +
                  4 * pad256((size_t)N * 4) + 2 * pad256((size_t)N * Tmax * 4) + pad256(ws) + 4096;
    int rc = g_arena.ensure(need);
    if (rc) return rc;
    cudaStream_t st = g_arena.stream;
    float *d_lp = g_arena.take<float>((size_t)N * per_window);
    int32_t *d_tg = g_arena.take<int32_t>((size_t)N * lcap);
    int32_t *d_il = g_arena.take<int32_t>(N);
    int32_t *d_tl = g_arena.take<int32_t>(N);
    int32_t *d_paths = g_arena.take<int32_t>((size_t)N * Tmax);
    float *d_scores = g_arena.take<float>((size_t)N * Tmax);
    int32_t *d_ts = g_arena.take<int32_t>((size_t)N * lcap);
    int32_t *d_te = g_arena.take<int32_t>((size_t)N * lcap);
    float *d_tp = g_arena.take<float>((size_t)N * lcap);
    float *d_total = g_arena.take<float>(N);
    int32_t *d_status = g_arena.take<int32_t>(N);
    void *d_ws = g_arena.take<unsigned char>(ws);
    if (Lmax > 0)
        IPFA_CUDA(cudaMemcpy2DAsync(d_tg, (size_t)Lmax * 4, targets, (size_t)tgt_stride * 4, (size_t)Lmax * 4,
                                    N, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_il, in_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_tl, tgt_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    const bool tok = tok_start != nullptr && Lmax > 0;
    const int C = windows_per_chunk(N, per_window);
    cudaStream_t cs = g_arena.compute;
    for (int w0 = 0, ci = 0; w0 < N; w0 += C, ++ci) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        const int64_t ot = (int64_t)w0 * Tmax, ol = (int64_t)w0 * Lmax;
        rc = upload_lp(d_lp + (int64_t)w0 * per_window, lp + (int64_t)w0 * stride_n, stride_n, n, per_window, st);
        if (rc) return rc;
        IPFA_CUDA(cudaEventRecord(g_arena.ready[ci], st));
        IPFA_CUDA(cudaStreamWaitEvent(cs, g_arena.ready[ci], 0));
        rc = ipfa_ctc_viterbi_device(d_lp + (int64_t)w0 * per_window, per_window, V, d_tg + ol, Lmax, d_il + w0,
                                     d_tl + w0, n, Tmax, Lmax, V, blank, d_paths + ot,
                                     scores_out ? d_scores + ot : nullptr, tok ? d_ts + ol : nullptr,
                                     tok ? d_te + ol : nullptr, (tok && tok_score) ? d_tp + ol : nullptr,
                                     d_total + w0, d_status + w0, d_ws, ws, cs);
        if (rc) return rc;
        IPFA_CUDA(cudaMemcpyAsync(paths_out + ot, d_paths + ot, (size_t)n * Tmax * 4, cudaMemcpyDeviceToHost, cs));
        if (scores_out)
            IPFA_CUDA(cudaMemcpyAsync(scores_out + ot, d_scores + ot, (size_t)n * Tmax * 4, cudaMemcpyDeviceToHost, cs));
    }
    if (tok) {
        IPFA_CUDA(cudaMemcpyAsync(tok_start, d_ts, (size_t)N * Lmax * 4, cudaMemcpyDeviceToHost, cs));
        IPFA_CUDA(cudaMemcpyAsync(tok_end, d_te, (size_t)N * Lmax * 4, cudaMemcpyDeviceToHost, cs));
        if (tok_score)
            IPFA_CUDA(cudaMemcpyAsync(tok_score, d_tp, (size_t)N * Lmax * 4, cudaMemcpyDeviceToHost, cs));
    }
    if (total_out) IPFA_CUDA(cudaMemcpyAsync(total_out, d_total, (size_t)N * 4, cudaMemcpyDeviceToHost, cs));
    IPFA_CUDA(cudaMemcpyAsync(status_out, d_status, (size_t)N * 4, cudaMemcpyDeviceToHost, cs));
    IPFA_CUDA(cudaStreamSynchronize(cs));
    return IPFA_OK;
}

extern "C" int ipfa_ctcseg_host(const float *lp, int64_t stride_n, int64_t stride_t, const int32_t *in_len,
                                const int32_t *gt, int64_t gt_stride, const int32_t *n_cols,
                                const int32_t *utt_begin, const int32_t *n_utts, int N, int Tmax, int Cmax,
                                int Kmax, int V, int blank, double index_duration, int score_len, int flags,
                                double *seg_out, int32_t *term_t_out, int32_t *timing_out,
                                float *char_prob_out, int32_t *state_out, int32_t *status_out) {
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !gt || !n_cols || !utt_begin || !n_utts || !seg_out || !term_t_out || !status_out ||
        N < 0 || Tmax <= 0 || Cmax <= 0 || Kmax <= 0 || stride_t != V)
        return IPFA_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(g_arena.mu);
    const int64_t per_window = (int64_t)Tmax * V;
    // audio longer than ctc-segmentation's min_window_size (8000 frames): windowed table mode, the
    // window doubled after IPFA_WIN_WINDOW_TOO_SMALL up to max_window_size (100000) like the
    // reference's `except IndexError` loop
    const bool windowed = Tmax > 8000;
    int window = 8000;
    size_t ws = windowed ? 0 : ipfa_ctcseg_workspace_bytes(N, Tmax, Cmax, Kmax, V);
    if (windowed)
        for (int wsz = window; wsz < 100000; wsz *= 2) {
            const size_t b = ipfa_ctcseg_windowed_workspace_bytes(N, Tmax, Cmax, Kmax, wsz, 1, flags);
            if (b > ws) ws = b;
            if (wsz >= Tmax) break;
        }
    const size_t n_seg = (size_t)N * Kmax * Kmax * 3;
    size_t need = pad256((size_t)N * per_window * 4) + pad256((size_t)N * Cmax * 4) +
                  pad256((size_t)N * (Kmax + 1) * 4) + 4 * pad256((size_t)N * 4) + pad256(n_seg * 8) +
                  pad256((size_t)N * Kmax * 4) + pad256((size_t)N * Kmax * (size_t)Cmax * 4) +
                  2 * pad256((size_t)N * Kmax * (size_t)Tmax * 4) + pad256(ws) + 8192;
    int rc = g_arena.ensure(need);
    if (rc) return rc;
    cudaStream_t st = g_arena.stream;
    float *d_lp = g_arena.take<float>((size_t)N * per_window);
    int32_t *d_gt = g_arena.take<int32_t>((size_t)N * Cmax);
    int32_t *d_ub = g_arena.take<int32_t>((size_t)N * (Kmax + 1));
    int32_t *d_il = g_arena.take<int32_t>(N);
    int32_t *d_nc = g_arena.take<int32_t>(N);
    int32_t *d_nu = g_arena.take<int32_t>(N);
    int32_t *d_status = g_arena.take<int32_t>(N);
    double *d_seg = g_arena.take<double>(n_seg);
    int32_t *d_term = g_arena.take<int32_t>((size_t)N * Kmax);
    int32_t *d_timing = timing_out ? g_arena.take<int32_t>((size_t)N * Kmax * Cmax) : nullptr;
    float *d_cprob = char_prob_out ? g_arena.take<float>((size_t)N * Kmax * Tmax) : nullptr;
    int32_t *d_state = state_out ? g_arena.take<int32_t>((size_t)N * Kmax * Tmax) : nullptr;
    void *d_ws = g_arena.take<unsigned char>(ws);
    rc = upload_lp(d_lp, lp, stride_n, N, per_window, st);
    if (rc) return rc;
    IPFA_CUDA(cudaMemcpy2DAsync(d_gt, (size_t)Cmax * 4, gt, (size_t)gt_stride * 4, (size_t)Cmax * 4, N,
                                cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_ub, utt_begin, (size_t)N * (Kmax + 1) * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_il, in_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_nc, n_cols, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_nu, n_utts, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemsetAsync(d_seg, 0xff, n_seg * 8, st));  // unfilled slots read as NaN
    if (!windowed) {
        rc = ipfa_ctcseg_device(d_lp, per_window, V, d_il, d_gt, Cmax, d_nc, d_ub, d_nu, N, Tmax, Cmax, Kmax, V,
                                blank, index_duration, score_len, flags, d_seg, d_term, d_timing, d_cprob, d_state,
                                d_status, d_ws, ws, st);
        if (rc) return rc;
    } else {
        while (true) {
            rc = ipfa_ctcseg_windowed_device(d_lp, nullptr, per_window, V, d_il, d_gt, Cmax, d_nc, d_ub, d_nu, N,
                                             Tmax, Cmax, Kmax, V, blank, index_duration, score_len, flags, window,
                                             1, d_seg, d_term, d_timing, d_cprob, d_state, d_status, d_ws, ws, st);
            if (rc) return rc;
            IPFA_CUDA(cudaMemcpyAsync(status_out, d_status, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
            IPFA_CUDA(cudaStreamSynchronize(st));
            bool too_small = false;
            for (int i = 0; i < N; ++i) too_small |= (status_out[i] & IPFA_WIN_WINDOW_TOO_SMALL) != 0;
            if (!too_small || window >= Tmax || window * 2 >= 100000) break;
            window *= 2;
        }
    }
    IPFA_CUDA(cudaMemcpyAsync(seg_out, d_seg, n_seg * 8, cudaMemcpyDeviceToHost, st));
    IPFA_CUDA(cudaMemcpyAsync(term_t_out, d_term, (size_t)N * Kmax * 4, cudaMemcpyDeviceToHost, st));
    if (timing_out)
        IPFA_CUDA(cudaMemcpyAsync(timing_out, d_timing, (size_t)N * Kmax * Cmax * 4, cudaMemcpyDeviceToHost, st));
    if (char_prob_out)
        IPFA_CUDA(cudaMemcpyAsync(char_prob_out, d_cprob, (size_t)N * Kmax * Tmax * 4, cudaMemcpyDeviceToHost, st));
    if (state_out)
        IPFA_CUDA(cudaMemcpyAsync(state_out, d_state, (size_t)N * Kmax * Tmax * 4, cudaMemcpyDeviceToHost, st));
    IPFA_CUDA(cudaMemcpyAsync(status_out, d_status, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    IPFA_CUDA(cudaStreamSynchronize(st));
    return IPFA_OK;
}
