// host_api.cu -- library state, status strings and the *_host entry points
// (host buffers in, host buffers out: H2D -> kernels -> D2H on an internal
// stream with a grow-only device arena).  See include/ipfa_b200.h.
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ipfa_common.cuh"

namespace ipfa {
cudaError_t g_last_cuda_error = cudaSuccess;
uint64_t g_launch_count = 0;

// ---- tuning switches: the IPFA_* environment variables are read ONCE, when the first entry point
// needs one (or again on ipfa_tuning_reload()); the compute calls look them up in this table.
namespace {
const char *const kTuningNames[] = {
    "IPFA_ALPHA_SHAPE", "IPFA_ALPHA_SMALL_SHAPE", "IPFA_ALPHA_LIN_SHAPE", "IPFA_ALPHA_LOG", "IPFA_NO_BUCKETS",
    "IPFA_VITERBI_SHAPE", "IPFA_VITERBI_SMALL_SHAPE", "IPFA_SEG_SHAPE", "IPFA_PIPE_TC",
    "IPFA_ALPHA_F32", "IPFA_SEG_SKEW", "IPFA_SWEEP_PHASES", "IPFA_SWEEP_KC", "IPFA_SWEEP_CTAS"};
constexpr int kTuningCount = sizeof(kTuningNames) / sizeof(kTuningNames[0]);
struct TuningTable {
    bool present[kTuningCount];
    char value[kTuningCount][32];
    void load() {
        for (int i = 0; i < kTuningCount; ++i) {
            const char *e = getenv(kTuningNames[i]);
            present[i] = e != nullptr;
            value[i][0] = 0;
            if (e) { strncpy(value[i], e, sizeof(value[i]) - 1); value[i][sizeof(value[i]) - 1] = 0; }
        }
    }
};
std::mutex g_tuning_mu;
TuningTable g_tuning;
bool g_tuning_loaded = false;
}  // namespace

const char *tuning(const char *name) {
    std::lock_guard<std::mutex> lock(g_tuning_mu);
    if (!g_tuning_loaded) { g_tuning.load(); g_tuning_loaded = true; }
    for (int i = 0; i < kTuningCount; ++i)
        if (strcmp(name, kTuningNames[i]) == 0) return g_tuning.present[i] ? g_tuning.value[i] : nullptr;
    return nullptr;
}

// ---- per-kernel timing for bench.py's roofline: when switched on, the launch sites of the T-serial
// lattice kernels bracket each launch with CUDA events on the launching stream (not while a graph is
// being captured); ipfa_profile_read_ms() returns the longest bracket since the last read -- the
// step's dominant kernel, timed alone instead of inferred from the step.
namespace {
constexpr int kProfileSlots = 16;
struct ProfileState {
    bool on = false;
    int used = 0;
    cudaEvent_t e0[kProfileSlots] = {}, e1[kProfileSlots] = {};
};
ProfileState g_profile;
}  // namespace

int profile_begin(cudaStream_t st) {
    if (!g_profile.on || g_profile.used >= kProfileSlots) return -1;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return -1;
    const int i = g_profile.used;
    if (!g_profile.e0[i]) {
        if (cudaEventCreate(&g_profile.e0[i]) != cudaSuccess || cudaEventCreate(&g_profile.e1[i]) != cudaSuccess) return -1;
    }
    if (cudaEventRecord(g_profile.e0[i], st) != cudaSuccess) return -1;
    ++g_profile.used;
    return i;
}
void profile_end(int slot, cudaStream_t st) {
    if (slot >= 0) cudaEventRecord(g_profile.e1[slot], st);
}

namespace {
constexpr int kMaxChunks = 32;
constexpr int kMaxDevices = 64;

// One arena per CUDA device: a grow-only device buffer, a copy-in stream, a compute stream and the
// events that order them.  A *_host call works on the arena of the device that is current when it
// is entered and holds that arena's mutex for its duration, so calls on different devices run
// concurrently and a process may switch devices between calls.
struct Arena {
    std::mutex mu;
    bool ready_ok = false;
    cudaStream_t stream = nullptr;    // copy-in stream (H2D)
    cudaStream_t compute = nullptr;   // kernels + D2H
    cudaEvent_t ready[kMaxChunks] = {};
    unsigned char *base = nullptr;
    size_t cap = 0, used = 0;

    void teardown() {
        for (int i = 0; i < kMaxChunks; ++i)
            if (ready[i]) { cudaEventDestroy(ready[i]); ready[i] = nullptr; }
        if (compute) { cudaStreamDestroy(compute); compute = nullptr; }
        if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
        ready_ok = false;
    }
    int ensure(size_t bytes) {
        if (!ready_ok) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking);
            for (int i = 0; i < kMaxChunks && e == cudaSuccess; ++i)
                e = cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming);
            if (e != cudaSuccess) { g_last_cuda_error = e; teardown(); return IPFA_ERR_CUDA; }
            ready_ok = true;
        }
        if (bytes > cap) {
            if (base) { cudaStreamSynchronize(stream); cudaStreamSynchronize(compute); cudaFree(base); base = nullptr; cap = 0; }
            size_t want = bytes + (bytes >> 3) + (1 << 20);
            cudaError_t e = cudaMalloc(&base, want);
            if (e != cudaSuccess) { g_last_cuda_error = e; base = nullptr; return IPFA_ERR_CUDA; }
            cap = want;
        }
        used = 0;
        return IPFA_OK;
    }
    // before ANY return that leaves copies in flight on the caller's buffers
    void quiesce() {
        if (stream) cudaStreamSynchronize(stream);
        if (compute) cudaStreamSynchronize(compute);
    }
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += bytes;
        return p;
    }
};
Arena g_arenas[kMaxDevices];

Arena *current_arena() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        g_last_cuda_error = (e != cudaSuccess) ? e : cudaErrorInvalidDevice;
        return nullptr;
    }
    return &g_arenas[dev];
}
inline size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

// inside a *_host function `A` is the locked arena: an error return first waits for the copies that
// were already queued on the caller's host buffers
#define IPFA_CUDA(call)                                              \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) { g_last_cuda_error = e__; A.quiesce(); return IPFA_ERR_CUDA; } \
    } while (0)
#define IPFA_RC(expr)                                                \
    do {                                                             \
        int rc__ = (expr);                                           \
        if (rc__) { A.quiesce(); return rc__; }                      \
    } while (0)

// copy a [N, rows, V] fp32 host block (row pitch V) with batch stride stride_n
int upload_lp(Arena &A, float *dst, const float *src, int64_t stride_n, int N, int64_t per_window, cudaStream_t st) {
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

extern "C" int ipfa_version(void) { return 202; }

extern "C" const char *ipfa_status_string(int status) {
    switch (status) {
        case IPFA_OK: return "ok";
        case IPFA_ERR_INVALID_ARG: return "invalid argument";
        case IPFA_ERR_UNSUPPORTED: return "lattice wider than the widest kernel instance";
        case IPFA_ERR_WORKSPACE: return "workspace too small";
        case IPFA_ERR_CUDA: return "CUDA runtime error";
        case IPFA_ERR_AUDIO_SHORTER_THAN_TEXT: return "Audio is shorter than text!";
        case IPFA_ERR_WINDOW: return "Maximum window size reached. Check data for large repetitions or noise.";
        default: return "unknown status";
    }
}

extern "C" void ipfa_profile_kernels(int enable) {
    g_profile.on = enable != 0;
    g_profile.used = 0;
}
extern "C" float ipfa_profile_read_ms(void) {
    float best = -1.0f;
    for (int i = 0; i < g_profile.used; ++i) {
        float ms = 0.0f;
        if (cudaEventSynchronize(g_profile.e1[i]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, g_profile.e0[i], g_profile.e1[i]) == cudaSuccess && ms > best)
            best = ms;
    }
    g_profile.used = 0;
    return best;
}

extern "C" void ipfa_tuning_reload(void) {
    std::lock_guard<std::mutex> lock(g_tuning_mu);
    g_tuning.load();
    g_tuning_loaded = true;
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
    Arena *arena = current_arena();
    if (!arena) return IPFA_ERR_CUDA;
    Arena &A = *arena;
    std::lock_guard<std::mutex> lock(A.mu);
    const int64_t per_window = (int64_t)Tmax * V;
    const size_t ws = ipfa_ctc_alpha_workspace_bytes(N, Tmax, Lmax, V);
    size_t need = pad256((size_t)N * per_window * 4) + pad256((size_t)N * (Lmax > 0 ? Lmax : 1) * 4) +
                  3 * pad256((size_t)N * 4) + pad256(ws) + 1024;
    int rc = A.ensure(need);
    if (rc) return rc;
    cudaStream_t st = A.stream;
    float *d_lp = A.take<float>((size_t)N * per_window);
    int32_t *d_tg = A.take<int32_t>((size_t)N * (Lmax > 0 ? Lmax : 1));
    int32_t *d_il = A.take<int32_t>(N);
    int32_t *d_tl = A.take<int32_t>(N);
    float *d_out = A.take<float>(N);
    void *d_ws = A.take<unsigned char>(ws);
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
    cudaStream_t cs = A.compute;
    // every copy is queued before the first kernel launch, so the copy engine never waits for the host
    int n_chunks = 0;
    for (int w0 = 0; w0 < N; w0 += C, ++n_chunks) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        rc = upload_lp(A, d_lp + (int64_t)w0 * per_window, lp + (int64_t)w0 * stride_n, stride_n, n, per_window, st);
        IPFA_RC(rc);
        IPFA_CUDA(cudaEventRecord(A.ready[n_chunks], st));
    }
    for (int w0 = 0, ci = 0; w0 < N; w0 += C, ++ci) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        IPFA_CUDA(cudaStreamWaitEvent(cs, A.ready[ci], 0));
        rc = ipfa_ctc_alpha_device(d_lp + (int64_t)w0 * per_window, per_window, V, d_tg + (int64_t)w0 * Lmax, Lmax,
                                   d_il + w0, d_tl + w0, n, Tmax, Lmax, V, blank, d_out + w0, d_ws, ws, cs);
        IPFA_RC(rc);
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
    Arena *arena = current_arena();
    if (!arena) return IPFA_ERR_CUDA;
    Arena &A = *arena;
    std::lock_guard<std::mutex> lock(A.mu);
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
    int rc = A.ensure(need);
    if (rc) return rc;
    cudaStream_t st = A.stream;
    float *d_lp = A.take<float>((size_t)N * per_window);
    int32_t *d_tg = A.take<int32_t>((size_t)N * lcap);
    int32_t *d_il = A.take<int32_t>(N);
    int32_t *d_tl = A.take<int32_t>(N);
    int32_t *d_paths = A.take<int32_t>((size_t)N * Tmax);
    float *d_scores = A.take<float>((size_t)N * Tmax);
    int32_t *d_ts = A.take<int32_t>((size_t)N * lcap);
    int32_t *d_te = A.take<int32_t>((size_t)N * lcap);
    float *d_tp = A.take<float>((size_t)N * lcap);
    float *d_total = A.take<float>(N);
    int32_t *d_status = A.take<int32_t>(N);
    void *d_ws = A.take<unsigned char>(ws);
    if (Lmax > 0)
        IPFA_CUDA(cudaMemcpy2DAsync(d_tg, (size_t)Lmax * 4, targets, (size_t)tgt_stride * 4, (size_t)Lmax * 4,
                                    N, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_il, in_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    IPFA_CUDA(cudaMemcpyAsync(d_tl, tgt_len, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    const bool tok = tok_start != nullptr && Lmax > 0;
    const int C = windows_per_chunk(N, per_window);
    cudaStream_t cs = A.compute;
    for (int w0 = 0, ci = 0; w0 < N; w0 += C, ++ci) {
        const int n = (N - w0 < C) ? (N - w0) : C;
        const int64_t ot = (int64_t)w0 * Tmax, ol = (int64_t)w0 * Lmax;
        rc = upload_lp(A, d_lp + (int64_t)w0 * per_window, lp + (int64_t)w0 * stride_n, stride_n, n, per_window, st);
        IPFA_RC(rc);
        IPFA_CUDA(cudaEventRecord(A.ready[ci], st));
        IPFA_CUDA(cudaStreamWaitEvent(cs, A.ready[ci], 0));
        rc = ipfa_ctc_viterbi_device(d_lp + (int64_t)w0 * per_window, per_window, V, d_tg + ol, Lmax, d_il + w0,
                                     d_tl + w0, n, Tmax, Lmax, V, blank, d_paths + ot,
                                     scores_out ? d_scores + ot : nullptr, tok ? d_ts + ol : nullptr,
                                     tok ? d_te + ol : nullptr, (tok && tok_score) ? d_tp + ol : nullptr,
                                     d_total + w0, d_status + w0, d_ws, ws, cs);
        IPFA_RC(rc);
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
    Arena *arena = current_arena();
    if (!arena) return IPFA_ERR_CUDA;
    Arena &A = *arena;
    std::lock_guard<std::mutex> lock(A.mu);
    const int64_t per_window = (int64_t)Tmax * V;
    // audio longer than ctc-segmentation's min_window_size (8000 frames): windowed table mode, the
    // window doubled after IPFA_WIN_WINDOW_TOO_SMALL up to max_window_size (100000) like the
    // reference's `except IndexError` loop
    const bool windowed = Tmax > 8000;
    bool window_exhausted = false;
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
    int rc = A.ensure(need);
    if (rc) return rc;
    cudaStream_t st = A.stream;
    float *d_lp = A.take<float>((size_t)N * per_window);
    int32_t *d_gt = A.take<int32_t>((size_t)N * Cmax);
    int32_t *d_ub = A.take<int32_t>((size_t)N * (Kmax + 1));
    int32_t *d_il = A.take<int32_t>(N);
    int32_t *d_nc = A.take<int32_t>(N);
    int32_t *d_nu = A.take<int32_t>(N);
    int32_t *d_status = A.take<int32_t>(N);
    double *d_seg = A.take<double>(n_seg);
    int32_t *d_term = A.take<int32_t>((size_t)N * Kmax);
    int32_t *d_timing = timing_out ? A.take<int32_t>((size_t)N * Kmax * Cmax) : nullptr;
    float *d_cprob = char_prob_out ? A.take<float>((size_t)N * Kmax * Tmax) : nullptr;
    int32_t *d_state = state_out ? A.take<int32_t>((size_t)N * Kmax * Tmax) : nullptr;
    void *d_ws = A.take<unsigned char>(ws);
    rc = upload_lp(A, d_lp, lp, stride_n, N, per_window, st);
    IPFA_RC(rc);
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
        IPFA_RC(rc);
    } else {
        while (true) {
            rc = ipfa_ctcseg_windowed_device(d_lp, nullptr, per_window, V, d_il, d_gt, Cmax, d_nc, d_ub, d_nu, N,
                                             Tmax, Cmax, Kmax, V, blank, index_duration, score_len, flags, window,
                                             1, d_seg, d_term, d_timing, d_cprob, d_state, d_status, d_ws, ws, st);
            IPFA_RC(rc);
            IPFA_CUDA(cudaMemcpyAsync(status_out, d_status, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
            IPFA_CUDA(cudaStreamSynchronize(st));
            bool too_small = false;
            for (int i = 0; i < N; ++i) too_small |= (status_out[i] & IPFA_WIN_WINDOW_TOO_SMALL) != 0;
            if (!too_small) break;
            // the reference doubles the window and gives up with IndexError at max_window_size
            if (window >= Tmax || window * 2 >= 100000) { window_exhausted = true; break; }
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
    return window_exhausted ? IPFA_ERR_WINDOW : IPFA_OK;
}
