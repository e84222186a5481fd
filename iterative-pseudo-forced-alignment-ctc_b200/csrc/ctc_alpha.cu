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
// 1 + 1 per blank state), not HBM bound; see DESIGN.md.
#include "emission_pipe.cuh"
#include "lattice_shapes.cuh"

namespace ipfa {

struct AlphaParams {
    const float *lp;
    int64_t stride_n, stride_t;
    const int32_t *targets;
    int64_t tgt_stride;
    const int32_t *in_len;
    const int32_t *tgt_len;
    const int32_t *order;  // optional window permutation (heaviest first), may be null
    int N, V, blank;
    int pitch, tc;         // pipe geometry
    int u_cap;             // capacity of the per-group column list (gather mode)
    int l_cap;             // Lmax the launch was sized for
    size_t group_smem;     // bytes of shared memory per group
    float *nll_out;
};

template <int P, int WARPS, bool DENSE>
__global__ void __launch_bounds__(WARPS == 1 ? 128 : 32 * WARPS)
ctc_alpha_kernel(const AlphaParams prm) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    constexpr int NT = 32 * WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int group = (WARPS == 1) ? (threadIdx.x >> 5) : 0;
    const int tid = (WARPS == 1) ? (threadIdx.x & 31) : threadIdx.x;
    int w = blockIdx.x * GROUPS + group;
    if (w >= prm.N) return;  // WARPS==1: whole warp leaves; WARPS>1: whole CTA leaves
    if (prm.order) w = prm.order[w];

    unsigned char *gsm = smem_raw + (size_t)group * prm.group_smem;
    float *ring = reinterpret_cast<float *>(gsm);
    float *xline = ring + (size_t)kStages * prm.tc * prm.pitch;  // [2][NT + 1] neighbour exchange (WARPS > 1)
    float *fin = xline + 2 * (NT + 1);                            // [2]
    int *cols = reinterpret_cast<int *>(fin + 2);                 // [u_cap] (gather mode)

    const int T = prm.in_len[w];
    const int L = max(0, min(prm.tgt_len[w], prm.l_cap));
    const int32_t *tg = prm.targets + (int64_t)w * prm.tgt_stride;
    const int blank = prm.blank;

    if (T <= 0) {
        if (tid == 0) prm.nll_out[w] = (L == 0) ? 0.0f : __int_as_float(0x7f800000);
        return;
    }

    // per-thread lattice constants
    int col[P];      // panel column of label_p
    bool skip[P];    // s-2 transition allowed into label_p
    bool bad = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;  // target index of this pair's label
        // States past the end of the target (j >= L) are left to run on garbage: they only
        // feed states further right, never a real one, and every value stays finite.
        const bool lab_ok = j < L;
        int lab = lab_ok ? tg[j] : blank;
        if (lab < 0 || lab >= prm.V) { bad = true; lab = blank; }
        const int prev = (j >= 1 && lab_ok) ? tg[j - 1] : -1;
        skip[p] = lab_ok && j >= 1 && prev != lab;
        col[p] = DENSE ? lab : (lab_ok ? j + 1 : 0);
    }
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
    if constexpr (WARPS > 1) {
        if (tid < 2) xline[tid * (NT + 1)] = kNegBig;  // left neighbour of thread 0: log(0)
    }
    group_sync<WARPS>();

    EmissionPipe<WARPS, DENSE> pipe;
    pipe.init(ring, cols, prm.lp + (int64_t)w * prm.stride_n, prm.stride_t, T, U, prm.V, prm.pitch,
              prm.tc, reinterpret_cast<uint64_t *>(gsm + prm.group_smem - 32), tid);
    pipe.prologue(tid);

    float ab[P], al[P];  // blank / label alphas (log2 domain)
#pragma unroll
    for (int p = 0; p < P; ++p) { ab[p] = kNegBig; al[p] = kNegBig; }

    // One frame of the recursion.  `rd`/`wr`: this frame's read / write lines of the
    // cross-warp exchange (WARPS > 1); inside a warp the neighbour comes by shuffle.
    auto frame = [&](const float *row, const float *rd, float *wr) {
        const float eb = row[colb];
        float el[P];
#pragma unroll
        for (int p = 0; p < P; ++p) el[p] = row[col[p]];
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
    };

    float *line0 = xline, *line1 = xline + NT + 1;
    for (int chunk = 0; chunk < pipe.nchunks; ++chunk) {
        float *panel = const_cast<float *>(pipe.acquire(chunk, tid));
        const int t0 = chunk * pipe.tc;
        const int rows = min(pipe.tc, T - t0);
        // in-place: natural log -> log2, clamp log(0) to the finite stand-in
        {
            float4 *p4 = reinterpret_cast<float4 *>(panel);
            const int n4 = (rows * prm.pitch) >> 2;
            for (int q = tid; q < n4; q += NT) {
                float4 v = p4[q];
                v.x = fmaxf(v.x * kLog2e, kNegBig); v.y = fmaxf(v.y * kLog2e, kNegBig);
                v.z = fmaxf(v.z * kLog2e, kNegBig); v.w = fmaxf(v.w * kLog2e, kNegBig);
                p4[q] = v;
            }
            group_sync<WARPS>();
        }
        int r = 0;
        if (chunk == 0) {  // frame 0: only states 0 and 1 are alive
            if (tid == 0) {
                ab[0] = panel[colb];
                if (L > 0) al[0] = panel[col[0]];
            }
            if constexpr (WARPS > 1) {
                line0[tid + 1] = al[P - 1];
                __syncthreads();
            }
            r = 1;
        }
        // frame t reads line[(t-1)&1] and writes line[t&1]; tc is even, so r has t's parity
        const float *row = panel + r * prm.pitch;
        if ((r & 1) && r < rows) { frame(row, line0, line1); row += prm.pitch; ++r; }
        for (; r + 1 < rows; r += 2) {
            frame(row, line1, line0);
            frame(row + prm.pitch, line0, line1);
            row += 2 * prm.pitch;
        }
        if (r < rows) frame(row, line1, line0);
    }

    // final states 2L (blank of pair L) and 2L-1 (label of pair L-1)
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int j = tid * P + p;
        if (j == L) fin[0] = ab[p];
        if (j == L - 1) fin[1] = al[p];
    }
    if (L == 0 && tid == 0) fin[1] = kNegBig;
    group_sync<WARPS>();
    if (tid == 0) {
        const float v = lse2_2(fin[0], fin[1]);
        float nll = -v * kLn2;
        if (v < kNegThreshold || bad) nll = __int_as_float(0x7f800000);
        prm.nll_out[w] = nll;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
extern cudaError_t g_last_cuda_error;
extern uint64_t g_launch_count;

template <int P, int WARPS, bool DENSE>
static int launch_alpha(AlphaParams prm, int Lmax, cudaStream_t stream) {
    constexpr int GROUPS = (WARPS == 1) ? 4 : 1;
    const int U = DENSE ? prm.V : (Lmax + 1);
    // shared-memory budget per group for the emission ring
    const size_t budget = (WARPS == 1) ? (16 * 1024) : (160 * 1024);
    PipeGeometry g = pipe_geometry(U, budget);
    prm.pitch = g.pitch;
    prm.tc = g.tc;
    prm.u_cap = DENSE ? 0 : ((Lmax + 1 + 3) & ~3);
    prm.l_cap = Lmax;
    size_t group_smem = g.ring_bytes + (2 * (32 * WARPS + 1) + 2) * sizeof(float) + (size_t)prm.u_cap * sizeof(int) + 40;
    group_smem = (group_smem + 15) & ~(size_t)15;
    prm.group_smem = group_smem;
    const size_t smem = group_smem * GROUPS;
    auto kern = ctc_alpha_kernel<P, WARPS, DENSE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    const int threads = (WARPS == 1) ? 128 : 32 * WARPS;
    const int blocks = (prm.N + GROUPS - 1) / GROUPS;
    kern<<<blocks, threads, smem, stream>>>(prm);
    ++g_launch_count;
    e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda_error = e; return IPFA_ERR_CUDA; }
    return IPFA_OK;
}

template <bool DENSE>
static int dispatch_alpha(const AlphaParams &prm, int Lmax, LatticeShape s, cudaStream_t stream) {
#define IPFA_X(P_, W_) \
    if (s.PER == P_ && s.WARPS == W_) return launch_alpha<P_, W_, DENSE>(prm, Lmax, stream);
    IPFA_FOR_EACH_SHAPE(IPFA_X)
#undef IPFA_X
    return IPFA_ERR_UNSUPPORTED;
}

bool use_dense_panel(int V, int Lmax) { return V <= 64 || V <= 2 * (Lmax + 1); }

}  // namespace ipfa

using namespace ipfa;

extern "C" size_t ipfa_ctc_alpha_workspace_bytes(int N, int, int, int) {
    return (size_t)(N > 0 ? N : 1) * sizeof(int32_t) + 256;
}

extern "C" int ipfa_ctc_alpha_device(const float *lp, int64_t stride_n, int64_t stride_t,
                                     const int32_t *targets, int64_t tgt_stride,
                                     const int32_t *in_len, const int32_t *tgt_len, int N, int Tmax,
                                     int Lmax, int V, int blank, float *nll_out, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    (void)workspace; (void)workspace_bytes; (void)Tmax;
    if (N == 0) return IPFA_OK;
    if (!lp || !in_len || !tgt_len || !nll_out || N < 0 || V <= 0 || Lmax < 0 || blank < 0 || blank >= V ||
        (Lmax > 0 && !targets))
        return IPFA_ERR_INVALID_ARG;
    LatticeShape s;
    if (!pick_lattice_shape(Lmax + 1, N, &s, "IPFA_ALPHA_SHAPE", use_dense_panel(V, Lmax) ? 3 : 6)) return IPFA_ERR_UNSUPPORTED;
    AlphaParams prm{};
    prm.lp = lp; prm.stride_n = stride_n; prm.stride_t = stride_t;
    prm.targets = targets; prm.tgt_stride = tgt_stride;
    prm.in_len = in_len; prm.tgt_len = tgt_len; prm.order = nullptr;
    prm.N = N; prm.V = V; prm.blank = blank; prm.nll_out = nll_out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (use_dense_panel(V, Lmax)) return dispatch_alpha<true>(prm, Lmax, s, st);
    return dispatch_alpha<false>(prm, Lmax, s, st);
}
