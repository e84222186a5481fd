// ipfa_common.cuh -- shared device helpers for the sm_100a alignment kernels.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include "../../include/ipfa_b200.h"

namespace ipfa {

// Value of an IPFA_* tuning switch, or nullptr when it is not set.  The environment is read once per
// process (host_api.cu; ipfa_tuning_reload() reads it again), never inside a compute call.
const char *tuning(const char *name);

// NVTX range over the launches of one phase (fill / backtrace / select ...): what a timeline tool
// groups the kernels by; a push / pop pair costs nothing measurable when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// bench.py's per-kernel timing (host_api.cu): event brackets around a launch, no-ops unless switched on
int profile_begin(cudaStream_t st);
void profile_end(int slot, cudaStream_t st);

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// Finite stand-in for log(0) in the log-sum-exp recursion: differences of two
// unreachable states stay finite (no inf-inf NaN) and adding an emission to it
// is absorbed (ulp(1e30) >> |emission|).
constexpr float kNegBig = -1.0e30f;
constexpr float kNegThreshold = -1.0e29f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// log2(2^a + 2^b)
__device__ __forceinline__ float lse2_2(float a, float b) {
    const float m = fmaxf(a, b);
    const float d = -fabsf(a - b);  // == min - max; the |.| and the sign fold into the MUFU operand
    return m + lg2_approx(1.0f + ex2_approx(d));
}
// log2(2^a + 2^b + 2^c): three MUFU ops (the max term is exp2(0) = 1)
__device__ __forceinline__ float lse2_3(float a, float b, float c) {
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    const float m = fmaxf(hi, c), mid = fminf(hi, c);
    return m + lg2_approx(1.0f + ex2_approx(mid - m) + ex2_approx(lo - m));
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async_4(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier + 1-D bulk copy (TMA engine, SASS: UBLKCP) ------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- thread-block clusters: distributed shared memory --------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
// shared::cluster address of the same smem variable in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t cluster_map(const void *smem_ptr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
// asynchronous 4-byte store into another CTA's shared memory that completes 4 bytes of an mbarrier there
__device__ __forceinline__ void st_async_u32(uint32_t addr, uint32_t v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(addr), "r"(v),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void fence_cluster() {  // orders this thread's accesses at cluster scope (release / acquire)
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// Length buckets (large ragged batches): windows whose lattice fits half the launch width go to
// a half-width instance -- half the lanes, half the instructions per frame.  Decided on the device
// (no host synchronisation): one pass writes the two index lists, both instances are launched over
// the full grid and the groups beyond their list's length leave at once.
static __global__ void length_bucket_kernel(const int32_t *__restrict__ tgt_len, int N, int small_units,
                                      int32_t *__restrict__ order, int32_t *__restrict__ count) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= N) return;
    const int cls = (max(tgt_len[w], 0) + 1 <= small_units) ? 0 : 1;
    // warp-aggregated append
    const unsigned active = __activemask();
    const unsigned same = __match_any_sync(active, cls);
    const int leader = __ffs(same) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(count + cls, __popc(same));
    base = __shfl_sync(same, base, leader);
    order[(int64_t)cls * N + base + __popc(same & ((1u << lane) - 1u))] = w;
}

constexpr int kBucketMinWindows = 4096;  // below this an extra launch costs more than it saves

}  // namespace ipfa
