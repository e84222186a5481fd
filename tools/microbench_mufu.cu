// Measures MUFU (ex2 / lg2) and shuffle issue rates per SM sub-partition on the box it runs on.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_mufu.cu && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CHAINS>
__global__ void k(float *out, long long *cyc, int iters) {
    float x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = -0.001f * (threadIdx.x + c + 1);
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (MODE == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
            if (MODE == 2) x[c] = __shfl_up_sync(0xffffffffu, x[c], 1);
            if (MODE == 3) asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(x[c]));
            if (MODE == 4) asm volatile("max.f32 %0, %0, 0fBF800000;" : "+f"(x[c]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ex2 issued with only the first `active` lanes of every warp participating: does the MUFU pipe
// charge a partially active warp less than a full one?
template <int CHAINS>
__global__ void k_partial(float *out, long long *cyc, int iters, int active) {
    float x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = -0.001f * (threadIdx.x + c + 1);
    __syncthreads();
    long long t0 = clock64();
    if ((threadIdx.x & 31) < active) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

void run_partial(int warps, int active) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    int iters = 4096;
    k_partial<8><<<148, warps * 32>>>(out, cyc, iters, active);
    k_partial<8><<<148, warps * 32>>>(out, cyc, iters, active);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("ex2    chains=8 warps/SM=%2d active lanes=%2d: %.2f cycles per warp-instruction per SMSP\n", warps, active,
           h / ((double)iters * 8 * (warps / 4.0)));
    cudaFree(out); cudaFree(cyc);
}

template <int MODE, int CHAINS>
void run(const char *name, int warps) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    int iters = 4096;
    k<MODE, CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    k<MODE, CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_smsp_warps = warps / 4.0;
    double inst = (double)iters * CHAINS * (per_smsp_warps < 1 ? 1 : per_smsp_warps);
    printf("%-6s chains=%d warps/SM=%2d: %.2f cycles per warp-instruction per SMSP (latency-bound if warps small)\n",
           name, CHAINS, warps, h / inst);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0, 8>("ex2", w); run<1, 8>("lg2", w); run<2, 8>("shfl", w); run<3, 8>("fadd", w); run<4, 8>("fmnmx", w);
    }
    for (int a : {32, 16, 8, 4, 1}) run_partial(16, a);
    run<0, 1>("ex2", 4); run<1, 1>("lg2", 4); run<2, 1>("shfl", 4); run<3, 1>("fadd", 4);
    return 0;
}
