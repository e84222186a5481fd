// Measures the FP64 pipe (DADD / DMUL) issue rate and latency per SM sub-partition, next to FADD,
// on the box it runs on (the linear-domain alpha recursion runs on this pipe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CHAINS>
__global__ void k(double *out, long long *cyc, int iters) {
    double x[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = 1.0 + 1e-9 * (threadIdx.x + c + 1); f[c] = (float)x[c]; }
    const double one = 1.0 + 1e-12 * threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) asm volatile("add.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(one));
            if (MODE == 1) asm volatile("mul.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(one));
            if (MODE == 2) asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(f[c]));
            if (MODE == 3) { asm volatile("add.f64 %0, %0, %1;" : "+d"(x[c]) : "d"(one));
                             asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(f[c])); }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c] + f[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int CHAINS>
void run(const char *name, int warps) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    int iters = 4096;
    k<MODE, CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    k<MODE, CHAINS><<<148, warps * 32>>>(out, cyc, iters);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_smsp_warps = warps / 4.0;
    double inst = (double)iters * CHAINS * (per_smsp_warps < 1 ? 1 : per_smsp_warps);
    printf("%-10s chains=%d warps/SM=%2d: %.2f cycles per warp-instruction%s per SMSP\n", name, CHAINS, warps,
           h / inst, MODE == 3 ? " pair" : "");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0, 8>("dadd", w); run<1, 8>("dmul", w); run<2, 8>("fadd", w); run<3, 8>("dadd+fadd", w);
    }
    run<0, 1>("dadd", 4); run<1, 1>("dmul", 4);
    return 0;
}
