// Measures the packed-fp32 instructions of sm_100 (add.f32x2 / mul.f32x2 / fma.rn.f32x2 -> SASS FADD2 /
// FMUL2 / FFMA2) next to scalar FADD: issue rate and dependent latency per SM sub-partition.  The third
// tier of the window scorer (ctc_alpha_f32.cu) keeps two windows in the halves of a 64-bit register.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_f32x2 tools/microbench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CHAINS>
__global__ void k(float *out, long long *cyc, int iters) {
    unsigned long long x[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { f[c] = 1.0f + 1e-3f * (threadIdx.x + c); x[c] = ((unsigned long long)__float_as_uint(f[c]) << 32) | __float_as_uint(f[c]); }
    const unsigned long long one = 0x3f8000003f800000ull;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (MODE == 0) asm volatile("add.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(one));
            if (MODE == 1) asm volatile("mul.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(one));
            if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(x[c]) : "l"(one));
            if (MODE == 3) asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(f[c]));
            if (MODE == 4) { asm volatile("add.f32x2 %0, %0, %1;" : "+l"(x[c]) : "l"(one));
                             asm volatile("add.f32 %0, %0, 0f3F800000;" : "+f"(f[c])); }
            if (MODE == 5) f[c] = __shfl_up_sync(0xffffffffu, f[c], 1);
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c] + __uint_as_float((unsigned)x[c]) + __uint_as_float((unsigned)(x[c] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
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
    printf("%-12s chains=%d warps/SM=%2d: %.2f cycles per warp-instruction%s per SMSP\n", name, CHAINS, warps,
           h / inst, MODE == 4 ? " pair" : "");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0, 8>("fadd2", w); run<1, 8>("fmul2", w); run<2, 8>("ffma2", w); run<3, 8>("fadd", w);
        run<4, 8>("fadd2+fadd", w); run<5, 8>("shfl", w);
    }
    run<0, 1>("fadd2 (lat)", 4); run<2, 1>("ffma2 (lat)", 4); run<3, 1>("fadd (lat)", 4); run<5, 1>("shfl (lat)", 4);
    return 0;
}
