// Microbenchmark: issue rate of FFMA (3 distinct registers), FFMA2 (fma.rn.f32x2) and the depthwise kernels' operand
// patterns on one SM sub-partition, to decide whether packed FMAs buy FP32 throughput or only issue slots on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/fma_rate scripts/ubench/fma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, const float* in, int iters, long long* cyc) {
    float2 a[8], x[8], w[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + 8 + i]);
        x[i] = make_float2(in[threadIdx.x + 16 + i], in[threadIdx.x + 24 + i]);
        w[i] = make_float2(in[threadIdx.x + 32 + i], in[threadIdx.x + 40 + i]);
    }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {          // 16 scalar FFMA, three distinct registers each
                    a[i].x = fmaf(x[i].x, w[i].x, a[i].x);
                    a[i].y = fmaf(x[i].y, w[i].y, a[i].y);
                } else if (MODE == 1) {   // 8 FFMA2, three distinct register pairs
                    a[i] = __ffma2_rn(x[i], w[i], a[i]);
                } else if (MODE == 2) {   // 8 FFMA2, one operand shared by all (the filter-gradient pattern)
                    a[i] = __ffma2_rn(x[i], w[0], a[i]);
                } else if (MODE == 3) {   // 16 scalar FFMA, one operand shared
                    a[i].x = fmaf(x[i].x, w[0].x, a[i].x);
                    a[i].y = fmaf(x[i].y, w[0].y, a[i].y);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
    float *out, *in; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&in, 4096 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 4);
    const int iters = 2000;
    k<MODE><<<148, threads>>>(out, in, iters, cyc);
    k<MODE><<<148, threads>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double fma_per_thread = (double)iters * 4 * 16;
    const double warps_per_smsp = threads / 32 / 4.0;
    // cycles per warp-level "32-lane FMA" per sub-partition
    printf("%-34s threads=%3d: %8lld cycles, %.3f cycles per 32 FMAs per SMSP (1.0 = 128 FMA/clk/SM)\n", name, threads, c,
           (double)c / (fma_per_thread * warps_per_smsp));
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<0>("FFMA  3 distinct regs", threads);
        run<1>("FFMA2 3 distinct pairs", threads);
        run<2>("FFMA2 shared multiplier", threads);
        run<3>("FFMA  shared multiplier", threads);
    }
    return 0;
}
