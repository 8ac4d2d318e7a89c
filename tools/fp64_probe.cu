// FP64 pipe probe for B200: dependent-DFMA latency and throughput vs (warps per SM, chains per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void chains(double* sink, int iters, long long* cyc) {
    double a[C];
#pragma unroll
    for (int c = 0; c < C; ++c) a[c] = threadIdx.x * 1e-9 + c;
    const double m = 0.999999, k = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < C; ++c) a[c] = fma(a[c], m, k);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) s += a[c];
    if (s == 12345.678) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int C>
void run(int warps_per_sm, int sms) {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    const int iters = 1 << 14;
    int block = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    int blocks_per_sm = warps_per_sm * 32 / block;
    chains<C><<<sms * blocks_per_sm, block>>>(sink, 256, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chains<C><<<sms * blocks_per_sm, block>>>(sink, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double dfma = (double)iters * C * warps_per_sm * 32 * sms;
    printf("warps/SM %2d chains %d : %7.2f TFLOP/s  cycles/iter(block0) %.2f  => %.2f cyc per dependent DFMA step, %.1f DFMA/clk/SM\n",
           warps_per_sm, C, 2 * dfma / (ms * 1e-3) / 1e12, (double)h / iters, (double)h / iters,
           (double)iters * C * warps_per_sm * 32 / (double)h);
    cudaFree(sink); cudaFree(cyc);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    for (int w : {1, 4, 8, 12, 16, 32, 64}) {
        run<1>(w, sms); run<2>(w, sms); run<4>(w, sms); run<8>(w, sms);
    }
    return 0;
}
