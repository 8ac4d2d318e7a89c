// Pipe micro-benchmarks for the roofline denominators and the issue model of the FP64-bound kernels (B200, sm_100a):
//   dfma          dependent-free DFMA chains                         -> FP64 vector peak (2 flop per DFMA)
//   dfma+int{1,2} the same with 1 / 2 independent integer instructions per DFMA: does integer work issue in the
//                 shadow of the 2-cycle DFMA dispatch?
//   dfma+lds      one conflict-free LDS.64 per 4 DFMA
//   dmma_*        mma.sync FP64 tensor-core shapes (m8n8k4, m16n8k4, m16n8k8, m16n8k16) -> DMMA peak
//   shfl          warp shuffles per clock per SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
// Prints one JSON object; bench.py / the profiles copy it to profiles/fp64_peaks_r2.json with the clock record.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// SM clock actually running during the probe: clock64 ticks of one block over a cudaEvent-timed kernel
__global__ void k_clock(long long* ticks, double* sink, int iters) {
    const long long t0 = clock64();
    double a = threadIdx.x * 1e-3;
    for (int it = 0; it < iters; ++it) a = fma(a, 1.0000001, 1e-9);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) ticks[0] = t1 - t0;
    if (a == 123.456) sink[0] = a;
}

template <int NINT, bool LDS>
__global__ void k_dfma(double* sink, int iters) {
    __shared__ double tab[256];
    tab[threadIdx.x & 255] = threadIdx.x * 1e-9;
    __syncthreads();
    double a[8];
    int u[4];
    const double b = 1.0000001, c = 1e-9;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = fma(a[j], b, c);
            if (NINT >= 1) u[j & 3] = (u[j & 3] ^ (it + j)) + 0x9e3779b9;          // LOP3 + IADD (1 "pair" ~ 2 instr)
            if (NINT >= 2) u[(j + 1) & 3] = __funnelshift_l(u[(j + 1) & 3], u[j & 3], 7) - j;
            if (LDS && (j & 3) == 0) a[j] += tab[(u[0] + j) & 255];
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456 || u[0] + u[1] + u[2] + u[3] == 0x7fffffff) sink[0] = s;
}

__global__ void k_shfl(double* sink, int iters) {
    int v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __shfl_xor_sync(0xffffffffu, v[j], 1 + (j & 15));
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 0x7fffffff) sink[0] = s;
}

template <int SHAPE>   // 0: m8n8k4, 1: m16n8k4, 2: m16n8k8, 3: m16n8k16
__global__ void k_dmma(double* sink, int iters) {
    constexpr int NACC = 4;   // independent accumulator tiles per warp
    double d[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = 1.0 - 1e-9 * (threadIdx.x + j);
#pragma unroll
    for (int t = 0; t < NACC; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[t][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < NACC; ++t) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(d[t][0]), "+d"(d[t][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(d[t][0]), "+d"(d[t][1]), "+d"(d[t][2]), "+d"(d[t][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(d[t][0]), "+d"(d[t][1]), "+d"(d[t][2]), "+d"(d[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(d[t][0]), "+d"(d[t][1]), "+d"(d[t][2]), "+d"(d[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < NACC; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += d[t][j];
    if (s == 123.456) sink[0] = s;
}

template <typename F>
double time_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double* sink;
    CK(cudaMalloc(&sink, 8));
    const int sm = p.multiProcessorCount, block = 256, grid = sm * 8, iters = 1 << 14;
    const double threads = (double)grid * block;
    printf("{\"gpu\": \"%s\", \"sm_count\": %d, \"clock_rate_mhz\": %.0f", p.name, sm, clk_khz / 1e3);
    {
        // warm the clocks with ~1 s of DFMA work, then measure the SM clock seen by a long dependent chain
        for (int rep = 0; rep < 200; ++rep) k_dfma<0, false><<<grid, block>>>(sink, iters);
        CK(cudaDeviceSynchronize());
        long long* ticks;
        CK(cudaMalloc(&ticks, 8));
        const double ms = time_ms([&] { k_clock<<<sm, block>>>(ticks, sink, 1 << 20); });   // one wave: block 0 spans the kernel
        long long h = 0;
        CK(cudaMemcpy(&h, ticks, 8, cudaMemcpyDeviceToHost));
        printf(", \"sm_clock_mhz_under_load\": %.0f", h / (ms * 1e-3) / 1e6);
    }
    {
        const double ms = time_ms([&] { k_dfma<0, false><<<grid, block>>>(sink, iters); });
        printf(", \"dfma_tflops\": %.3f", 2.0 * 8 * iters * threads / (ms * 1e-3) / 1e12);
        const double ms1 = time_ms([&] { k_dfma<1, false><<<grid, block>>>(sink, iters); });
        const double ms2 = time_ms([&] { k_dfma<2, false><<<grid, block>>>(sink, iters); });
        const double ms3 = time_ms([&] { k_dfma<0, true><<<grid, block>>>(sink, iters); });
        const double ms4 = time_ms([&] { k_dfma<2, true><<<grid, block>>>(sink, iters); });
        printf(", \"dfma_ms\": %.4f, \"dfma_int1_ms\": %.4f, \"dfma_int2_ms\": %.4f, \"dfma_lds_ms\": %.4f, \"dfma_int2_lds_ms\": %.4f",
               ms, ms1, ms2, ms3, ms4);
    }
    {
        const double ms = time_ms([&] { k_shfl<<<grid, block>>>(sink, iters); });
        printf(", \"shfl_warp_instr_per_clk_per_sm\": %.3f",
               8.0 * iters * threads / 32 / (ms * 1e-3) / (clk_khz * 1e3) / sm);
    }
    {
        const double warps = threads / 32;
        const double fl[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
        const char* nm[4] = {"dmma_m8n8k4_tflops", "dmma_m16n8k4_tflops", "dmma_m16n8k8_tflops", "dmma_m16n8k16_tflops"};
        double ms[4];
        ms[0] = time_ms([&] { k_dmma<0><<<grid, block>>>(sink, iters / 4); });
        ms[1] = time_ms([&] { k_dmma<1><<<grid, block>>>(sink, iters / 4); });
        ms[2] = time_ms([&] { k_dmma<2><<<grid, block>>>(sink, iters / 4); });
        ms[3] = time_ms([&] { k_dmma<3><<<grid, block>>>(sink, iters / 4); });
        for (int s = 0; s < 4; ++s) printf(", \"%s\": %.3f", nm[s], fl[s] * 4 * (iters / 4) * warps / (ms[s] * 1e-3) / 1e12);
    }
    printf("}\n");
    return 0;
}
