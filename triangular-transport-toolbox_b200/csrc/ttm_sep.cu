// Separable-monotonicity kernels: K-sep-eval (S_k, d_k S_k), K-gram (DMMA), K-sepobj.
//
// Reference: s (separable arm) transport_map.py:2550-2558; log-determinants :2618-2641, :2686-2709;
// worker_task_monotone :2903-3172 (Gram-type contractions :2966-2975, :3031-3050; reduced objective
// fun_mon_objective :2978-3018).

#include "ttm_common.cuh"
#include "ttm_kernels.h"
#include "ttm_sweep.cuh"

namespace {

constexpr int T_SEP = 128;

// S_i = sum_j a_j psi^non_j(Xt_i) + sum_j b_j psi^mon_j(Xt_i);  dS_i = sum_j b_j dpsi^mon_j(Xd_i)
__global__ void __launch_bounds__(T_SEP) sep_eval_kernel(const PlanView P, const double* __restrict__ Xt, int64_t ld,
                                                         int64_t N, const double* __restrict__ coeffs,
                                                         double* __restrict__ S_out, const double* __restrict__ Xd,
                                                         int64_t ldd, double* __restrict__ dS_out) {
    extern __shared__ double s_coef[];
    const int m = P.m_non + P.m_mon;
    for (int j = threadIdx.x; j < m; j += T_SEP) s_coef[j] = coeffs[j];
    __syncthreads();
    const double* acoef = s_coef;
    const double* bcoef = s_coef + P.m_non;
    const int64_t rows = (N + T_SEP - 1) / T_SEP;
    for (int64_t row0 = (int64_t)blockIdx.x * R_OBJ; row0 < rows; row0 += (int64_t)gridDim.x * R_OBJ) {
        int64_t idx[R_OBJ];
        bool ok[R_OBJ];
        double S[R_OBJ];
#pragma unroll
        for (int r = 0; r < R_OBJ; ++r) {
            const int64_t i = (row0 + r) * T_SEP + threadIdx.x;
            ok[r] = (row0 + r < rows) && (i < N);
            idx[r] = ok[r] ? i : N - 1;
            S[r] = 0.0;
        }
        if (S_out) {
            nonmon_sweep_rt<false>(P, Xt, ld, idx, acoef, S, nullptr, 0);
            for (int j = 0; j < P.m_mon; ++j) {
                const double b = bcoef[j];
#pragma unroll
                for (int r = 0; r < R_OBJ; ++r)
                    S[r] = fma(b, plan_term(P, P.o_mon_ptr, P.o_mon_fac, j, Xt, ld, idx[r]), S[r]);
            }
#pragma unroll
            for (int r = 0; r < R_OBJ; ++r)
                if (ok[r]) S_out[idx[r]] = S[r];
        }
        if (dS_out) {
            double dS[R_OBJ];
#pragma unroll
            for (int r = 0; r < R_OBJ; ++r) dS[r] = 0.0;
            for (int j = 0; j < P.m_dmon; ++j) {
                const double b = bcoef[j];
#pragma unroll
                for (int r = 0; r < R_OBJ; ++r)
                    dS[r] = fma(b, plan_term(P, P.o_dmon_ptr, P.o_dmon_fac, j, Xd, ldd, idx[r]), dS[r]);
            }
#pragma unroll
            for (int r = 0; r < R_OBJ; ++r)
                if (ok[r]) dS_out[idx[r]] = dS[r];
        }
    }
}

// -------------------------------------------------------------------------------------------------
// K-gram: G = Psi^T Psi with Psi = [Psi_non | Psi_mon] (N x M), generated on the fly per sample tile
// (no HBM round trip of Psi) and contracted with FP64 tensor-core DMMA (mma.sync m8n8k4).
// tcgen05 has no f64 kind, so this legacy-style warp MMA is the sm_100a tensor path for FP64.
// Each block owns a range of samples and the full M x M output; partial Grams are written to
// scratch[block] and summed in fixed block order by gram_reduce_kernel (deterministic).
// -------------------------------------------------------------------------------------------------
constexpr int G_KS = 32;   // samples per staged tile
constexpr int T_GRAM = 256;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Mp = M rounded up to a multiple of 8.  smem tile: [G_KS][Mp+?]
__global__ void __launch_bounds__(T_GRAM) gram_kernel(const PlanView P, const double* __restrict__ Xt, int64_t ld,
                                                      int64_t N, double* __restrict__ scratch, int Mp) {
    extern __shared__ double tile[];  // [G_KS][Mp + 4]
    const int M = P.m_non + P.m_mon;
    const int ldt = Mp + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = T_GRAM / 32;
    const int nt = Mp / 8;                 // 8x8 output tiles per dimension
    const int ntile = nt * (nt + 1) / 2;   // upper triangle (tj >= ti)
    // accumulators: each warp owns tiles w, w+nwarp, ... (<= MAXT per warp, else loop in passes)
    constexpr int MAXT = 24;
    const int64_t chunk = (N + gridDim.x - 1) / gridDim.x;
    const int64_t s_lo = (int64_t)blockIdx.x * chunk;
    const int64_t s_hi = (s_lo + chunk < N) ? s_lo + chunk : N;
    double* out = scratch + (int64_t)blockIdx.x * Mp * Mp;

    for (int pass0 = 0; pass0 < ntile; pass0 += nwarp * MAXT) {
        double c0[MAXT], c1[MAXT];
#pragma unroll
        for (int q = 0; q < MAXT; ++q) c0[q] = c1[q] = 0.0;
        for (int64_t s0 = s_lo; s0 < s_hi; s0 += G_KS) {
            __syncthreads();
            // stage Psi rows s0..s0+G_KS-1: thread (s, j-stride)
            for (int e = threadIdx.x; e < G_KS * Mp; e += T_GRAM) {
                const int s = e % G_KS, j = e / G_KS;
                double v = 0.0;
                const int64_t i = s0 + s;
                if (i < s_hi && j < M)
                    v = (j < P.m_non) ? plan_term(P, P.o_non_ptr, P.o_non_fac, j, Xt, ld, i)
                                      : plan_term(P, P.o_mon_ptr, P.o_mon_fac, j - P.m_non, Xt, ld, i);
                tile[s * ldt + j] = v;
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < MAXT; ++q) {
                const int t = pass0 + warp + q * nwarp;
                if (t < ntile) {
                    // unrank upper-triangular tile index t -> (ti, tj), tj >= ti
                    int ti = 0, rem = t;
                    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
                    const int tj = ti + rem;
                    const int arow = ti * 8 + (lane >> 2), bcol = tj * 8 + (lane >> 2), kk = lane & 3;
#pragma unroll
                    for (int k0 = 0; k0 < G_KS; k0 += 4)
                        dmma_m8n8k4(c0[q], c1[q], tile[(k0 + kk) * ldt + arow], tile[(k0 + kk) * ldt + bcol]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < MAXT; ++q) {
            const int t = pass0 + warp + q * nwarp;
            if (t < ntile) {
                int ti = 0, rem = t;
                while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
                const int tj = ti + rem;
                const int row = ti * 8 + (lane >> 2), col = tj * 8 + 2 * (lane & 3);
                out[(int64_t)row * Mp + col] = c0[q];
                out[(int64_t)row * Mp + col + 1] = c1[q];
            }
        }
    }
}

// G[i][j] = sum_b scratch[b][min][max] (fixed order), symmetric fill, M x M output
__global__ void gram_reduce_kernel(const double* __restrict__ scratch, int nblocks, int Mp, int M,
                                   double* __restrict__ G) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * M) return;
    const int i = e / M, j = e % M;
    const int r = min(i, j), c = max(i, j);
    // element (r, c) lives in tile (r/8, c/8) with tj >= ti: always stored
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += scratch[(int64_t)b * Mp * Mp + (int64_t)r * Mp + c];
    G[e] = s;
}

// -------------------------------------------------------------------------------------------------
// K-sepobj: out[0] = sum_i log dS_i, out[1+j] = sum_i dpsi_ij / dS_i,
//           dS_i = sum_j (b_j + delta) dpsi_ij      (transport_map.py:2990-3006)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T_SEP) sepobj_kernel(const PlanView P, const double* __restrict__ Xt, int64_t ld,
                                                       int64_t N, const double* __restrict__ b, double delta,
                                                       double* __restrict__ partials, unsigned int* counter,
                                                       double* __restrict__ out) {
    extern __shared__ double sm[];
    const int mm = P.m_dmon;
    double* s_b = sm;                 // [mm]
    double* s_acc = sm + mm;          // [mm][T_SEP]
    double* s_red = s_acc + mm * T_SEP;  // [T_SEP/32][1+mm]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < mm; j += T_SEP) s_b[j] = b[j] + delta;
    for (int j = 0; j < mm; ++j) s_acc[j * T_SEP + tid] = 0.0;
    __syncthreads();
    double lacc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * T_SEP + tid; i < N; i += (int64_t)gridDim.x * T_SEP) {
        double dS = 0.0;
        for (int j = 0; j < mm; ++j) dS = fma(s_b[j], plan_term(P, P.o_dmon_ptr, P.o_dmon_fac, j, Xt, ld, i), dS);
        lacc += log(dS);
        const double inv = 1.0 / dS;
        for (int j = 0; j < mm; ++j)
            s_acc[j * T_SEP + tid] += plan_term(P, P.o_dmon_ptr, P.o_dmon_fac, j, Xt, ld, i) * inv;
    }
    lacc = warp_sum(lacc);
    if (lane == 0) s_red[warp * (1 + mm)] = lacc;
    for (int j = 0; j < mm; ++j) {
        const double v = warp_sum(s_acc[j * T_SEP + tid]);
        if (lane == 0) s_red[warp * (1 + mm) + 1 + j] = v;
    }
    __syncthreads();
    double* part = partials + (int64_t)blockIdx.x * (1 + mm);
    for (int j = tid; j < 1 + mm; j += T_SEP) {
        double v = 0.0;
        for (int w = 0; w < T_SEP / 32; ++w) v += s_red[w * (1 + mm) + j];
        part[j] = v;
    }
    __threadfence();
    __shared__ unsigned int s_last;
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int j = tid; j < 1 + mm; j += T_SEP) {
            double v = 0.0;
            for (unsigned int bk = 0; bk < gridDim.x; ++bk) v += __ldcg(partials + (int64_t)bk * (1 + mm) + j);
            out[j] = v;
        }
        if (tid == 0) *counter = 0u;
    }
}

// change-of-variables bookkeeping of the density evaluators (transport_map.py:2618-2644, :2680-2712)
__global__ void density_acc_kernel(double* __restrict__ acc, const double* __restrict__ S,
                                   const double* __restrict__ dS, double sigma, int mode, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double ld = log(dS[i] / sigma);
    if (mode == 0) acc[i] += -0.5 * S[i] * S[i] - 0.91893853320467274178 + ld;   // log N(z;0,1) + log dS/sigma
    else acc[i] -= ld;
}

__global__ void density_finish_kernel(const double* __restrict__ acc, const double* __restrict__ logt,
                                      double* __restrict__ out, int64_t N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    out[i] = exp(acc[i] + (logt ? logt[i] : 0.0));
}

}  // namespace

cudaError_t ttm_launch_density_acc(double* acc, const double* S, const double* dS, double sigma, int mode, int64_t N,
                                   cudaStream_t st) {
    density_acc_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, S, dS, sigma, mode, N);
    return cudaGetLastError();
}

cudaError_t ttm_launch_density_finish(const double* acc, const double* logt, double* out, int64_t N, cudaStream_t st) {
    density_finish_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, logt, out, N);
    return cudaGetLastError();
}

cudaError_t ttm_launch_sep_eval(const PlanView& P, const double* Xt, int64_t ld, int64_t N, const double* coeffs,
                                double* S_out, const double* Xd, int64_t ldd, double* dS_out, cudaStream_t st) {
    if (N == 0) return cudaSuccess;
    const int64_t rows = (N + T_SEP - 1) / T_SEP;
    int64_t grid = (rows + R_OBJ - 1) / R_OBJ;
    if (grid > 148 * 16) grid = 148 * 16;
    const size_t smem = sizeof(double) * (size_t)(P.m_non + P.m_mon);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sep_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    sep_eval_kernel<<<(unsigned)grid, T_SEP, smem, st>>>(P, Xt, ld, N, coeffs, S_out, Xd, ldd, dS_out);
    return cudaGetLastError();
}

cudaError_t ttm_launch_gram(const PlanView& P, const double* Xt, int64_t ld, int64_t N, double* G, double* scratch,
                            int64_t scratch_doubles, int sm_count, cudaStream_t st) {
    const int M = P.m_non + P.m_mon;
    if (M == 0) return cudaSuccess;
    const int Mp = (M + 7) / 8 * 8;
    int64_t grid = scratch_doubles / ((int64_t)Mp * Mp);
    if (grid > sm_count) grid = sm_count;
    const int64_t max_by_n = (N + G_KS - 1) / G_KS;
    if (grid > max_by_n) grid = max_by_n;
    if (grid < 1) return cudaErrorInvalidValue;
    const size_t smem = sizeof(double) * (size_t)G_KS * (Mp + 4);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    gram_kernel<<<(unsigned)grid, T_GRAM, smem, st>>>(P, Xt, ld, N, scratch, Mp);
    gram_reduce_kernel<<<(M * M + 255) / 256, 256, 0, st>>>(scratch, (int)grid, Mp, M, G);
    return cudaGetLastError();
}

cudaError_t ttm_launch_sepobj(const PlanView& P, const double* Xt, int64_t ld, int64_t N, const double* b,
                              double delta, double* partials, unsigned int* counter, double* out, int max_grid,
                              int sm_count, cudaStream_t st) {
    const int mm = P.m_dmon;
    int64_t grid = (N + T_SEP - 1) / T_SEP;
    if (grid > (int64_t)sm_count * 8) grid = (int64_t)sm_count * 8;
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    const size_t smem = sizeof(double) * (size_t)(mm + mm * T_SEP + (T_SEP / 32) * (1 + mm));
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(sepobj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    sepobj_kernel<<<(unsigned)grid, T_SEP, smem, st>>>(P, Xt, ld, N, b, delta, partials, counter, out);
    return cudaGetLastError();
}
