// exp(x) for the FP64-bound kernels (tile K-objgrad, fused inverse): branch-free, interleavable, and cheap on
// SHARED MEMORY as well as on the FP64 pipe.
//
//   exp(x) = 2^n * 2^(j/32) * e^r,   32 x / ln 2 = 32 n + j + f,   r = x - (32 n + j) ln2/32,  |r| <= ln2/64
//
//   * degree-6 Taylor polynomial of e^r (truncation 3.5e-18 relative),
//   * 32-entry table of correctly rounded 2^(j/32), pre-scaled by 2^-1021 and split into low/high words
//     (ttm_exp_tab32.h): a look-up is two conflict-free 32-bit LDS.  ncu showed that a 64-bit table with more than 16
//     entries costs ~6 shared-memory wavefronts per look-up (random 8-byte words, 2.9-way conflicts per half warp),
//     which made a 1024-entry/degree-3 variant shared-memory bound: 3 fewer FP64 instructions, no faster;
//   * ONE fused reduction step: r carries the representation error of ln2/32, i.e. the result is exp of an argument
//     perturbed by a relative 2^-54 (half an ulp of the argument): relative error <= 1 ulp + |x| 2^-54;
//   * the rounding constant carries the offset 1021*32, so the low word of t is 32 (n + 1021) + j, non-negative in
//     range: one clamp (VIMNMX[.RELU]), one shift and one IMAD insert the binary exponent.
// 10 FP64 + 7 other instructions.  Domain: |x| < 4.6e7 (the rounded multiple must fit 31 bits); results saturate at
// 2^-1021 / 2^1023 instead of underflowing / overflowing (the objective is inf or the integrand below 4.5e-308 long
// before); NaN in -> NaN out.
#pragma once

#include <cuda_runtime.h>

#include "ttm_exp_tab32.h"

namespace ttm_exp32 {

#define TTM_E32_K 46.16624130844683
#define TTM_E32_C 0.02166084939249829
#define TTM_E32_OFF 32672                                   /* 1021 * 32 */
#define TTM_E32_MAGIC (6755399441055744.0 + 32672.0)
#define TTM_E32_TOP (2044 * 32 + 31)

// shared-memory copy of the table: [0..32) low words, [32..64) high words
__device__ __forceinline__ void stage_table(unsigned int* s_tab, int tid, int nthreads) {
    for (int j = tid; j < 64; j += nthreads) s_tab[j] = (j < 32) ? g_ttm_exp2_tab32_lo[j] : g_ttm_exp2_tab32_hi[j - 32];
}

// hi + (c >> 5) * 2^20 as SHF + IMAD (the compiler's own strength reduction takes three instructions)
__device__ __forceinline__ int insert_exponent(int hi, int c) {
    int n, out;
    asm("shr.s32 %0, %1, 5;" : "=r"(n) : "r"(c));
    asm("mad.lo.s32 %0, %1, 1048576, %2;" : "=r"(out) : "r"(n), "r"(hi));
    return out;
}

// p = 2^(j/32 - 1021) e^r, m = 32 (n + 1021) + j   (written over L independent arguments so that the dependent
// FP64 chains are interleaved at source level)
template <int L>
__device__ __forceinline__ void core_v(const double (&x)[L], double (&p)[L], int (&m)[L],
                                       const unsigned int* __restrict__ tab) {
    double t[L], nf[L], r[L], tb[L];
#pragma unroll
    for (int l = 0; l < L; ++l) t[l] = fma(x[l], TTM_E32_K, TTM_E32_MAGIC);
#pragma unroll
    for (int l = 0; l < L; ++l) {
        m[l] = __double2loint(t[l]);
        nf[l] = t[l] - TTM_E32_MAGIC;
    }
#pragma unroll
    for (int l = 0; l < L; ++l) tb[l] = __hiloint2double((int)tab[32 + (m[l] & 31)], (int)tab[m[l] & 31]);
#pragma unroll
    for (int l = 0; l < L; ++l) r[l] = fma(nf[l], -TTM_E32_C, x[l]);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(r[l], 1.0 / 720.0, 1.0 / 120.0);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(p[l], r[l], 1.0 / 24.0);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(p[l], r[l], 1.0 / 6.0);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(p[l], r[l], 0.5);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(p[l], r[l], 1.0);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] = fma(p[l], r[l], 1.0);
#pragma unroll
    for (int l = 0; l < L; ++l) p[l] *= tb[l];
}

// argument <= 0 (Gaussian weight e^{-x^2/4})
template <int L>
__device__ __forceinline__ void exp_neg_v(const double (&x)[L], double (&out)[L], const unsigned int* __restrict__ tab) {
    double p[L];
    int m[L];
    core_v<L>(x, p, m, tab);
#pragma unroll
    for (int l = 0; l < L; ++l)
        out[l] = __hiloint2double(insert_exponent(__double2hiint(p[l]), max(m[l], 0)), __double2loint(p[l]));
}

// general argument (rectifier)
template <int L>
__device__ __forceinline__ void exp_gen_v(const double (&x)[L], double (&out)[L], const unsigned int* __restrict__ tab) {
    double p[L];
    int m[L];
    core_v<L>(x, p, m, tab);
#pragma unroll
    for (int l = 0; l < L; ++l)
        out[l] = __hiloint2double(insert_exponent(__double2hiint(p[l]), __vimin_s32_relu(m[l], TTM_E32_TOP)),
                                  __double2loint(p[l]));
}

__device__ __forceinline__ double exp_neg_1(double x, const unsigned int* __restrict__ tab) {
    double a[1] = {x}, o[1];
    exp_neg_v<1>(a, o, tab);
    return o[0];
}
__device__ __forceinline__ double exp_gen_1(double x, const unsigned int* __restrict__ tab) {
    double a[1] = {x}, o[1];
    exp_gen_v<1>(a, o, tab);
    return o[0];
}

}  // namespace ttm_exp32
