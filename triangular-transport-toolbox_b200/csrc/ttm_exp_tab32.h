// 2^(j/32 - 1021), j = 0..31: correctly rounded 2^(j/32) (80-digit decimal arithmetic) scaled by 2^-1021, split into
// low and high 32-bit words.  Two 128-byte arrays: 32 lanes reading 32-bit words of one array touch 32 distinct
// banks (equal indices broadcast), so a look-up costs two conflict-free shared-memory wavefronts; a 64-bit table
// with >16 entries costs ~6 (measured: the 1024-entry variant made the kernel shared-memory bound).
#pragma once
__device__ const unsigned int g_ttm_exp2_tab32_lo[32] = {
    0x00000000u, 0xd3158574u, 0x6cf9890fu, 0xd0125b51u, 0x3c7d517bu, 0x3168b9aau, 0x6e756238u, 0xf51fdee1u, 0x0a31b715u, 0x373aa9cbu, 0x4c123422u, 0x6061892du, 0xd5362a27u, 0x569d4f82u, 0xdd485429u, 0xb03a5585u, 0x667f3bcdu, 0xe8ec5f74u, 0x73eb0187u, 0x994cce13u, 0x422aa0dbu, 0xb0cdc5e5u, 0x82a3f090u, 0xb23e255du, 0x995ad3adu, 0xf2fb5e47u, 0xdd85529cu, 0xdcef9069u, 0xdcfba487u, 0x337b9b5fu, 0xa2a490dau, 0x5b6e4540u};
__device__ const unsigned int g_ttm_exp2_tab32_hi[32] = {
    0x00200000u, 0x002059b0u, 0x0020b558u, 0x00211301u, 0x002172b8u, 0x0021d487u, 0x0022387au, 0x00229e9du, 0x002306feu, 0x002371a7u, 0x0023dea6u, 0x00244e08u, 0x0024bfdau, 0x0025342bu, 0x0025ab07u, 0x0026247eu, 0x0026a09eu, 0x00271f75u, 0x0027a114u, 0x00282589u, 0x0028ace5u, 0x00293737u, 0x0029c491u, 0x002a5503u, 0x002ae89fu, 0x002b7f76u, 0x002c199bu, 0x002cb720u, 0x002d5818u, 0x002dfc97u, 0x002ea4afu, 0x002f5076u};
