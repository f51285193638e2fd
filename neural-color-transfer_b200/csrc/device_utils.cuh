// Small device-side building blocks shared by several translation units.
#pragma once
#include "nct_internal.h"

__host__ __device__ __forceinline__ uint32_t nct_xy_to_int(int x, int y) { return ((uint32_t)y << 12) | (uint32_t)x; }
__host__ __device__ __forceinline__ int nct_int_to_x(uint32_t v) { return (int)(v & 0xFFFu); }
__host__ __device__ __forceinline__ int nct_int_to_y(uint32_t v) { return (int)((v >> 12) & 0xFFFu); }

__device__ __forceinline__ float nct_butterfly(float acc)
{
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return acc;
}

// Exclusive prefix sum of n ints: out[i] = sum_{j<i} in[j], out[n] = total (out has n+1 entries).
// in and out may alias only if identical pointers are NOT used (out is n+1 long).
int nct_exclusive_scan_i32(nct_ctx *ctx, const int *in_dev, int *out_dev, int n);

// Inverse lists of an NNF: for every target pixel t in [0, n_tgt) the ascending list of source
// pixels s in [0, n_src) with nnf[s] == t.  start has n_tgt + 1 entries, list has n_src entries.
// Both live in ctx scratch ("inv_start", "inv_list").
int nct_build_inverse_nnf(nct_ctx *ctx, const uint32_t *nnf_dev, int n_src, int tgt_w, int n_tgt, const int **start_dev,
                          const int **list_dev);
