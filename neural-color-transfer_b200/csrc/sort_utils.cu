// Stable key-value radix sort used to build deterministic reverse adjacency lists.
// This is plumbing, not a hot op: CUB's DeviceRadixSort (header-only, ships with the CUDA toolkit).
#include "nct_internal.h"
#include <cub/device/device_radix_sort.cuh>

int nct_sort_pairs_u32(nct_ctx *ctx, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out,
                       int n, int end_bit)
{
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, ctx->stream);
    if (e != cudaSuccess) return nct_fail(ctx, NCT_ERR_CUDA, "cub SortPairs (size query): %s", cudaGetErrorString(e));
    void *tmp = nct_scratch(ctx, "sort_tmp", tmp_bytes + 16);
    if (!tmp) return NCT_ERR_NOMEM;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, n, 0, end_bit, ctx->stream);
    if (e != cudaSuccess) return nct_fail(ctx, NCT_ERR_CUDA, "cub SortPairs: %s", cudaGetErrorString(e));
    ctx->launches += 3;
    return NCT_OK;
}
