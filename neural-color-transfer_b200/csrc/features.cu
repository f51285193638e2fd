// Feature-volume helpers: layout transposes and the fused L2 normalisation.
//
// nct_l2norm replaces norm() of NCT/GeneralizedPatchMatch.cu:237-283, which is
// caffe_gpu_mul + gemv(ones) + powx(0.5) + gemm(ones x dis) + div with four
// cudaMalloc/cudaFree per call (3 calls per level).  Here: one warp per pixel reads the
// C-float row once (coalesced LDG.128), reduces sum-of-squares in the canonical 32-slot
// order of oracle/pm_oracle.c (decision D2/D5) and writes the normalised row.
// HBM-bound: 8 bytes per element (read + write).
#include "nct_internal.h"
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ float butterfly(float acc)
{
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return acc;
}

// four normalised values -> the output row: FP32 as they are, or rounded to FP16 (round to nearest even: the FP16
// feature store of the PatchMatch volumes, nct_l2norm_f16)
__device__ __forceinline__ void store4(float *row, int vi, float4 o) { reinterpret_cast<float4 *>(row)[vi] = o; }
__device__ __forceinline__ void store4(__half *row, int vi, float4 o)
{
    const __half2 lo = __floats2half2_rn(o.x, o.y), hi = __floats2half2_rn(o.z, o.w);
    uint2 u;
    u.x = *reinterpret_cast<const unsigned *>(&lo);
    u.y = *reinterpret_cast<const unsigned *>(&hi);
    reinterpret_cast<uint2 *>(row)[vi] = u;
}

// one warp per pixel; V = C/4 vectors, lane handles vectors lane, lane+32, ...
template <typename O>
__global__ void __launch_bounds__(256) l2norm_kernel(const float *__restrict__ src, O *__restrict__ dst, int npix,
                                                     int C)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    const int V = C >> 2;
    for (int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < npix; p += warps_per_grid) {
        const float4 *x = reinterpret_cast<const float4 *>(src + (size_t)p * C);
        O *y = dst + (size_t)p * C;
        float4 v[4];  // C <= 512 -> at most 4 vectors per lane
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int vi = lane + 32 * k;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vi < V) v[k] = x[vi];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int vi = lane + 32 * k;
            if (vi < V) {
                acc = __fmaf_rn(v[k].x, v[k].x, acc);
                acc = __fmaf_rn(v[k].y, v[k].y, acc);
                acc = __fmaf_rn(v[k].z, v[k].z, acc);
                acc = __fmaf_rn(v[k].w, v[k].w, acc);
            }
        }
        const float ss = butterfly(acc);
        const float nrm = __fsqrt_rn(ss);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int vi = lane + 32 * k;
            if (vi < V) {
                float4 o;
                if (ss > 0.f) {
                    o.x = __fdiv_rn(v[k].x, nrm);
                    o.y = __fdiv_rn(v[k].y, nrm);
                    o.z = __fdiv_rn(v[k].z, nrm);
                    o.w = __fdiv_rn(v[k].w, nrm);
                } else {
                    o = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                store4(y, vi, o);
            }
        }
    }
}

// tiled transpose of a [rows][cols] matrix -> [cols][rows]; 32x32 tile through padded smem
template <bool ROWS_ON_X>
__global__ void transpose_kernel(const float *__restrict__ src, float *__restrict__ dst, int rows, int cols)
{
    __shared__ float tile[32][33];
    const int bx = (ROWS_ON_X ? blockIdx.y : blockIdx.x) * 32, by = (ROWS_ON_X ? blockIdx.x : blockIdx.y) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) dst[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

}  // namespace

extern "C" {

int nct_l2norm(nct_ctx *ctx, const float *src, float *dst, int C, int H, int W)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src && dst, "null device pointer");
    NCT_REQUIRE(ctx, C > 0 && C % 4 == 0 && C <= 512, "channel count %d must be a multiple of 4, <= 512", C);
    NCT_REQUIRE(ctx, H > 0 && W > 0, "bad size");
    const int npix = H * W;
    int blocks = nct_div_up(npix, 8);
    const int max_blocks = ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    l2norm_kernel<float><<<blocks, 256, 0, ctx->stream>>>(src, dst, npix, C);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

// the same normalisation (bit-identical FP32 values), each value then rounded to FP16 (RN-even): the PatchMatch volumes
// of the FP16 feature store.  dst: C*H*W halves (uint16 storage).
int nct_l2norm_f16(nct_ctx *ctx, const float *src, uint16_t *dst, int C, int H, int W)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src && dst, "null device pointer");
    NCT_REQUIRE(ctx, C > 0 && C % 4 == 0 && C <= 512, "channel count %d must be a multiple of 4, <= 512", C);
    NCT_REQUIRE(ctx, H > 0 && W > 0, "bad size");
    const int npix = H * W;
    int blocks = nct_div_up(npix, 8);
    const int max_blocks = ctx->num_sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    l2norm_kernel<__half><<<blocks, 256, 0, ctx->stream>>>(src, reinterpret_cast<__half *>(dst), npix, C);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_chw_to_hwc(nct_ctx *ctx, const float *src, float *dst, int C, int H, int W)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src && dst && src != dst && C > 0 && H > 0 && W > 0, "bad arguments");
    const int rows = C, cols = H * W;  // [C][HW] -> [HW][C]
    dim3 block(32, 8), grid(nct_div_up(cols, 32), nct_div_up(rows, 32));
    transpose_kernel<false><<<grid, block, 0, ctx->stream>>>(src, dst, rows, cols);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_hwc_to_chw(nct_ctx *ctx, const float *src, float *dst, int C, int H, int W)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src && dst && src != dst && C > 0 && H > 0 && W > 0, "bad arguments");
    const int rows = H * W, cols = C;  // [HW][C] -> [C][HW]
    dim3 block(32, 8), grid(nct_div_up(rows, 32), nct_div_up(cols, 32));
    transpose_kernel<true><<<grid, block, 0, ctx->stream>>>(src, dst, rows, cols);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // extern "C"
