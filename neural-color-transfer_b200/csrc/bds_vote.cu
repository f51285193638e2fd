// Bidirectional-similarity (BDS) votes on the GPU.
//
// Replaces, per level (NCT/main.cu:286-318):
//   * the four D2H copies of ann/annd/bnn/bnnd + the serial host loops of reconstruct_bds
//     (NCT/GeneralizedPatchMatch.cu:122-235)                      -> nct_reconstruct_bds
//   * avg_vote_bds_a + avg_vote_bds_b (float atomics) + avg_vote_bds + norm() + feature_distance
//     (NCT/GeneralizedPatchMatch.cu:1074-1202, 237-283, 833-855)   -> nct_bds_feature_error
//
// Design: the completeness vote (each B pixel scatters its 3x3 patch through bnn) is turned into a
// GATHER through the inverse lists of bnn (ascending B index per A pixel), so the result is
// deterministic -- the reference's float atomicAdd order is not.  The feature vote, the division by
// the weight, the L2 normalisation and the dot product with the content feature are one kernel: one
// warp per A pixel, lanes own channels (coalesced LDG.128 rows of S), nothing but err[p] is written
// (the reference materialises vote_Ndata_C1, copy_Ndata_C1 and a normalised copy: 3 x C x HW floats).
// Bound: HBM/L2 gather of <= 18 + (inverse list length) rows of C floats per A pixel.
#include "device_utils.cuh"

namespace {

// ---------------------------------------------------------------- exclusive scan (3 phases)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int base = w > 0 ? warp_sums[w - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + x - v;
}

__global__ void scan_tiles_kernel(const int *__restrict__ in, int *__restrict__ out, int *__restrict__ tile_sums, int n)
{
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int total;
    int ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place, grand total to tile_sums[ntiles]
__global__ void scan_sums_kernel(int *__restrict__ tile_sums, int ntiles)
{
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int start = 0; start < ntiles; start += SCAN_THREADS) {
        const int i = start + threadIdx.x;
        const int v = i < ntiles ? tile_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total);
        const int c = carry;
        if (i < ntiles) tile_sums[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = carry;
}

__global__ void scan_add_kernel(int *__restrict__ out, const int *__restrict__ tile_sums, int n, int ntiles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sums[i / SCAN_TILE];
    if (i == 0) out[n] = tile_sums[ntiles];
}

// ---------------------------------------------------------------- inverse lists
__global__ void inv_count_kernel(const uint32_t *__restrict__ nnf, int n_src, int tgt_w, int *__restrict__ count)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_src) {
        const uint32_t v = nnf[s];
        atomicAdd(&count[nct_int_to_y(v) * tgt_w + nct_int_to_x(v)], 1);
    }
}

__global__ void inv_fill_kernel(const uint32_t *__restrict__ nnf, int n_src, int tgt_w, const int *__restrict__ start,
                                int *__restrict__ cursor, int *__restrict__ list)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_src) {
        const uint32_t v = nnf[s];
        const int t = nct_int_to_y(v) * tgt_w + nct_int_to_x(v);
        const int pos = atomicAdd(&cursor[t], 1);
        list[start[t] + pos] = s;
    }
}

// ascending order inside every list (the fill order depends on atomics): insertion sort, lists are short
__global__ void inv_sort_kernel(const int *__restrict__ start, int *__restrict__ list, int n_tgt)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tgt) return;
    const int b = start[t], e = start[t + 1];
    for (int i = b + 1; i < e; ++i) {
        const int key = list[i];
        int j = i - 1;
        while (j >= b && list[j] > key) {
            list[j + 1] = list[j];
            --j;
        }
        list[j + 1] = key;
    }
}

// ---------------------------------------------------------------- 8-bit colour reconstruction
__global__ void reconstruct_bds_kernel(const uint8_t *__restrict__ b_img, const uint32_t *__restrict__ ann,
                                       const int *__restrict__ inv_start, const int *__restrict__ inv_list, int ah, int aw,
                                       int bh, int bw, double wa, double wb, uint8_t *__restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ah * aw) return;
    const int ax = p % aw, ay = p / aw;
    int asum[3] = {0, 0, 0}, bsum[3] = {0, 0, 0}, na = 0, nb = 0;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const int nx = ax + dx, ny = ay + dy;
            if (nx < aw && nx >= 0 && ny < ah && ny >= 0) {
                const uint32_t vp = ann[ny * aw + nx];
                const int xp = nct_int_to_x(vp) - dx, yp = nct_int_to_y(vp) - dy;
                if (xp < bw && xp >= 0 && yp < bh && yp >= 0) {
                    const uint8_t *bv = b_img + ((size_t)yp * bw + xp) * 3;
                    asum[0] += bv[0]; asum[1] += bv[1]; asum[2] += bv[2];
                    na++;
                }
            }
            // completeness: B pixels whose match is (ax-dx, ay-dy) put their pixel at offset (dx,dy) here
            const int x0 = ax - dx, y0 = ay - dy;
            if (x0 >= 0 && x0 < aw && y0 >= 0 && y0 < ah) {
                const int a0 = y0 * aw + x0;
                for (int t = inv_start[a0]; t < inv_start[a0 + 1]; ++t) {
                    const int b = inv_list[t];
                    const int xb = b % bw + dx, yb = b / bw + dy;
                    if (xb < bw && xb >= 0 && yb < bh && yb >= 0) {
                        const uint8_t *bv = b_img + ((size_t)yb * bw + xb) * 3;
                        bsum[0] += bv[0]; bsum[1] += bv[1]; bsum[2] += bv[2];
                        nb++;
                    }
                }
            }
        }
    const double aw_ = __dmul_rn((double)na, wa), bw_ = __dmul_rn((double)nb, wb);
    const double den = __dadd_rn(aw_, bw_);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double num = __dadd_rn(__dmul_rn((double)asum[c], wa), __dmul_rn((double)bsum[c], wb));
        out[(size_t)p * 3 + c] = (uint8_t)(int)__ddiv_rn(num, den);
    }
}

// ---------------------------------------------------------------- fused feature vote -> err
// VPL = float4 vectors per lane (C = 128 * VPL for C >= 128; for C < 128 only lanes < C/4 are active)
template <int VPL>
__global__ void __launch_bounds__(256) bds_feature_error_kernel(
    const float *__restrict__ c_norm, const float *__restrict__ s_raw, const uint32_t *__restrict__ ann,
    const int *__restrict__ inv_start, const int *__restrict__ inv_list, int C, int ah, int aw, int bh, int bw, double wa,
    double wb, float *__restrict__ err, float *__restrict__ vote_out)
{
    const int lane = threadIdx.x & 31;
    const int p = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (p >= ah * aw) return;
    const int V = C >> 2;
    const int ax = p % aw, ay = p / aw;
    const bool active = lane < V;  // all lanes for C >= 128
    float4 out[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) out[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    float pw = 0.f;
    const float wbf = (float)wb;

    // coherence vote (avg_vote_bds_a): double-precision accumulate, rounded to float each step
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const int nx = ax + dx, ny = ay + dy;
            if (nx < aw && nx >= 0 && ny < ah && ny >= 0) {
                const uint32_t vp = ann[ny * aw + nx];
                const int xp = nct_int_to_x(vp) - dx, yp = nct_int_to_y(vp) - dy;
                if (xp < bw && xp >= 0 && yp < bh && yp >= 0) {
                    pw = (float)__dadd_rn((double)pw, wa);
                    if (active) {
                        const float4 *row = reinterpret_cast<const float4 *>(s_raw + ((size_t)yp * bw + xp) * C);
#pragma unroll
                        for (int k = 0; k < VPL; ++k) {
                            const float4 s = __ldg(row + lane + 32 * k);
                            out[k].x = (float)__dadd_rn((double)out[k].x, __dmul_rn((double)s.x, wa));
                            out[k].y = (float)__dadd_rn((double)out[k].y, __dmul_rn((double)s.y, wa));
                            out[k].z = (float)__dadd_rn((double)out[k].z, __dmul_rn((double)s.z, wa));
                            out[k].w = (float)__dadd_rn((double)out[k].w, __dmul_rn((double)s.w, wa));
                        }
                    }
                }
            }
        }
    // completeness vote (avg_vote_bds_b) as a gather through the inverse lists
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy) {
            const int x0 = ax - dx, y0 = ay - dy;
            if (x0 < 0 || x0 >= aw || y0 < 0 || y0 >= ah) continue;
            const int a0 = y0 * aw + x0;
            const int t1 = inv_start[a0 + 1];
            for (int t = inv_start[a0]; t < t1; ++t) {
                const int b = inv_list[t];
                const int xb = b % bw + dx, yb = b / bw + dy;
                if (xb < bw && xb >= 0 && yb < bh && yb >= 0) {
                    pw = __fadd_rn(pw, wbf);
                    if (active) {
                        const float4 *row = reinterpret_cast<const float4 *>(s_raw + ((size_t)yb * bw + xb) * C);
#pragma unroll
                        for (int k = 0; k < VPL; ++k) {
                            const float4 s = __ldg(row + lane + 32 * k);
                            out[k].x = __fadd_rn(out[k].x, (float)__dmul_rn(wb, (double)s.x));
                            out[k].y = __fadd_rn(out[k].y, (float)__dmul_rn(wb, (double)s.y));
                            out[k].z = __fadd_rn(out[k].z, (float)__dmul_rn(wb, (double)s.z));
                            out[k].w = __fadd_rn(out[k].w, (float)__dmul_rn(wb, (double)s.w));
                        }
                    }
                }
            }
        }
    // avg_vote_bds: divide by the accumulated weight
    if (pw > 0.f) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            out[k].x = __fdiv_rn(out[k].x, pw);
            out[k].y = __fdiv_rn(out[k].y, pw);
            out[k].z = __fdiv_rn(out[k].z, pw);
            out[k].w = __fdiv_rn(out[k].w, pw);
        }
    }
    if (vote_out && active) {
        float4 *row = reinterpret_cast<float4 *>(vote_out + (size_t)p * C);
#pragma unroll
        for (int k = 0; k < VPL; ++k) row[lane + 32 * k] = out[k];
    }
    // norm(): L2 norm of the voted vector, canonical order
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        acc = __fmaf_rn(out[k].x, out[k].x, acc);
        acc = __fmaf_rn(out[k].y, out[k].y, acc);
        acc = __fmaf_rn(out[k].z, out[k].z, acc);
        acc = __fmaf_rn(out[k].w, out[k].w, acc);
    }
    const float ss = nct_butterfly(acc);
    const float nrm = __fsqrt_rn(ss);
    // feature_distance: -<c_hat, vote_hat>
    acc = 0.f;
    if (active) {
        const float4 *crow = reinterpret_cast<const float4 *>(c_norm + (size_t)p * C);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const float4 c = __ldg(crow + lane + 32 * k);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ss > 0.f) {
                v.x = __fdiv_rn(out[k].x, nrm);
                v.y = __fdiv_rn(out[k].y, nrm);
                v.z = __fdiv_rn(out[k].z, nrm);
                v.w = __fdiv_rn(out[k].w, nrm);
            }
            acc = __fmaf_rn(c.x, v.x, acc);
            acc = __fmaf_rn(c.y, v.y, acc);
            acc = __fmaf_rn(c.z, v.z, acc);
            acc = __fmaf_rn(c.w, v.w, acc);
        }
    }
    const float dot = nct_butterfly(acc);
    if (lane == 0) err[p] = -dot;
}

}  // namespace

int nct_exclusive_scan_i32(nct_ctx *ctx, const int *in, int *out, int n)
{
    const int ntiles = nct_div_up(n, SCAN_TILE);
    int *tile_sums = (int *)nct_scratch(ctx, "scan_tile_sums", sizeof(int) * ((size_t)ntiles + 1));
    if (!tile_sums) return NCT_ERR_NOMEM;
    scan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, tile_sums, n);
    NCT_CHECK_LAUNCH(ctx);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(tile_sums, ntiles);
    NCT_CHECK_LAUNCH(ctx);
    scan_add_kernel<<<nct_div_up(n, 256), 256, 0, ctx->stream>>>(out, tile_sums, n, ntiles);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

__global__ void inv_keys_kernel(const uint32_t *__restrict__ nnf, int n_src, int tgt_w, uint32_t *__restrict__ keys,
                                uint32_t *__restrict__ vals, int *__restrict__ count)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_src) {
        const uint32_t v = nnf[s];
        const int t = nct_int_to_y(v) * tgt_w + nct_int_to_x(v);
        keys[s] = (uint32_t)t;
        vals[s] = (uint32_t)s;
        atomicAdd(&count[t], 1);
    }
}

int nct_sort_pairs_u32(nct_ctx *ctx, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out,
                       int n, int end_bit);

// Inverse lists by a stable radix sort of (target, source) pairs: ascending source order inside every list, no
// data-dependent serial work (a first version insertion-sorted each list in one thread: 1.5 ms per call on hub targets).
int nct_build_inverse_nnf(nct_ctx *ctx, const uint32_t *nnf, int n_src, int tgt_w, int n_tgt, const int **start_dev,
                          const int **list_dev)
{
    int *count = (int *)nct_scratch(ctx, "inv_count", sizeof(int) * ((size_t)n_tgt + 1));
    int *start = (int *)nct_scratch(ctx, "inv_start", sizeof(int) * ((size_t)n_tgt + 1));
    uint32_t *buf = (uint32_t *)nct_scratch(ctx, "inv_sort", sizeof(uint32_t) * (size_t)n_src * 4);
    if (!count || !start || !buf) return NCT_ERR_NOMEM;
    uint32_t *keys = buf, *keys_out = buf + n_src, *vals = buf + 2 * (size_t)n_src, *vals_out = buf + 3 * (size_t)n_src;
    NCT_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)n_tgt + 1), ctx->stream));
    inv_keys_kernel<<<nct_div_up(n_src, 256), 256, 0, ctx->stream>>>(nnf, n_src, tgt_w, keys, vals, count);
    NCT_CHECK_LAUNCH(ctx);
    int rc = nct_exclusive_scan_i32(ctx, count, start, n_tgt);
    if (rc) return rc;
    int bits = 1;
    while ((1 << bits) < n_tgt) bits++;
    rc = nct_sort_pairs_u32(ctx, keys, keys_out, vals, vals_out, n_src, bits);
    if (rc) return rc;
    *start_dev = start;
    *list_dev = reinterpret_cast<const int *>(vals_out);
    return NCT_OK;
}

extern "C" {

int nct_reconstruct_bds(nct_ctx *ctx, const uint8_t *a_bgr, const uint8_t *b_bgr, const uint32_t *ann, const uint32_t *bnn,
                        int ah, int aw, int bh, int bw, double w_cohen, double w_complete, uint8_t *out_bgr)
{
    NCT_ENTER(ctx);
    (void)a_bgr;  // the reference only uses a's size (Mat::zeros(a.size(), ...))
    NCT_REQUIRE(ctx, b_bgr && ann && bnn && out_bgr && ah > 0 && aw > 0 && bh > 0 && bw > 0, "bad arguments");
    const int *start, *list;
    int rc = nct_build_inverse_nnf(ctx, bnn, bh * bw, aw, ah * aw, &start, &list);
    if (rc) return rc;
    const double wa = w_cohen / (double)(aw * ah), wb = w_complete / (double)(bw * bh);
    reconstruct_bds_kernel<<<nct_div_up(ah * aw, 128), 128, 0, ctx->stream>>>(b_bgr, ann, start, list, ah, aw, bh, bw, wa,
                                                                              wb, out_bgr);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_bds_feature_error(nct_ctx *ctx, const float *c_norm_hwc, const float *s_raw_hwc, const uint32_t *ann,
                          const uint32_t *bnn, int C, int ah, int aw, int bh, int bw, float w_cohen, float w_complete,
                          float *err_dev, float *vote_out_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, c_norm_hwc && s_raw_hwc && ann && bnn && err_dev, "null device pointer");
    NCT_REQUIRE(ctx, C == 16 || C == 32 || C == 64 || C == 128 || C == 256 || C == 384 || C == 512,
                "unsupported channel count %d", C);
    const int *start, *list;
    int rc = nct_build_inverse_nnf(ctx, bnn, bh * bw, aw, ah * aw, &start, &list);
    if (rc) return rc;
    const double wa = (double)w_cohen / (double)(aw * ah), wb = (double)w_complete / (double)(bw * bh);
    const int blocks = nct_div_up(ah * aw, 8);
#define LAUNCH(VPL)                                                                                               \
    bds_feature_error_kernel<VPL><<<blocks, 256, 0, ctx->stream>>>(c_norm_hwc, s_raw_hwc, ann, start, list, C, ah, \
                                                                   aw, bh, bw, wa, wb, err_dev, vote_out_dev)
    if (C <= 128) LAUNCH(1);
    else if (C == 256) LAUNCH(2);
    else if (C == 384) LAUNCH(3);
    else LAUNCH(4);
#undef LAUNCH
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // extern "C"
