// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05, sm_100a), fed by TMA.
//
// Replaces cudnnConvolutionForward + cudnnAddTensor + in-place ReLU (caffe/layers/cudnn_conv_layer.cu:20-37,
// cudnn_relu_layer.cu:19) for the 3x3 / pad 1 / stride 1 layers of the VGG-19 trunk with Cin >= 64.
//
//   D[128 pixels][BN couts] (FP32, in TMEM) += A[128 pixels][32 k] x B[BN couts][32 k]^T     (kind::tf32)
//
// * M tile = a 16 x 8 spatial box of output pixels.  For filter tap (dy, dx) and a 32-channel chunk the A operand is
//   the SAME box shifted by (dx, dy): one 3-D TMA load (cp.async.bulk.tensor.3d, box {32 ch, 16 x, 8 y}) out of the
//   NHWC activation; the halo / image border is TMA's out-of-bounds zero fill (coordinates may be negative), so no
//   im2col buffer and no padding pass exist.  128-byte rows, SWIZZLE_128B = the canonical K-major UMMA layout.
// * B operand = weights re-laid out once to [Cout][tap][Cin] (K-major), 2-D TMA box {32 k, BN}.
// * Warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//   warps 2-5 = epilogue (tcgen05.ld -> +bias -> ReLU -> NHWC store).  Multi-stage smem ring with full/empty
//   mbarriers; tcgen05.commit releases a stage when the MMAs reading it have retired.
// * Precision.  kind::tf32 TRUNCATES the FP32 operands to a 10-bit mantissa (measured: tools/tf32_probe.py) and
//   accumulates in FP32.  X3 = false: plain TF32 (~9e-3 of the feature range after 13 layers).  X3 = true ("3xTF32"):
//   every operand x is split exactly into hi = rne_tf32(x) and lo = x - hi (weights once on the host, activations by the
//   producing layer's epilogue); a_lo*b_hi + a_hi*b_lo + a_hi*b_hi are three MMAs into the same accumulator.  hi is
//   exactly representable, so the hardware truncation only touches the last bit of the lo parts: FP32-level accuracy
//   at tensor-core speed.  (A first version that let the hardware truncate the full-precision operand and stored
//   only the truncation remainder was biased: 1.9e-4 of the feature range after 13 layers.)
#include "nct_internal.h"
#include <cuda.h>
#include <cstring>

namespace {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;   // 128 output pixels per CTA
constexpr int BK = 32;                                         // 32 fp32 = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 4;                           // 16 KB
constexpr int NTHREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B, bits [61,64)
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// round-to-nearest-even to TF32's 11 significant bits (finite inputs)
__host__ __device__ __forceinline__ float tf32_rne(float v)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(v);
    b += 0x00000FFFu + ((b >> 13) & 1u);
    return __uint_as_float(b & 0xFFFFE000u);
#else
    uint32_t b;
    memcpy(&b, &v, 4);
    b += 0x00000FFFu + ((b >> 13) & 1u);
    b &= 0xFFFFE000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
#endif
}

struct ConvMaps {
    CUtensorMap a, a_lo, b, b_lo;
};

template <int BN, bool X3>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvMaps maps, const float *__restrict__ bias, float *__restrict__ out,
                  float *__restrict__ out_hi, float *__restrict__ out_lo, int H, int W, int Cin, int Cout, int tiles_x)
{
    constexpr int STAGES = X3 ? 3 : 4;
    constexpr int B_BYTES = BN * BK * 4;
    constexpr int STAGE_BYTES = (X3 ? 2 : 1) * (A_BYTES + B_BYTES);
    // stage layout: A | B | (A_lo | B_lo)
    constexpr int OFF_B = A_BYTES, OFF_ALO = A_BYTES + B_BYTES, OFF_BLO = 2 * A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;            // full[STAGES], empty[STAGES], tmem_full, tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 1);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_y0 = (blockIdx.x / tiles_x) * TILE_H, tile_x0 = (blockIdx.x % tiles_x) * TILE_W;
    const int n0 = blockIdx.y * BN;
    const int kchunks = Cin / BK;
    const int KB = 9 * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation: BN fp32 accumulator columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(empty_bar(s), (uint32_t)(((kb / STAGES) - 1) & 1));
                const int tap = kb / kchunks, kc = (kb % kchunks) * BK;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                const uint32_t st = base + s * STAGE_BYTES;
                mbar_expect_tx(full_bar(s), (uint32_t)STAGE_BYTES);
                tma_load_3d(st, &maps.a, full_bar(s), kc, tile_x0 + dx, tile_y0 + dy);
                tma_load_2d(st + OFF_B, &maps.b, full_bar(s), kb * BK, n0);
                if (X3) {
                    tma_load_3d(st + OFF_ALO, &maps.a_lo, full_bar(s), kc, tile_x0 + dx, tile_y0 + dy);
                    tma_load_2d(st + OFF_BLO, &maps.b_lo, full_bar(s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full_bar(s), (uint32_t)((kb / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE_BYTES;
                const uint64_t adesc = make_desc_sw128(st), bdesc = make_desc_sw128(st + OFF_B);
                const uint64_t alo = make_desc_sw128(st + OFF_ALO), blo = make_desc_sw128(st + OFF_BLO);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {  // UMMA_K = 8 for tf32: advance 32 bytes inside the 128-byte swizzle row
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    if (X3) {  // small terms first, then the main product
                        umma_tf32(tmem_base, alo + adv, bdesc + adv, idesc, (kb | k) != 0);
                        umma_tf32(tmem_base, adesc + adv, blo + adv, idesc, 1u);
                        umma_tf32(tmem_base, adesc + adv, bdesc + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, adesc + adv, bdesc + adv, idesc, (kb | k) != 0);
                    }
                }
                umma_commit(empty_bar(s));                     // frees the smem stage once these MMAs have read it
                if (kb == KB - 1) umma_commit(tmem_full_bar);  // accumulator complete
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp (id % 4) owns TMEM lanes [32*(id%4), +32) =====
        const int lg = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = lg * 32 + lane;                      // pixel index inside the tile = TMEM lane
        const int x = tile_x0 + (m % TILE_W), y = tile_y0 + (m / TILE_W);
        const bool valid = x < W && y < H;
        const size_t off = ((size_t)y * W + x) * Cout + n0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + n0 + c0 + j));
                    float4 o;
                    o.x = fmaxf(__uint_as_float(r[j]) + bb.x, 0.f);
                    o.y = fmaxf(__uint_as_float(r[j + 1]) + bb.y, 0.f);
                    o.z = fmaxf(__uint_as_float(r[j + 2]) + bb.z, 0.f);
                    o.w = fmaxf(__uint_as_float(r[j + 3]) + bb.w, 0.f);
                    if (out) *reinterpret_cast<float4 *>(out + off + c0 + j) = o;
                    if (X3) {  // exact split o = hi + lo, hi = o rounded to TF32 (so the tensor core reads it unchanged)
                        float4 hh, l;
                        hh.x = tf32_rne(o.x); l.x = o.x - hh.x;
                        hh.y = tf32_rne(o.y); l.y = o.y - hh.y;
                        hh.z = tf32_rne(o.z); l.z = o.z - hh.z;
                        hh.w = tf32_rne(o.w); l.w = o.w - hh.w;
                        *reinterpret_cast<float4 *>(out_hi + off + c0 + j) = hh;
                        *reinterpret_cast<float4 *>(out_lo + off + c0 + j) = l;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// exact split x = hi + lo with hi = rne_tf32(x) (activations that did not come out of the tensor-core epilogue)
__global__ void tf32_split_kernel(const float *__restrict__ x, float *__restrict__ hi, float *__restrict__ lo, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = x[i], h = tf32_rne(v);
        hi[i] = h;
        lo[i] = v - h;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_act(nct_ctx *ctx, EncodeTiledFn encode, CUtensorMap *m, const float *ptr, int H, int W, int C)
{
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4};
    cuuint32_t box[3] = {BK, TILE_W, TILE_H};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: %d", (int)r);
    return NCT_OK;
}

int encode_wgt(nct_ctx *ctx, EncodeTiledFn encode, CUtensorMap *m, const float *ptr, int Cin, int Cout, int BN)
{
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin * 4};
    cuuint32_t box[2] = {BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return NCT_OK;
}

template <int BN, bool X3>
int launch(nct_ctx *ctx, const ConvMaps &maps, const float *bias, float *out, float *out_hi, float *out_lo, int H, int W, int Cin, int Cout)
{
    constexpr int STAGES = X3 ? 3 : 4;
    const int tiles_x = nct_div_up(W, TILE_W), tiles_y = nct_div_up(H, TILE_H);
    dim3 grid(tiles_x * tiles_y, Cout / BN);
    const size_t smem = (size_t)STAGES * (X3 ? 2 : 1) * (A_BYTES + BN * BK * 4) + 8 * (2 * STAGES + 2) + 1024;
    NCT_CUDA(ctx, cudaFuncSetAttribute(conv3x3_tc_kernel<BN, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_tc_kernel<BN, X3><<<grid, NTHREADS, smem, ctx->stream>>>(maps, bias, out, out_hi, out_lo, H, W, Cin, Cout, tiles_x);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // namespace

int nct_tf32_split(nct_ctx *ctx, const float *x, float *hi, float *lo, size_t n)
{
    tf32_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(x, hi, lo, n);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

void nct_tf32_split_host(const float *x, float *hi, float *lo, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        hi[i] = tf32_rne(x[i]);
        lo[i] = x[i] - hi[i];
    }
}

// in: NHWC FP32 [H][W][Cin]; w_kmajor: [Cout][9*Cin] (k = tap*Cin + c); out: NHWC [H][W][Cout] = relu(conv + bias).
// Plain TF32: in / w are the FP32 tensors (truncated by the hardware), in_lo = w_lo = out_hi = out_lo = NULL.
// 3xTF32: in / w are the TF32-rounded hi parts, in_lo / w_lo the exact remainders; out_hi / out_lo receive the split of the
// result, out (the full FP32 result) may be NULL when nobody needs it.
int nct_conv3x3_tensorcore(nct_ctx *ctx, const float *in, const float *in_lo, const float *w_kmajor, const float *w_lo, const float *bias,
                           float *out, float *out_hi, float *out_lo, int H, int W, int Cin, int Cout)
{
    NCT_REQUIRE(ctx, Cin % BK == 0 && Cin >= 64, "tensor-core conv needs Cin %% 32 == 0 and >= 64 (got %d)", Cin);
    NCT_REQUIRE(ctx, Cout % 64 == 0, "tensor-core conv needs Cout %% 64 == 0 (got %d)", Cout);
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const bool x3 = in_lo && w_lo && out_lo && out_hi;
    NCT_REQUIRE(ctx, x3 || out, "plain TF32 conv needs an output buffer");
    const int BN = (Cout % 128 == 0) ? 128 : 64;
    ConvMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = encode_act(ctx, encode, &maps.a, in, H, W, Cin);
    if (!rc) rc = encode_wgt(ctx, encode, &maps.b, w_kmajor, Cin, Cout, BN);
    if (!rc && x3) rc = encode_act(ctx, encode, &maps.a_lo, in_lo, H, W, Cin);
    if (!rc && x3) rc = encode_wgt(ctx, encode, &maps.b_lo, w_lo, Cin, Cout, BN);
    if (rc) return rc;
    if (BN == 128) return x3 ? launch<128, true>(ctx, maps, bias, out, out_hi, out_lo, H, W, Cin, Cout) : launch<128, false>(ctx, maps, bias, out, out_hi, out_lo, H, W, Cin, Cout);
    return x3 ? launch<64, true>(ctx, maps, bias, out, out_hi, out_lo, H, W, Cin, Cout) : launch<64, false>(ctx, maps, bias, out, out_hi, out_lo, H, W, Cin, Cout);
}
