// Exact fixed-point 3x3 convolution on the 5th-generation tensor cores: tcgen05.mma kind::i8 (INT8 x INT8 -> INT32 in
// TMEM), operands staged by TMA.  "Engine 3" of the VGG-19 trunk; replaces cudnnConvolutionForward + cudnnAddTensor +
// in-place ReLU (caffe/layers/cudnn_conv_layer.cu:20-37, cudnn_relu_layer.cu:19) for the layers with Cin >= 64.
//
// Why integers.  A floating-point tensor-core accumulation has no specified order, so its result can only be compared
// within a tolerance -- and this pipeline amplifies a 1-ulp feature difference into a 40 dB image difference (DESIGN.md
// section 6).  INT32 accumulation of INT8 products is EXACT, hence independent of any hardware order: the result below
// is a pure function of the inputs and oracle/vgg.py::q_conv3x3_relu reproduces it bit for bit.
//
//   E     : max(X) < 2^E   (X >= 0, the tensor's maximum is produced by the previous layer's epilogue, atomicMax)
//   xq    = floor(x * 2^(31-E))            = d0 256^3 + d1 256^2 + d2 256 + d3     d0 in [0,128], d1..d3 in [-128,127]
//   wq    = rint(w * 2^(22-Ew[cout]))      = e0 256^2 + e1 256 + e2                e0 in [-64,64], e1,e2 in [-128,127]
//   acc_d = sum_{taps, channels} sum_{i+j=d} d_i e_j      d = 0..3   (9 MMAs per K step into 4 TMEM accumulators;
//                                                                    the d >= 4 cross terms are dropped: zero-mean, 2^-32)
//   S     = acc_0 2^24 + acc_1 2^16 + acc_2 2^8 + acc_3   (|S| < 2^53: exact in FP64)
//   out   = max(fl32(S * 2^(E+Ew-37)) + bias, 0)
//
// Accuracy against an FP64 convolution: 3-6e-7 of the layer's range (the FP32 sequential order: 1-2.5e-6).
//
// Kernel shape (same skeleton as conv_tc.cu): M tile = 16 x 8 output pixels, the A operand of filter tap (dy, dx) is the
// same box shifted, loaded by one 3-D TMA per digit plane out of the NHWC digit planes (image border = TMA zero fill);
// B = digit planes of the weights [Cout][tap][Cin]; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
// K chunk = KB bytes (= channels) of one tap: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B, needed for Cin = 64).
#include "nct_internal.h"
#include <cuda.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace {

constexpr int TILE_W = 16, TILE_H = 8, BM = TILE_W * TILE_H;   // 128 output pixels per CTA
constexpr int NTHREADS = 192;
constexpr int NXD = 4, NWD = 3;                                // digit planes of activations / weights
constexpr int SMEM_LIMIT = 227 * 1024;


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

// K-major shared-memory matrix descriptor: rows of KB bytes (KB = 128: SWIZZLE_128B, 64: SWIZZLE_64B), 8-row groups
// 8 * KB bytes apart
template <int KB>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, 16-byte units, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * KB) >> 4) << 32;               // stride byte offset between 8-row groups, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell)
    d |= (uint64_t)(KB == 128 ? 2 : 4) << 61;           // layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}

// instruction descriptor, kind::i8: D = S32 (2 at [4,6)), A / B format at [7,10) / [10,13): 0 = unsigned, 1 = signed 8 bit;
// both K-major; N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N, int a_signed, int b_signed)
{
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// exponent E with v < 2^E, from the FP32 exponent field of a non-negative maximum (oracle/vgg.py::q_exponent)
__host__ __device__ __forceinline__ int q_exponent_bits(uint32_t bits) { return (int)((bits >> 23) & 0xFFu) - 126; }

struct ConvMapsI8 {
    CUtensorMap a[NXD], b[NWD];
};

template <int BN, int KB>
struct I8Cfg {
    static constexpr int A_BYTES = BM * KB, B_BYTES = BN * KB;
    static constexpr int STAGE_BYTES = NXD * A_BYTES + NWD * B_BYTES;
    static constexpr int MAX_STAGES = (SMEM_LIMIT - 2048) / STAGE_BYTES;
    static constexpr int STAGES = MAX_STAGES > 6 ? 6 : MAX_STAGES;
    static constexpr int TMEM_COLS = 4 * BN;   // 256 or 512: powers of two
    static_assert(STAGES >= 2, "tile does not fit in shared memory");
};

template <int BN, int KB, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_i8_kernel(const __grid_constant__ ConvMapsI8 maps, const float *__restrict__ bias, const int *__restrict__ wexp,
                  const uint32_t *__restrict__ in_max_bits, float *__restrict__ out, uint32_t *__restrict__ out_max_bits,
                  int *__restrict__ dbg_acc, int H, int W, int Cin, int Cout, int tiles_x)
{
    using Cfg = I8Cfg<BN, KB>;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int OFF_B = NXD * A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
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
    const int kchunks = Cin / KB;
    const int NKB = 9 * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: four INT32 accumulators of BN columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < NKB; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(empty_bar(s), (uint32_t)(((kb / STAGES) - 1) & 1));
                const int tap = kb / kchunks, kc = (kb % kchunks) * KB;
                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                const uint32_t st = base + s * STAGE_BYTES;
                mbar_expect_tx(full_bar(s), (uint32_t)STAGE_BYTES);
#pragma unroll
                for (int i = 0; i < NXD; ++i) tma_load_3d(st + i * A_BYTES, &maps.a[i], full_bar(s), kc, tile_x0 + dx, tile_y0 + dy);
#pragma unroll
                for (int j = 0; j < NWD; ++j) tma_load_2d(st + OFF_B + j * B_BYTES, &maps.b[j], full_bar(s), tap * Cin + kc, n0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): 9 digit products per K step into the accumulators d = i + j =====
        if (lane == 0) {
            constexpr uint32_t idesc_u = make_idesc_i8(BM, BN, 0, 1);   // leading activation digit: unsigned [0, 128]
            constexpr uint32_t idesc_s = make_idesc_i8(BM, BN, 1, 1);
            for (int kb = 0; kb < NKB; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full_bar(s), (uint32_t)((kb / STAGES) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = base + s * STAGE_BYTES;
                uint64_t ad[NXD], bd[NWD];
#pragma unroll
                for (int i = 0; i < NXD; ++i) ad[i] = make_desc<KB>(st + i * A_BYTES);
#pragma unroll
                for (int j = 0; j < NWD; ++j) bd[j] = make_desc<KB>(st + OFF_B + j * B_BYTES);
#pragma unroll
                for (int k = 0; k < KB / 32; ++k) {  // UMMA_K = 32 for 8-bit operands: 32 bytes inside the swizzle row
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);
                    const uint32_t acc = (kb | k) != 0;
                    umma_i8(tmem_base + 0 * BN, ad[0] + adv, bd[0] + adv, idesc_u, acc);
                    umma_i8(tmem_base + 1 * BN, ad[0] + adv, bd[1] + adv, idesc_u, acc);
                    umma_i8(tmem_base + 1 * BN, ad[1] + adv, bd[0] + adv, idesc_s, 1u);
                    umma_i8(tmem_base + 2 * BN, ad[0] + adv, bd[2] + adv, idesc_u, acc);
                    umma_i8(tmem_base + 2 * BN, ad[1] + adv, bd[1] + adv, idesc_s, 1u);
                    umma_i8(tmem_base + 2 * BN, ad[2] + adv, bd[0] + adv, idesc_s, 1u);
                    umma_i8(tmem_base + 3 * BN, ad[1] + adv, bd[2] + adv, idesc_s, acc);
                    umma_i8(tmem_base + 3 * BN, ad[2] + adv, bd[1] + adv, idesc_s, 1u);
                    umma_i8(tmem_base + 3 * BN, ad[3] + adv, bd[0] + adv, idesc_s, 1u);
                }
                umma_commit(empty_bar(s));                      // frees the smem stage once these MMAs have read it
                if (kb == NKB - 1) umma_commit(tmem_full_bar);  // accumulators complete
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp (id % 4) owns TMEM lanes [32*(id%4), +32) =====
        const int lg = warp & 3;
        const int E = q_exponent_bits(__ldg(in_max_bits));
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = lg * 32 + lane;                      // pixel index inside the tile = TMEM lane
        const int x = tile_x0 + (m % TILE_W), y = tile_y0 + (m / TILE_W);
        const bool valid = x < W && y < H;
        const size_t pix = (size_t)y * W + x;
        const size_t off = pix * Cout + n0;
        float vmax = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            int32_t a0[16], a1[16], a2[16], a3[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
            tmem_ld16(taddr + 0 * BN, a0);
            tmem_ld16(taddr + 1 * BN, a1);
            tmem_ld16(taddr + 2 * BN, a2);
            tmem_ld16(taddr + 3 * BN, a3);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
                if (dbg_acc) {
                    const size_t plane = (size_t)H * W * Cout;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        dbg_acc[0 * plane + off + c0 + j] = a0[j];
                        dbg_acc[1 * plane + off + c0 + j] = a1[j];
                        dbg_acc[2 * plane + off + c0 + j] = a2[j];
                        dbg_acc[3 * plane + off + c0 + j] = a3[j];
                    }
                }
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // S = a0 2^24 + a1 2^16 + a2 2^8 + a3, exact (|S| < 2^53); one rounding to FP32, then + bias, ReLU
                    const long long S = ((long long)a0[j] << 24) + ((long long)a1[j] << 16) + ((long long)a2[j] << 8) + (long long)a3[j];
                    const int e = E + __ldg(wexp + n0 + c0 + j) - 37;
                    const double scale = __longlong_as_double((long long)(e + 1023) << 52);
                    const float v = __double2float_rn(__ll2double_rn(S) * scale);
                    o[j] = fmaxf(__fadd_rn(v, __ldg(bias + n0 + c0 + j)), 0.f);
                    vmax = fmaxf(vmax, o[j]);
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4 *>(out + off + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
        }
        // tensor maximum for the next layer's quantisation (non-negative floats order like their bit patterns)
        uint32_t mb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(vmax));
        if (lane == 0 && out_max_bits && mb > *reinterpret_cast<volatile uint32_t *>(out_max_bits)) atomicMax(out_max_bits, mb);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ persistent variant
// One CTA per SM walks the (pixel tile, N tile) list with a stride of gridDim.x.  What the one-tile-per-CTA kernel pays per
// tile -- barrier set-up, TMEM allocation, the first TMA round trip with an empty pipeline, the epilogue with the tensor
// pipe idle, CTA teardown and relaunch -- is paid once, and across tiles the three roles overlap: the TMA warp runs ahead
// into the next tile's stages while the epilogue drains the current one, and with NBUF = 2 accumulator sets in TMEM
// (4 x BN x 2 <= 512 columns, i.e. BN = 64) the MMA warp starts the next tile while the epilogue warps still read the
// previous set.  Shallow layers (K = 576: conv1_2, conv2_1) are epilogue-bound in the one-tile kernel (25 - 37 % tensor
// pipe, profiles/r2_conv_i8_ncu.md); results are bit-identical (same integer sums, same epilogue).
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int BN, int KB, int STAGES, int NBUF>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_i8_persistent_kernel(const __grid_constant__ ConvMapsI8 maps, const float *__restrict__ bias, const int *__restrict__ wexp,
                             const uint32_t *__restrict__ in_max_bits, float *__restrict__ out, uint32_t *__restrict__ out_max_bits,
                             int *__restrict__ dbg_acc, int H, int W, int Cin, int Cout, int tiles_x, int total_tiles)
{
    using Cfg = I8Cfg<BN, KB>;
    constexpr int A_BYTES = Cfg::A_BYTES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int OFF_B = NXD * A_BYTES;
    constexpr int TMEM_COLS = NBUF * 4 * BN;
    static_assert(TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two <= 512 columns");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;   // full[STAGES], empty[STAGES], tfull[NBUF], tempty[NBUF], tmem_ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + NBUF + b); };
    const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 2 * NBUF);
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));
    volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2 * NBUF));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pix_tiles = total_tiles / (Cout / BN);
    const int kchunks = Cin / KB;
    const int NKB = 9 * kchunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_gen;

    if (warp == 0) {
        // ===== TMA producer: the stage ring runs on across tile boundaries =====
        if (lane == 0) {
            int g = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int pt = t % pix_tiles, n0 = (t / pix_tiles) * BN;   // pixel tile fastest: concurrent CTAs share the weight tile
                const int tile_y0 = (pt / tiles_x) * TILE_H, tile_x0 = (pt % tiles_x) * TILE_W;
                for (int kb = 0; kb < NKB; ++kb, ++g) {
                    const int s = g % STAGES;
                    if (g >= STAGES) mbar_wait(empty_bar(s), (uint32_t)(((g / STAGES) - 1) & 1));
                    const int tap = kb / kchunks, kc = (kb % kchunks) * KB;
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    const uint32_t st = base + s * STAGE_BYTES;
                    mbar_expect_tx(full_bar(s), (uint32_t)STAGE_BYTES);
#pragma unroll
                    for (int i = 0; i < NXD; ++i) tma_load_3d(st + i * A_BYTES, &maps.a[i], full_bar(s), kc, tile_x0 + dx, tile_y0 + dy);
#pragma unroll
                    for (int j = 0; j < NWD; ++j) tma_load_2d(st + OFF_B + j * B_BYTES, &maps.b[j], full_bar(s), tap * Cin + kc, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc_u = make_idesc_i8(BM, BN, 0, 1);
            constexpr uint32_t idesc_s = make_idesc_i8(BM, BN, 1, 1);
            int g = 0, it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int buf = it % NBUF;
                if (it >= NBUF) {   // the epilogue has drained this accumulator set
                    mbar_wait(tempty_bar(buf), (uint32_t)(((it / NBUF) - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t tacc = tmem_base + (uint32_t)(buf * 4 * BN);
                for (int kb = 0; kb < NKB; ++kb, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(full_bar(s), (uint32_t)((g / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = base + s * STAGE_BYTES;
                    uint64_t ad[NXD], bd[NWD];
#pragma unroll
                    for (int i = 0; i < NXD; ++i) ad[i] = make_desc<KB>(st + i * A_BYTES);
#pragma unroll
                    for (int j = 0; j < NWD; ++j) bd[j] = make_desc<KB>(st + OFF_B + j * B_BYTES);
#pragma unroll
                    for (int k = 0; k < KB / 32; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        const uint32_t acc = (kb | k) != 0;
                        umma_i8(tacc + 0 * BN, ad[0] + adv, bd[0] + adv, idesc_u, acc);
                        umma_i8(tacc + 1 * BN, ad[0] + adv, bd[1] + adv, idesc_u, acc);
                        umma_i8(tacc + 1 * BN, ad[1] + adv, bd[0] + adv, idesc_s, 1u);
                        umma_i8(tacc + 2 * BN, ad[0] + adv, bd[2] + adv, idesc_u, acc);
                        umma_i8(tacc + 2 * BN, ad[1] + adv, bd[1] + adv, idesc_s, 1u);
                        umma_i8(tacc + 2 * BN, ad[2] + adv, bd[0] + adv, idesc_s, 1u);
                        umma_i8(tacc + 3 * BN, ad[1] + adv, bd[2] + adv, idesc_s, acc);
                        umma_i8(tacc + 3 * BN, ad[2] + adv, bd[1] + adv, idesc_s, 1u);
                        umma_i8(tacc + 3 * BN, ad[3] + adv, bd[0] + adv, idesc_s, 1u);
                    }
                    umma_commit(empty_bar(s));
                    if (kb == NKB - 1) umma_commit(tfull_bar(buf));
                }
            }
        }
    } else {
        // ===== epilogue: 4 warps, warp (id % 4) owns TMEM lanes [32*(id%4), +32) =====
        const int lg = warp & 3;
        const int E = q_exponent_bits(__ldg(in_max_bits));
        const int m = lg * 32 + lane;
        float vmax = 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int buf = it % NBUF;
            const int pt = t % pix_tiles, n0 = (t / pix_tiles) * BN;
            const int tile_y0 = (pt / tiles_x) * TILE_H, tile_x0 = (pt % tiles_x) * TILE_W;
            const int x = tile_x0 + (m % TILE_W), y = tile_y0 + (m / TILE_W);
            const bool valid = x < W && y < H;
            const size_t off = ((size_t)y * W + x) * Cout + n0;
            mbar_wait(tfull_bar(buf), (uint32_t)((it / NBUF) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                int32_t a0[16], a1[16], a2[16], a3[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * 4 * BN + c0);
                tmem_ld16(taddr + 0 * BN, a0);
                tmem_ld16(taddr + 1 * BN, a1);
                tmem_ld16(taddr + 2 * BN, a2);
                tmem_ld16(taddr + 3 * BN, a3);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (valid) {
                    if (dbg_acc) {
                        const size_t plane = (size_t)H * W * Cout;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            dbg_acc[0 * plane + off + c0 + j] = a0[j];
                            dbg_acc[1 * plane + off + c0 + j] = a1[j];
                            dbg_acc[2 * plane + off + c0 + j] = a2[j];
                            dbg_acc[3 * plane + off + c0 + j] = a3[j];
                        }
                    }
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const long long S = ((long long)a0[j] << 24) + ((long long)a1[j] << 16) + ((long long)a2[j] << 8) + (long long)a3[j];
                        const int e = E + __ldg(wexp + n0 + c0 + j) - 37;
                        const double scale = __longlong_as_double((long long)(e + 1023) << 52);
                        const float v = __double2float_rn(__ll2double_rn(S) * scale);
                        o[j] = fmaxf(__fadd_rn(v, __ldg(bias + n0 + c0 + j)), 0.f);
                        vmax = fmaxf(vmax, o[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4 *>(out + off + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                }
            }
            // this warp has read its lanes of the accumulator set: hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
        uint32_t mb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(vmax));
        if (lane == 0 && out_max_bits && mb > *reinterpret_cast<volatile uint32_t *>(out_max_bits)) atomicMax(out_max_bits, mb);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ activation digit planes
__device__ __forceinline__ void q_digits(float v, double s, int &d0, int &d1, int &d2, int &d3)
{
    const uint32_t q = __double2uint_rd((double)fmaxf(v, 0.f) * s);   // floor(x 2^(31-E)) < 2^31
    uint32_t r = q;   // (r - d) is in [0, 2^31]: unsigned arithmetic, q + 128 may reach 2^31
    d3 = (int)(int8_t)(r & 0xFFu); r = (r - (uint32_t)d3) >> 8;
    d2 = (int)(int8_t)(r & 0xFFu); r = (r - (uint32_t)d2) >> 8;
    d1 = (int)(int8_t)(r & 0xFFu); r = (r - (uint32_t)d1) >> 8;
    d0 = (int)r;
}

// x: FP32 [n] (NHWC flattened) -> planes[i][n] (i = 0..3, plane stride `pstride` bytes); 16 elements per thread
__global__ void __launch_bounds__(256) act_digits_kernel(const float *__restrict__ x, const uint32_t *__restrict__ max_bits,
                                                         uint8_t *__restrict__ planes, size_t pstride, size_t n)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    const int E = q_exponent_bits(__ldg(max_bits));
    const double s = __longlong_as_double((long long)(31 - E + 1023) << 52);
    uint32_t w[NXD][4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + i0) + q4);
        const float vv[4] = {v.x, v.y, v.z, v.w};
        uint32_t p[NXD] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int d[4];
            q_digits(vv[e], s, d[0], d[1], d[2], d[3]);
#pragma unroll
            for (int i = 0; i < NXD; ++i) p[i] |= (uint32_t)(d[i] & 0xFF) << (8 * e);
        }
#pragma unroll
        for (int i = 0; i < NXD; ++i) w[i][q4] = p[i];
    }
#pragma unroll
    for (int i = 0; i < NXD; ++i)
        *reinterpret_cast<uint4 *>(planes + (size_t)i * pstride + i0) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
}

// 2x2 / stride 2 ceil-mode MAX pooling (caffe/layers/pooling_layer.cpp:90-93) fused with the digit split: the pooled
// FP32 tensor itself is never needed (no feature level is a pooling output).  4 channels per thread.
__global__ void __launch_bounds__(256) pool_digits_kernel(const float *__restrict__ in, const uint32_t *__restrict__ max_bits,
                                                          uint8_t *__restrict__ planes, size_t pstride, int H, int W, int C, int Ho, int Wo)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4n = C / 4;
    if (t >= (long long)Ho * Wo * c4n) return;
    const int c4 = (int)(t % c4n);
    const int po = (int)(t / c4n);
    const int xo = po % Wo, yo = po / Wo;
    float4 m = make_float4(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int y = 2 * yo + dy, x = 2 * xo + dx;
            if (y < H && x < W) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(in + ((size_t)y * W + x) * C) + c4);
                m.x = v.x > m.x ? v.x : m.x;
                m.y = v.y > m.y ? v.y : m.y;
                m.z = v.z > m.z ? v.z : m.z;
                m.w = v.w > m.w ? v.w : m.w;
            }
        }
    const int E = q_exponent_bits(__ldg(max_bits));
    const double s = __longlong_as_double((long long)(31 - E + 1023) << 52);
    const float vv[4] = {m.x, m.y, m.z, m.w};
    uint32_t p[NXD] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int d[4];
        q_digits(vv[e], s, d[0], d[1], d[2], d[3]);
#pragma unroll
        for (int i = 0; i < NXD; ++i) p[i] |= (uint32_t)(d[i] & 0xFF) << (8 * e);
    }
#pragma unroll
    for (int i = 0; i < NXD; ++i) *reinterpret_cast<uint32_t *>(planes + (size_t)i * pstride + (size_t)po * C + c4 * 4) = p[i];
}

// maximum of a non-negative FP32 tensor (used when the producer is not one of this file's kernels)
__global__ void __launch_bounds__(256) tensor_max_kernel(const float *__restrict__ x, size_t n, uint32_t *__restrict__ max_bits)
{
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, x[i]);
    const uint32_t mb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(fmaxf(m, 0.f)));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, mb);
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

int encode_act(nct_ctx *ctx, EncodeTiledFn encode, CUtensorMap *m, const uint8_t *ptr, int H, int W, int C, int KB)
{
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {(cuuint64_t)C, (cuuint64_t)W * C};
    cuuint32_t box[3] = {(cuuint32_t)KB, TILE_W, TILE_H};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        KB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled(activation digits) failed: %d", (int)r);
    return NCT_OK;
}

int encode_wgt(nct_ctx *ctx, EncodeTiledFn encode, CUtensorMap *m, const int8_t *ptr, int Cin, int Cout, int BN, int KB)
{
    cuuint64_t dims[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)9 * Cin};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        KB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled(weight digits) failed: %d", (int)r);
    return NCT_OK;
}

int env_int(const char *name, int dflt);

template <int BN, int KB>
int launch(nct_ctx *ctx, const ConvMapsI8 &maps, const float *bias, const int *wexp, const uint32_t *in_max, float *out, uint32_t *out_max,
           int *dbg, int H, int W, int Cin, int Cout)
{
    using Cfg = I8Cfg<BN, KB>;
    constexpr int STAGES = Cfg::STAGES;
    const int tiles_x = nct_div_up(W, TILE_W), tiles_y = nct_div_up(H, TILE_H);
    // persistent tile loop for the shallow layers (K = 9 Cin <= 1152: conv1_2 ... conv3_1), where the one-tile kernel is
    // epilogue-bound; the deep layers are tensor-bound either way and measured 5 - 10 % slower with the single accumulator set
    // they can afford (profiles/r2_conv_i8_ncu.md).  NCT_I8_PERSIST=0 / 1 forces one kernel for every layer (A/B, tests).
    const int persist = env_int("NCT_I8_PERSIST", -1);
    if (persist == 1 || (persist < 0 && Cin <= 128)) {
        constexpr int NBUF = (BN == 64) ? 2 : 1;   // two accumulator sets fit TMEM's 512 columns only at BN = 64
        const int total = tiles_x * tiles_y * (Cout / BN);
        const int grid = total < ctx->num_sms ? total : ctx->num_sms;
        const size_t smem_p = (size_t)STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 2 * NBUF + 1) + 1024;
        NCT_CUDA(ctx, cudaFuncSetAttribute(conv3x3_i8_persistent_kernel<BN, KB, STAGES, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
        conv3x3_i8_persistent_kernel<BN, KB, STAGES, NBUF><<<grid, NTHREADS, smem_p, ctx->stream>>>(maps, bias, wexp, in_max, out, out_max, dbg, H, W,
                                                                                                 Cin, Cout, tiles_x, total);
        NCT_CHECK_LAUNCH(ctx);
        return NCT_OK;
    }
    dim3 grid(tiles_x * tiles_y, Cout / BN);
    const size_t smem = (size_t)STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 2) + 1024;
    NCT_CUDA(ctx, cudaFuncSetAttribute(conv3x3_i8_kernel<BN, KB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_i8_kernel<BN, KB, STAGES><<<grid, NTHREADS, smem, ctx->stream>>>(maps, bias, wexp, in_max, out, out_max, dbg, H, W, Cin, Cout, tiles_x);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// ---- host: digit planes of the weights.  w_oihw -> planes[j][cout][tap*cin + c] (int8), wexp[cout]
void nct_q_weight_digits_host(const float *w_oihw, int cin, int cout, int8_t *planes, int *wexp)
{
    const size_t K = (size_t)9 * cin, plane = K * cout;
    for (int o = 0; o < cout; ++o) {
        float mx = 0.f;
        for (size_t i = 0; i < K; ++i) mx = std::fmax(mx, std::fabs(w_oihw[(size_t)o * K + i]));
        uint32_t bits;
        memcpy(&bits, &mx, 4);
        const int Ew = q_exponent_bits(bits);
        wexp[o] = Ew;
        const double s = std::ldexp(1.0, 22 - Ew);
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < 9; ++t) {
                long long r = (long long)std::nearbyint((double)w_oihw[((size_t)o * cin + c) * 9 + t] * s);   // ties to even
                const int d2 = (int)(int8_t)(r & 0xFF); r = (r - d2) >> 8;
                const int d1 = (int)(int8_t)(r & 0xFF); r = (r - d1) >> 8;
                const int d0 = (int)r;
                const size_t k = (size_t)o * K + (size_t)t * cin + c;
                planes[0 * plane + k] = (int8_t)d0;
                planes[1 * plane + k] = (int8_t)d1;
                planes[2 * plane + k] = (int8_t)d2;
            }
    }
}

int nct_q_tensor_max(nct_ctx *ctx, const float *x, size_t n, uint32_t *max_bits)
{
    NCT_CUDA(ctx, cudaMemsetAsync(max_bits, 0, 4, ctx->stream));
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
    tensor_max_kernel<<<blocks, 256, 0, ctx->stream>>>(x, n, max_bits);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

// x FP32 NHWC [H][W][C] -> 4 digit planes (plane stride pstride bytes >= H*W*C), scaled by the tensor maximum in *max_bits
int nct_q_act_digits(nct_ctx *ctx, const float *x, const uint32_t *max_bits, uint8_t *planes, size_t pstride, size_t n)
{
    NCT_REQUIRE(ctx, n % 16 == 0 && pstride % 16 == 0, "digit planes need a multiple of 16 elements");
    act_digits_kernel<<<(unsigned)((n / 16 + 255) / 256), 256, 0, ctx->stream>>>(x, max_bits, planes, pstride, n);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_q_pool_digits(nct_ctx *ctx, const float *x, const uint32_t *max_bits, uint8_t *planes, size_t pstride, int H, int W, int C, int Ho, int Wo)
{
    const long long threads = (long long)Ho * Wo * (C / 4);
    pool_digits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(x, max_bits, planes, pstride, H, W, C, Ho, Wo);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

// planes: 4 activation digit planes [H][W][Cin] (stride pstride); wplanes: 3 weight digit planes [Cout][9*Cin]; out = relu(Q conv + bias)
int nct_conv3x3_i8(nct_ctx *ctx, const uint8_t *planes, size_t pstride, const uint32_t *in_max_bits, const int8_t *wplanes, const int *wexp,
                   const float *bias, float *out, uint32_t *out_max_bits, int *dbg_acc, int H, int W, int Cin, int Cout)
{
    NCT_REQUIRE(ctx, Cin % 64 == 0 && Cin >= 64, "fixed-point tensor-core conv needs Cin %% 64 == 0 (got %d)", Cin);
    NCT_REQUIRE(ctx, Cout % 64 == 0, "fixed-point tensor-core conv needs Cout %% 64 == 0 (got %d)", Cout);
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return nct_fail(ctx, NCT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    int KB = (Cin % 128 == 0) ? env_int("NCT_I8_KB", 128) : 64;
    if (KB != 64 && KB != 128) KB = 128;
    int BN = (Cout % 128 == 0) ? env_int("NCT_I8_BN", 128) : 64;
    if (BN != 64 && BN != 128) BN = 128;
    ConvMapsI8 maps;
    memset(&maps, 0, sizeof(maps));
    const size_t wplane = (size_t)9 * Cin * Cout;
    for (int i = 0; i < NXD; ++i) {
        int rc = encode_act(ctx, encode, &maps.a[i], planes + (size_t)i * pstride, H, W, Cin, KB);
        if (rc) return rc;
    }
    for (int j = 0; j < NWD; ++j) {
        int rc = encode_wgt(ctx, encode, &maps.b[j], wplanes + (size_t)j * wplane, Cin, Cout, BN, KB);
        if (rc) return rc;
    }
    if (BN == 128 && KB == 128) return launch<128, 128>(ctx, maps, bias, wexp, in_max_bits, out, out_max_bits, dbg_acc, H, W, Cin, Cout);
    if (BN == 128 && KB == 64) return launch<128, 64>(ctx, maps, bias, wexp, in_max_bits, out, out_max_bits, dbg_acc, H, W, Cin, Cout);
    if (BN == 64 && KB == 128) return launch<64, 128>(ctx, maps, bias, wexp, in_max_bits, out, out_max_bits, dbg_acc, H, W, Cin, Cout);
    return launch<64, 64>(ctx, maps, bias, wexp, in_max_bits, out, out_max_bits, dbg_acc, H, W, Cin, Cout);
}

extern "C" {

// One layer through the exact fixed-point engine, from plain FP32 tensors (diagnostic / unit-test entry point; the trunk
// keeps its digit planes resident instead): in_dev NHWC FP32 >= 0, w_oihw_host / bias_host as Caffe stores them.
// acc_dbg_dev (optional): int32 [4][H*W][Cout], the raw TMEM accumulators.
int nct_conv3x3_fixedpoint(nct_ctx *ctx, const float *in_dev, const float *w_oihw_host, const float *bias_host, float *out_dev,
                           int *acc_dbg_dev, int H, int W, int Cin, int Cout)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, in_dev && w_oihw_host && bias_host && out_dev && H > 0 && W > 0, "bad arguments");
    NCT_REQUIRE(ctx, Cin % 64 == 0 && Cout % 64 == 0, "Cin and Cout must be multiples of 64");
    const size_t n = (size_t)H * W * Cin, wn = (size_t)9 * Cin * Cout;
    std::vector<int8_t> wd(3 * wn);
    std::vector<int> we(Cout);
    nct_q_weight_digits_host(w_oihw_host, Cin, Cout, wd.data(), we.data());
    uint8_t *planes = (uint8_t *)nct_scratch(ctx, "q1_planes", 4 * n);
    int8_t *wpl = (int8_t *)nct_scratch(ctx, "q1_wplanes", 3 * wn);
    int *wexp = (int *)nct_scratch(ctx, "q1_wexp", sizeof(int) * Cout);
    float *bias = (float *)nct_scratch(ctx, "q1_bias", sizeof(float) * Cout);
    uint32_t *slots = (uint32_t *)nct_scratch(ctx, "q1_max", 8);
    if (!planes || !wpl || !wexp || !bias || !slots) return NCT_ERR_NOMEM;
    NCT_CUDA(ctx, cudaMemcpyAsync(wpl, wd.data(), 3 * wn, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(wexp, we.data(), sizeof(int) * Cout, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(bias, bias_host, sizeof(float) * Cout, cudaMemcpyHostToDevice, ctx->stream));
    int rc = nct_q_tensor_max(ctx, in_dev, n, slots);
    if (rc) return rc;
    NCT_CUDA(ctx, cudaMemsetAsync(slots + 1, 0, 4, ctx->stream));
    rc = nct_q_act_digits(ctx, in_dev, slots, planes, n, n);
    if (rc) return rc;
    rc = nct_conv3x3_i8(ctx, planes, n, slots, wpl, wexp, bias, out_dev, slots + 1, acc_dbg_dev, H, W, Cin, Cout);
    if (rc) return rc;
    NCT_CUDA(ctx, nct_stream_wait(ctx));   // the host vectors above must outlive the copies
    return NCT_OK;
}

}  // extern "C"
