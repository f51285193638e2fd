// VGG-19 trunk (conv1_1 .. conv5_1) feature extractor.
//
// Replaces Classifier::Predict / Preprocess (NCT/Classifier.cpp:59-143, 185-275) and the Caffe Net::Forward it
// drives (caffe/net.cpp:554-586; cudnn_conv_layer.cu:20-37 conv + bias, in-place ReLU, 2x2/2 ceil-mode MAX pooling
// caffe/layers/pooling_layer.cpp:90-93, pooling_layer.cu:10-40) for the one graph the reference ever runs
// (demo/model/vgg19/VGG_ILSVRC_19_layers_deploy.prototxt), truncated after conv5_1 -- the reference also runs
// conv5_2..pool5 although nothing reads them -- and, on re-forwards, after the deepest layer still needed.
//
// Layout: activations are pixel-major FP32 (NHWC, N = 1), so every feature map is directly the HWC volume
// PatchMatch consumes (the reference keeps planar blobs and pays strided gathers).  Weights are re-laid out
// once to [tap][Cin][Cout].
//
// This file holds the FP32 CUDA-core implicit-GEMM path (exact FP32 products, fixed summation order:
// tap-major, channel-minor).  It is the numerically conservative engine: FP32 like the reference's cuDNN path.
#include "nct_internal.h"
#include <vector>
#include <cstring>

namespace {

struct ConvSpec { const char *name; int cin, cout; bool pool_before; int level; };
// level: index into the 5 feature maps (0 = conv5_1 ... 4 = conv1_1), -1 = not exported
const ConvSpec kTrunk[13] = {
    {"conv1_1", 3, 64, false, 4},   {"conv1_2", 64, 64, false, -1},
    {"conv2_1", 64, 128, true, 3},  {"conv2_2", 128, 128, false, -1},
    {"conv3_1", 128, 256, true, 2}, {"conv3_2", 256, 256, false, -1}, {"conv3_3", 256, 256, false, -1}, {"conv3_4", 256, 256, false, -1},
    {"conv4_1", 256, 512, true, 1}, {"conv4_2", 512, 512, false, -1}, {"conv4_3", 512, 512, false, -1}, {"conv4_4", 512, 512, false, -1},
    {"conv5_1", 512, 512, true, 0},
};

}  // namespace

struct VggState {
    float *w[13] = {nullptr};   // [9][cin][cout]      (CUDA-core engine)
    float *wk[13] = {nullptr};  // [cout][9][cin]      (K-major, tensor-core engines)
    float *wk_hi[13] = {nullptr};  // rne_tf32(wk)         (3xTF32 engine)
    float *wk_lo[13] = {nullptr};  // wk - wk_hi
    float *b[13] = {nullptr};
    int8_t *wq[13] = {nullptr};    // [3][cout][9][cin] balanced base-256 digits of the weights  (exact fixed-point engine)
    int *wexp[13] = {nullptr};     // per-output-channel weight exponents
    uint32_t *max_slots = nullptr; // [16] FP32 bit patterns of the layer-output maxima (engine 3)
    bool have[13] = {false};
    // 0 = FP32 CUDA cores, 1 = tcgen05 kind::tf32, 2 = tcgen05 3xTF32 (FP32-accurate), 3 = tcgen05 kind::i8 exact fixed point
    int engine = 0;
};

int nct_conv3x3_tensorcore(nct_ctx *ctx, const float *in, const float *in_lo, const float *w_kmajor, const float *w_lo, const float *bias,
                           float *out, float *out_hi, float *out_lo, int H, int W, int Cin, int Cout);
int nct_tf32_split(nct_ctx *ctx, const float *x, float *hi, float *lo, size_t n);
void nct_tf32_split_host(const float *x, float *hi, float *lo, size_t n);
// conv_i8.cu
void nct_q_weight_digits_host(const float *w_oihw, int cin, int cout, int8_t *planes, int *wexp);
int nct_q_act_digits(nct_ctx *ctx, const float *x, const uint32_t *max_bits, uint8_t *planes, size_t pstride, size_t n);
int nct_q_pool_digits(nct_ctx *ctx, const float *x, const uint32_t *max_bits, uint8_t *planes, size_t pstride, int H, int W, int C, int Ho, int Wo);
int nct_conv3x3_i8(nct_ctx *ctx, const uint8_t *planes, size_t pstride, const uint32_t *in_max_bits, const int8_t *wplanes, const int *wexp,
                   const float *bias, float *out, uint32_t *out_max_bits, int *dbg_acc, int H, int W, int Cin, int Cout);

namespace {

// ------------------------------------------------------------------ preprocessing (Classifier::Preprocess)
// 8-bit BGR -> float, minus the BGR mean (103.939, 116.779, 123.68), kept HWC (3 channels)
__global__ void preprocess_kernel(const uint8_t *__restrict__ bgr, float *__restrict__ out, int npix)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    out[(size_t)p * 3 + 0] = __fsub_rn((float)bgr[(size_t)p * 3 + 0], 103.939f);
    out[(size_t)p * 3 + 1] = __fsub_rn((float)bgr[(size_t)p * 3 + 1], 116.779f);
    out[(size_t)p * 3 + 2] = __fsub_rn((float)bgr[(size_t)p * 3 + 2], 123.68f);
}

// ------------------------------------------------------------------ conv1_1: Cin = 3 (K = 27), direct
__global__ void __launch_bounds__(256) conv_first_kernel(const float *__restrict__ in, const float *__restrict__ wt,
                                                         const float *__restrict__ bias, float *__restrict__ out, int H, int W,
                                                         uint32_t *__restrict__ max_bits)
{
    // weights in shared memory as [27][o4 = 0..3][group = 0..3] float4: the four channel groups a warp's lanes belong to read
    // ONE contiguous 64-byte line per instruction (the plain [27][64] layout put groups 0 / 2 and 1 / 3 on the same banks:
    // 2-way conflicts, 27.7 M of them per launch, and the kernel was shared-memory bound at 96 % L1 throughput)
    __shared__ float4 sw4[27 * 16];
    __shared__ float sb[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) {
        const int tc = i / 64, o = i % 64;            // source: wt[tap * 3 + c][cout]
        const int g = o / 16, o4 = (o % 16) / 4, e = o % 4;
        reinterpret_cast<float *>(sw4)[((tc * 4 + o4) * 4 + g) * 4 + e] = wt[i];
    }
    if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    // one thread = TWO horizontally adjacent pixels x 16 output channels: every weight vector read from shared memory feeds
    // eight FMAs instead of four (the kernel is shared-memory-bandwidth bound); per output the FMA sequence is unchanged
    // (tap-major, channel-minor, out-of-image taps skipped)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs_x = (W + 1) >> 1;
    const int pp = t >> 2, cg = (t & 3) * 16;
    __shared__ uint32_t smax[8];
    float vmax = 0.f;
    if (pp < H * pairs_x) {   // (every thread stays alive for the block-wide maximum below)
        const int x0 = (pp % pairs_x) * 2, y = pp / pairs_x;
        const bool has1 = x0 + 1 < W;
        float acc[2][16];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int o = 0; o < 16; ++o) acc[q][o] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xa = x0 + kx - 1, xb = xa + 1;
                const bool oka = xa >= 0 && xa < W, okb = has1 && xb < W;   // (xb >= 0 always)
                const float *ipa = in + ((size_t)yy * W + (oka ? xa : 0)) * 3;
                const float *ipb = in + ((size_t)yy * W + (okb ? xb : 0)) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float va = ipa[c], vb = ipb[c];
                    const float4 *wp = sw4 + ((ky * 3 + kx) * 3 + c) * 16 + (cg >> 4);
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const float4 wv = wp[o4 * 4];
                        if (oka) {
                            acc[0][o4 * 4 + 0] = __fmaf_rn(va, wv.x, acc[0][o4 * 4 + 0]);
                            acc[0][o4 * 4 + 1] = __fmaf_rn(va, wv.y, acc[0][o4 * 4 + 1]);
                            acc[0][o4 * 4 + 2] = __fmaf_rn(va, wv.z, acc[0][o4 * 4 + 2]);
                            acc[0][o4 * 4 + 3] = __fmaf_rn(va, wv.w, acc[0][o4 * 4 + 3]);
                        }
                        if (okb) {
                            acc[1][o4 * 4 + 0] = __fmaf_rn(vb, wv.x, acc[1][o4 * 4 + 0]);
                            acc[1][o4 * 4 + 1] = __fmaf_rn(vb, wv.y, acc[1][o4 * 4 + 1]);
                            acc[1][o4 * 4 + 2] = __fmaf_rn(vb, wv.z, acc[1][o4 * 4 + 2]);
                            acc[1][o4 * 4 + 3] = __fmaf_rn(vb, wv.w, acc[1][o4 * 4 + 3]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            if (q == 1 && !has1) break;
            float4 *op = reinterpret_cast<float4 *>(out + ((size_t)y * W + x0 + q) * 64 + cg);
#pragma unroll
            for (int o = 0; o < 16; o += 4) {
                const float4 r = make_float4(fmaxf(acc[q][o] + sb[cg + o], 0.f), fmaxf(acc[q][o + 1] + sb[cg + o + 1], 0.f),
                                             fmaxf(acc[q][o + 2] + sb[cg + o + 2], 0.f), fmaxf(acc[q][o + 3] + sb[cg + o + 3], 0.f));
                op[o / 4] = r;
                vmax = fmaxf(fmaxf(vmax, fmaxf(r.x, r.y)), fmaxf(r.z, r.w));
            }
        }
    }
    if (max_bits) {
        // tensor maximum for the fixed-point engine's quantisation of the next layer's input: ONE atomic per block, and only
        // if it can raise the value
        const uint32_t mb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(vmax));
        if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mb;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t m = smax[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) m = max(m, smax[i]);
            if (m > *reinterpret_cast<volatile uint32_t *>(max_bits)) atomicMax(max_bits, m);
        }
    }
}

// ------------------------------------------------------------------ generic 3x3 conv, implicit GEMM on CUDA cores
// C[M = H*W pixels][N = Cout] = sum over (tap, cin) A[pixel shifted by tap][cin] * Wt[tap][cin][cout]
// block tile 128 (pixels) x 64 (couts), K chunk 16 channels of one tap, 256 threads, 8 x 4 outputs per thread.
constexpr int BM = 128, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) conv3x3_kernel(const float *__restrict__ in, const float *__restrict__ wt,
                                                      const float *__restrict__ bias, float *__restrict__ out, int H, int W,
                                                      int Cin, int Cout)
{
    __shared__ float As[2][BK][BM + 4];
    __shared__ float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int M = H * W;
    // loader roles: A tile = 128 pixels x 16 channels = 512 float4 -> 2 per thread; B tile = 16 x 64 = 256 float4 -> 1 per thread
    const int a_pix = tid >> 1;          // 0..127
    const int a_c4 = (tid & 1) * 8;      // channel offset 0 or 8 (two float4 each)
    const int pm = m0 + a_pix;
    const int px = pm % W, py = pm / W;
    const bool pvalid = pm < M;
    const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;
    // compute roles
    const int tx = tid & 15, ty = tid >> 4;  // tx -> 4 couts, ty -> 8 pixels
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int kchunks = Cin / BK;
    const int total = 9 * kchunks;
    float4 ra0, ra1, rb;
    auto load_global = [&](int it) {
        const int tap = it / kchunks, kc = (it % kchunks) * BK;
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        ra0 = ra1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pvalid && yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float4 *ip = reinterpret_cast<const float4 *>(in + ((size_t)yy * W + xx) * Cin + kc + a_c4);
            ra0 = __ldg(ip);
            ra1 = __ldg(ip + 1);
        }
        rb = __ldg(reinterpret_cast<const float4 *>(wt + ((size_t)tap * Cin + kc + b_k) * Cout + n0 + b_n4));
    };
    auto store_smem = [&](int buf) {
        As[buf][a_c4 + 0][a_pix] = ra0.x; As[buf][a_c4 + 1][a_pix] = ra0.y; As[buf][a_c4 + 2][a_pix] = ra0.z; As[buf][a_c4 + 3][a_pix] = ra0.w;
        As[buf][a_c4 + 4][a_pix] = ra1.x; As[buf][a_c4 + 5][a_pix] = ra1.y; As[buf][a_c4 + 6][a_pix] = ra1.z; As[buf][a_c4 + 7][a_pix] = ra1.w;
        *reinterpret_cast<float4 *>(&Bs[buf][b_k][b_n4]) = rb;
    };
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) load_global(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < total) {
            store_smem(buf ^ 1);
            __syncthreads();
        }
    }
    const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + n0 + tx * 4));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m < M) {
            float4 o;
            o.x = fmaxf(acc[i][0] + bb.x, 0.f);
            o.y = fmaxf(acc[i][1] + bb.y, 0.f);
            o.z = fmaxf(acc[i][2] + bb.z, 0.f);
            o.w = fmaxf(acc[i][3] + bb.w, 0.f);
            *reinterpret_cast<float4 *>(out + (size_t)m * Cout + n0 + tx * 4) = o;
        }
    }
}

// ------------------------------------------------------------------ 2x2 / stride 2 max pooling, ceil mode, NHWC
__global__ void maxpool_kernel(const float *__restrict__ in, float *__restrict__ out, int H, int W, int C, int Ho, int Wo)
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
    reinterpret_cast<float4 *>(out + (size_t)po * C)[c4] = m;
}

int pooled(int n) { return (n - 2 + 1) / 2 + 1; }  // ceil((n - 2) / 2) + 1

}  // namespace

void nct_vgg_free(nct_ctx *ctx)
{
    if (!ctx || !ctx->vgg) return;
    for (int i = 0; i < 13; ++i) {
        if (ctx->vgg->w[i]) cudaFree(ctx->vgg->w[i]);
        if (ctx->vgg->wk[i]) cudaFree(ctx->vgg->wk[i]);
        if (ctx->vgg->wk_lo[i]) cudaFree(ctx->vgg->wk_lo[i]);
        if (ctx->vgg->wk_hi[i]) cudaFree(ctx->vgg->wk_hi[i]);
        if (ctx->vgg->b[i]) cudaFree(ctx->vgg->b[i]);
        if (ctx->vgg->wq[i]) cudaFree(ctx->vgg->wq[i]);
        if (ctx->vgg->wexp[i]) cudaFree(ctx->vgg->wexp[i]);
    }
    if (ctx->vgg->max_slots) cudaFree(ctx->vgg->max_slots);
    delete ctx->vgg;
    ctx->vgg = nullptr;
}

extern "C" {

int nct_vgg19_num_layers(void) { return 13; }

const char *nct_vgg19_layer_name(int layer) { return (layer >= 0 && layer < 13) ? kTrunk[layer].name : nullptr; }

int nct_vgg19_layer_shape(int layer, int *cin, int *cout)
{
    if (layer < 0 || layer >= 13) return NCT_ERR_ARG;
    if (cin) *cin = kTrunk[layer].cin;
    if (cout) *cout = kTrunk[layer].cout;
    return NCT_OK;
}

int nct_vgg19_set_weights(nct_ctx *ctx, int layer, const float *w_oihw_host, const float *bias_host)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, layer >= 0 && layer < 13 && w_oihw_host && bias_host, "bad arguments");
    if (!ctx->vgg) ctx->vgg = new VggState();
    const int cin = kTrunk[layer].cin, cout = kTrunk[layer].cout;
    std::vector<float> wt((size_t)9 * cin * cout);
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i)
            for (int t = 0; t < 9; ++t) wt[((size_t)t * cin + i) * cout + o] = w_oihw_host[((size_t)o * cin + i) * 9 + t];
    std::vector<float> wkm((size_t)9 * cin * cout);
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i)
            for (int t = 0; t < 9; ++t) wkm[((size_t)o * 9 + t) * cin + i] = w_oihw_host[((size_t)o * cin + i) * 9 + t];
    VggState *v = ctx->vgg;
    if (!v->wk[layer]) NCT_CUDA(ctx, cudaMalloc(&v->wk[layer], wkm.size() * sizeof(float)));
    NCT_CUDA(ctx, cudaMemcpy(v->wk[layer], wkm.data(), wkm.size() * sizeof(float), cudaMemcpyHostToDevice));
    {   // exact hi/lo split for the 3xTF32 engine
        std::vector<float> hi(wkm.size()), lo(wkm.size());
        nct_tf32_split_host(wkm.data(), hi.data(), lo.data(), wkm.size());
        if (!v->wk_hi[layer]) NCT_CUDA(ctx, cudaMalloc(&v->wk_hi[layer], wkm.size() * sizeof(float)));
        if (!v->wk_lo[layer]) NCT_CUDA(ctx, cudaMalloc(&v->wk_lo[layer], wkm.size() * sizeof(float)));
        NCT_CUDA(ctx, cudaMemcpy(v->wk_hi[layer], hi.data(), wkm.size() * sizeof(float), cudaMemcpyHostToDevice));
        NCT_CUDA(ctx, cudaMemcpy(v->wk_lo[layer], lo.data(), wkm.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (cin >= 64) {   // balanced base-256 digit planes for the exact fixed-point engine
        std::vector<int8_t> wd(3 * wkm.size());
        std::vector<int> we(cout);
        nct_q_weight_digits_host(w_oihw_host, cin, cout, wd.data(), we.data());
        if (!v->wq[layer]) NCT_CUDA(ctx, cudaMalloc(&v->wq[layer], wd.size()));
        if (!v->wexp[layer]) NCT_CUDA(ctx, cudaMalloc(&v->wexp[layer], cout * sizeof(int)));
        NCT_CUDA(ctx, cudaMemcpy(v->wq[layer], wd.data(), wd.size(), cudaMemcpyHostToDevice));
        NCT_CUDA(ctx, cudaMemcpy(v->wexp[layer], we.data(), cout * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (!v->max_slots) NCT_CUDA(ctx, cudaMalloc(&v->max_slots, 16 * sizeof(uint32_t)));
    if (!v->w[layer]) NCT_CUDA(ctx, cudaMalloc(&v->w[layer], wt.size() * sizeof(float)));
    if (!v->b[layer]) NCT_CUDA(ctx, cudaMalloc(&v->b[layer], cout * sizeof(float)));
    NCT_CUDA(ctx, cudaMemcpy(v->w[layer], wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
    NCT_CUDA(ctx, cudaMemcpy(v->b[layer], bias_host, cout * sizeof(float), cudaMemcpyHostToDevice));
    v->have[layer] = true;
    return NCT_OK;
}

int nct_vgg19_set_engine(nct_ctx *ctx, int engine)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, engine >= 0 && engine <= 3, "engine must be 0 (FP32 CUDA cores), 1 (tcgen05 TF32), 2 (tcgen05 3xTF32) or 3 (tcgen05 INT8 exact fixed point)");
    if (!ctx->vgg) ctx->vgg = new VggState();
    ctx->vgg->engine = engine;
    return NCT_OK;
}

int nct_vgg19_level_dims(int h, int w, int dims[5][3])
{
    if (!dims || h <= 0 || w <= 0) return NCT_ERR_ARG;
    const int ch[5] = {512, 512, 256, 128, 64};
    int hh = h, ww = w;
    for (int l = 4; l >= 0; --l) {
        dims[l][0] = ch[l];
        dims[l][1] = hh;
        dims[l][2] = ww;
        hh = pooled(hh);
        ww = pooled(ww);
    }
    return NCT_OK;
}

int nct_vgg19_features(nct_ctx *ctx, const uint8_t *bgr_dev, int h, int w, int deepest_level, float *feat_dev[5])
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, bgr_dev && feat_dev && h >= 16 && w >= 16, "bad arguments (image must be at least 16 x 16)");
    NCT_REQUIRE(ctx, deepest_level >= 0 && deepest_level <= 4, "deepest_level must be in [0, 4] (0 = conv5_1)");
    VggState *v = ctx->vgg;
    NCT_REQUIRE(ctx, v != nullptr, "VGG-19 weights not loaded");
    int last_layer = 0;
    for (int i = 0; i < 13; ++i)
        if (kTrunk[i].level >= deepest_level && kTrunk[i].level >= 0) last_layer = i > last_layer ? i : last_layer;
    for (int i = 0; i <= last_layer; ++i)
        if (!v->have[i]) return nct_fail(ctx, NCT_ERR_STATE, "weights of %s not loaded", kTrunk[i].name);
    for (int l = deepest_level; l <= 4; ++l) NCT_REQUIRE(ctx, feat_dev[l] != nullptr, "feat_dev[%d] is null", l);

    const size_t max_act = (size_t)h * w * 64;  // largest activation: conv1_x
    float *buf0 = (float *)nct_scratch(ctx, "vgg_act0", sizeof(float) * max_act);
    float *buf1 = (float *)nct_scratch(ctx, "vgg_act1", sizeof(float) * max_act);
    float *inp = (float *)nct_scratch(ctx, "vgg_input", sizeof(float) * (size_t)h * w * 3);
    if (!buf0 || !buf1 || !inp) return NCT_ERR_NOMEM;
    if (v->engine == 3) {
        // ---- exact fixed-point engine (conv_i8.cu): conv1_1 on CUDA cores in the canonical FP32 order, every other layer
        // as 9 INT8 digit products on the tensor cores; activations travel as 4 digit planes scaled by the tensor maximum
        uint8_t *planes = (uint8_t *)nct_scratch(ctx, "vgg_qplanes", 4 * max_act);
        if (!planes) return NCT_ERR_NOMEM;
        NCT_CUDA(ctx, cudaMemsetAsync(v->max_slots, 0, 16 * sizeof(uint32_t), ctx->stream));
        preprocess_kernel<<<nct_div_up(h * w, 256), 256, 0, ctx->stream>>>(bgr_dev, inp, h * w);
        NCT_CHECK_LAUNCH(ctx);
        int H = h, W = w;
        float *bufs[2] = {buf0, buf1};
        int flip = 0;
        const float *cur = inp;
        for (int i = 0; i <= last_layer; ++i) {
            const ConvSpec &L = kTrunk[i];
            float *dst = (L.level >= 0) ? feat_dev[L.level] : bufs[flip];
            if (L.level < 0) flip ^= 1;
            if (dst == cur) return nct_fail(ctx, NCT_ERR_STATE, "internal: aliasing activation buffers");
            if (i == 0) {
                conv_first_kernel<<<nct_div_up(H * ((W + 1) / 2) * 4, 256), 256, 0, ctx->stream>>>(cur, v->w[0], v->b[0], dst, H, W, v->max_slots + 0);
                NCT_CHECK_LAUNCH(ctx);
            } else {
                int rc;
                if (L.pool_before) {
                    const int Ho = pooled(H), Wo = pooled(W);
                    rc = nct_q_pool_digits(ctx, cur, v->max_slots + (i - 1), planes, max_act, H, W, L.cin, Ho, Wo);
                    H = Ho;
                    W = Wo;
                } else {
                    rc = nct_q_act_digits(ctx, cur, v->max_slots + (i - 1), planes, max_act, (size_t)H * W * L.cin);
                }
                if (rc) return rc;
                rc = nct_conv3x3_i8(ctx, planes, max_act, v->max_slots + (i - 1), v->wq[i], v->wexp[i], v->b[i], dst, v->max_slots + i, nullptr,
                                    H, W, L.cin, L.cout);
                if (rc) return rc;
            }
            cur = dst;
        }
        return NCT_OK;
    }
    // 3xTF32: every activation travels with its truncation residual
    const bool x3 = v->engine == 2;
    float *lo_bufs[2] = {nullptr, nullptr}, *lo_feat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *hi_bufs[2] = {nullptr, nullptr}, *hi_feat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (x3) {
        lo_bufs[0] = (float *)nct_scratch(ctx, "vgg_act0_lo", sizeof(float) * max_act);
        lo_bufs[1] = (float *)nct_scratch(ctx, "vgg_act1_lo", sizeof(float) * max_act);
        hi_bufs[0] = (float *)nct_scratch(ctx, "vgg_act0_hi", sizeof(float) * max_act);
        hi_bufs[1] = (float *)nct_scratch(ctx, "vgg_act1_hi", sizeof(float) * max_act);
        if (!hi_bufs[0] || !hi_bufs[1]) return NCT_ERR_NOMEM;
        int dims[5][3];
        nct_vgg19_level_dims(h, w, dims);
        char name[32];
        for (int l = deepest_level; l <= 4; ++l) {
            snprintf(name, sizeof(name), "vgg_feat_lo%d", l);
            lo_feat[l] = (float *)nct_scratch(ctx, name, sizeof(float) * (size_t)dims[l][0] * dims[l][1] * dims[l][2]);
            snprintf(name, sizeof(name), "vgg_feat_hi%d", l);
            hi_feat[l] = (float *)nct_scratch(ctx, name, sizeof(float) * (size_t)dims[l][0] * dims[l][1] * dims[l][2]);
            if (!lo_feat[l] || !hi_feat[l]) return NCT_ERR_NOMEM;
        }
        if (!lo_bufs[0] || !lo_bufs[1]) return NCT_ERR_NOMEM;
    }
    const float *cur_lo = nullptr, *cur_hi = nullptr;
    preprocess_kernel<<<nct_div_up(h * w, 256), 256, 0, ctx->stream>>>(bgr_dev, inp, h * w);
    NCT_CHECK_LAUNCH(ctx);

    int H = h, W = w;
    const float *cur = inp;
    float *bufs[2] = {buf0, buf1};
    int flip = 0;
    for (int i = 0; i <= last_layer; ++i) {
        const ConvSpec &L = kTrunk[i];
        if (L.pool_before) {
            const int Ho = pooled(H), Wo = pooled(W);
            float *dst = bufs[flip];
            flip ^= 1;
            const long long threads = (long long)Ho * Wo * (L.cin / 4);
            maxpool_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(cur, dst, H, W, L.cin, Ho, Wo);
            NCT_CHECK_LAUNCH(ctx);
            if (x3) {
                float *dlo = lo_bufs[flip ^ 1], *dhi = hi_bufs[flip ^ 1];
                int rc = nct_tf32_split(ctx, dst, dhi, dlo, (size_t)Ho * Wo * L.cin);
                if (rc) return rc;
                cur_lo = dlo;
                cur_hi = dhi;
            }
            cur = dst;
            H = Ho;
            W = Wo;
        }
        float *dst = (L.level >= 0) ? feat_dev[L.level] : bufs[flip];
        float *dst_lo = x3 ? ((L.level >= 0) ? lo_feat[L.level] : lo_bufs[flip]) : nullptr;
        float *dst_hi = x3 ? ((L.level >= 0) ? hi_feat[L.level] : hi_bufs[flip]) : nullptr;
        if (L.level < 0) flip ^= 1;
        if (dst == cur) return nct_fail(ctx, NCT_ERR_STATE, "internal: aliasing activation buffers");
        if (i == 0) {
            conv_first_kernel<<<nct_div_up(H * ((W + 1) / 2) * 4, 256), 256, 0, ctx->stream>>>(cur, v->w[0], v->b[0], dst, H, W, nullptr);
        } else if (v->engine >= 1) {
            int rc = nct_conv3x3_tensorcore(ctx, x3 ? cur_hi : cur, x3 ? cur_lo : nullptr, x3 ? v->wk_hi[i] : v->wk[i],
                                            x3 ? v->wk_lo[i] : nullptr, v->b[i], dst, dst_hi, dst_lo, H, W, L.cin, L.cout);
            if (rc) return rc;
            cur = dst;
            cur_lo = dst_lo;
            cur_hi = dst_hi;
            continue;
        } else {
            dim3 grid(nct_div_up(H * W, BM), L.cout / BN);
            conv3x3_kernel<<<grid, 256, 0, ctx->stream>>>(cur, v->w[i], v->b[i], dst, H, W, L.cin, L.cout);
        }
        NCT_CHECK_LAUNCH(ctx);
        if (x3 && i == 0) {  // conv1_1 runs on CUDA cores: split its output here
            int rc = nct_tf32_split(ctx, dst, dst_hi, dst_lo, (size_t)H * W * L.cout);
            if (rc) return rc;
            cur_lo = dst_lo;
            cur_hi = dst_hi;
        }
        cur = dst;
    }
    return NCT_OK;
}

}  // extern "C"
