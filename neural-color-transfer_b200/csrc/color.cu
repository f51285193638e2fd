// Colour-space and per-pixel colour-fit kernels.
//
// Replaces, on the device and without host round trips:
//   cvtColor BGR2Lab / Lab2BGR (8U)      NCT/main.cu:352,371; CT/ColorTransfer.h:58; CT/ColorTransfer.cpp:1469
//   resize INTER_LINEAR 8UC3 / 64FC3     NCT/main.cu:106-107; CT/ColorTransfer.cpp:462-463
//   build_accumTable_downsample + the 3x3 mean/std gain-bias fit   CT/ColorTransfer.cpp:425-455, 1194-1265
//   confidence weights                   CT/ColorTransfer.cpp:1302-1340
//   roughness map                        CT/ColorTransfer.cpp:466-489
//   apply + convertTo(8U) + Lab2BGR      CT/ColorTransfer.cpp:1436-1469
//
// The OpenCV arithmetic is third-party (OpenCV 2.4.10, not in the reference tree); it is restated
// here from OpenCV's published fixed-point algorithms (integer Lab with gamma / cube-root / inverse
// gamma lookup tables, 11-bit fixed-point bilinear resize with the 2x2-area shortcut for exact 2:1
// downscales, float-coefficient bilinear for 64F) and verified bit-for-bit against cv2 over all
// 2^24 colours in tests/.  All kernels here are trivially HBM-bound streaming kernels.
#include <mutex>
#include "device_utils.cuh"
#include <cmath>
#include <cstring>
#include <vector>

namespace {

// ------------------------------------------------------------------ lookup tables (host-built)
constexpr int LAB_SHIFT = 12, GAMMA_SHIFT = 3, LAB_SHIFT2 = LAB_SHIFT + GAMMA_SHIFT;
constexpr int CBRT_TAB_SIZE = 256 * 3 / 2 * (1 << GAMMA_SHIFT);  // 3072
constexpr int BASE_SHIFT = 14, BASE = 1 << BASE_SHIFT;
constexpr int INV_GAMMA_SHIFT = 12, INV_GAMMA_TAB_SIZE = 1 << INV_GAMMA_SHIFT;

struct LabTables {
    uint16_t gamma[256];             // sRGB -> linear, scaled by 255 * 8
    uint16_t cbrt[CBRT_TAB_SIZE];    // f(t) of CIE Lab, scaled by 2^15
    uint16_t l2y[256], l2fy[256];    // L -> Y, L -> fy (scaled by 2^14)
    uint16_t inv_gamma[INV_GAMMA_TAB_SIZE];  // linear -> sRGB 8 bit
};

// round-half-to-even of a float (cvRound)
inline int cv_round(float v) { return (int)lrintf(v); }

// OpenCV's cube root: quartic rational approximation evaluated in double on the mantissa,
// result mantissa truncated to 23 bits.
float cv_cbrt(float x)
{
    uint32_t bits;
    memcpy(&bits, &x, 4);
    if ((bits & 0x7fffffffu) == 0) return 0.f;
    int ex = (int)((bits >> 23) & 0xff) - 127;
    const uint32_t frac = bits & ((1u << 23) - 1);
    int shx = ex % 3;
    shx -= shx >= 0 ? 3 : 0;
    ex = (ex - shx) / 3 - 1;
    uint64_t db = ((uint64_t)(shx + 1023) << 52) | ((uint64_t)frac << 29);
    double fr;
    memcpy(&fr, &db, 8);
    const double num = ((((45.2548339756803022511987494 * fr + 192.2798368355061050458134625) * fr +
                          119.1654824285581628956914143) * fr + 13.43250139086239872172837314) * fr +
                        0.1636161226585754240958355063);
    const double den = ((((14.80884093219134573786480845 * fr + 151.9714051044435648658557668) * fr +
                          168.5254414101568283957668343) * fr + 33.9905941350215598754191872) * fr + 1.0);
    const double r = num / den;
    uint64_t rb;
    memcpy(&rb, &r, 8);
    const int rexp = (int)((rb >> 52) & 0x7ff) - 1023;
    const uint32_t out = ((uint32_t)((ex + 127 + (rexp + 1)) & 0xff) << 23) | (uint32_t)((rb & ((1ull << 52) - 1)) >> 29);
    float y;
    memcpy(&y, &out, 4);
    return y;
}

void build_lab_tables(LabTables &t)
{
    for (int i = 0; i < 256; ++i) {
        const float x = (float)i / 255.0f;
        const double xd = (double)x;
        const double v = xd <= 809.0 / 20000.0 ? xd / 12.92 : pow((xd + 11.0 / 200.0) / (1.0 + 11.0 / 200.0), 2.4);
        t.gamma[i] = (uint16_t)cv_round((float)(255 * (1 << GAMMA_SHIFT)) * (float)v);
    }
    const float lthresh = 216.0f / 24389.0f, lscale = 841.0f / 108.0f, lbias = 16.0f / 116.0f;
    for (int i = 0; i < CBRT_TAB_SIZE; ++i) {
        const float x = (float)i / (float)(255 * (1 << GAMMA_SHIFT));
        float v;
        if (x < lthresh) v = (float)((double)x * (double)lscale + (double)lbias);  // fused multiply-add in float
        else v = cv_cbrt(x);
        t.cbrt[i] = (uint16_t)cv_round((float)(1 << LAB_SHIFT2) * v);
    }
    for (int i = 0; i < 256; ++i) {
        int y, ify;
        if (i <= 20) {
            y = cv_round((float)(i * BASE * 20 * 9) / (float)(17 * 29 * 29 * 29));
            ify = cv_round((float)BASE * (16.0f / 116.0f + (float)(i * 5) / (float)(3 * 17 * 29)));
        } else {
            const float fy = (float)(i * 100 * BASE) / (float)(255 * 116) + (float)(16 * BASE) / 116.0f;
            ify = cv_round(fy);
            const float fy2 = fy * fy;
            const float fy3 = fy2 * fy;
            y = cv_round(fy3 / (float)((double)BASE * BASE));
        }
        t.l2y[i] = (uint16_t)y;
        t.l2fy[i] = (uint16_t)ify;
    }
    for (int i = 0; i < INV_GAMMA_TAB_SIZE; ++i) {
        const float x = (float)i / (float)INV_GAMMA_TAB_SIZE;
        const double xd = (double)x;
        const float v = xd <= 7827.0 / 2500000.0 ? (float)(xd * 12.92) : (float)(pow(xd, 5.0 / 12.0) * (1.0 + 11.0 / 200.0) - 11.0 / 200.0);
        t.inv_gamma[i] = (uint16_t)cv_round(255.0f * v);
    }
}

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

__device__ __forceinline__ void bgr2lab_px(const LabTables *__restrict__ t, int Bv, int Gv, int Rv, uint8_t &L, uint8_t &a,
                                           uint8_t &b)
{
    const int R = t->gamma[Rv], G = t->gamma[Gv], B = t->gamma[Bv];
    const int fX = t->cbrt[descale(R * 1777 + G * 1541 + B * 778, LAB_SHIFT)];
    const int fY = t->cbrt[descale(R * 871 + G * 2929 + B * 296, LAB_SHIFT)];
    const int fZ = t->cbrt[descale(R * 73 + G * 448 + B * 3575, LAB_SHIFT)];
    const int Lscale = (116 * 255 + 50) / 100;
    const int Lshift = -((16 * 255 * (1 << LAB_SHIFT2) + 50) / 100);
    L = sat_u8(descale(Lscale * fY + Lshift, LAB_SHIFT2));
    a = sat_u8(descale(500 * (fX - fY) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2));
    b = sat_u8(descale(200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2));
}

__device__ __forceinline__ int ab_to_xz(int i)
{
    if (i <= 3390) return i * 108 / 841 - BASE * 16 / 116 * 108 / 841;
    return i * i / BASE * i / BASE;
}

__device__ __forceinline__ void lab2bgr_px(const LabTables *__restrict__ t, int LL, int aa, int bb, uint8_t &Bo, uint8_t &Go,
                                           uint8_t &Ro)
{
    const int y = t->l2y[LL], ify = t->l2fy[LL];
    const int adiv = ((5 * aa * 53687 + (1 << 7)) >> 13) - 128 * BASE / 500;
    const int bdiv = ((bb * 41943 + (1 << 4)) >> 9) - 128 * BASE / 200 + 1;
    const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
    const int shift = LAB_SHIFT + (BASE_SHIFT - INV_GAMMA_SHIFT);
    int ro = descale(12615 * x + -6296 * y + -2223 * z, shift);
    int go = descale(-3773 * x + 7684 * y + 185 * z, shift);
    int bo = descale(217 * x + -836 * y + 4715 * z, shift);
    ro = min(max(ro, 0), INV_GAMMA_TAB_SIZE - 1);
    go = min(max(go, 0), INV_GAMMA_TAB_SIZE - 1);
    bo = min(max(bo, 0), INV_GAMMA_TAB_SIZE - 1);
    Ro = (uint8_t)t->inv_gamma[ro];
    Go = (uint8_t)t->inv_gamma[go];
    Bo = (uint8_t)t->inv_gamma[bo];
}

__global__ void bgr2lab_kernel(const LabTables *__restrict__ t, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                               int npix)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    uint8_t L, a, b;
    bgr2lab_px(t, src[p * 3], src[p * 3 + 1], src[p * 3 + 2], L, a, b);
    dst[p * 3] = L;
    dst[p * 3 + 1] = a;
    dst[p * 3 + 2] = b;
}

__global__ void lab2bgr_kernel(const LabTables *__restrict__ t, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                               int npix)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    uint8_t B, G, R;
    lab2bgr_px(t, src[p * 3], src[p * 3 + 1], src[p * 3 + 2], B, G, R);
    dst[p * 3] = B;
    dst[p * 3 + 1] = G;
    dst[p * 3 + 2] = R;
}

// ------------------------------------------------------------------ bilinear resize
// coefficient of destination index d for a sn -> dn resize (float maths like cv::resize)
__device__ __forceinline__ void lin_coef(int d, double scale, int sn, bool clamp_ofs, int &s0, float &f)
{
    float fx = (float)(((double)d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (clamp_ofs) {
        if (sx < 0) { fx = 0.f; sx = 0; }
        if (sx >= sn - 1) { fx = 0.f; sx = sn - 1; }
    }
    s0 = sx;
    f = fx;
}

__device__ __forceinline__ int rne_short(float v) { return min(max(__float2int_rn(v), -32768), 32767); }

__global__ void resize_u8c3_kernel(const uint8_t *__restrict__ src, int sh, int sw, uint8_t *__restrict__ dst, int dh, int dw,
                                   double scale_x, double scale_y)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    if (sw == 2 * dw && sh == 2 * dh) {  // exact 2:1 -> INTER_AREA fast path
        const uint8_t *r0 = src + ((size_t)(2 * dy) * sw + 2 * dx) * 3, *r1 = r0 + (size_t)sw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[((size_t)dy * dw + dx) * 3 + c] = (uint8_t)((r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2);
        return;
    }
    int sx, sy;
    float fx, fy;
    lin_coef(dx, scale_x, sw, true, sx, fx);
    lin_coef(dy, scale_y, sh, false, sy, fy);
    const int a0 = rne_short(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = rne_short(__fmul_rn(fx, 2048.f));
    const int b0 = rne_short(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = rne_short(__fmul_rn(fy, 2048.f));
    const int x1 = min(sx + 1, sw - 1);
    const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int S0 = src[((size_t)y0 * sw + sx) * 3 + c] * a0 + src[((size_t)y0 * sw + x1) * 3 + c] * a1;
        const int S1 = src[((size_t)y1 * sw + sx) * 3 + c] * a0 + src[((size_t)y1 * sw + x1) * 3 + c] * a1;
        const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        dst[((size_t)dy * dw + dx) * 3 + c] = sat_u8(v);
    }
}

__global__ void resize_f64c3_kernel(const double *__restrict__ src, int sh, int sw, double *__restrict__ dst, int dh, int dw,
                                    double scale_x, double scale_y)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    int sx, sy;
    float fx, fy;
    lin_coef(dx, scale_x, sw, true, sx, fx);
    lin_coef(dy, scale_y, sh, false, sy, fy);
    const double a0 = (double)__fsub_rn(1.f, fx), a1 = (double)fx;
    const double b0 = (double)__fsub_rn(1.f, fy), b1 = (double)fy;
    const int x1 = min(sx + 1, sw - 1);
    const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double S0 = __dadd_rn(__dmul_rn(src[((size_t)y0 * sw + sx) * 3 + c], a0), __dmul_rn(src[((size_t)y0 * sw + x1) * 3 + c], a1));
        const double S1 = __dadd_rn(__dmul_rn(src[((size_t)y1 * sw + sx) * 3 + c], a0), __dmul_rn(src[((size_t)y1 * sw + x1) * 3 + c], a1));
        dst[((size_t)dy * dw + dx) * 3 + c] = __dadd_rn(__dmul_rn(S0, b0), __dmul_rn(S1, b1));
    }
}

// ------------------------------------------------------------------ local gain / bias fit
__global__ void local_fit_kernel(const uint8_t *__restrict__ cnt, const uint8_t *__restrict__ stl, int h, int w, double eps,
                                 double *__restrict__ a, double *__restrict__ b)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const int x0 = max(x - 1, 0), x1 = min(x + 2, w), y0 = max(y - 1, 0), y1 = min(y + 2, h);
    long long c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0}, s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            const uint8_t *pc = cnt + ((size_t)yy * w + xx) * 3, *ps = stl + ((size_t)yy * w + xx) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int vc = pc[c], vs = ps[c];
                c1[c] += vc; c2[c] += vc * vc;
                s1[c] += vs; s2[c] += vs * vs;
            }
        }
    const double n = (double)((x1 - x0) * (y1 - y0));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double cm = __ddiv_rn((double)c1[c], n);
        const double cv = __dsqrt_rn(fmax(__dsub_rn(__ddiv_rn((double)c2[c], n), __dmul_rn(cm, cm)), 0.0));
        const double sm = __ddiv_rn((double)s1[c], n);
        const double sv = __dsqrt_rn(fmax(__dsub_rn(__ddiv_rn((double)s2[c], n), __dmul_rn(sm, sm)), 0.0));
        const double av = __ddiv_rn(sv, __dadd_rn(cv, eps));
        a[((size_t)y * w + x) * 3 + c] = av;
        b[((size_t)y * w + x) * 3 + c] = __dmul_rn(__dsub_rn(sm, __dmul_rn(cm, av)), 1.0 / 255.0);
    }
}

// ------------------------------------------------------------------ min / max + confidence weights
__global__ void minmax_partial_kernel(const float *__restrict__ v, int n, float *__restrict__ part)
{
    __shared__ float smin[256], smax[256];
    float lo = INFINITY, hi = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = v[i];
        lo = fminf(lo, x);
        hi = fmaxf(hi, x);
    }
    smin[threadIdx.x] = lo;
    smax[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]);
            smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = smin[0];
        part[2 * blockIdx.x + 1] = smax[0];
    }
}

__global__ void minmax_final_kernel(const float *__restrict__ part, int nparts, float *__restrict__ out)
{
    __shared__ float smin[256], smax[256];
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        lo = fminf(lo, part[2 * i]);
        hi = fmaxf(hi, part[2 * i + 1]);
    }
    smin[threadIdx.x] = lo;
    smax[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]);
            smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = smin[0];
        out[1] = smax[0];
    }
}

__global__ void confidence_kernel(const float *__restrict__ err, const float *__restrict__ mm, int n, double *__restrict__ wgt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lo = (double)mm[0], hi = (double)mm[1];
    const double e = __ddiv_rn(__dsub_rn((double)err[i], lo), __dsub_rn(hi, lo));
    const double v = __dsub_rn(1.0, e);
    wgt[i] = v < 1e-6 ? 1e-6 : v;  // std::max(1.0 - err, 1e-6)
}

// ------------------------------------------------------------------ roughness / apply
__global__ void roughness_kernel(const uint8_t *__restrict__ lab, const double *__restrict__ a, const double *__restrict__ b,
                                 int n, double *__restrict__ rough)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double col = __dmul_rn((double)lab[(size_t)i * 3 + 2], 1.0 / 255.0);
    const double nc = __dadd_rn(__dmul_rn(col, a[(size_t)i * 3 + 2]), b[(size_t)i * 3 + 2]);
    rough[i] = (nc < 0.0 || nc > 1.0) ? 1e-6 : 1.0;
}

__global__ void apply_kernel(const LabTables *__restrict__ t, const uint8_t *__restrict__ lab, const double *__restrict__ a,
                             const double *__restrict__ b, int n, uint8_t *__restrict__ out_bgr, uint8_t *__restrict__ out_lab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double col = __dmul_rn((double)lab[(size_t)i * 3 + c], 1.0 / 255.0);
        double v = __dadd_rn(__dmul_rn(col, a[(size_t)i * 3 + c]), b[(size_t)i * 3 + c]);
        v = fmin(fmax(v, 0.0), 1.0);
        q[c] = min(max(__double2int_rn(__dmul_rn(v, 255.0)), 0), 255);  // convertTo(CV_8U, 255): cvRound, saturate
    }
    if (out_lab) {
        out_lab[(size_t)i * 3] = (uint8_t)q[0];
        out_lab[(size_t)i * 3 + 1] = (uint8_t)q[1];
        out_lab[(size_t)i * 3 + 2] = (uint8_t)q[2];
    }
    uint8_t B, G, R;
    lab2bgr_px(t, q[0], q[1], q[2], B, G, R);
    out_bgr[(size_t)i * 3] = B;
    out_bgr[(size_t)i * 3 + 1] = G;
    out_bgr[(size_t)i * 3 + 2] = R;
}

const LabTables *lab_tables(nct_ctx *ctx)
{
    auto it = ctx->scratch.find("lab_tables");
    if (it != ctx->scratch.end() && it->second.ptr) return (const LabTables *)it->second.ptr;
    void *d = nct_scratch(ctx, "lab_tables", sizeof(LabTables));
    if (!d) return nullptr;
    static LabTables host;  // identical for every ctx; contexts may be driven from several host threads at once
    static std::once_flag built;
    std::call_once(built, [] { build_lab_tables(host); });
    if (cudaMemcpyAsync(d, &host, sizeof(LabTables), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return nullptr;
    cudaStreamSynchronize(ctx->stream);
    return (const LabTables *)d;
}

}  // namespace

extern "C" {

int nct_bgr2lab_u8(nct_ctx *ctx, const uint8_t *bgr_dev, uint8_t *lab_dev, int npix)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, bgr_dev && lab_dev && npix > 0, "bad arguments");
    const LabTables *t = lab_tables(ctx);
    if (!t) return nct_fail(ctx, NCT_ERR_CUDA, "Lab tables upload failed");
    bgr2lab_kernel<<<nct_div_up(npix, 256), 256, 0, ctx->stream>>>(t, bgr_dev, lab_dev, npix);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_lab2bgr_u8(nct_ctx *ctx, const uint8_t *lab_dev, uint8_t *bgr_dev, int npix)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, bgr_dev && lab_dev && npix > 0, "bad arguments");
    const LabTables *t = lab_tables(ctx);
    if (!t) return nct_fail(ctx, NCT_ERR_CUDA, "Lab tables upload failed");
    lab2bgr_kernel<<<nct_div_up(npix, 256), 256, 0, ctx->stream>>>(t, lab_dev, bgr_dev, npix);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_resize_linear_u8c3(nct_ctx *ctx, const uint8_t *src_dev, int sh, int sw, uint8_t *dst_dev, int dh, int dw)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src_dev && dst_dev && sh > 0 && sw > 0 && dh > 0 && dw > 0, "bad arguments");
    dim3 block(32, 8), grid(nct_div_up(dw, 32), nct_div_up(dh, 8));
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    resize_u8c3_kernel<<<grid, block, 0, ctx->stream>>>(src_dev, sh, sw, dst_dev, dh, dw, 1.0 / inv_x, 1.0 / inv_y);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_resize_linear_f64c3(nct_ctx *ctx, const double *src_dev, int sh, int sw, double *dst_dev, int dh, int dw)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, src_dev && dst_dev && src_dev != dst_dev && sh > 0 && sw > 0 && dh > 0 && dw > 0, "bad arguments");
    dim3 block(32, 8), grid(nct_div_up(dw, 32), nct_div_up(dh, 8));
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    resize_f64c3_kernel<<<grid, block, 0, ctx->stream>>>(src_dev, sh, sw, dst_dev, dh, dw, 1.0 / inv_x, 1.0 / inv_y);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_local_fit(nct_ctx *ctx, const uint8_t *cnt_lab_dev, const uint8_t *stl_lab_dev, int h, int w, double eps, double *a_dev,
                  double *b_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, cnt_lab_dev && stl_lab_dev && a_dev && b_dev && h > 0 && w > 0, "bad arguments");
    dim3 block(32, 8), grid(nct_div_up(w, 32), nct_div_up(h, 8));
    local_fit_kernel<<<grid, block, 0, ctx->stream>>>(cnt_lab_dev, stl_lab_dev, h, w, eps, a_dev, b_dev);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_confidence_weights(nct_ctx *ctx, const float *err_dev, int n, double *weight_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, err_dev && weight_dev && n > 0, "bad arguments");
    int blocks = nct_div_up(n, 256);
    if (blocks > 1024) blocks = 1024;
    float *part = (float *)nct_scratch(ctx, "minmax_part", sizeof(float) * (2 * 1024 + 2));
    if (!part) return NCT_ERR_NOMEM;
    float *mm = part + 2 * 1024;
    minmax_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(err_dev, n, part);
    NCT_CHECK_LAUNCH(ctx);
    minmax_final_kernel<<<1, 256, 0, ctx->stream>>>(part, blocks, mm);
    NCT_CHECK_LAUNCH(ctx);
    confidence_kernel<<<nct_div_up(n, 256), 256, 0, ctx->stream>>>(err_dev, mm, n, weight_dev);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_upsample_coefficients(nct_ctx *ctx, const double *a_lvl_dev, const double *b_lvl_dev, int h, int w,
                              const uint8_t *cnt_lab_full_dev, int H, int W, double *a_full_dev, double *b_full_dev,
                              double *rough_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, a_lvl_dev && b_lvl_dev && cnt_lab_full_dev && a_full_dev && b_full_dev && rough_dev, "null pointer");
    if (W > w || H > h) {
        int rc = nct_resize_linear_f64c3(ctx, a_lvl_dev, h, w, a_full_dev, H, W);
        if (rc) return rc;
        rc = nct_resize_linear_f64c3(ctx, b_lvl_dev, h, w, b_full_dev, H, W);
        if (rc) return rc;
    } else {
        if (a_full_dev != a_lvl_dev)
            NCT_CUDA(ctx, cudaMemcpyAsync(a_full_dev, a_lvl_dev, sizeof(double) * 3 * (size_t)H * W, cudaMemcpyDeviceToDevice, ctx->stream));
        if (b_full_dev != b_lvl_dev)
            NCT_CUDA(ctx, cudaMemcpyAsync(b_full_dev, b_lvl_dev, sizeof(double) * 3 * (size_t)H * W, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    roughness_kernel<<<nct_div_up(H * W, 256), 256, 0, ctx->stream>>>(cnt_lab_full_dev, a_full_dev, b_full_dev, H * W, rough_dev);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_apply_coefficients(nct_ctx *ctx, const uint8_t *cnt_lab_full_dev, const double *a_dev, const double *b_dev, int H, int W,
                           uint8_t *out_bgr_dev, uint8_t *out_lab_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, cnt_lab_full_dev && a_dev && b_dev && out_bgr_dev && H > 0 && W > 0, "bad arguments");
    const LabTables *t = lab_tables(ctx);
    if (!t) return nct_fail(ctx, NCT_ERR_CUDA, "Lab tables upload failed");
    apply_kernel<<<nct_div_up(H * W, 256), 256, 0, ctx->stream>>>(t, cnt_lab_full_dev, a_dev, b_dev, H * W, out_bgr_dev, out_lab_dev);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // extern "C"
