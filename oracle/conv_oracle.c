/* oracle/conv_oracle.c -- canonical-order FP32 restatement of the VGG-19 trunk's layers.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); never linked into the product.
 *
 * Follows
 *   conv   caffe/layers/base_conv_layer.cpp:257-282 (3x3, pad 1, stride 1 cross-correlation + bias; weights OIHW)
 *          and the naive reference loop of caffe/test/test_convolution_layer.cpp:22-139
 *   relu   caffe/layers/relu_layer.cpp:9-19 (in place after every conv in the deploy prototxt)
 *   pool   caffe/layers/pooling_layer.cpp:86-170 (2x2 / 2 MAX, ceil mode, window clipped at the border,
 *          init -FLT_MAX, strict >)
 *
 * Parity unpinned by the reference (its conv numerics depend on the cuDNN algorithm picked at run time; Caffe's own
 * test tolerance is 1e-4, test_convolution_layer.cpp:231-265).  What this file adds over oracle/vgg.py (torch-CPU,
 * summation order unspecified) is a DEFINED summation order, so that the product's FP32 engine can be compared
 * bit for bit and the whole pipeline can be checked end to end without feature rounding noise:
 *
 *   acc = +0
 *   for tap = ky*3 + kx in 0..8 (ky outer), skipping taps whose input pixel is outside the image (zero padding):
 *       for c = 0 .. Cin-1 ascending:   acc = fmaf(in[y+ky-1][x+kx-1][c], w[o][c][ky][kx], acc)
 *   out = max(acc + bias[o], 0)
 *
 * (A padded tap would contribute fmaf(0, w, acc) == acc for finite w, so skipping it is the same thing.)
 * Layout: HWC activations, OIHW weights as Caffe stores them.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_conv3x3_relu_canon(const float *in, const float *w_oihw, const float *bias, float *out, int H, int W, int Cin, int Cout)
{
    /* re-lay the weights as [tap][cin][cout] so the inner loop runs over independent outputs (vectorisable without
     * changing any single output's operation order) */
    float *wt = (float *)malloc(sizeof(float) * 9 * (size_t)Cin * Cout);
    for (int o = 0; o < Cout; ++o)
        for (int c = 0; c < Cin; ++c)
            for (int t = 0; t < 9; ++t) wt[((size_t)t * Cin + c) * Cout + o] = w_oihw[((size_t)o * Cin + c) * 9 + t];
#pragma omp parallel
    {
        float *acc = (float *)malloc(sizeof(float) * Cout);
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                for (int o = 0; o < Cout; ++o) acc[o] = 0.f;
                for (int t = 0; t < 9; ++t) {
                    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const float *ip = in + ((size_t)yy * W + xx) * Cin;
                    const float *wp = wt + (size_t)t * Cin * Cout;
                    for (int c = 0; c < Cin; ++c) {
                        const float v = ip[c];
                        const float *wr = wp + (size_t)c * Cout;
#pragma omp simd
                        for (int o = 0; o < Cout; ++o) acc[o] = fmaf(v, wr[o], acc[o]);
                    }
                }
                float *op = out + ((size_t)y * W + x) * Cout;
                for (int o = 0; o < Cout; ++o) {
                    const float s = acc[o] + bias[o];
                    op[o] = s > 0.f ? s : 0.f;
                }
            }
        free(acc);
    }
    free(wt);
}

void orc_maxpool2x2_ceil(const float *in, float *out, int H, int W, int C)
{
    const int Ho = (H - 2 + 1) / 2 + 1, Wo = (W - 2 + 1) / 2 + 1;
#pragma omp parallel for
    for (int yo = 0; yo < Ho; ++yo)
        for (int xo = 0; xo < Wo; ++xo)
            for (int c = 0; c < C; ++c) {
                float m = -FLT_MAX;
                for (int dy = 0; dy < 2; ++dy)
                    for (int dx = 0; dx < 2; ++dx) {
                        const int y = 2 * yo + dy, x = 2 * xo + dx;
                        if (y < H && x < W) {
                            const float v = in[((size_t)y * W + x) * C + c];
                            if (v > m) m = v;
                        }
                    }
                out[((size_t)yo * Wo + xo) * C + c] = m;
            }
}

/* Classifier::Preprocess (NCT/Classifier.cpp:211-275): 8U BGR -> float minus the BGR mean, kept HWC */
void orc_preprocess_bgr(const unsigned char *bgr, float *out, int npix)
{
    const float mean[3] = {103.939f, 116.779f, 123.68f};
    for (int p = 0; p < npix; ++p)
        for (int c = 0; c < 3; ++c) out[(size_t)p * 3 + c] = (float)bgr[(size_t)p * 3 + c] - mean[c];
}
