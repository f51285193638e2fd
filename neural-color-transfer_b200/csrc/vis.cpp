// Per-level debug artefacts: the reference's ENABLE_VIS build (NCT/main.cu:169-173, 333-347, 361-364, 382-422;
// reconstruct_flow NCT/GeneralizedPatchMatch.cu:337-353; CT/ColorTransfer.cpp:222-253 visualizeClusterRandom,
// :1127-1178 getHeat, :1267-1300 / :1382-1415 / :1450-1462 the a / b maps, :333-351 the dilated cluster map).
// Off by default, as in the reference (CT/Config.h:8 has the define commented out); switched on per context with
// nct_set_vis (CLI: -vis 1).  Same file names: <out>/<prefix>_{aFlow,bFlow,tCnt,tStl,knn,errMap,refine_init,
// refine_nonlocal,aVis,aVis_init,aVis_nonlocal,bVis,bVis_init,bVis_nonlocal}_<level>.png and <prefix>_cluster_small.png.
// Not written: patchVis (a 2x-height mosaic of every patch pair).  The reference colours clusters from a fixed table of
// 260 random colours (CT/Config.h:17-51); here the palette is a hash of the cluster index -- the maps are for the eye.
// Everything is rendered on the host from copies of the device arrays (a debugging path: the reference renders on the
// host too), with the reference's arithmetic (float ratios for the flow, the five-segment heat ramp, int(a * 50),
// int(b * 255 + 127)).
#include "nct_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

extern "C" int nct_png_write(const char *path, const uint8_t *bgr, int h, int w);

namespace {

std::string vis_path(nct_ctx *ctx, const char *what, int level)
{
    char tail[96];
    if (level >= 0) snprintf(tail, sizeof(tail), "_%s_%d.png", what, level);
    else snprintf(tail, sizeof(tail), "_%s.png", what);
    return ctx->vis_dir + "/" + ctx->vis_prefix + tail;
}

int write_png(nct_ctx *ctx, const char *what, int level, const std::vector<uint8_t> &bgr, int h, int w)
{
    const std::string p = vis_path(ctx, what, level);
    if (nct_png_write(p.c_str(), bgr.data(), h, w) != NCT_OK) return nct_fail(ctx, NCT_ERR_IO, "cannot write %s", p.c_str());
    return NCT_OK;
}

template <class T>
int fetch(nct_ctx *ctx, const T *dev, size_t n, std::vector<T> &host)
{
    host.resize(n);
    NCT_CUDA(ctx, cudaMemcpyAsync(host.data(), dev, sizeof(T) * n, cudaMemcpyDeviceToHost, ctx->stream));
    NCT_CUDA(ctx, nct_stream_wait(ctx));
    return NCT_OK;
}

void cluster_colour(int l, uint8_t &b, uint8_t &g, uint8_t &r)
{
    uint32_t h = (uint32_t)(l + 1) * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    b = (uint8_t)(64 + (h & 0xBF)); g = (uint8_t)(64 + ((h >> 8) & 0xBF)); r = (uint8_t)(64 + ((h >> 16) & 0xBF));
}

// getHeat (CT/ColorTransfer.cpp:1127-1178), v already in [0, 1]
void heat(float v, uint8_t &r, uint8_t &g, uint8_t &b)
{
    v = std::min(std::max(v, 0.f), 1.f);
    double dr, dg, db;
    if (v < 0.1242) { db = 0.504 + ((1. - 0.504) / 0.1242) * v; dg = dr = 0.; }
    else if (v < 0.3747) { db = 1.; dr = 0.; dg = (v - 0.1242) * (1. / (0.3747 - 0.1242)); }
    else if (v < 0.6253) { db = (0.6253 - v) * (1. / (0.6253 - 0.3747)); dg = 1.; dr = (v - 0.3747) * (1. / (0.6253 - 0.3747)); }
    else if (v < 0.8758) { db = 0.; dr = 1.; dg = (0.8758 - v) * (1. / (0.8758 - 0.6253)); }
    else { db = 0.; dg = 0.; dr = 1. - (v - 0.8758) * ((1. - 0.504) / (1. - 0.8758)); }
    r = (uint8_t)std::min(255, (int)(255 * dr));
    g = (uint8_t)std::min(255, (int)(255 * dg));
    b = (uint8_t)std::min(255, (int)(255 * db));
}

inline uint8_t clamp255(int v) { return (uint8_t)std::min(std::max(v, 0), 255); }

}  // namespace

extern "C" int nct_set_vis(nct_ctx *ctx, const char *dir, const char *prefix)
{
    if (!ctx) return NCT_ERR_ARG;
    ctx->vis_dir = dir ? dir : "";
    ctx->vis_prefix = prefix ? prefix : "pair";
    return NCT_OK;
}

// <prefix>_cluster_small.png: the k-means labels at conv5_1 resolution (NCT/main.cu:169-173)
int nct_vis_cluster_small(nct_ctx *ctx, const int *labels_dev, int lh, int lw)
{
    std::vector<int> lab;
    int rc = fetch(ctx, labels_dev, (size_t)lh * lw, lab);
    if (rc) return rc;
    std::vector<uint8_t> img((size_t)lh * lw * 3);
    for (size_t i = 0; i < lab.size(); ++i) cluster_colour(lab[i], img[i * 3], img[i * 3 + 1], img[i * 3 + 2]);
    return write_png(ctx, "cluster_small", -1, img, lh, lw);
}

// aFlow / bFlow (reconstruct_flow) and the level's two images
int nct_vis_flows(nct_ctx *ctx, int level, const uint32_t *ann_dev, const uint32_t *bnn_dev, const uint8_t *cnt_dev, const uint8_t *stl_dev,
                  int ah, int aw, int bh, int bw)
{
    std::vector<uint32_t> nnf;
    std::vector<uint8_t> img;
    for (int dir = 0; dir < 2; ++dir) {
        const int h = dir ? bh : ah, w = dir ? bw : aw, oh = dir ? ah : bh, ow = dir ? aw : bw;
        int rc = fetch(ctx, dir ? bnn_dev : ann_dev, (size_t)h * w, nnf);
        if (rc) return rc;
        img.assign((size_t)h * w * 3, 0);
        for (size_t p = 0; p < nnf.size(); ++p) {
            const int xbest = (int)(nnf[p] & 0xFFF), ybest = (int)((nnf[p] >> 12) & 0xFFF);
            img[p * 3] = (uint8_t)(255 * ((float)xbest / ow));
            img[p * 3 + 2] = (uint8_t)(255 * ((float)ybest / oh));
        }
        rc = write_png(ctx, dir ? "bFlow" : "aFlow", level, img, h, w);
        if (rc) return rc;
        rc = fetch(ctx, dir ? stl_dev : cnt_dev, (size_t)h * w * 3, img);
        if (rc) return rc;
        rc = write_png(ctx, dir ? "tStl" : "tCnt", level, img, h, w);
        if (rc) return rc;
    }
    return NCT_OK;
}

// knn_<level>.png: which (dilated) cluster a pixel's neighbours are searched in (getClusters, CT/ColorTransfer.cpp:273-351):
// a label cell belongs to its own cluster and to the cluster of every 4-neighbour cell with a different label; the
// reference paints the clusters in index order, so the highest index wins
int nct_vis_knn_clusters(nct_ctx *ctx, int level, const int *labels_dev, int lh, int lw, int h, int w, int samples)
{
    std::vector<int> lab;
    int rc = fetch(ctx, labels_dev, (size_t)lh * lw, lab);
    if (rc) return rc;
    std::vector<uint8_t> img((size_t)h * w * 3, 0);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int cx = x / samples, cy = y / samples;
            if (cx >= lw || cy >= lh) continue;
            int best = lab[(size_t)cy * lw + cx];
            if (cx + 1 < lw) best = std::max(best, lab[(size_t)cy * lw + cx + 1]);
            if (cx > 0) best = std::max(best, lab[(size_t)cy * lw + cx - 1]);
            if (cy + 1 < lh) best = std::max(best, lab[(size_t)(cy + 1) * lw + cx]);
            if (cy > 0) best = std::max(best, lab[(size_t)(cy - 1) * lw + cx]);
            const size_t p = ((size_t)y * w + x) * 3;
            cluster_colour(best, img[p], img[p + 1], img[p + 2]);
        }
    return write_png(ctx, "knn", level, img, h, w);
}

// errMap_<level>.png: heat map of the normalised matching error (CT/ColorTransfer.cpp:1302-1337)
int nct_vis_error_map(nct_ctx *ctx, int level, const float *err_dev, int h, int w)
{
    std::vector<float> err;
    int rc = fetch(ctx, err_dev, (size_t)h * w, err);
    if (rc) return rc;
    double mn = 1e8, mx = -1e8;
    for (float e : err) { mn = std::min(mn, (double)e); mx = std::max(mx, (double)e); }
    std::vector<uint8_t> img((size_t)h * w * 3);
    for (size_t p = 0; p < err.size(); ++p) {
        const double e = ((double)err[p] - mn) / (mx - mn);
        uint8_t r, g, b;
        heat((float)e, r, g, b);
        img[p * 3] = b; img[p * 3 + 1] = g; img[p * 3 + 2] = r;
    }
    return write_png(ctx, "errMap", level, img, h, w);
}

// aVis*/bVis* of a coefficient pair given at (mh x mw); `samples` > 0: nearest-neighbour look-up at (y / samples, x / samples)
// into the level-size maps (the "init" maps, CT/ColorTransfer.cpp:1267-1300), else full-size maps.  Optionally returns the
// nearest-upsampled maps (for the refine_init image).
int nct_vis_coefficients(nct_ctx *ctx, int level, const char *suffix, const double *a_dev, const double *b_dev, int mh, int mw, int H, int W,
                         int samples, std::vector<double> *a_up, std::vector<double> *b_up)
{
    std::vector<double> a, b;
    int rc = fetch(ctx, a_dev, (size_t)mh * mw * 3, a);
    if (!rc) rc = fetch(ctx, b_dev, (size_t)mh * mw * 3, b);
    if (rc) return rc;
    std::vector<uint8_t> ia((size_t)H * W * 3), ib((size_t)H * W * 3);
    if (a_up) a_up->resize((size_t)H * W * 3);
    if (b_up) b_up->resize((size_t)H * W * 3);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int sy = samples > 0 ? std::min(y / samples, mh - 1) : y, sx = samples > 0 ? std::min(x / samples, mw - 1) : x;
            const size_t s = ((size_t)sy * mw + sx) * 3, p = ((size_t)y * W + x) * 3;
            for (int c = 0; c < 3; ++c) {
                ia[p + c] = clamp255((int)(a[s + c] * 50));
                ib[p + c] = clamp255((int)(b[s + c] * 255 + 127));
                if (a_up) (*a_up)[p + c] = a[s + c];
                if (b_up) (*b_up)[p + c] = b[s + c];
            }
        }
    char name[48];
    snprintf(name, sizeof(name), "aVis%s", suffix);
    rc = write_png(ctx, name, level, ia, H, W);
    if (rc) return rc;
    snprintf(name, sizeof(name), "bVis%s", suffix);
    return write_png(ctx, name, level, ib, H, W);
}

// refine_init / refine_nonlocal: the image the given full-size coefficient maps produce (CT/ColorTransfer.cpp:1471-1477)
int nct_vis_refine(nct_ctx *ctx, int level, const char *what, const uint8_t *cnt_lab_full_dev, const double *a_full_dev, const double *b_full_dev,
                   int H, int W)
{
    uint8_t *img_dev = (uint8_t *)nct_scratch(ctx, "vis_img", (size_t)H * W * 3);
    if (!img_dev) return NCT_ERR_NOMEM;
    int rc = nct_apply_coefficients(ctx, cnt_lab_full_dev, a_full_dev, b_full_dev, H, W, img_dev, nullptr);
    if (rc) return rc;
    std::vector<uint8_t> img;
    rc = fetch(ctx, img_dev, (size_t)H * W * 3, img);
    if (rc) return rc;
    return write_png(ctx, what, level, img, H, W);
}
