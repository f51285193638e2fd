// Per-pair orchestrator: the L = 5 -> 1 progressive colour transfer.
//
// Replaces transfer_color_single_bds (NCT/main.cu:47-454): same level loop, same constants (layers conv5_1..conv1_1,
// 10 PatchMatch iterations, search ranges maxLen/16, /32, /64, 32, 32, patch 3), same order of operations -- but every
// step runs on the device on one stream: no per-level cudaMalloc/cudaFree (workspace arena), no D2H/H2D round trips
// (the reference makes >= 6 per level plus three CSR uploads), truncated VGG re-forwards.  The only host
// synchronisations are the k-means set-up (once per pair) and the data-dependent WLS iteration count.
#include "device_utils.cuh"
#include <algorithm>
#include <cstring>


extern "C" {

void nct_config_default(nct_config *cfg)
{
    if (!cfg) return;
    // CT/Config.h:58-72 (the constructor defaults win over the -h help strings of NCT/main.cu:40-43)
    cfg->bds_weight = 2.0;
    cfg->var_eps = 0.60;
    cfg->nonlocal_weight = 2.0;
    cfg->local_weight = 0.125;
    cfg->wls_lambda_init = 0.024;
    cfg->cluster_num = 10;
    cfg->k_num = 8;
    cfg->patch_size = 3;
    cfg->wls_alpha = 1.2;
    cfg->pm_iters = 10;            // NCT/main.cu:65
    cfg->kmeans_iters = 11;        // CT/ColorTransfer.cpp:373
    cfg->wls_rel_tol = 1e-8;       // relative residual; the maps then agree with a direct solve to ~1e-10
    cfg->stop_after_level = 4;
    cfg->feature_store = 0;        // FP32 PatchMatch volumes, the reference's storage
}

int nct_transfer_pair_dev(nct_ctx *ctx, const uint8_t *cnt_bgr_dev, int ch, int cw, const uint8_t *stl_bgr_dev, int sh, int sw,
                          const nct_config *cfg_in, uint8_t *out_bgr_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, cnt_bgr_dev && stl_bgr_dev && out_bgr_dev, "null device pointer");
    NCT_REQUIRE(ctx, ch >= 32 && cw >= 32 && sh >= 32 && sw >= 32, "images must be at least 32 x 32");
    NCT_REQUIRE(ctx, ch <= 4095 && cw <= 4095 && sh <= 4095 && sw <= 4095, "image sides above 4095 do not fit the 12-bit NNF packing");
    nct_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else nct_config_default(&cfg);
    NCT_REQUIRE(ctx, cfg.patch_size == 3 && cfg.k_num == 8, "patch_size must be 3 and k_num 8 (CT/Config.h:68-70)");
    NCT_REQUIRE(ctx, cfg.feature_store == 0 || cfg.feature_store == 1, "feature_store must be 0 (FP32) or 1 (FP16)");
    const int L = 5;
    int dc[5][3], ds[5][3];
    nct_vgg19_level_dims(ch, cw, dc);
    nct_vgg19_level_dims(sh, sw, ds);
    const int maxLen = std::max(std::max(cw, ch), std::max(sw, sh));
    const int range[5] = {maxLen / 16, maxLen / 32, maxLen / 64, 32, 32};  // NCT/main.cu:77-83
    const size_t nC = (size_t)ch * cw, nS = (size_t)sh * sw;

    // ---- workspace
    float *featC[5], *featS[5];
    uint8_t *cntImg[5], *stlImg[5];
    char name[64];
    for (int l = 0; l < L; ++l) {
        snprintf(name, sizeof(name), "pipe_featC%d", l);
        featC[l] = (float *)nct_scratch(ctx, name, sizeof(float) * (size_t)dc[l][0] * dc[l][1] * dc[l][2]);
        snprintf(name, sizeof(name), "pipe_featS%d", l);
        featS[l] = (float *)nct_scratch(ctx, name, sizeof(float) * (size_t)ds[l][0] * ds[l][1] * ds[l][2]);
        snprintf(name, sizeof(name), "pipe_cntImg%d", l);
        cntImg[l] = (uint8_t *)nct_scratch(ctx, name, (size_t)dc[l][1] * dc[l][2] * 3);
        snprintf(name, sizeof(name), "pipe_stlImg%d", l);
        stlImg[l] = (uint8_t *)nct_scratch(ctx, name, (size_t)ds[l][1] * ds[l][2] * 3);
        if (!featC[l] || !featS[l] || !cntImg[l] || !stlImg[l]) return NCT_ERR_NOMEM;
    }
    const size_t maxFeat = std::max((size_t)dc[4][0] * nC, (size_t)ds[4][0] * nS);  // 64 channels at full size dominate...
    size_t maxFeatAll = maxFeat;
    for (int l = 0; l < L; ++l) {
        maxFeatAll = std::max(maxFeatAll, (size_t)dc[l][0] * dc[l][1] * dc[l][2]);
        maxFeatAll = std::max(maxFeatAll, (size_t)ds[l][0] * ds[l][1] * ds[l][2]);
    }
    float *normC = (float *)nct_scratch(ctx, "pipe_normC", sizeof(float) * maxFeatAll);
    float *normS = (float *)nct_scratch(ctx, "pipe_normS", sizeof(float) * maxFeatAll);
    uint16_t *normC16 = nullptr, *normS16 = nullptr;   // FP16 feature store: the volumes PatchMatch gathers from
    if (cfg.feature_store == 1) {
        normC16 = (uint16_t *)nct_scratch(ctx, "pipe_normC16", sizeof(uint16_t) * maxFeatAll);
        normS16 = (uint16_t *)nct_scratch(ctx, "pipe_normS16", sizeof(uint16_t) * maxFeatAll);
        if (!normC16 || !normS16) return NCT_ERR_NOMEM;
    }
    uint32_t *ann = (uint32_t *)nct_scratch(ctx, "pipe_ann", sizeof(uint32_t) * nC);
    uint32_t *bnn = (uint32_t *)nct_scratch(ctx, "pipe_bnn", sizeof(uint32_t) * nS);
    uint32_t *ann_prev = (uint32_t *)nct_scratch(ctx, "pipe_ann_prev", sizeof(uint32_t) * nC);
    uint32_t *bnn_prev = (uint32_t *)nct_scratch(ctx, "pipe_bnn_prev", sizeof(uint32_t) * nS);
    float *annd = (float *)nct_scratch(ctx, "pipe_annd", sizeof(float) * nC);
    float *bnnd = (float *)nct_scratch(ctx, "pipe_bnnd", sizeof(float) * nS);
    float *err = (float *)nct_scratch(ctx, "pipe_err", sizeof(float) * nC);
    uint8_t *cntLabFull = (uint8_t *)nct_scratch(ctx, "pipe_cntLabFull", nC * 3);
    uint8_t *smlRes = (uint8_t *)nct_scratch(ctx, "pipe_smlRes", nC * 3);
    uint8_t *cntLab = (uint8_t *)nct_scratch(ctx, "pipe_cntLab", nC * 3);
    uint8_t *stlLab = (uint8_t *)nct_scratch(ctx, "pipe_stlLab", nC * 3);
    uint8_t *refine = (uint8_t *)nct_scratch(ctx, "pipe_refine", nC * 3);
    int *labels = (int *)nct_scratch(ctx, "pipe_labels", sizeof(int) * (size_t)dc[0][1] * dc[0][2]);
    int *knn_id = (int *)nct_scratch(ctx, "pipe_knn_id", sizeof(int) * nC * 8);
    double *knn_w = (double *)nct_scratch(ctx, "pipe_knn_w", sizeof(double) * nC * 8);
    double *a_lvl = (double *)nct_scratch(ctx, "pipe_a_lvl", sizeof(double) * nC * 3);
    double *b_lvl = (double *)nct_scratch(ctx, "pipe_b_lvl", sizeof(double) * nC * 3);
    double *a_full = (double *)nct_scratch(ctx, "pipe_a_full", sizeof(double) * nC * 3);
    double *b_full = (double *)nct_scratch(ctx, "pipe_b_full", sizeof(double) * nC * 3);
    double *rough = (double *)nct_scratch(ctx, "pipe_rough", sizeof(double) * nC);
    double *weight = (double *)nct_scratch(ctx, "pipe_weight", sizeof(double) * nC);
    if (!normC || !normS || !ann || !bnn || !ann_prev || !bnn_prev || !annd || !bnnd || !err || !cntLabFull || !smlRes || !cntLab ||
        !stlLab || !refine || !labels || !knn_id || !knn_w || !a_lvl || !b_lvl || !a_full || !b_full || !rough || !weight)
        return NCT_ERR_NOMEM;

    int rc;
#define STEP(call)            \
    do {                      \
        rc = (call);          \
        if (rc) return rc;    \
    } while (0)

    // ---- ColorTransfer ctor: m_cntLab (CT/ColorTransfer.h:54-60)
    STEP(nct_bgr2lab_u8(ctx, cnt_bgr_dev, cntLabFull, (int)nC));
    // ---- features of both images (NCT/main.cu:94,102)
    {
        NctStageTimer t(ctx, ST_VGG);
        STEP(nct_vgg19_features(ctx, cnt_bgr_dev, ch, cw, 0, featC));
        STEP(nct_vgg19_features(ctx, stl_bgr_dev, sh, sw, 0, featS));
    }
    // ---- image pyramids from the ORIGINAL images (NCT/main.cu:104-108)
    NCT_CUDA(ctx, cudaMemcpyAsync(cntImg[4], cnt_bgr_dev, nC * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(stlImg[4], stl_bgr_dev, nS * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int l = L - 2; l >= 0; --l) {
        STEP(nct_resize_linear_u8c3(ctx, cntImg[l + 1], dc[l + 1][1], dc[l + 1][2], cntImg[l], dc[l][1], dc[l][2]));
        STEP(nct_resize_linear_u8c3(ctx, stlImg[l + 1], ds[l + 1][1], ds[l + 1][2], stlImg[l], ds[l][1], ds[l][2]));
    }
    // ---- cluster the normalised conv5_1 features of the content image (NCT/main.cu:139-168)
    STEP(nct_l2norm(ctx, featC[0], normC, dc[0][0], dc[0][1], dc[0][2]));
    { NctStageTimer t(ctx, ST_KMEANS);
    STEP(nct_cluster_features(ctx, normC, dc[0][1], dc[0][2], dc[0][0], cfg.cluster_num, cfg.kmeans_iters, labels)); }
    const bool vis = !ctx->vis_dir.empty();   // ENABLE_VIS artefacts (vis.cpp)
    if (vis) STEP(nct_vis_cluster_small(ctx, labels, dc[0][1], dc[0][2]));

    const uint8_t *result = cnt_bgr_dev;
    for (int l = 0; l < L; ++l) {
        const int C = dc[l][0], ah = dc[l][1], aw = dc[l][2], bh = ds[l][1], bw = ds[l][2];
        // NNF initialisation / upsampling (NCT/main.cu:230-251)
        if (l == 0) {
            STEP(nct_nnf_init(ctx, ann, ah, aw, bh, bw));
            STEP(nct_nnf_init(ctx, bnn, bh, bw, ah, aw));
        } else {
            std::swap(ann, ann_prev);
            std::swap(bnn, bnn_prev);
            STEP(nct_nnf_upsample(ctx, ann_prev, dc[l - 1][1], dc[l - 1][2], ann, ah, aw, bh, bw));
            STEP(nct_nnf_upsample(ctx, bnn_prev, ds[l - 1][1], ds[l - 1][2], bnn, bh, bw, ah, aw));
        }
        // normalise both feature volumes (:254-275)
        STEP(nct_l2norm(ctx, featS[l], normS, C, bh, bw));
        STEP(nct_l2norm(ctx, featC[l], normC, C, ah, aw));
        // bidirectional PatchMatch (:283-284)
        const int params[11] = {C, ah, aw, bh, bw, cfg.patch_size, cfg.pm_iters, range[l], 0, 10, 1};
        if (cfg.feature_store == 1) {
            // FP16 feature store: PatchMatch reads half-precision copies (half the bytes); the BDS feature error below still
            // uses the FP32 content volume and the un-normalised style features
            STEP(nct_l2norm_f16(ctx, featS[l], normS16, C, bh, bw));
            STEP(nct_l2norm_f16(ctx, featC[l], normC16, C, ah, aw));
            NctStageTimer t(ctx, ST_PM);
            STEP(nct_patchmatch_bidir_f16(ctx, normC16, normS16, ann, annd, bnn, bnnd, params));
        } else {
            NctStageTimer t(ctx, ST_PM);
            STEP(nct_patchmatch_bidir(ctx, normC, normS, ann, annd, bnn, bnnd, params));
        }
        if (vis) STEP(nct_vis_flows(ctx, l, ann, bnn, cntImg[l], stlImg[l], ah, aw, bh, bw));
        // BDS colour reconstruction at level size (:291) and BDS feature error (:297-318)
        { NctStageTimer t(ctx, ST_BDS);
        STEP(nct_reconstruct_bds(ctx, cntImg[l], stlImg[l], ann, bnn, ah, aw, bh, bw, 1.0, (double)(float)cfg.bds_weight, smlRes));
        STEP(nct_bds_feature_error(ctx, normC, featS[l], ann, bnn, C, ah, aw, bh, bw, 1.0f, (float)cfg.bds_weight, err, nullptr)); }
        // Lab images of the level (:351-375)
        STEP(nct_bgr2lab_u8(ctx, cntImg[l], cntLab, ah * aw));
        STEP(nct_bgr2lab_u8(ctx, smlRes, stlLab, ah * aw));
        // non-local neighbours (:359); label cells are 2^l pixels wide
        { NctStageTimer t(ctx, ST_KNN); STEP(nct_find_knns(ctx, labels, dc[0][2], dc[0][1], cfg.cluster_num, cntLab, ah, aw, 1 << l, knn_id, knn_w)); }
        if (vis) STEP(nct_vis_knn_clusters(ctx, l, labels, dc[0][1], dc[0][2], ah, aw, 1 << l));
        // transfer_color_downsample (CT/ColorTransfer.cpp:1180-1478)
        STEP(nct_local_fit(ctx, cntLab, stlLab, ah, aw, cfg.var_eps, a_lvl, b_lvl));
        STEP(nct_confidence_weights(ctx, err, ah * aw, weight));
        if (vis) {
            // the local fit, looked up at (y / samples, x / samples) with samples = 2^(4 - level) (NCT/main.cu:365)
            std::vector<double> au, bu;
            STEP(nct_vis_error_map(ctx, l, err, ah, aw));
            STEP(nct_vis_coefficients(ctx, l, "_init", a_lvl, b_lvl, ah, aw, ch, cw, 1 << (L - 1 - l), &au, &bu));
            NCT_CUDA(ctx, cudaMemcpyAsync(a_full, au.data(), sizeof(double) * nC * 3, cudaMemcpyHostToDevice, ctx->stream));
            NCT_CUDA(ctx, cudaMemcpyAsync(b_full, bu.data(), sizeof(double) * nC * 3, cudaMemcpyHostToDevice, ctx->stream));
            STEP(nct_vis_refine(ctx, l, "refine_init", cntLabFull, a_full, b_full, ch, cw));   // (waits for the copies: au / bu stay valid)
        }
        const double normFactor = (double)(cw * ch) / (double)(aw * ah);
        double lam = cfg.wls_lambda_init * normFactor;
        { NctStageTimer t(ctx, ST_CG);
        STEP(nct_solve_nonlocal(ctx, a_lvl, b_lvl, weight, cntLab, stlLab, knn_id, knn_w, ah, aw, l, cfg.local_weight, cfg.wls_alpha,
                                cfg.nonlocal_weight, cfg.k_num, normFactor, nullptr)); }
        STEP(nct_upsample_coefficients(ctx, a_lvl, b_lvl, ah, aw, cntLabFull, ch, cw, a_full, b_full, rough));
        if (ah == ch && aw == cw) lam = lam * 4;
        if (vis) {
            STEP(nct_vis_coefficients(ctx, l, "_nonlocal", a_full, b_full, ch, cw, ch, cw, 0, nullptr, nullptr));
            STEP(nct_vis_refine(ctx, l, "refine_nonlocal", cntLabFull, a_full, b_full, ch, cw));
        }
        { NctStageTimer t(ctx, ST_WLS);
        static const bool warm = !(getenv("NCT_WLS_WARM") && atoi(getenv("NCT_WLS_WARM")) == 0);  // default on
        ctx->wls_warm = warm ? 1 : 0;
        if (l == 0) ctx->wls_prev_n = 0;  // never carry a solution from one pair into the next
        rc = nct_solve_wls(ctx, a_full, b_full, rough, cntLabFull, ch, cw, lam, cfg.wls_alpha, cfg.wls_rel_tol, 0, nullptr, nullptr);
        ctx->wls_warm = 0;
        if (rc) return rc; }
        if (vis) STEP(nct_vis_coefficients(ctx, l, "", a_full, b_full, ch, cw, ch, cw, 0, nullptr, nullptr));
        STEP(nct_apply_coefficients(ctx, cntLabFull, a_full, b_full, ch, cw, refine, nullptr));
        result = refine;
        if (l >= cfg.stop_after_level) break;
        // re-extract the content features from the intermediate result (:424-427), only as deep as still needed
        if (l < L - 1) { NctStageTimer t(ctx, ST_VGG); STEP(nct_vgg19_features(ctx, refine, ch, cw, l + 1, featC)); }
    }
#undef STEP
    NCT_CUDA(ctx, cudaMemcpyAsync(out_bgr_dev, result, nC * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    return NCT_OK;
}

int nct_transfer_pair(nct_ctx *ctx, const uint8_t *cnt_bgr_host, int ch, int cw, const uint8_t *stl_bgr_host, int sh, int sw,
                      const nct_config *cfg, uint8_t *out_bgr_host)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, cnt_bgr_host && stl_bgr_host && out_bgr_host && ch > 0 && cw > 0 && sh > 0 && sw > 0, "bad arguments");
    const size_t nC = (size_t)ch * cw * 3, nS = (size_t)sh * sw * 3;
    uint8_t *dC = (uint8_t *)nct_scratch(ctx, "pipe_in_cnt", nC);
    uint8_t *dS = (uint8_t *)nct_scratch(ctx, "pipe_in_stl", nS);
    uint8_t *dO = (uint8_t *)nct_scratch(ctx, "pipe_out", nC);
    if (!dC || !dS || !dO) return NCT_ERR_NOMEM;
    NCT_CUDA(ctx, cudaMemcpyAsync(dC, cnt_bgr_host, nC, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(dS, stl_bgr_host, nS, cudaMemcpyHostToDevice, ctx->stream));
    int rc = nct_transfer_pair_dev(ctx, dC, ch, cw, dS, sh, sw, cfg, dO);
    if (rc) return rc;
    NCT_CUDA(ctx, cudaMemcpyAsync(out_bgr_host, dO, nC, cudaMemcpyDeviceToHost, ctx->stream));
    NCT_CUDA(ctx, nct_stream_wait(ctx));
    return NCT_OK;
}

}  // extern "C"
