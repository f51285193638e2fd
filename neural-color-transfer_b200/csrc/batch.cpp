// pairs.txt batch driver: the reference's transfer_single (NCT/main.cu:456-543) behind the C ABI.
//
// Same surface: <input_dir>/pairs.txt with lines "content style bds\n", images relative to input_dir, longer side
// clamped to MAX_SIZE = 1000 (CT/Config.h:5, NCT/main.cu:500-522), output "<out>/<cnt>_<stl>_<bds %2.2f>.png",
// same stdout lines.  Added: (rank, world) sharding -- rank r processes the lines i with i % world == r -- so one
// process per GPU covers a pair list with no communication (pairs are independent); nct_run_pairs_ex adds
// resume-by-existing-output, the ENABLE_VIS artefacts (8f-3) and counts of failed / skipped pairs.
// Images are PNG only (no libjpeg / libtiff in this image; the reference's imread also takes JPEG, BMP, ...).
#include "nct_internal.h"
#include <sys/stat.h>
#include <cstdio>
#include <cstring>
#include <string>

extern "C" {
int nct_png_read(const char *path, uint8_t **bgr_out, int *h_out, int *w_out);
void nct_png_free(uint8_t *p);
int nct_png_write(const char *path, const uint8_t *bgr, int h, int w);
int nct_set_vis(nct_ctx *ctx, const char *dir, const char *prefix);
}

namespace {

const int MAX_SIZE = 1000;

std::string base_name(const std::string &path)
{
    const size_t pos = path.find_last_of("\\/") + 1;  // npos + 1 == 0
    const size_t dot = path.find_last_of('.');
    return path.substr(pos, dot == std::string::npos || dot < pos ? std::string::npos : dot - pos);
}

// resize(cnt, cnt, Size(cw, ch)) of NCT/main.cu:500-522 (INTER_LINEAR default), on the device
int clamp_size(nct_ctx *ctx, uint8_t *&img, int &h, int &w)
{
    if (w <= MAX_SIZE && h <= MAX_SIZE) return NCT_OK;
    int cw = MAX_SIZE;
    int chh = (int)(cw / (float)w * h);
    if (w < h) {
        chh = MAX_SIZE;
        cw = (int)(chh / (float)h * w);
    }
    uint8_t *d_src = (uint8_t *)nct_scratch(ctx, "batch_resize_src", (size_t)h * w * 3);
    uint8_t *d_dst = (uint8_t *)nct_scratch(ctx, "batch_resize_dst", (size_t)chh * cw * 3);
    if (!d_src || !d_dst) return NCT_ERR_NOMEM;
    NCT_CUDA(ctx, cudaMemcpyAsync(d_src, img, (size_t)h * w * 3, cudaMemcpyHostToDevice, ctx->stream));
    int rc = nct_resize_linear_u8c3(ctx, d_src, h, w, d_dst, chh, cw);
    if (rc) return rc;
    uint8_t *out = (uint8_t *)malloc((size_t)chh * cw * 3);
    if (!out) return NCT_ERR_NOMEM;
    NCT_CUDA(ctx, cudaMemcpyAsync(out, d_dst, (size_t)chh * cw * 3, cudaMemcpyDeviceToHost, ctx->stream));
    NCT_CUDA(ctx, nct_stream_wait(ctx));
    nct_png_free(img);
    img = out;
    h = chh;
    w = cw;
    return NCT_OK;
}

}  // namespace

extern "C" {

int nct_run_pairs_ex(nct_ctx *ctx, const char *input_dir, const char *output_dir, const nct_config *cfg_in, int rank, int world,
                     int flags, int *pairs_done, int *pairs_failed, int *pairs_skipped);

int nct_run_pairs(nct_ctx *ctx, const char *input_dir, const char *output_dir, const nct_config *cfg_in, int rank, int world,
                  int *pairs_done)
{
    return nct_run_pairs_ex(ctx, input_dir, output_dir, cfg_in, rank, world, 0, pairs_done, nullptr, nullptr);
}

int nct_run_pairs_ex(nct_ctx *ctx, const char *input_dir, const char *output_dir, const nct_config *cfg_in, int rank, int world,
                     int flags, int *pairs_done, int *pairs_failed, int *pairs_skipped)
{
    if (!ctx || !input_dir || !output_dir) return NCT_ERR_ARG;
    cudaSetDevice(ctx->device);  // the calling thread may be a fresh worker thread (CLI -ngpu / -inflight)
    NCT_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world, "bad rank/world %d/%d", rank, world);
    nct_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else nct_config_default(&cfg);
    mkdir(output_dir, 0775);
    const std::string in_dir(input_dir), out_dir(output_dir);
    const std::string pairs_file = in_dir + "/pairs.txt";
    FILE *fp = fopen(pairs_file.c_str(), "r");
    if (fp == NULL) {
        printf("Error: File %s does not exist in the input directory.\n", pairs_file.c_str());
        return nct_fail(ctx, NCT_ERR_IO, "cannot open %s", pairs_file.c_str());  // the reference crashes here
    }
    char cntFile[260], stlFile[260];
    float bdsWeight = 0.f;
    int line = 0, done = 0, failed = 0, skipped = 0;
    while (fscanf(fp, "%259s %259s %f\n", cntFile, stlFile, &bdsWeight) == 3) {
        const int idx = line++;
        if (idx % world != rank) continue;
        cfg.bds_weight = bdsWeight;
        printf("-----------------***********************----------------------\n");
        printf("Content: %s, style: %s, BDS weight: %f.\n", cntFile, stlFile, cfg.bds_weight);
        const std::string cntStr = in_dir + "/" + cntFile, stlStr = in_dir + "/" + stlFile;
        char stem[512], fileName[1024];
        snprintf(stem, sizeof(stem), "%s_%s_%2.2f", base_name(cntStr).c_str(), base_name(stlStr).c_str(), cfg.bds_weight);
        snprintf(fileName, sizeof(fileName), "%s/%s.png", out_dir.c_str(), stem);
        if (flags & NCT_RUN_RESUME) {   // resume an interrupted pair list: a result that is already there is kept
            struct stat st;
            if (stat(fileName, &st) == 0 && st.st_size > 0) {
                printf("Output %s exists, skipped (resume).\n\n", fileName);
                skipped++;
                continue;
            }
        }
        uint8_t *cnt = nullptr, *stl = nullptr;
        int ch = 0, cw = 0, sh = 0, sw = 0;
        if (nct_png_read(cntStr.c_str(), &cnt, &ch, &cw) != NCT_OK) {
            printf("Error: Fail reading content image: %s\n", cntStr.c_str());
            failed++;
            continue;
        }
        printf("\n**Read content file: %s, w = %d, h = %d\n", cntStr.c_str(), cw, ch);
        if (nct_png_read(stlStr.c_str(), &stl, &sh, &sw) != NCT_OK) {
            printf("Error: Fail reading style image: %s\n", stlStr.c_str());
            nct_png_free(cnt);
            failed++;
            continue;
        }
        printf("Read style file: %s, w = %d, h = %d\n", stlStr.c_str(), sw, sh);
        int rc = clamp_size(ctx, cnt, ch, cw);
        if (!rc) rc = clamp_size(ctx, stl, sh, sw);
        uint8_t *out = (uint8_t *)malloc((size_t)ch * cw * 3);
        if (!rc && !out) rc = NCT_ERR_NOMEM;
        if (flags & NCT_RUN_VIS) nct_set_vis(ctx, out_dir.c_str(), stem);   // ENABLE_VIS artefacts next to the result
        if (!rc) rc = nct_transfer_pair(ctx, cnt, ch, cw, stl, sh, sw, &cfg, out);
        if (flags & NCT_RUN_VIS) nct_set_vis(ctx, nullptr, nullptr);
        if (!rc) {
            if (nct_png_write(fileName, out, ch, cw) != NCT_OK) rc = nct_fail(ctx, NCT_ERR_IO, "cannot write %s", fileName);
            else printf("Final output file: %s.\n\n", fileName);
        }
        nct_png_free(cnt);
        nct_png_free(stl);
        free(out);
        if (rc) {
            printf("Error: pair %d failed: %s\n", idx, nct_last_error(ctx));
            failed++;
            continue;  // the reference has no error path; keep going with the next pair
        }
        done++;
    }
    fclose(fp);
    if (pairs_done) *pairs_done = done;
    if (pairs_failed) *pairs_failed = failed;
    if (pairs_skipped) *pairs_skipped = skipped;
    return NCT_OK;
}

}  // extern "C"
