/*
 * nct.h -- C ABI of libnct.so, the B200-native (sm_100a) implementation of the
 * Neural-Color-Transfer hot path.
 *
 * Every entry point cites the reference interface it replaces.  Shorthand:
 *   NCT/ = code/windows/neural_color_transfer/source/      (under the reference root)
 *   CT/  = NCT/ColorTransfer/
 *
 * Conventions
 *   - plain C types only; `dev` pointers are CUDA device pointers on the ctx's GPU,
 *     `host` pointers are ordinary (preferably pinned) host memory;
 *   - every function returns 0 on success and a negative nct_status on error;
 *     nct_last_error(ctx) returns a human-readable message for the last failure;
 *   - a ctx is single-threaded; all device work of a ctx is issued on one CUDA
 *     stream (nct_set_stream) and is ASYNCHRONOUS unless stated otherwise;
 *   - feature volumes are pixel-major FP32 ("HWC": f[(y*W + x)*C + c]).  The
 *     reference keeps Caffe's planar CHW; nct_chw_to_hwc converts for callers
 *     that still hold Caffe blobs;
 *   - NNF entries are the reference's packed uint32 (y << 12) | x
 *     (NCT/GeneralizedPatchMatch.cu:24-34).
 *   - there is NO CPU fallback: without a CUDA device nct_create fails.
 */
#ifndef NCT_H
#define NCT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nct_ctx nct_ctx;

typedef enum nct_status {
    NCT_OK = 0,
    NCT_ERR_CUDA = -1,      /* a CUDA runtime call failed */
    NCT_ERR_ARG = -2,       /* invalid argument / unsupported shape */
    NCT_ERR_STATE = -3,     /* call order (e.g. weights not loaded) */
    NCT_ERR_IO = -4,        /* file could not be read / parsed / written */
    NCT_ERR_NOMEM = -5
} nct_status;

/* ---------------------------------------------------------------- context */

/* Replaces cudaSetDevice/cudaDeviceReset + the two Classifier constructions of
 * NCT/main.cu:562-582 (weights are loaded separately, see nct_vgg19_*). */
int nct_create(int gpu_id, nct_ctx **out);
int nct_destroy(nct_ctx *ctx);
const char *nct_last_error(const nct_ctx *ctx);
/* cudaStream_t passed as void*; NULL = the ctx's own non-blocking stream. */
int nct_set_stream(nct_ctx *ctx, void *cuda_stream);
void *nct_get_stream(const nct_ctx *ctx);
int nct_synchronize(nct_ctx *ctx);
/* library/ABI version: major*10000 + minor*100 + patch */
int nct_version(void);
/* number of kernels this library launched on `ctx` since creation / last reset */
long long nct_launch_count(const nct_ctx *ctx);
void nct_reset_launch_count(nct_ctx *ctx);

/* ---------------------------------------------------------------- layout / normalise */

/* Caffe blob (planar C x H x W) -> pixel-major H x W x C.  No reference counterpart
 * (the reference consumes CHW in place); needed only by callers holding Caffe blobs. */
int nct_chw_to_hwc(nct_ctx *ctx, const float *src_chw_dev, float *dst_hwc_dev, int C, int H, int W);
int nct_hwc_to_chw(nct_ctx *ctx, const float *src_hwc_dev, float *dst_chw_dev, int C, int H, int W);

/* Replaces norm(dst, src, smooth=NULL, dim), NCT/GeneralizedPatchMatch.cu:237-283:
 * dst[p][c] = src[p][c] / sqrt(sum_c src[p][c]^2); zero-norm pixels give zeros.
 * One fused kernel instead of 5 launches + 4 cudaMalloc/Free.  In-place allowed. */
int nct_l2norm(nct_ctx *ctx, const float *src_hwc_dev, float *dst_hwc_dev, int C, int H, int W);

/* ---------------------------------------------------------------- NNF */

/* Replaces init_Ann_kernel<<<>>>(ann, params), NCT/GeneralizedPatchMatch.cu:527-544. */
int nct_nnf_init(nct_ctx *ctx, uint32_t *ann_dev, int ah, int aw, int bh, int bw);

/* Replaces upSample_kernel<<<>>>(ann, ann_tmp, params, aw_half, ah_half) + the D2D copy,
 * NCT/GeneralizedPatchMatch.cu:546-580 and NCT/main.cu:238-250.  ann_half_dev and
 * ann_dev must not alias. */
int nct_nnf_upsample(nct_ctx *ctx, const uint32_t *ann_half_dev, int ah_half, int aw_half,
                     uint32_t *ann_dev, int ah, int aw, int bh, int bw);

/* Replaces patchmatch_single<<<>>>(a1, b1, NULL, ann, annd, params),
 * NCT/GeneralizedPatchMatch.cu:677-831.  `params` is the reference's 11-int HOST array
 * {C, ah, aw, bh, bw, patch_w(=3), iters, rs_max, flag_constraint(=0), constraint, energy}
 * (NCT/main.cu:204-214).  a/b are L2-normalised HWC volumes.  ann in/out, annd out.
 * Deterministic jump-flood schedule: bit-exact vs oracle/pm_oracle.c. */
int nct_patchmatch(nct_ctx *ctx, const float *a_hwc_dev, const float *b_hwc_dev,
                   uint32_t *ann_dev, float *annd_dev, const int params[11]);

/* Both directions of NCT/main.cu:283-284 in the same launches (A->B into ann/annd,
 * B->A into bnn/bnnd); params_ab as above, the B->A params are derived by swapping. */
int nct_patchmatch_bidir(nct_ctx *ctx, const float *a_hwc_dev, const float *b_hwc_dev,
                         uint32_t *ann_dev, float *annd_dev, uint32_t *bnn_dev, float *bnnd_dev,
                         const int params_ab[11]);

/* The per-column XORWOW uniforms the reference draws (curand_init(seed = column, 0, 0),
 * NCT/GeneralizedPatchMatch.cu:54-66): out[col*ndraws + k].  Exposed for tests. */
int nct_xorwow_table(nct_ctx *ctx, float *out_dev, int ncols, int ndraws);

/* Candidate evaluations performed by the last nct_patchmatch* call on this ctx
 * (summed over directions; synchronises the stream).  stats[0] = evaluated after
 * de-duplication, stats[1] = reference-semantics count. */
int nct_patchmatch_stats(nct_ctx *ctx, long long stats[2]);
/* enable (1) / disable (0) evaluation counting inside the PatchMatch kernels */
int nct_patchmatch_count_evals(nct_ctx *ctx, int enable);

/* ---------------------------------------------------------------- BDS votes */

/* Replaces the host function reconstruct_bds(a, b, ann, bnn, patch_w=3, wCohen, wComplete),
 * NCT/GeneralizedPatchMatch.cu:122-235 (called at NCT/main.cu:291 after four D2H copies):
 * BDS-voted B colours in A's layout.  8-bit BGR images (H x W x 3, device), out has A's size.
 * Bit-exact vs the oracle (integer sums, IEEE double divide, truncating store). */
int nct_reconstruct_bds(nct_ctx *ctx, const uint8_t *a_bgr_dev, const uint8_t *b_bgr_dev,
                        const uint32_t *ann_dev, const uint32_t *bnn_dev, int ah, int aw, int bh, int bw,
                        double w_cohen, double w_complete, uint8_t *out_bgr_dev);

/* Replaces avg_vote_bds_a + avg_vote_bds_b + avg_vote_bds (NCT/GeneralizedPatchMatch.cu:1074-1202),
 * norm() of the voted volume and feature_distance (:833-855), i.e. NCT/main.cu:297-318:
 * err[p] = -< c_norm[p], normalise(BDS vote of s_raw)[p] >.  c_norm: L2-normalised A features,
 * s_raw: UN-normalised B features, both HWC.  vote_out_dev may be NULL; if given it receives the
 * voted (weight-divided, un-normalised) features in A's layout (HWC). */
int nct_bds_feature_error(nct_ctx *ctx, const float *c_norm_hwc_dev, const float *s_raw_hwc_dev,
                          const uint32_t *ann_dev, const uint32_t *bnn_dev, int C, int ah, int aw, int bh, int bw,
                          float w_cohen, float w_complete, float *err_dev, float *vote_out_dev);

#ifdef __cplusplus
}
#endif
#endif /* NCT_H */
