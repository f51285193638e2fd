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

/* Stage timing of nct_transfer_pair* (CUDA events recorded on the ctx stream around each stage; the spans are
 * resolved when nct_profile_get is called, which synchronises).  Stages: 0 vgg, 1 patchmatch, 2 bds, 3 knn,
 * 4 nonlocal_cg, 5 wls, 6 misc, 7 kmeans.  ms_out = accumulated device milliseconds, spans_out = timed spans.
 * enable: 0 = off, 1 = every stage, 2 = the PatchMatch stage only.  A measuring aid: a profiled context runs slower than
 * its neighbours when several contexts share a GPU (bench.py keeps it out of the headline region). */
int nct_profile_enable(nct_ctx *ctx, int enable);
int nct_profile_reset(nct_ctx *ctx);
int nct_profile_get(nct_ctx *ctx, int stage, double *ms_out, long long *spans_out);
const char *nct_profile_stage_name(int stage);

/* Test hook: copy `bytes` of the named internal scratch buffer to host (synchronises).  Lets parity tests feed
 * the kernel's own intermediate arrays (e.g. "nl_d2", "nl_wx2", "nl_wy2", "nl_kw2") to the oracle. */
int nct_debug_read_scratch(nct_ctx *ctx, const char *name, void *host_dst, size_t bytes);

/* ---------------------------------------------------------------- layout / normalise */

/* Caffe blob (planar C x H x W) -> pixel-major H x W x C.  No reference counterpart
 * (the reference consumes CHW in place); needed only by callers holding Caffe blobs. */
int nct_chw_to_hwc(nct_ctx *ctx, const float *src_chw_dev, float *dst_hwc_dev, int C, int H, int W);
int nct_hwc_to_chw(nct_ctx *ctx, const float *src_hwc_dev, float *dst_chw_dev, int C, int H, int W);

/* Replaces norm(dst, src, smooth=NULL, dim), NCT/GeneralizedPatchMatch.cu:237-283:
 * dst[p][c] = src[p][c] / sqrt(sum_c src[p][c]^2); zero-norm pixels give zeros.
 * One fused kernel instead of 5 launches + 4 cudaMalloc/Free.  In-place allowed. */
int nct_l2norm(nct_ctx *ctx, const float *src_hwc_dev, float *dst_hwc_dev, int C, int H, int W);
/* The same normalisation, each value then rounded to FP16 (round to nearest even): dst = C*H*W halves. */
int nct_l2norm_f16(nct_ctx *ctx, const float *src_dev, uint16_t *dst_f16_dev, int C, int H, int W);

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
 * (NCT/main.cu:204-214).  a/b are L2-normalised HWC volumes.  ann in/out, annd out.  iters in [0, 31] (the
 * reference uses 10), C in {16, 32, 64, 128, 256, 512}, sides <= 4096 (12-bit packing).
 * Deterministic jump-flood schedule (decisions D1-D4): bit-exact vs oracle/pm_oracle.c. */
int nct_patchmatch(nct_ctx *ctx, const float *a_hwc_dev, const float *b_hwc_dev,
                   uint32_t *ann_dev, float *annd_dev, const int params[11]);

/* Both directions of NCT/main.cu:283-284 in the same launches (A->B into ann/annd,
 * B->A into bnn/bnnd); params_ab as above, the B->A params are derived by swapping. */
int nct_patchmatch_bidir(nct_ctx *ctx, const float *a_hwc_dev, const float *b_hwc_dev,
                         uint32_t *ann_dev, float *annd_dev, uint32_t *bnn_dev, float *bnnd_dev,
                         const int params_ab[11]);
/* FP16 feature store (throughput mode, no counterpart in the reference): the same kernels gathering from volumes stored as
 * IEEE half (uint16 storage, produced by nct_l2norm_f16).  Every half is converted to FP32 exactly and the arithmetic is
 * unchanged, so the result is bit-identical to nct_patchmatch* on the FP16-rounded volumes.  C >= 64, iters >= 1. */
int nct_patchmatch_f16(nct_ctx *ctx, const uint16_t *a_f16_dev, const uint16_t *b_f16_dev, uint32_t *ann_dev, float *annd_dev,
                       const int params[11]);
int nct_patchmatch_bidir_f16(nct_ctx *ctx, const uint16_t *a_f16_dev, const uint16_t *b_f16_dev, uint32_t *ann_dev, float *annd_dev,
                             uint32_t *bnn_dev, float *bnnd_dev, const int params[11]);

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

/* ---------------------------------------------------------------- colour space / resize
 * OpenCV 2.4.10 arithmetic (third-party, not in the reference tree), restated bit-exactly
 * (verified against cv2's plain C++ paths over all 2^24 colours).  8-bit images are H x W x 3. */

/* cvtColor(src, dst, CV_BGR2Lab) on 8UC3: NCT/main.cu:352,371; CT/ColorTransfer.h:58 */
int nct_bgr2lab_u8(nct_ctx *ctx, const uint8_t *bgr_dev, uint8_t *lab_dev, int npix);
/* cvtColor(src, dst, CV_Lab2BGR) on 8UC3: CT/ColorTransfer.cpp:1469 */
int nct_lab2bgr_u8(nct_ctx *ctx, const uint8_t *lab_dev, uint8_t *bgr_dev, int npix);
/* resize(src, dst, Size(dw, dh), 0, 0, CV_INTER_LINEAR) on 8UC3: NCT/main.cu:106-107, 509, 521 */
int nct_resize_linear_u8c3(nct_ctx *ctx, const uint8_t *src_dev, int sh, int sw, uint8_t *dst_dev, int dh, int dw);
/* the same on 64FC3: CT/ColorTransfer.cpp:462-463 */
int nct_resize_linear_f64c3(nct_ctx *ctx, const double *src_dev, int sh, int sw, double *dst_dev, int dh, int dw);

/* ---------------------------------------------------------------- colour fit (ColorTransfer) */

/* build_accumTable_downsample x2 + the local fit loop of transfer_color_downsample
 * (CT/ColorTransfer.cpp:425-455, 1194-1265): a = sigma_S / (sigma_C + eps), b = (mu_S - mu_C a)/255 from the
 * clipped 3x3 window of the 8-bit Lab images.  a_dev / b_dev: h*w*3 doubles (Vec3d layout). */
int nct_local_fit(nct_ctx *ctx, const uint8_t *cnt_lab_dev, const uint8_t *stl_lab_dev, int h, int w, double eps,
                  double *a_dev, double *b_dev);

/* m_weight = max(1 - (err - min)/(max - min), 1e-6): CT/ColorTransfer.cpp:1302-1340 */
int nct_confidence_weights(nct_ctx *ctx, const float *err_dev, int n, double *weight_dev);

/* solve_nonlocal_downsample_gpu_gradient + solve_ls_cg_gpu x3 (CT/ColorTransfer.cpp:548-949,
 * CT/SparseSolver_GPU.cu:3-198), matrix-free.  a/b in: local fit (start vector), out: refined.  knn_id_dev:
 * [h*w][8] int pixel ids (-1 = none); knn_w_dev: [h*w][8] NN.w values.  iters_out (host, may be NULL; passing it
 * synchronises) receives the CG iterations per channel. */
int nct_solve_nonlocal(nct_ctx *ctx, double *a_dev, double *b_dev, const double *weight_dev,
                       const uint8_t *cnt_lab_dev, const uint8_t *stl_lab_dev, const int *knn_id_dev,
                       const double *knn_w_dev, int h, int w, int layer, double local_weight, double alpha,
                       double nonlocal_weight, int knum, double d_weight, int iters_out[3]);

/* solve_ls_cg_gpu with its own argument list (CT/SparseSolver_GPU.cuh:12, CT/SparseSolver_GPU.cu:3-198), for callers
 * that keep the reference's host assembly (CT/ColorTransfer.cpp:548-949): A is constraints x size in CSR with ONE-based
 * rowindex / columns, every array on the HOST; x holds the start vector and receives the result; plain CG on
 * A^T A x = A^T b, `while (r1 > tolerance^2 && k <= maxitrs)`.  A^T A is applied as A^T (A p), never formed.
 * Synchronous (x is a host result).  iters_out (may be NULL) receives the number of iterations done. */
int nct_solve_ls_cg(nct_ctx *ctx, int size, int constraints, const double *A, const int *columns, const int *rowindex,
                    double *x, const double *b, int nonzeros, double tolerance, int maxitrs, int *iters_out);

/* upsample_color_coefficients_bilinear (CT/ColorTransfer.cpp:457-490): level -> full size + roughness map */
int nct_upsample_coefficients(nct_ctx *ctx, const double *a_lvl_dev, const double *b_lvl_dev, int h, int w,
                              const uint8_t *cnt_lab_full_dev, int H, int W, double *a_full_dev, double *b_full_dev,
                              double *rough_dev);

/* solve_direct_cpu(aRes, bRes, nonZeroNum, eleNum, oneBased, A, rowIndex, columns, Ba0, Xa0, ..., Bb2, Xb2)
 * (CT/SparseSolver_CPU.h:35-43, CT/SparseSolver_CPU.cpp:104-286: MKL PARDISO, real SPD, UPPER triangle in CSR, six right-hand
 * sides) with the same numerical arguments: host arrays, one-based CSR when one_based != 0, B[k] / X[k] = the six
 * right-hand sides / solutions (a0 a1 a2 b0 b1 b2).  Any SPD matrix in that format: Jacobi-preconditioned CG in FP64 on the
 * GPU to the relative residual rel_tol (<= 0: 1e-10; PARDISO is exact to rounding).  The pipeline does not use it: for the WLS
 * system nct_solve_wls (multigrid, matrix-free) is ~40x faster; this is for callers that keep the reference's assembly
 * (CT/ColorTransfer.cpp:951-1099). */
int nct_solve_direct(nct_ctx *ctx, int nnz, int n, int one_based, const double *A, const int *row_index, const int *columns,
                     const double *const B[6], double *const X[6], double rel_tol, int max_iters, int *iters_out,
                     double *rel_res_out);

/* solve_WLS_roughness_cpu + solve_direct_cpu (CT/ColorTransfer.cpp:951-1125, CT/SparseSolver_CPU.cpp:104-286):
 * (diag(rough) + L_g) x = diag(rough) x0 for the six maps, in place.  rel_tol <= 0 -> 1e-10, max_iters <= 0 -> default.
 * Synchronises the stream (the iteration count is data dependent). */
int nct_solve_wls(nct_ctx *ctx, double *a_dev, double *b_dev, const double *rough_dev, const uint8_t *cnt_lab_full_dev,
                  int H, int W, double lam, double alpha, double rel_tol, int max_iters, int *iters_out,
                  double *rel_res_out);

/* Same system solved by Jacobi-preconditioned CG (thousands of iterations; kept as an independent cross-check of
 * the multigrid solver above). */
int nct_solve_wls_jacobi(nct_ctx *ctx, double *a_dev, double *b_dev, const double *rough_dev,
                         const uint8_t *cnt_lab_full_dev, int H, int W, double lam, double alpha, double rel_tol,
                         int max_iters, int *iters_out, double *rel_res_out);

/* res = clamp(Lab a + b, 0, 1) -> convertTo(8U, 255) -> Lab2BGR (CT/ColorTransfer.cpp:1436-1469).
 * out_lab_dev may be NULL. */
int nct_apply_coefficients(nct_ctx *ctx, const uint8_t *cnt_lab_full_dev, const double *a_dev, const double *b_dev,
                           int H, int W, uint8_t *out_bgr_dev, uint8_t *out_lab_dev);

/* Measurement aid (no reference counterpart): read bandwidth in GB/s of coalesced 16-byte loads over
 * a `bytes` buffer read `passes` times in one launch (best of 5 launches).  A buffer that fits the L2 (e.g. 48 MB)
 * measures the L2 read roofline PatchMatch's candidate-row gathers run against; one far larger than L2 the HBM one. */
int nct_probe_read_bandwidth(nct_ctx *ctx, size_t bytes, int passes, double *gbps_out);

/* ---------------------------------------------------------------- VGG-19 features
 * Replaces Classifier (NCT/Classifier.h:51-61, NCT/Classifier.cpp:5-143) + caffe::Net<float> for the fixed graph of
 * demo/model/vgg19/VGG_ILSVRC_19_layers_deploy.prototxt, truncated after conv5_1.  Trunk layer index 0..12 =
 * conv1_1, conv1_2, conv2_1, conv2_2, conv3_1..conv3_4, conv4_1..conv4_4, conv5_1.  Feature level l = 0..4 =
 * conv5_1, conv4_1, conv3_1, conv2_1, conv1_1 (post-ReLU: the prototxt's ReLUs are in place), the order of
 * params.layers in NCT/main.cu:55-59. */
int nct_vgg19_num_layers(void);
const char *nct_vgg19_layer_name(int layer);
int nct_vgg19_layer_shape(int layer, int *cin, int *cout);
/* Caffe blob layout: weights O x I x 3 x 3, bias O (host pointers). Replaces Net::CopyTrainedLayersFrom
 * (caffe/net.cpp:798) for one layer. */
int nct_vgg19_set_weights(nct_ctx *ctx, int layer, const float *w_oihw_host, const float *bias_host);
/* Convolution engine: 0 = FP32 on CUDA cores (exact FP32 products, fixed summation order: bit-exact vs
 * oracle/conv_oracle.c; the default of a new context), 1 = tcgen05 tensor cores, kind::tf32 operands from TMA-staged
 * shared memory, FP32 accumulation in TMEM (~1e-2 of the feature range after 13 layers), 2 = the same with the
 * 3xTF32 hi/lo operand split (FP32-accurate, ~2e-4 of the feature range), 3 = tcgen05 kind::i8 EXACT fixed point
 * (conv_i8.cu: activations and weights as balanced base-256 digit planes, 9 INT8 MMAs per K step into four INT32 TMEM
 * accumulators, combined exactly -- no unspecified accumulation order anywhere, bit-exact vs oracle/vgg.py::
 * features_fixedpoint, 3-6e-7 of the layer range against an FP64 convolution; the CLI's and bench.py's default).
 * Tensor-core engines cover the layers with Cin >= 64; conv1_1 always runs on CUDA cores. */
int nct_vgg19_set_engine(nct_ctx *ctx, int engine);
/* One convolution layer (3x3, pad 1, stride 1, + bias, ReLU: cudnnConvolutionForward + cudnnAddTensor + ReLU,
 * caffe/layers/cudnn_conv_layer.cu:20-37) through the exact fixed-point tensor-core engine, from plain FP32 tensors:
 * in_dev NHWC FP32 (values >= 0), w_oihw_host / bias_host in Caffe's blob layout (host), out_dev NHWC FP32.  Cin and
 * Cout multiples of 64.  acc_dbg_dev (optional, device int32 [4][H*W][Cout]) receives the raw INT32 accumulators.
 * Diagnostic / unit-test entry point: the trunk (nct_vgg19_features) keeps digit planes and weights resident. */
int nct_conv3x3_fixedpoint(nct_ctx *ctx, const float *in_dev, const float *w_oihw_host, const float *bias_host, float *out_dev,
                           int *acc_dbg_dev, int H, int W, int Cin, int Cout);
/* Feature-map sizes {C, H, W} per level for an h x w image under Caffe's ceil-mode pooling
 * (caffe/layers/pooling_layer.cpp:90-93); replaces the Dim outputs of Classifier::Predict (NCT/Classifier.h:30-43). */
int nct_vgg19_level_dims(int h, int w, int dims[5][3]);
/* Classifier::Predict(img, layers, data_s): 8-bit BGR device image -> post-ReLU feature maps, HWC FP32, written to
 * the caller's device buffers feat_dev[l] for l = deepest_level..4 (sizes from nct_vgg19_level_dims).  The forward
 * stops after the layer of `deepest_level` (0 = conv5_1 = full trunk), which is all a re-forward needs
 * (NCT/main.cu:424-427 re-runs the whole net). */
int nct_vgg19_features(nct_ctx *ctx, const uint8_t *bgr_dev, int h, int w, int deepest_level, float *feat_dev[5]);

/* Net::CopyTrainedLayersFrom(trained_file) (caffe/net.cpp:798-815; NCT/Classifier.cpp:20): reads the conv1_1..conv5_1
 * blobs from a binary .caffemodel (V1 `layers` or V2 `layer` records) with a built-in protobuf wire reader. */
int nct_vgg19_load_caffemodel(nct_ctx *ctx, const char *path);

/* ---------------------------------------------------------------- clustering / non-local neighbours */

/* ColorTransfer::clusterFeastures (CT/ColorTransfer.cpp:355-395) = root split of cvflann's hierarchical k-means
 * (CT/Flann/kmeans_index.h:700-880; branching k = 10, 11 iterations, FLANN_CENTERS_RANDOM after srand(1)).
 * feat: L2-normalised conv5_1 of the content image, HWC (h*w rows of C floats).  labels_dev: h*w ints in [0, k).
 * If the root cannot be split (fewer than k distinct points) all labels are 0.  Synchronises the stream. */
int nct_cluster_features(nct_ctx *ctx, const float *feat_norm_hwc_dev, int h, int w, int C, int k, int iterations,
                         int *labels_dev);

/* ColorTransfer::findKnns (CT/ColorTransfer.cpp:397-423 with getClusters :273-353, findSubKNNs :136-195,
 * sortMergeComputeWeight :60-110): for every pixel of the level-size 8-bit Lab image the 8 nearest other pixels
 * (in Lab) among the pixels sharing one of its (4-neighbour dilated) clusters; label cell (cx, cy) covers a
 * samples x samples pixel block.  knn_id_dev: [h*w][8] pixel ids (-1 = none), knn_w_dev: [h*w][8] exp(1 - d/3). */
int nct_find_knns(nct_ctx *ctx, const int *labels_dev, int lw, int lh, int nlabels, const uint8_t *lab_dev, int h, int w,
                  int samples, int *knn_id_dev, double *knn_w_dev);

/* ---------------------------------------------------------------- per-pair pipeline */

/* Config (CT/Config.h:55-98) + the constants hard-coded in the orchestrator (NCT/main.cu:64-83). */
typedef struct nct_config {
    double bds_weight;       /* m_reverseWeight  (-bds, overridden per pair by pairs.txt)  default 2.0   */
    double var_eps;          /* m_varEpslon      (-eps)                                    default 0.6   */
    double nonlocal_weight;  /* m_nonlocalWeight (-nl)                                     default 2.0   */
    double local_weight;     /* m_localWeight    (-l)                                      default 0.125 */
    double wls_lambda_init;  /* m_wlsLamdaInit   (-w)                                      default 0.024 */
    int cluster_num;         /* m_clusterNum  10 */
    int k_num;               /* m_kNum         8 */
    int patch_size;          /* m_patchSize    3 */
    double wls_alpha;        /* m_wlsAlpha   1.2 */
    int pm_iters;            /* params.iter   10 (NCT/main.cu:65) */
    int kmeans_iters;        /* 11 (CT/ColorTransfer.cpp:373) */
    double wls_rel_tol;      /* relative residual at which the WLS PCG stops (stands in for PARDISO's exact solve) */
    int stop_after_level;    /* 4 = full pyramid; smaller values stop early (test hook: intermediates stay in scratch) */
    int feature_store;       /* 0 = FP32 PatchMatch volumes (the reference's storage; default), 1 = FP16: the L2-normalised
                              * volumes PatchMatch gathers from are rounded to FP16 (half the bytes of the L2-bound kernel;
                              * throughput mode beyond the reference).  The result then equals the oracle
                              * run with FP16-rounded volumes (oracle/pipeline.py feature_store="f16") bit for bit, not the
                              * FP32 one. */
} nct_config;

void nct_config_default(nct_config *cfg);

/* transfer_color_single_bds(refineCS, classifier_C, classifier_S, config, cnt, stl, preName), NCT/main.cu:47-454:
 * 8-bit BGR content (ch x cw) and style (sh x sw) -> 8-bit BGR result of the content's size.  VGG-19 weights must
 * have been loaded.  _dev: all three images are device buffers, asynchronous except for the data-dependent WLS
 * iteration counts; the host variant copies in, runs, copies out and synchronises.  cfg == NULL -> defaults. */
int nct_transfer_pair_dev(nct_ctx *ctx, const uint8_t *cnt_bgr_dev, int ch, int cw, const uint8_t *stl_bgr_dev, int sh, int sw,
                          const nct_config *cfg, uint8_t *out_bgr_dev);
int nct_transfer_pair(nct_ctx *ctx, const uint8_t *cnt_bgr_host, int ch, int cw, const uint8_t *stl_bgr_host, int sh, int sw,
                      const nct_config *cfg, uint8_t *out_bgr_host);

/* The same search by exhaustive comparison inside each cluster (O(sum of squared cluster sizes)); independent
 * cross-check of the grid search above. */
int nct_find_knns_brute(nct_ctx *ctx, const int *labels_dev, int lw, int lh, int nlabels, const uint8_t *lab_dev, int h,
                        int w, int samples, int *knn_id_dev, double *knn_w_dev);

/* ---------------------------------------------------------------- image files / pair lists */

/* imread(path) (NCT/main.cu:483,491): PNG -> malloc'ed 8-bit BGR (alpha dropped, grey / palette expanded); free with
 * nct_png_free.  imwrite(path, bgr) (NCT/main.cu:538): 8-bit BGR -> RGB PNG. */
int nct_png_read(const char *path, uint8_t **bgr_out, int *h_out, int *w_out);
void nct_png_free(uint8_t *p);
int nct_png_write(const char *path, const uint8_t *bgr, int h, int w);

/* transfer_single (NCT/main.cu:456-543): runs every line `content style bds` of <input_dir>/pairs.txt whose index i
 * satisfies i % world == rank and writes <output_dir>/<content>_<style>_<bds %2.2f>.png.  Images with a side above
 * 1000 are shrunk first (MAX_SIZE, CT/Config.h:5).  A pair that cannot be read is reported and skipped. */
int nct_run_pairs(nct_ctx *ctx, const char *input_dir, const char *output_dir, const nct_config *cfg, int rank, int world,
                  int *pairs_done);
/* The same with options and counters.  flags: NCT_RUN_RESUME = skip a pair whose output file already exists (restart of an
 * interrupted list; the reference recomputes everything, NCT/main.cu:471-540), NCT_RUN_VIS = also write the per-level debug
 * artefacts of the reference's ENABLE_VIS build next to the result (nct_set_vis).  pairs_failed counts the owned pairs
 * that could not be read, processed or written (each is reported on stdout and skipped, as nct_run_pairs does). */
#define NCT_RUN_RESUME 1
#define NCT_RUN_VIS 2
int nct_run_pairs_ex(nct_ctx *ctx, const char *input_dir, const char *output_dir, const nct_config *cfg, int rank, int world,
                     int flags, int *pairs_done, int *pairs_failed, int *pairs_skipped);
/* Debug artefacts of the reference's ENABLE_VIS build (NCT/main.cu:169-173, 333-347, 361-364, 382-422; off by default there
 * too, CT/Config.h:8): while `dir` is set, nct_transfer_pair(_dev) on this context writes, per pyramid level l,
 * <dir>/<prefix>_{aFlow,bFlow,tCnt,tStl,knn,errMap,refine_init,refine_nonlocal,aVis,aVis_init,aVis_nonlocal,bVis,bVis_init,
 * bVis_nonlocal}_<l>.png and <dir>/<prefix>_cluster_small.png (host rendering, slow: a debugging aid).  dir = NULL: off. */
int nct_set_vis(nct_ctx *ctx, const char *dir, const char *prefix);


#ifdef __cplusplus
}
#endif
#endif /* NCT_H */
