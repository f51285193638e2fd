// Internal definitions shared by the libnct translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <map>
#include "../../include/nct.h"

struct NctBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
};

// A captured launch sequence, replayed with one cudaGraphLaunch (the solvers' iteration blocks: thousands of tiny
// dependent launches per pair otherwise).  `key` = everything the captured kernels received by value (sizes, device
// pointers): a replay is only valid while it is unchanged, otherwise the sequence is captured again.
struct NctGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<unsigned long long> key;
    long long nodes = 0;                // kernel launches of one pass through the captured sequence
    long long launches_at_begin = 0;
    bool capturing = false;
};

struct nct_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string last_error;
    long long launches = 0;
    cudaEvent_t wait_event = nullptr;  // cudaEventBlockingSync, see nct_stream_wait

    // named, grow-only device scratch buffers ("workspace arena"): no per-level
    // cudaMalloc/cudaFree on the hot path (the reference does ~12 per level,
    // NCT/main.cu:238-257,293-299).
    std::map<std::string, NctBuffer> scratch;

    // PatchMatch bookkeeping
    int pm_count_evals = 0;
    unsigned long long *pm_counters = nullptr;  // device, 2 x u64

    // optional stage timing (CUDA events on ctx->stream), see nct_profile_*
    int profile = 0;
    struct ProfSpan { int stage; cudaEvent_t e0, e1; };
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {0};
    long long prof_calls[16] = {0};

    // transcendental tables evaluated by the HOST C library (the same libm the oracle uses), so that the gradient and
    // neighbour weights entering the un-converged CG are bit-identical to the oracle's (a 1-ulp
    // difference there changes the final image by ~43 dB)
    std::map<unsigned long long, std::vector<double>> pow_host;  // keyed by the exponent's bit pattern
    std::vector<double> knnw_host;

    // WLS warm start (set by the pair orchestrator around its per-level solves; off for direct nct_solve_wls callers)
    int wls_warm = 0;
    int wls_prev_n = 0;
    int wls_last_iters = 0;  // iteration count of the previous solve: the size of the first batch queued without a host check

    std::map<std::string, NctGraph> graphs;

    // ENABLE_VIS artefacts (vis.cpp): written per level into vis_dir when it is not empty
    std::string vis_dir, vis_prefix;

    // opaque sub-module states (owned, freed in nct_destroy)
    struct VggState *vgg = nullptr;
};

int nct_fail(nct_ctx *ctx, int code, const char *fmt, ...);
// Host wait for everything queued on ctx->stream WITHOUT spinning: a blocking-sync event, so that the many host
// threads of a multi-pair / multi-GPU run (one per context) sleep instead of burning a core each while they wait.
cudaError_t nct_stream_wait(nct_ctx *ctx);
// returns device pointer of at least `bytes` bytes, stable until a larger request under the same name
void *nct_scratch(nct_ctx *ctx, const char *name, size_t bytes);

// ---- captured launch sequences (context.cu).  Usage:
//   if (!nct_graph_cached(ctx, "name", key)) { nct_graph_begin(ctx, "name", key, &handle_or_null); ...launches on ctx->stream...;
//                                               nct_graph_end(ctx, "name"); }
//   nct_graph_launch(ctx, "name");
// Nothing may allocate (nct_scratch) or synchronise between begin and end.  With `while_handle` != nullptr the sequence
// becomes the body of a device-side WHILE loop (CUDA conditional graph node): it repeats until a kernel of the body calls
// cudaGraphSetConditional(*while_handle, 0); the handle's value is reset to 1 at every launch.
bool nct_graph_cached(nct_ctx *ctx, const char *name, const std::vector<unsigned long long> &key);
int nct_graph_begin(nct_ctx *ctx, const char *name, const std::vector<unsigned long long> &key, cudaGraphConditionalHandle *while_handle);
int nct_graph_end(nct_ctx *ctx, const char *name);
int nct_graph_abort(nct_ctx *ctx, const char *name, int rc);   // a launch failed while capturing: ends the capture, drops the graph, returns rc
int nct_graph_launch(nct_ctx *ctx, const char *name);   // adds the sequence's launch count to ctx->launches once
long long nct_graph_nodes(nct_ctx *ctx, const char *name);
void nct_graphs_free(nct_ctx *ctx);

// [256 * 256] pow(|l1 * (1/255) - l0 * (1/255)|, alpha) at index l0 * 256 + l1 (compute_gradientMat, CT/ColorTransfer.cpp:519-546)
const double *nct_pow_table(nct_ctx *ctx, double alpha);
// [3 * 255^2 + 1] exp(1 - sqrt(D2) / 255 / 3) indexed by the integer squared Lab distance (sortMergeComputeWeight, CT/ColorTransfer.cpp:60-110)
const double *nct_knn_weight_table(nct_ctx *ctx);

// Every public entry point starts with this: null check, and the context's device made current for the CALLING thread
// (the current device is per host thread: a worker thread that did not create the context starts on device 0).
#define NCT_ENTER(ctx)                                                                               \
    do {                                                                                             \
        if (!(ctx)) return NCT_ERR_ARG;                                                              \
        cudaSetDevice((ctx)->device);                                                                \
    } while (0)

#define NCT_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return nct_fail((ctx), NCT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                            cudaGetErrorString(e__));                                                \
    } while (0)

#define NCT_CHECK_LAUNCH(ctx)                                                                        \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                        \
        if (e__ != cudaSuccess)                                                                      \
            return nct_fail((ctx), NCT_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,    \
                            cudaGetErrorString(e__));                                                \
    } while (0)

#define NCT_REQUIRE(ctx, cond, ...)                                                                  \
    do {                                                                                             \
        if (!(cond)) return nct_fail((ctx), NCT_ERR_ARG, __VA_ARGS__);                               \
    } while (0)

enum NctStage { ST_VGG = 0, ST_PM, ST_BDS, ST_KNN, ST_CG, ST_WLS, ST_MISC, ST_KMEANS, ST_COUNT };
// RAII stage timer: records two events on the ctx stream when profiling is on
struct NctStageTimer {
    nct_ctx *ctx;
    int idx = -1;
    NctStageTimer(nct_ctx *c, int stage);
    ~NctStageTimer();
};

// vis.cpp (debug artefacts of the reference's ENABLE_VIS build)
int nct_vis_cluster_small(nct_ctx *ctx, const int *labels_dev, int lh, int lw);
int nct_vis_flows(nct_ctx *ctx, int level, const uint32_t *ann_dev, const uint32_t *bnn_dev, const uint8_t *cnt_dev, const uint8_t *stl_dev,
                  int ah, int aw, int bh, int bw);
int nct_vis_knn_clusters(nct_ctx *ctx, int level, const int *labels_dev, int lh, int lw, int h, int w, int samples);
int nct_vis_error_map(nct_ctx *ctx, int level, const float *err_dev, int h, int w);
int nct_vis_coefficients(nct_ctx *ctx, int level, const char *suffix, const double *a_dev, const double *b_dev, int mh, int mw, int H, int W,
                         int samples, std::vector<double> *a_up, std::vector<double> *b_up);
int nct_vis_refine(nct_ctx *ctx, int level, const char *what, const uint8_t *cnt_lab_full_dev, const double *a_full_dev, const double *b_full_dev,
                   int H, int W);

static inline int nct_div_up(int a, int b) { return (a + b - 1) / b; }
