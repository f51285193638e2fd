// Context, error reporting and workspace arena of libnct.
// Replaces the device setup of NCT/main.cu:562-570 (cudaSetDevice/cudaDeviceReset/
// cudaMemGetInfo) and the per-level cudaMalloc/cudaFree churn of NCT/main.cu:238-326.
#include "nct_internal.h"
#include <cmath>
#include <cstring>

int nct_fail(nct_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    else fprintf(stderr, "libnct: %s\n", buf);
    return code;
}

cudaError_t nct_stream_wait(nct_ctx *ctx)
{
    if (!ctx->wait_event) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->wait_event, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->wait_event, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->wait_event);
}

void *nct_scratch(nct_ctx *ctx, const char *name, size_t bytes)
{
    NctBuffer &b = ctx->scratch[name];
    if (b.bytes >= bytes && b.ptr) return b.ptr;
    if (b.ptr) {
        // the old buffer may still be in use by queued work on the stream
        cudaStreamSynchronize(ctx->stream);
        cudaFree(b.ptr);
        b.ptr = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    void *p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        nct_fail(ctx, NCT_ERR_NOMEM, "cudaMalloc(%zu) for scratch '%s' failed", want, name);
        return nullptr;
    }
    b.ptr = p;
    b.bytes = want;
    return p;
}

// Both tables are uploaded with stream-ordered copies on ctx->stream (a synchronous cudaMemcpy from pageable memory
// may return before the DMA has landed, and ctx->stream does not synchronise with the legacy default stream).
const double *nct_pow_table(nct_ctx *ctx, double alpha)
{
    // one table per exponent: the CG stage uses (double)(float)alpha, the WLS stage alpha itself
    // (CT/ColorTransfer.cpp:548-550 takes floats, :951 doubles); tables are never rebuilt in steady state
    char name[48];
    unsigned long long bits;
    memcpy(&bits, &alpha, sizeof(bits));
    snprintf(name, sizeof(name), "tab_pow_%016llx", bits);
    auto it = ctx->scratch.find(name);
    if (it != ctx->scratch.end() && it->second.ptr) return (const double *)it->second.ptr;
    double *dev = (double *)nct_scratch(ctx, name, sizeof(double) * 65536);
    if (!dev) return nullptr;
    std::vector<double> &host = ctx->pow_host[bits];
    host.resize(65536);
    for (int l0 = 0; l0 < 256; ++l0)
        for (int l1 = 0; l1 < 256; ++l1) {
            volatile double a = (double)l1 * (1.0 / 255.0), b = (double)l0 * (1.0 / 255.0);  // volatile: no contraction
            volatile double d = a - b;
            host[l0 * 256 + l1] = pow(fabs(d), alpha);
        }
    if (cudaMemcpyAsync(dev, host.data(), sizeof(double) * 65536, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        nct_fail(ctx, NCT_ERR_CUDA, "upload of the pow table failed");
        return nullptr;
    }
    return dev;
}

const double *nct_knn_weight_table(nct_ctx *ctx)
{
    const int n = 3 * 255 * 255 + 1;
    auto it = ctx->scratch.find("tab_knnw");
    if (it != ctx->scratch.end() && it->second.ptr) return (const double *)it->second.ptr;
    double *dev = (double *)nct_scratch(ctx, "tab_knnw", sizeof(double) * n);
    if (!dev) return nullptr;
    ctx->knnw_host.resize(n);
    for (int i = 0; i < n; ++i) {
        volatile double d = sqrt((double)i) / 255.0;
        volatile double q = d / 3.0;
        ctx->knnw_host[i] = exp(1.0 - q);
    }
    if (cudaMemcpyAsync(dev, ctx->knnw_host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
        nct_fail(ctx, NCT_ERR_CUDA, "upload of the k-NN weight table failed");
        return nullptr;
    }
    return dev;
}

// ---------------------------------------------------------------- captured launch sequences
bool nct_graph_cached(nct_ctx *ctx, const char *name, const std::vector<unsigned long long> &key)
{
    auto it = ctx->graphs.find(name);
    return it != ctx->graphs.end() && it->second.exec && it->second.key == key;
}

static void graph_release(NctGraph &g)
{
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr;
    g.graph = nullptr;
    g.capturing = false;
}

int nct_graph_begin(nct_ctx *ctx, const char *name, const std::vector<unsigned long long> &key, cudaGraphConditionalHandle *while_handle)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    NCT_CUDA(ctx, cudaStreamIsCapturing(ctx->stream, &st));
    if (st != cudaStreamCaptureStatusNone) {   // an earlier capture was abandoned by an error return: discard it
        cudaGraph_t dead = nullptr;
        cudaStreamEndCapture(ctx->stream, &dead);
        if (dead) cudaGraphDestroy(dead);
        cudaGetLastError();
    }
    NctGraph &g = ctx->graphs[name];
    graph_release(g);
    g.key = key;
    g.launches_at_begin = ctx->launches;
    // thread-local capture mode: other host threads (other contexts) keep allocating / synchronising freely
    if (while_handle) {
        NCT_CUDA(ctx, cudaGraphCreate(&g.graph, 0));
        NCT_CUDA(ctx, cudaGraphConditionalHandleCreate(while_handle, g.graph, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams p = {};
        p.type = cudaGraphNodeTypeConditional;
        p.conditional.handle = *while_handle;
        p.conditional.type = cudaGraphCondTypeWhile;
        p.conditional.size = 1;
        cudaGraphNode_t node;
        NCT_CUDA(ctx, cudaGraphAddNode(&node, g.graph, nullptr, 0, &p));
        NCT_CUDA(ctx, cudaStreamBeginCaptureToGraph(ctx->stream, p.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    } else {
        NCT_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    }
    g.capturing = true;
    return NCT_OK;
}

int nct_graph_end(nct_ctx *ctx, const char *name)
{
    auto it = ctx->graphs.find(name);
    if (it == ctx->graphs.end() || !it->second.capturing) return nct_fail(ctx, NCT_ERR_STATE, "graph '%s' is not being captured", name);
    NctGraph &g = it->second;
    g.capturing = false;
    cudaGraph_t captured = nullptr;
    NCT_CUDA(ctx, cudaStreamEndCapture(ctx->stream, &captured));
    if (!g.graph) g.graph = captured;   // plain capture; (capture into a WHILE body returns the body graph, owned by g.graph)
    g.nodes = ctx->launches - g.launches_at_begin;
    ctx->launches = g.launches_at_begin;   // nothing has run yet
    NCT_CUDA(ctx, cudaGraphInstantiate(&g.exec, g.graph, 0));
    return NCT_OK;
}

int nct_graph_abort(nct_ctx *ctx, const char *name, int rc)
{
    // leave the stream usable (a stream left in capture mode fails every later synchronise / copy on it) and keep the
    // message of the call that failed
    cudaGraph_t dead = nullptr;
    cudaStreamEndCapture(ctx->stream, &dead);
    auto it = ctx->graphs.find(name);
    if (it != ctx->graphs.end()) {
        if (dead && !it->second.graph) cudaGraphDestroy(dead);   // (a WHILE body belongs to its parent graph)
        graph_release(it->second);
        ctx->launches = it->second.launches_at_begin;
        ctx->graphs.erase(it);
    } else if (dead) {
        cudaGraphDestroy(dead);
    }
    cudaGetLastError();
    return rc;
}

int nct_graph_launch(nct_ctx *ctx, const char *name)
{
    auto it = ctx->graphs.find(name);
    if (it == ctx->graphs.end() || !it->second.exec) return nct_fail(ctx, NCT_ERR_STATE, "graph '%s' has not been captured", name);
    NCT_CUDA(ctx, cudaGraphLaunch(it->second.exec, ctx->stream));
    ctx->launches += it->second.nodes;
    return NCT_OK;
}

long long nct_graph_nodes(nct_ctx *ctx, const char *name)
{
    auto it = ctx->graphs.find(name);
    return it == ctx->graphs.end() ? 0 : it->second.nodes;
}

void nct_graphs_free(nct_ctx *ctx)
{
    for (auto &kv : ctx->graphs) graph_release(kv.second);
    ctx->graphs.clear();
}

NctStageTimer::NctStageTimer(nct_ctx *c, int stage) : ctx(c)
{
    if (!c || !c->profile) return;
    if (c->profile == 2 && stage != ST_PM) return;   // level 2: PatchMatch spans only (the roofline kernel)
    nct_ctx::ProfSpan sp;
    sp.stage = stage;
    auto get = [&]() {
        cudaEvent_t e;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    sp.e0 = get();
    sp.e1 = get();
    cudaEventRecord(sp.e0, c->stream);
    idx = (int)c->prof_spans.size();
    c->prof_spans.push_back(sp);
}

NctStageTimer::~NctStageTimer()
{
    if (idx >= 0) cudaEventRecord(ctx->prof_spans[idx].e1, ctx->stream);
}

extern "C" {

int nct_profile_enable(nct_ctx *ctx, int enable)
{
    NCT_ENTER(ctx);
    ctx->profile = enable == 2 ? 2 : (enable ? 1 : 0);   // 1 = every stage, 2 = the PatchMatch stage only
    if (enable) {
        // events for ~4 pairs up front: creating them inside the measured region (cudaEventCreate takes driver-wide locks the
        // other contexts' launching threads contend for) slowed the profiled context measurably
        while (ctx->prof_pool.size() < 256) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) break;
            ctx->prof_pool.push_back(e);
        }
    }
    return NCT_OK;
}

int nct_profile_reset(nct_ctx *ctx)
{
    NCT_ENTER(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &sp : ctx->prof_spans) { ctx->prof_pool.push_back(sp.e0); ctx->prof_pool.push_back(sp.e1); }
    ctx->prof_spans.clear();
    for (int i = 0; i < 16; ++i) { ctx->prof_ms[i] = 0; ctx->prof_calls[i] = 0; }
    return NCT_OK;
}

int nct_profile_get(nct_ctx *ctx, int stage, double *ms_out, long long *spans_out)
{
    if (!ctx || stage < 0 || stage >= ST_COUNT) return NCT_ERR_ARG;
    NCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &sp : ctx->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { ctx->prof_ms[sp.stage] += ms; ctx->prof_calls[sp.stage] += 1; }
        ctx->prof_pool.push_back(sp.e0);
        ctx->prof_pool.push_back(sp.e1);
    }
    ctx->prof_spans.clear();
    if (ms_out) *ms_out = ctx->prof_ms[stage];
    if (spans_out) *spans_out = ctx->prof_calls[stage];
    return NCT_OK;
}

const char *nct_profile_stage_name(int stage)
{
    static const char *names[ST_COUNT] = {"vgg", "patchmatch", "bds", "knn", "nonlocal_cg", "wls", "misc", "kmeans"};
    return (stage >= 0 && stage < ST_COUNT) ? names[stage] : nullptr;
}

int nct_version(void) { return 100; }

int nct_create(int gpu_id, nct_ctx **out)
{
    if (!out) return NCT_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return nct_fail(nullptr, NCT_ERR_CUDA, "no CUDA device available (%s); libnct has no CPU fallback",
                        cudaGetErrorString(e));
    if (gpu_id < 0 || gpu_id >= count)
        return nct_fail(nullptr, NCT_ERR_ARG, "gpu_id %d out of range (device count %d)", gpu_id, count);
    if (cudaSetDevice(gpu_id) != cudaSuccess) return nct_fail(nullptr, NCT_ERR_CUDA, "cudaSetDevice(%d) failed", gpu_id);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, gpu_id) != cudaSuccess)
        return nct_fail(nullptr, NCT_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return nct_fail(nullptr, NCT_ERR_CUDA, "device %d is sm_%d%d; libnct is built for sm_100a only", gpu_id,
                        prop.major, prop.minor);
    nct_ctx *ctx = new nct_ctx();
    ctx->device = gpu_id;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return nct_fail(nullptr, NCT_ERR_CUDA, "cudaStreamCreate failed");
    }
    ctx->stream = ctx->own_stream;
    if (cudaMalloc(&ctx->pm_counters, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return nct_fail(nullptr, NCT_ERR_NOMEM, "cudaMalloc failed");
    }
    cudaMemset(ctx->pm_counters, 0, 2 * sizeof(unsigned long long));
    *out = ctx;
    return NCT_OK;
}

// defined by vgg19.cu
extern "C++" void nct_vgg_free(nct_ctx *ctx);

int nct_destroy(nct_ctx *ctx)
{
    if (!ctx) return NCT_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nct_vgg_free(ctx);
    nct_graphs_free(ctx);
    for (auto &sp : ctx->prof_spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (auto &e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
    for (auto &kv : ctx->scratch)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (ctx->pm_counters) cudaFree(ctx->pm_counters);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return NCT_OK;
}

const char *nct_last_error(const nct_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null ctx"; }

int nct_set_stream(nct_ctx *ctx, void *cuda_stream)
{
    NCT_ENTER(ctx);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return NCT_OK;
}

void *nct_get_stream(const nct_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int nct_synchronize(nct_ctx *ctx)
{
    NCT_ENTER(ctx);
    NCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NCT_OK;
}

int nct_debug_read_scratch(nct_ctx *ctx, const char *name, void *host_dst, size_t bytes)
{
    if (!ctx || !name || !host_dst) return NCT_ERR_ARG;
    cudaSetDevice(ctx->device);
    auto it = ctx->scratch.find(name);
    if (it == ctx->scratch.end() || !it->second.ptr) return nct_fail(ctx, NCT_ERR_ARG, "no scratch buffer named '%s'", name);
    if (it->second.bytes < bytes) return nct_fail(ctx, NCT_ERR_ARG, "scratch '%s' holds %zu bytes < %zu", name, it->second.bytes, bytes);
    NCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NCT_CUDA(ctx, cudaMemcpy(host_dst, it->second.ptr, bytes, cudaMemcpyDeviceToHost));
    return NCT_OK;
}

long long nct_launch_count(const nct_ctx *ctx) { return ctx ? ctx->launches : 0; }
void nct_reset_launch_count(nct_ctx *ctx)
{
    if (ctx) ctx->launches = 0;
}

}  // extern "C"
