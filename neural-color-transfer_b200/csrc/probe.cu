// Read-bandwidth probe: the roofline denominators bench.py reports next to the PatchMatch kernel are measured in the same
// process on the same GPU -- a buffer that fits the L2 (default 48 MB: "l2") or one far larger than it ("hbm").
// Not part of the reference's surface (measurement aid).
#include "device_utils.cuh"

namespace {

// every thread streams 16-byte words `passes` times over the buffer, the same coalesced LDG.128 pattern the PatchMatch
// row reads use; the sum keeps the loads alive
__global__ void __launch_bounds__(256) probe_read_kernel(const uint4 *__restrict__ buf, size_t n16, int passes, unsigned *__restrict__ sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
            const uint4 v = __ldcg(buf + i);   // L2-coherent load: bypasses L1, so a small buffer measures L2, not L1
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x9e3779b9u) *sink = acc;
}

}  // namespace

extern "C" int nct_probe_read_bandwidth(nct_ctx *ctx, size_t bytes, int passes, double *gbps_out)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, bytes >= (1u << 20) && passes > 0 && gbps_out, "bad arguments");
    bytes &= ~(size_t)15;
    uint4 *buf = (uint4 *)nct_scratch(ctx, "probe_buf", bytes);
    unsigned *sink = (unsigned *)nct_scratch(ctx, "probe_sink", 16);
    if (!buf || !sink) return NCT_ERR_NOMEM;
    NCT_CUDA(ctx, cudaMemsetAsync(buf, 1, bytes, ctx->stream));
    const int blocks = ctx->num_sms * 8;
    cudaEvent_t e0, e1;
    NCT_CUDA(ctx, cudaEventCreate(&e0));
    NCT_CUDA(ctx, cudaEventCreate(&e1));
    probe_read_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, bytes / 16, 1, sink);   // warm: brings the buffer into L2
    NCT_CHECK_LAUNCH(ctx);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        NCT_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        probe_read_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, bytes / 16, passes, sink);
        NCT_CHECK_LAUNCH(ctx);
        NCT_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        NCT_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        NCT_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double g = (double)bytes * passes / 1e9 / (ms / 1e3);
        if (g > best) best = g;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *gbps_out = best;
    return NCT_OK;
}
