// Feature clustering (k-means, once per pair) and in-cluster 8-NN search in Lab (once per level) on the GPU.
//
// Replaces (all host code in the reference, after a D2H copy of conv5_1):
//   ColorTransfer::clusterFeastures + cvflann KMeansIndex root split   CT/ColorTransfer.cpp:355-395,
//                                                                       CT/Flann/kmeans_index.h:108-137, 700-880
//   ColorTransfer::findKnns (getClusters, per-cluster nanoflann KD-trees, sortMergeComputeWeight)
//                                                                       CT/ColorTransfer.cpp:60-110, 136-195, 273-423
// Specification = oracle/cluster_oracle.c (decisions K1-K5): MSVC rand/random_shuffle for the initial centres,
// float L2 in groups of four, exact integer Lab distances ranked by (d^2, pixel id).
//
// k-means works on 1936 x 512 floats: tiny, launch-latency bound; the kernels only have to be deterministic.
// The 8-NN search is an exact brute force inside each (dilated) cluster: a block owns 256 query pixels of one
// cluster and streams that cluster's members through shared memory; a candidate costs ~6 integer instructions
// (vabsdiff4 + dp4a on the packed 8-bit Lab triple, 64-bit key compare); per-thread top-8 lives in registers.
#include "device_utils.cuh"
#include <vector>

int nct_sort_pairs_u32(nct_ctx *ctx, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out,
                       int n, int end_bit);

namespace {

// ================================================================== k-means
struct KmState {
    int ok;        // initial centres found (otherwise all labels stay 0)
    int done;      // converged or not ok: remaining iterations are no-ops
    int changed;
};

__device__ float l2_ff(const float *__restrict__ a, const float *__restrict__ b, int size)
{
    float result = 0.f;
    int i = 0;
    for (; i + 3 < size; i += 4) {
        const float d0 = __fsub_rn(a[i], b[i]), d1 = __fsub_rn(a[i + 1], b[i + 1]), d2 = __fsub_rn(a[i + 2], b[i + 2]),
                    d3 = __fsub_rn(a[i + 3], b[i + 3]);
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
        result = __fadd_rn(result, s);
    }
    for (; i < size; ++i) {
        const float d0 = __fsub_rn(a[i], b[i]);
        result = __fadd_rn(result, __fmul_rn(d0, d0));
    }
    return result;
}

__device__ float l2_fd(const float *__restrict__ a, const double *__restrict__ b, int size)
{
    float result = 0.f;
    int i = 0;
    for (; i + 3 < size; i += 4) {
        const float d0 = (float)__dsub_rn((double)a[i], b[i]), d1 = (float)__dsub_rn((double)a[i + 1], b[i + 1]),
                    d2 = (float)__dsub_rn((double)a[i + 2], b[i + 2]), d3 = (float)__dsub_rn((double)a[i + 3], b[i + 3]);
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
        result = __fadd_rn(result, s);
    }
    for (; i < size; ++i) {
        const float d0 = (float)__dsub_rn((double)a[i], b[i]);
        result = __fadd_rn(result, __fmul_rn(d0, d0));
    }
    return result;
}

// chooseCentersRandom: walk the shuffled index list, skip points closer than 1e-16 to an already chosen centre
__global__ void km_pick_centers_kernel(const float *__restrict__ f, int n, int dim, int k, const int *__restrict__ shuffled,
                                       int *__restrict__ centers, KmState *st, int *__restrict__ labels)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int counter = 0, index = 0;
    bool ok = n >= k;
    for (index = 0; ok && index < k; ++index) {
        bool duplicate = true;
        while (duplicate) {
            duplicate = false;
            if (counter == n) { ok = false; break; }
            const int rnd = shuffled[counter++];
            centers[index] = rnd;
            for (int j = 0; j < index; ++j)
                if (l2_ff(f + (size_t)centers[index] * dim, f + (size_t)centers[j] * dim, dim) < 1e-16f) duplicate = true;
        }
    }
    st->ok = ok ? 1 : 0;
    st->done = ok ? 0 : 1;
    st->changed = 0;
}

__global__ void km_init_kernel(const float *__restrict__ f, int dim, int k, const int *__restrict__ centers, const KmState *st,
                               double *__restrict__ dc)
{
    if (!st->ok) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < k * dim) dc[t] = (double)f[(size_t)centers[t / dim] * dim + t % dim];
}

__global__ void km_dist_kernel(const float *__restrict__ f, int n, int dim, int k, const double *__restrict__ dc,
                               const KmState *st, float *__restrict__ dist)
{
    if (st->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int i = t / k, j = t % k;
    dist[t] = l2_fd(f + (size_t)i * dim, dc + (size_t)j * dim, dim);
}

// nearest centre (first minimum), radius of every cluster (max distance of its members), change flag
__global__ void km_assign_kernel(const float *__restrict__ dist, int n, int k, KmState *st, int *__restrict__ belongs,
                                 unsigned *__restrict__ radius_bits, int first)
{
    if (st->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sq = dist[(size_t)i * k];
    int nc = 0;
    for (int j = 1; j < k; ++j) {
        const float nsq = dist[(size_t)i * k + j];
        if (sq > nsq) { nc = j; sq = nsq; }
    }
    atomicMax(&radius_bits[nc], __float_as_uint(sq));  // distances are >= 0: uint order == float order
    if (first || nc != belongs[i]) {
        belongs[i] = nc;
        if (!first) st->changed = 1;
    }
}

__global__ void km_count_kernel(const int *__restrict__ belongs, int n, const KmState *st, int *__restrict__ count)
{
    if (st->done) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&count[belongs[i]], 1);
}

// member lists of the clusters in ascending point order (one warp per cluster: ballot + prefix count), so that the centre
// sums below walk ~n / k points instead of testing all n
__global__ void km_members_kernel(const int *__restrict__ belongs, int n, int k, const KmState *st, int *__restrict__ members)
{
    if (st->done) return;
    const int c = blockIdx.x, lane = threadIdx.x;
    int off = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool m = i < n && belongs[i] == c;
        const unsigned b = __ballot_sync(0xffffffffu, m);
        if (m) members[(size_t)c * n + off + __popc(b & ((1u << lane) - 1u))] = i;
        off += __popc(b);
    }
}

// new centres: double sums in point-index order (one thread per (cluster, dim))
__global__ void km_centers_kernel(const float *__restrict__ f, int n, int dim, int k, const int *__restrict__ members,
                                  const int *__restrict__ count, const KmState *st, double *__restrict__ dc)
{
    if (st->done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k * dim) return;
    const int c = t / dim, d = t % dim;
    const int m = count[c];
    const int *list = members + (size_t)c * n;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s = __dadd_rn(s, (double)f[(size_t)list[j] * dim + d]);
    dc[t] = __ddiv_rn(s, (double)m);
}

__global__ void km_reset_kernel(const KmState *st, unsigned *__restrict__ radius_bits, int *__restrict__ count, int k)
{
    if (st->done) return;
    const int t = threadIdx.x;
    if (t < k) {
        radius_bits[t] = 0;
        count[t] = 0;
    }
}

// empty-cluster repair + convergence bookkeeping (CT/Flann/kmeans_index.h:808-834), single thread
__global__ void km_fixup_kernel(const float *__restrict__ f, int n, int dim, int k, const double *__restrict__ dc,
                                int *__restrict__ belongs, int *__restrict__ count, const unsigned *__restrict__ radius_bits,
                                KmState *st)
{
    if (blockIdx.x != 0 || threadIdx.x != 0 || st->done) return;
    int changed = st->changed;
    for (int i = 0; i < k; ++i) {
        if (count[i] == 0) {
            int j = (i + 1) % k;
            while (count[j] <= 1) j = (j + 1) % k;
            const float rj = __uint_as_float(radius_bits[j]);
            for (int q = 0; q < n; ++q) {
                if (belongs[q] == j && l2_fd(f + (size_t)q * dim, dc + (size_t)j * dim, dim) == rj) {
                    belongs[q] = i;
                    count[j]--;
                    count[i]++;
                    break;
                }
            }
            changed = 1;
        }
    }
    st->done = changed ? 0 : 1;
    st->changed = 0;
}

// K1: MSVC rand() + VS2013 std::random_shuffle, seeded with srand(1)
void msvc_shuffled_indices(int n, std::vector<int> &v)
{
    uint32_t seed = 1;
    auto rnd = [&]() -> unsigned long {
        seed = seed * 214013u + 2531011u;
        return (unsigned long)((seed >> 16) & 0x7fff);
    };
    v.resize(n);
    for (int i = 0; i < n; ++i) v[i] = i;
    const unsigned long RBITS = 15, RMAX = (1UL << 15) - 1;
    for (unsigned long index = 2; (long)index <= n; ++index) {
        unsigned long rm = RMAX, rn = rnd() & RMAX;
        for (; rm < index && rm != ~0UL; rm = rm << RBITS | RMAX) rn = rn << RBITS | (rnd() & RMAX);
        std::swap(v[index - 1], v[rn % index]);
    }
}

// ================================================================== in-cluster 8-NN
constexpr int KNN_TPB = 256;
constexpr int KNN_CHUNK = 1024;
constexpr int MAXK = 16;  // cluster labels are bits of a 16-bit mask

__global__ void cell_masks_kernel(const int *__restrict__ labels, int lw, int lh, uint32_t *__restrict__ mask)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= lw * lh) return;
    const int x = id % lw, y = id / lw;
    uint32_t m = 1u << labels[id];
    if (x < lw - 1) m |= 1u << labels[id + 1];
    if (x > 0) m |= 1u << labels[id - 1];
    if (y < lh - 1) m |= 1u << labels[id + lw];
    if (y > 0) m |= 1u << labels[id - lw];
    mask[id] = m;
}

__global__ void pair_flags_kernel(const uint32_t *__restrict__ mask, int lw, int w, int n, int samples, int K,
                                  int *__restrict__ flags)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)K * n) return;
    const int l = (int)(t / n), p = (int)(t % n);
    const int cx = (p % w) / samples, cy = (p / w) / samples;
    flags[t] = (mask[cy * lw + cx] >> l) & 1u;
}

// per-cluster offsets / counts / first block of the padded block layout
__global__ void cluster_layout_kernel(const int *__restrict__ pos, int n, int K, int *__restrict__ off /*K+1*/,
                                      int *__restrict__ blk_start /*K+1*/)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int b = 0;
    for (int l = 0; l <= K; ++l) off[l] = pos[(size_t)l * n];
    for (int l = 0; l < K; ++l) {
        blk_start[l] = b;
        b += (off[l + 1] - off[l] + KNN_TPB - 1) / KNN_TPB;
    }
    blk_start[K] = b;
}

__global__ void members_kernel(const int *__restrict__ flags, const int *__restrict__ pos, const uint8_t *__restrict__ lab, int n,
                               int K, uint32_t *__restrict__ mem_lab, int *__restrict__ mem_id)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)K * n) return;
    if (!flags[t]) return;
    const int p = (int)(t % n);
    const int o = pos[t];
    mem_lab[o] = (uint32_t)lab[(size_t)p * 3] | ((uint32_t)lab[(size_t)p * 3 + 1] << 8) | ((uint32_t)lab[(size_t)p * 3 + 2] << 16);
    mem_id[o] = p;
}

struct Top8 {
    unsigned long long key[8];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) key[i] = ~0ull;
    }
    // keys are unique per candidate (they contain the pixel id); equal keys = the same candidate seen twice
    __device__ __forceinline__ void insert(unsigned long long k)
    {
        if (k >= key[7]) return;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (key[i] == k) return;
#pragma unroll
        for (int i = 7; i > 0; --i) {
            if (key[i - 1] > k) key[i] = key[i - 1];
            else if (key[i] > k) { key[i] = k; k = ~0ull; }
        }
        if (k != ~0ull && key[0] > k) key[0] = k;
    }
};

__global__ void __launch_bounds__(KNN_TPB) knn_cluster_kernel(const uint32_t *__restrict__ mem_lab, const int *__restrict__ mem_id,
                                                              const int *__restrict__ off, const int *__restrict__ blk_start, int K,
                                                              unsigned long long *__restrict__ pair_top)
{
    __shared__ uint2 cand[KNN_CHUNK];
    const int b = blockIdx.x;
    if (b >= blk_start[K]) return;
    int l = 0;
    while (l + 1 < K && b >= blk_start[l + 1]) ++l;
    const int base = off[l], cnt = off[l + 1] - off[l];
    const int qi = (b - blk_start[l]) * KNN_TPB + threadIdx.x;
    const bool active = qi < cnt;
    const uint32_t qlab = active ? mem_lab[base + qi] : 0u;
    const int qid = active ? mem_id[base + qi] : -1;
    Top8 top;
    top.init();
    for (int c0 = 0; c0 < cnt; c0 += KNN_CHUNK) {
        const int m = min(KNN_CHUNK, cnt - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < m; t += KNN_TPB) cand[t] = make_uint2(mem_lab[base + c0 + t], (uint32_t)mem_id[base + c0 + t]);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int t = 0; t < m; ++t) {
                const uint2 c = cand[t];
                const uint32_t ad = __vabsdiffu4(qlab, c.x);
                const uint32_t d2 = __dp4a(ad, ad, 0u);
                const unsigned long long key = ((unsigned long long)d2 << 32) | c.y;
                if (key < top.key[7] && (int)c.y != qid) top.insert(key);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) pair_top[(size_t)(base + qi) * 8 + i] = top.key[i];
    }
}

__global__ void knn_merge_kernel(const uint32_t *__restrict__ mask, const int *__restrict__ pos, const unsigned long long *__restrict__ pair_top,
                                 int lw, int w, int n, int samples, int K, int *__restrict__ knn_id, double *__restrict__ knn_w,
                                 const double *__restrict__ wtab)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int cx = (p % w) / samples, cy = (p / w) / samples;
    const uint32_t m = mask[cy * lw + cx];
    Top8 top;
    top.init();
    for (int l = 0; l < K; ++l) {
        if (!((m >> l) & 1u)) continue;
        const int o = pos[(size_t)l * n + p];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const unsigned long long k = pair_top[(size_t)o * 8 + i];
            if (k != ~0ull) top.insert(k);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned long long k = top.key[i];
        if (k != ~0ull) {
            knn_id[(size_t)p * 8 + i] = (int)(uint32_t)(k & 0xffffffffull);
            knn_w[(size_t)p * 8 + i] = wtab[(uint32_t)(k >> 32)];  // exp(1 - sqrt(D2)/255/3), host-libm table
        } else {
            knn_id[(size_t)p * 8 + i] = -1;
            knn_w[(size_t)p * 8 + i] = 0.0;
        }
    }
}

}  // namespace

extern "C" {

int nct_cluster_features(nct_ctx *ctx, const float *feat_norm_hwc_dev, int h, int w, int C, int k, int iterations, int *labels_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, feat_norm_hwc_dev && labels_dev && h > 0 && w > 0 && C > 0, "bad arguments");
    NCT_REQUIRE(ctx, k >= 2 && k <= MAXK, "cluster count %d out of range [2, %d]", k, MAXK);
    const int n = h * w;
    std::vector<int> shuf;
    msvc_shuffled_indices(n, shuf);
    int *d_shuf = (int *)nct_scratch(ctx, "km_shuffle", sizeof(int) * (size_t)n);
    char *misc = (char *)nct_scratch(ctx, "km_misc", 4096);
    double *dc = (double *)nct_scratch(ctx, "km_centers", sizeof(double) * (size_t)k * C);
    float *dist = (float *)nct_scratch(ctx, "km_dist", sizeof(float) * (size_t)n * k);
    int *members = (int *)nct_scratch(ctx, "km_members", sizeof(int) * (size_t)n * k);
    if (!d_shuf || !misc || !dc || !dist || !members) return NCT_ERR_NOMEM;
    KmState *st = (KmState *)misc;
    int *centers = (int *)(misc + 64);
    unsigned *radius = (unsigned *)(misc + 256);
    int *count = (int *)(misc + 512);
    // pageable -> device copy of a few KB: the staging copy completes before the call returns
    NCT_CUDA(ctx, cudaMemcpyAsync(d_shuf, shuf.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemsetAsync(labels_dev, 0, sizeof(int) * (size_t)n, ctx->stream));
    NCT_CUDA(ctx, cudaMemsetAsync(misc, 0, 4096, ctx->stream));
    km_pick_centers_kernel<<<1, 32, 0, ctx->stream>>>(feat_norm_hwc_dev, n, C, k, d_shuf, centers, st, labels_dev);
    NCT_CHECK_LAUNCH(ctx);
    km_init_kernel<<<nct_div_up(k * C, 256), 256, 0, ctx->stream>>>(feat_norm_hwc_dev, C, k, centers, st, dc);
    NCT_CHECK_LAUNCH(ctx);
    for (int it = 0; it <= iterations; ++it) {
        if (it > 0) {
            km_members_kernel<<<k, 32, 0, ctx->stream>>>(labels_dev, n, k, st, members);
            NCT_CHECK_LAUNCH(ctx);
            km_centers_kernel<<<nct_div_up(k * C, 128), 128, 0, ctx->stream>>>(feat_norm_hwc_dev, n, C, k, members, count, st, dc);
            NCT_CHECK_LAUNCH(ctx);
        }
        km_reset_kernel<<<1, 32, 0, ctx->stream>>>(st, radius, count, k);
        NCT_CHECK_LAUNCH(ctx);
        km_dist_kernel<<<nct_div_up(n * k, 128), 128, 0, ctx->stream>>>(feat_norm_hwc_dev, n, C, k, dc, st, dist);
        NCT_CHECK_LAUNCH(ctx);
        km_assign_kernel<<<nct_div_up(n, 128), 128, 0, ctx->stream>>>(dist, n, k, st, labels_dev, radius, it == 0 ? 1 : 0);
        NCT_CHECK_LAUNCH(ctx);
        km_count_kernel<<<nct_div_up(n, 128), 128, 0, ctx->stream>>>(labels_dev, n, st, count);
        NCT_CHECK_LAUNCH(ctx);
        if (it > 0) {
            km_fixup_kernel<<<1, 32, 0, ctx->stream>>>(feat_norm_hwc_dev, n, C, k, dc, labels_dev, count, radius, st);
            NCT_CHECK_LAUNCH(ctx);
        }
    }
    // the shuffled index vector must outlive the async copy
    NCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NCT_OK;
}

int nct_find_knns_brute(nct_ctx *ctx, const int *labels_dev, int lw, int lh, int nlabels, const uint8_t *lab_dev, int h, int w,
                  int samples, int *knn_id_dev, double *knn_w_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, labels_dev && lab_dev && knn_id_dev && knn_w_dev, "null pointer");
    NCT_REQUIRE(ctx, nlabels >= 1 && nlabels <= MAXK && samples >= 1, "bad cluster count / samples");
    NCT_REQUIRE(ctx, (long long)lw * samples >= w && (long long)lh * samples >= h, "label grid %dx%d x %d does not cover the %dx%d image", lw, lh, samples, w, h);
    const int n = h * w, K = nlabels;
    const double *wtab = nct_knn_weight_table(ctx);
    if (!wtab) return NCT_ERR_NOMEM;
    const size_t KN = (size_t)K * n;
    NCT_REQUIRE(ctx, KN < (1ull << 31), "image too large");
    uint32_t *mask = (uint32_t *)nct_scratch(ctx, "knn_mask", sizeof(uint32_t) * (size_t)lw * lh);
    int *flags = (int *)nct_scratch(ctx, "knn_flags", sizeof(int) * (KN + 1));
    int *pos = (int *)nct_scratch(ctx, "knn_pos", sizeof(int) * (KN + 1));
    int *layout = (int *)nct_scratch(ctx, "knn_layout", sizeof(int) * 2 * (MAXK + 1));
    // a pixel belongs to at most 5 clusters (own label + 4 neighbouring cells' labels)
    const size_t max_pairs = (size_t)n * (K < 5 ? K : 5);
    uint32_t *mem_lab = (uint32_t *)nct_scratch(ctx, "knn_mem_lab", sizeof(uint32_t) * max_pairs);
    int *mem_id = (int *)nct_scratch(ctx, "knn_mem_id", sizeof(int) * max_pairs);
    unsigned long long *pair_top = (unsigned long long *)nct_scratch(ctx, "knn_pair_top", sizeof(unsigned long long) * 8 * max_pairs);
    if (!mask || !flags || !pos || !layout || !mem_lab || !mem_id || !pair_top) return NCT_ERR_NOMEM;
    int *off = layout, *blk_start = layout + MAXK + 1;
    cell_masks_kernel<<<nct_div_up(lw * lh, 256), 256, 0, ctx->stream>>>(labels_dev, lw, lh, mask);
    NCT_CHECK_LAUNCH(ctx);
    pair_flags_kernel<<<(unsigned)((KN + 255) / 256), 256, 0, ctx->stream>>>(mask, lw, w, n, samples, K, flags);
    NCT_CHECK_LAUNCH(ctx);
    int rc = nct_exclusive_scan_i32(ctx, flags, pos, (int)KN);
    if (rc) return rc;
    cluster_layout_kernel<<<1, 32, 0, ctx->stream>>>(pos, n, K, off, blk_start);
    NCT_CHECK_LAUNCH(ctx);
    members_kernel<<<(unsigned)((KN + 255) / 256), 256, 0, ctx->stream>>>(flags, pos, lab_dev, n, K, mem_lab, mem_id);
    NCT_CHECK_LAUNCH(ctx);
    const int max_blocks = (int)(max_pairs / KNN_TPB) + K + 1;
    knn_cluster_kernel<<<max_blocks, KNN_TPB, 0, ctx->stream>>>(mem_lab, mem_id, off, blk_start, K, pair_top);
    NCT_CHECK_LAUNCH(ctx);
    knn_merge_kernel<<<nct_div_up(n, 128), 128, 0, ctx->stream>>>(mask, pos, pair_top, lw, w, n, samples, K, knn_id_dev, knn_w_dev, wtab);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // extern "C"
