// In-cluster 8-NN search in 8-bit Lab space with a uniform colour grid (exact).
//
// Replaces ColorTransfer::findKnns (CT/ColorTransfer.cpp:397-423: per-cluster nanoflann KD-trees on the host,
// OpenMP over <= 10 clusters, then sortMergeComputeWeight).  Specification: oracle/cluster_oracle.c (K4/K5):
// 8 nearest other pixels by exact integer squared Lab distance, ties by pixel id, among the pixels that share
// one of the query's (4-neighbour dilated) clusters.
//
// Method: every (pixel, cluster) membership is binned into a [cluster][L/2][a/2][b/2] table of 2x2x2-colour cells
// (counting sort: histogram, exclusive scan, scatter).  A query walks Chebyshev shells of cells around its own
// colour in each of its clusters; cells adjacent along b are contiguous in the sorted arrays, so a shell is a
// handful of contiguous ranges.  After shell R every unvisited point is at least LB away (distance to the visited
// box), so the search stops as soon as the 8th best distance is < LB^2 -- typically after the 3x3x3 shell.  Sparse
// outliers fall back to scanning their whole cluster once the shell radius exceeds a threshold, which bounds the
// cost by the brute force.  The result does not depend on the visiting order (keys are (d^2, id)), so the scatter
// order inside a cell does not matter: deterministic output.
#include "device_utils.cuh"
#include <cstdlib>

namespace {

// The cell size is chosen per call from the number of pixels (template parameter CS = log2 of the cell edge): the
// finest level (490k pixels) uses 2-unit cells, the coarse levels -- whose few thousand points would leave a fine grid
// almost empty and send every query through many empty shells -- use 4, 8 or 16-unit cells.
constexpr int MAXK = 16;
constexpr int FALLBACK_R = 6;               // shell radius after which the whole cluster is scanned instead

template <int CS>
__device__ __forceinline__ int cell_of(uint32_t lab)
{
    constexpr int CD = 256 >> CS;
    const int L = lab & 255, a = (lab >> 8) & 255, b = (lab >> 16) & 255;
    return ((L >> CS) * CD + (a >> CS)) * CD + (b >> CS);
}

__device__ __forceinline__ uint32_t pack_lab(const uint8_t *__restrict__ lab, int p)
{
    return (uint32_t)lab[(size_t)p * 3] | ((uint32_t)lab[(size_t)p * 3 + 1] << 8) | ((uint32_t)lab[(size_t)p * 3 + 2] << 16);
}

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

template <int CS>
__global__ void grid_count_kernel(const uint32_t *__restrict__ mask, const uint8_t *__restrict__ lab, int lw, int w, int n,
                                  int samples, int K, int *__restrict__ table)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t m = mask[((p / w) / samples) * lw + (p % w) / samples];
    constexpr int CELLS = (256 >> CS) * (256 >> CS) * (256 >> CS);
    const int c = cell_of<CS>(pack_lab(lab, p));
    for (int l = 0; l < K; ++l)
        if ((m >> l) & 1u) atomicAdd(&table[(size_t)l * CELLS + c], 1);
}

template <int CS>
__global__ void grid_fill_kernel(const uint32_t *__restrict__ mask, const uint8_t *__restrict__ lab, int lw, int w, int n,
                                 int samples, int K, const int *__restrict__ start, int *__restrict__ cursor,
                                 uint32_t *__restrict__ s_lab, int *__restrict__ s_id)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t m = mask[((p / w) / samples) * lw + (p % w) / samples];
    constexpr int CELLS = (256 >> CS) * (256 >> CS) * (256 >> CS);
    const uint32_t v = pack_lab(lab, p);
    const int c = cell_of<CS>(v);
    for (int l = 0; l < K; ++l)
        if ((m >> l) & 1u) {
            const size_t key = (size_t)l * CELLS + c;
            const int pos = start[key] + atomicAdd(&cursor[key], 1);
            s_lab[pos] = v;
            s_id[pos] = p;
        }
}

struct Top8 {
    unsigned long long key[8];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) key[i] = ~0ull;
    }
    __device__ __forceinline__ void insert(unsigned long long k)
    {
        if (k >= key[7]) return;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (key[i] == k) return;  // the same candidate reached through another cluster
#pragma unroll
        for (int i = 7; i > 0; --i) {
            if (key[i - 1] > k) key[i] = key[i - 1];
            else if (key[i] > k) { key[i] = k; k = ~0ull; }
        }
        if (k != ~0ull && key[0] > k) key[0] = k;
    }
};

__device__ __forceinline__ void scan_range(const uint32_t *__restrict__ s_lab, const int *__restrict__ s_id, int i0, int i1,
                                           uint32_t qlab, int qid, Top8 &top)
{
    for (int i = i0; i < i1; ++i) {
        const uint32_t c = __ldg(s_lab + i);
        const uint32_t ad = __vabsdiffu4(qlab, c);
        const uint32_t d2 = __dp4a(ad, ad, 0u);
        if (d2 <= (uint32_t)(top.key[7] >> 32)) {
            const int id = __ldg(s_id + i);
            if (id != qid) top.insert(((unsigned long long)d2 << 32) | (uint32_t)id);
        }
    }
}

template <int CS>
__global__ void __launch_bounds__(128) knn_grid_kernel(const uint32_t *__restrict__ mask, const uint8_t *__restrict__ lab,
                                                       const int *__restrict__ start, const uint32_t *__restrict__ s_lab,
                                                       const int *__restrict__ s_id, int lw, int w, int n, int samples, int K,
                                                       int *__restrict__ knn_id, double *__restrict__ knn_w,
                                                       const double *__restrict__ wtab)
{
    constexpr int CD = 256 >> CS, CELLS = CD * CD * CD;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t m = mask[((p / w) / samples) * lw + (p % w) / samples];
    const uint32_t q = pack_lab(lab, p);
    const int qc[3] = {(int)(q & 255), (int)((q >> 8) & 255), (int)((q >> 16) & 255)};
    const int c0 = qc[0] >> CS, c1 = qc[1] >> CS, c2 = qc[2] >> CS;
    Top8 top;
    top.init();
    for (int l = 0; l < K; ++l) {
        if (!((m >> l) & 1u)) continue;
        const int *st = start + (size_t)l * CELLS;
        const int cl_begin = st[0], cl_end = st[CELLS];
        if (cl_end - cl_begin <= 512) {  // small cluster: scan it
            scan_range(s_lab, s_id, cl_begin, cl_end, q, p, top);
            continue;
        }
        for (int R = 0;; ++R) {
            if (R > FALLBACK_R) {  // sparse outlier: bounded by the brute force over this cluster
                scan_range(s_lab, s_id, cl_begin, cl_end, q, p, top);
                break;
            }
            // visit the cells at Chebyshev distance exactly R
            const int lo2 = max(c2 - R, 0), hi2 = min(c2 + R, CD - 1);
            for (int d0 = -R; d0 <= R; ++d0) {
                const int x0 = c0 + d0;
                if (x0 < 0 || x0 >= CD) continue;
                for (int d1 = -R; d1 <= R; ++d1) {
                    const int x1 = c1 + d1;
                    if (x1 < 0 || x1 >= CD) continue;
                    const int row = (x0 * CD + x1) * CD;
                    if (max(abs(d0), abs(d1)) == R) {
                        scan_range(s_lab, s_id, st[row + lo2], st[row + hi2 + 1], q, p, top);
                    } else {
                        if (c2 - R >= 0) scan_range(s_lab, s_id, st[row + c2 - R], st[row + c2 - R + 1], q, p, top);
                        if (c2 + R < CD) scan_range(s_lab, s_id, st[row + c2 + R], st[row + c2 + R + 1], q, p, top);
                    }
                }
            }
            // lower bound on the distance of anything outside the visited box
            int lb = 1 << 20;
            bool covers_all = true;
            const int cc[3] = {c0, c1, c2};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                const int lo = (cc[ax] - R) << CS, hi = ((cc[ax] + R + 1) << CS) - 1;
                if (lo > 0) { lb = min(lb, qc[ax] - lo + 1); covers_all = false; }
                if (hi < 255) { lb = min(lb, hi + 1 - qc[ax]); covers_all = false; }
            }
            if (covers_all) break;
            const unsigned long long k8 = top.key[7];
            if (k8 != ~0ull && (uint32_t)(k8 >> 32) < (uint32_t)(lb * lb)) break;
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

template <int CS>
int run_find_knns(nct_ctx *ctx, const int *labels_dev, int lw, int lh, int K, const uint8_t *lab_dev, int h, int w, int samples,
                  int *knn_id_dev, double *knn_w_dev, const double *wtab)
{
    constexpr int CELLS = (256 >> CS) * (256 >> CS) * (256 >> CS);
    const int n = h * w;
    const size_t T = (size_t)K * CELLS;
    const size_t max_pairs = (size_t)n * (K < 5 ? K : 5);
    uint32_t *mask = (uint32_t *)nct_scratch(ctx, "knn_mask", sizeof(uint32_t) * (size_t)lw * lh);
    int *count = (int *)nct_scratch(ctx, "knng_count", sizeof(int) * (T + 1));
    int *start = (int *)nct_scratch(ctx, "knng_start", sizeof(int) * (T + 1));
    uint32_t *s_lab = (uint32_t *)nct_scratch(ctx, "knng_lab", sizeof(uint32_t) * max_pairs);
    int *s_id = (int *)nct_scratch(ctx, "knng_id", sizeof(int) * max_pairs);
    if (!mask || !count || !start || !s_lab || !s_id) return NCT_ERR_NOMEM;
    cell_masks_kernel<<<nct_div_up(lw * lh, 256), 256, 0, ctx->stream>>>(labels_dev, lw, lh, mask);
    NCT_CHECK_LAUNCH(ctx);
    NCT_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (T + 1), ctx->stream));
    grid_count_kernel<CS><<<nct_div_up(n, 256), 256, 0, ctx->stream>>>(mask, lab_dev, lw, w, n, samples, K, count);
    NCT_CHECK_LAUNCH(ctx);
    int rc = nct_exclusive_scan_i32(ctx, count, start, (int)T);
    if (rc) return rc;
    NCT_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (T + 1), ctx->stream));
    grid_fill_kernel<CS><<<nct_div_up(n, 256), 256, 0, ctx->stream>>>(mask, lab_dev, lw, w, n, samples, K, start, count, s_lab, s_id);
    NCT_CHECK_LAUNCH(ctx);
    knn_grid_kernel<CS><<<nct_div_up(n, 128), 128, 0, ctx->stream>>>(mask, lab_dev, start, s_lab, s_id, lw, w, n, samples, K, knn_id_dev, knn_w_dev, wtab);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

}  // namespace

extern "C" {

int nct_find_knns(nct_ctx *ctx, const int *labels_dev, int lw, int lh, int nlabels, const uint8_t *lab_dev, int h, int w,
                  int samples, int *knn_id_dev, double *knn_w_dev)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, labels_dev && lab_dev && knn_id_dev && knn_w_dev, "null pointer");
    NCT_REQUIRE(ctx, nlabels >= 1 && nlabels <= MAXK && samples >= 1, "bad cluster count / samples");
    NCT_REQUIRE(ctx, (long long)lw * samples >= w && (long long)lh * samples >= h, "label grid %dx%d x %d does not cover the %dx%d image", lw, lh, samples, w, h);
    const double *wtab = nct_knn_weight_table(ctx);
    if (!wtab) return NCT_ERR_NOMEM;
    const int n = h * w;
    // cell edge 2, 4, 8, 16 colour units (the result does not depend on it, only the search cost)
    static const int force_cs = getenv("NCT_KNN_CS") ? atoi(getenv("NCT_KNN_CS")) : 0;
    const int cs = force_cs ? force_cs : (n >= 300000 ? 1 : (n >= 60000 ? 2 : (n >= 10000 ? 3 : 4)));
    switch (cs) {
    case 1: return run_find_knns<1>(ctx, labels_dev, lw, lh, nlabels, lab_dev, h, w, samples, knn_id_dev, knn_w_dev, wtab);
    case 2: return run_find_knns<2>(ctx, labels_dev, lw, lh, nlabels, lab_dev, h, w, samples, knn_id_dev, knn_w_dev, wtab);
    case 3: return run_find_knns<3>(ctx, labels_dev, lw, lh, nlabels, lab_dev, h, w, samples, knn_id_dev, knn_w_dev, wtab);
    default: return run_find_knns<4>(ctx, labels_dev, lw, lh, nlabels, lab_dev, h, w, samples, knn_id_dev, knn_w_dev, wtab);
    }
}

}  // extern "C"
