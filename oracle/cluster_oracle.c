/*
 * oracle/cluster_oracle.c -- CPU restatement of the reference's feature clustering and in-cluster 8-NN search.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Follows
 *   ColorTransfer::clusterFeastures      CT/ColorTransfer.cpp:355-395
 *   cvflann KMeansIndex (root split)     CT/Flann/kmeans_index.h:108-137 (chooseCentersRandom), 368-381 (buildIndex),
 *                                        700-880 (computeClustering), 487-540 + 1067-1105 (the 10-way root split is the
 *                                        result: getMinVarianceClusters splits the root once and stops)
 *   UniqueRandom                         CT/Flann/random.h:103-140 (std::random_shuffle after srand(1))
 *   L2 distance                          CT/Flann/dist.h:136-181 (float accumulator, groups of four)
 *   ColorTransfer::findKnns              CT/ColorTransfer.cpp:397-423 -> getClusters :273-353, insertClusterPixel :255-271,
 *                                        findSubKNNs :136-195 (nanoflann, 9-NN then drop self),
 *                                        sortMergeComputeWeight :60-110 (sort (dist,id), unique, keep 8, w = exp(1 - d/3))
 *
 * Parity status: "parity unpinned" -- the reference has no test for this stage, and its results depend on the MSVC
 * C runtime (rand / random_shuffle), on /fp:fast float summation and on nanoflann's traversal order for ties.
 *
 * Spec decisions
 *   K1  rand() is the MSVC LCG (seed*214013+2531011, (seed>>16)&0x7fff) and random_shuffle the VS2013 algorithm
 *       (15 random bits per draw, widened while RAND_MAX < index); both restated from public knowledge of the MSVC CRT,
 *       not from the reference tree.
 *   K2  L2 distances: float accumulator, `result += d0*d0 + d1*d1 + d2*d2 + d3*d3` per group of four, left to right,
 *       no contraction (the reference compiles with /fp:fast, which leaves the order unspecified).
 *   K3  features handed to k-means are the L2-normalised conv5_1 rows produced by the canonical norm (pm_oracle D2/D5);
 *       the reference normalises them with a sequential host loop (NCT/main.cu:145-165).
 *   K4  8-NN ranking uses the exact integer squared distance of the 8-bit Lab triples, ties broken by the smaller pixel id
 *       (the reference: double Euclidean distance of u8/255 values, ties by KD-tree traversal order of a shuffled point
 *       set).  The stored weight is exp(1 - d/3) with d = sqrt(D2)/255 in double.
 *   K5  a pixel with fewer than 8 candidates gets id = -1, w = 0 padding (the reference: vector::resize(k) with the
 *       default NN(), CT/ColorTransfer.cpp:108).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- K1: MSVC rand / random_shuffle ---- */
static uint32_t g_seed = 1;
static void msvc_srand(uint32_t s) { g_seed = s; }
static int msvc_rand(void)
{
    g_seed = g_seed * 214013u + 2531011u;
    return (int)((g_seed >> 16) & 0x7fff);
}
static void msvc_random_shuffle(int *v, int n)
{
    const unsigned long RBITS = 15, RMAX = (1UL << 15) - 1;
    for (unsigned long index = 2; (long)index <= n; ++index) {
        unsigned long rm = RMAX;
        unsigned long rn = (unsigned long)msvc_rand() & RMAX;
        for (; rm < index && rm != ~0UL; rm = rm << RBITS | RMAX) rn = rn << RBITS | ((unsigned long)msvc_rand() & RMAX);
        unsigned long off = rn % index;
        int t = v[index - 1];
        v[index - 1] = v[off];
        v[off] = t;
    }
}

void orc_msvc_shuffle(int n, int *out)
{
    msvc_srand(1);
    for (int i = 0; i < n; ++i) out[i] = i;
    msvc_random_shuffle(out, n);
}

/* ---- K2: cvflann L2 ---- */
static float l2_ff(const float *a, const float *b, int size)
{
    float result = 0.f;
    int i = 0;
    for (; i + 3 < size; i += 4) {
        float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < size; ++i) { float d0 = a[i] - b[i]; result += d0 * d0; }
    return result;
}
static float l2_fd(const float *a, const double *b, int size)
{
    float result = 0.f;
    int i = 0;
    for (; i + 3 < size; i += 4) {
        float d0 = (float)(a[i] - b[i]), d1 = (float)(a[i + 1] - b[i + 1]), d2 = (float)(a[i + 2] - b[i + 2]),
              d3 = (float)(a[i + 3] - b[i + 3]);
        result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    for (; i < size; ++i) { float d0 = (float)(a[i] - b[i]); result += d0 * d0; }
    return result;
}

/* Root-level k-means of the KMeansIndex (branching = k, `iterations` Lloyd steps, random distinct initial centres).
 * features: n x dim floats.  labels_out[n].  Returns the number of clusters (k, or 1 if the root cannot be split). */
int orc_kmeans_labels(const float *features, int n, int dim, int k, int iterations, int *labels_out, int *centers_idx_out)
{
    for (int i = 0; i < n; ++i) labels_out[i] = 0;
    if (n < k) return 1;
    msvc_srand(1);
    int *vals = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) vals[i] = i;
    msvc_random_shuffle(vals, n);
    int *centers = (int *)malloc(sizeof(int) * (size_t)k);
    int counter = 0, index;
    for (index = 0; index < k; ++index) {
        int duplicate = 1;
        while (duplicate) {
            duplicate = 0;
            if (counter == n) goto done_centers;
            int rnd = vals[counter++];
            centers[index] = rnd;
            for (int j = 0; j < index; ++j)
                if (l2_ff(features + (size_t)centers[index] * dim, features + (size_t)centers[j] * dim, dim) < 1e-16f) duplicate = 1;
        }
    }
done_centers:
    free(vals);
    if (index < k) { free(centers); return 1; }
    if (centers_idx_out) memcpy(centers_idx_out, centers, sizeof(int) * (size_t)k);

    double *dc = (double *)malloc(sizeof(double) * (size_t)k * dim);
    for (int i = 0; i < k; ++i)
        for (int d = 0; d < dim; ++d) dc[(size_t)i * dim + d] = (double)features[(size_t)centers[i] * dim + d];
    free(centers);
    float *radiuses = (float *)calloc((size_t)k, sizeof(float));
    int *count = (int *)calloc((size_t)k, sizeof(int));
    int *belongs = labels_out;
    for (int i = 0; i < n; ++i) {
        float sq = l2_fd(features + (size_t)i * dim, dc, dim);
        belongs[i] = 0;
        for (int j = 1; j < k; ++j) {
            float nsq = l2_fd(features + (size_t)i * dim, dc + (size_t)j * dim, dim);
            if (sq > nsq) { belongs[i] = j; sq = nsq; }
        }
        if (sq > radiuses[belongs[i]]) radiuses[belongs[i]] = sq;
        count[belongs[i]]++;
    }
    int converged = 0, iteration = 0;
    while (!converged && iteration < iterations) {
        converged = 1;
        iteration++;
        memset(dc, 0, sizeof(double) * (size_t)k * dim);
        for (int i = 0; i < k; ++i) radiuses[i] = 0;
        for (int i = 0; i < n; ++i) {
            double *c = dc + (size_t)belongs[i] * dim;
            const float *v = features + (size_t)i * dim;
            for (int d = 0; d < dim; ++d) c[d] += v[d];
        }
        for (int i = 0; i < k; ++i)
            for (int d = 0; d < dim; ++d) dc[(size_t)i * dim + d] /= count[i];
        for (int i = 0; i < n; ++i) {
            float sq = l2_fd(features + (size_t)i * dim, dc, dim);
            int nc = 0;
            for (int j = 1; j < k; ++j) {
                float nsq = l2_fd(features + (size_t)i * dim, dc + (size_t)j * dim, dim);
                if (sq > nsq) { nc = j; sq = nsq; }
            }
            if (sq > radiuses[nc]) radiuses[nc] = sq;
            if (nc != belongs[i]) {
                count[belongs[i]]--;
                count[nc]++;
                belongs[i] = nc;
                converged = 0;
            }
        }
        for (int i = 0; i < k; ++i) {
            if (count[i] == 0) {
                int j = (i + 1) % k;
                while (count[j] <= 1) j = (j + 1) % k;
                for (int q = 0; q < n; ++q) {
                    if (belongs[q] == j && l2_fd(features + (size_t)q * dim, dc + (size_t)j * dim, dim) == radiuses[j]) {
                        belongs[q] = i;
                        count[j]--;
                        count[i]++;
                        break;
                    }
                }
                converged = 0;
            }
        }
    }
    free(dc); free(radiuses); free(count);
    return k;
}

/* cluster membership bit mask per label cell: own label + the labels of the 4-neighbours (getClusters :288-315) */
void orc_cluster_masks(const int *labels, int lw, int lh, uint32_t *mask)
{
    for (int y = 0; y < lh; ++y)
        for (int x = 0; x < lw; ++x) {
            int id = y * lw + x;
            uint32_t m = 1u << labels[id];
            if (x < lw - 1) m |= 1u << labels[id + 1];
            if (x > 0) m |= 1u << labels[id - 1];
            if (y < lh - 1) m |= 1u << labels[id + lw];
            if (y > 0) m |= 1u << labels[id - lw];
            mask[id] = m;
        }
}

typedef struct { uint64_t key[8]; int n; } top8_t;
static inline void top8_insert(top8_t *t, uint64_t key)
{
    int n = t->n;
    if (n == 8 && key >= t->key[7]) return;
    for (int i = 0; i < n; ++i) if (t->key[i] == key) return;
    int pos = n < 8 ? n : 7;
    while (pos > 0 && t->key[pos - 1] > key) { t->key[pos] = t->key[pos - 1]; pos--; }
    t->key[pos] = key;
    if (n < 8) t->n = n + 1;
}

typedef struct { int L, id; uint32_t lab; } member_t;
static int cmp_member(const void *a, const void *b)
{
    const member_t *x = (const member_t *)a, *y = (const member_t *)b;
    if (x->L != y->L) return x->L - y->L;
    return x->id - y->id;
}

/* findKnns: labels on the lw x lh grid (k-means of the coarsest level); lab = level-size 8-bit Lab image (h x w x 3);
 * cell (cx, cy) covers pixels [cx*samples, min((cx+1)*samples, w)) x [...] (insertClusterPixel).
 * Output knn_id[n][8] (pixel ids, -1 = none), knn_w[n][8] = exp(1 - d/3). */
void orc_find_knns(const int *labels, int lw, int lh, int nlabels, const uint8_t *lab, int h, int w, int samples, int *knn_id,
                   double *knn_w)
{
    const int n = h * w;
    uint32_t *mask = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)lw * lh);
    orc_cluster_masks(labels, lw, lh, mask);
    top8_t *top = (top8_t *)calloc((size_t)n, sizeof(top8_t));
    member_t *mem = (member_t *)malloc(sizeof(member_t) * (size_t)n);
    for (int l = 0; l < nlabels; ++l) {
        int cnt = 0;
        for (int p = 0; p < n; ++p) {
            int x = p % w, y = p / w;
            int cx = x / samples, cy = y / samples;
            if (cx >= lw || cy >= lh) continue;
            if (mask[cy * lw + cx] & (1u << l)) {
                mem[cnt].L = lab[(size_t)p * 3];
                mem[cnt].id = p;
                mem[cnt].lab = (uint32_t)lab[(size_t)p * 3] | ((uint32_t)lab[(size_t)p * 3 + 1] << 8) | ((uint32_t)lab[(size_t)p * 3 + 2] << 16);
                cnt++;
            }
        }
        qsort(mem, (size_t)cnt, sizeof(member_t), cmp_member);
#pragma omp parallel for schedule(dynamic, 256)
        for (int i = 0; i < cnt; ++i) {
            top8_t *t = &top[mem[i].id];
            const int qL = mem[i].L, qa = (mem[i].lab >> 8) & 255, qb = (mem[i].lab >> 16) & 255;
            /* sweep outwards in L; stop a side when dL^2 exceeds the current 8th best */
            int lo = i - 1, hi = i + 1;
            while (lo >= 0 || hi < cnt) {
                uint64_t worst = t->n == 8 ? (t->key[7] >> 32) : UINT64_MAX;
                int go_lo = lo >= 0, go_hi = hi < cnt;
                if (go_lo) { int d = qL - mem[lo].L; if ((uint64_t)(d * d) > worst) { lo = -1; go_lo = 0; } }
                if (go_hi) { int d = mem[hi].L - qL; if ((uint64_t)(d * d) > worst) { hi = cnt; go_hi = 0; } }
                if (!go_lo && !go_hi) break;
                int j = (go_lo && (!go_hi || (qL - mem[lo].L) <= (mem[hi].L - qL))) ? lo-- : hi++;
                int dL = qL - mem[j].L, da = qa - (int)((mem[j].lab >> 8) & 255), db = qb - (int)((mem[j].lab >> 16) & 255);
                uint64_t d2 = (uint64_t)(dL * dL + da * da + db * db);
                top8_insert(t, (d2 << 32) | (uint32_t)mem[j].id);
            }
        }
    }
    for (int p = 0; p < n; ++p)
        for (int k = 0; k < 8; ++k) {
            if (k < top[p].n) {
                uint64_t key = top[p].key[k];
                double d = sqrt((double)(key >> 32)) / 255.0;
                knn_id[(size_t)p * 8 + k] = (int)(key & 0xffffffffu);
                knn_w[(size_t)p * 8 + k] = exp(1.0 - d / 3.0);
            } else {
                knn_id[(size_t)p * 8 + k] = -1;
                knn_w[(size_t)p * 8 + k] = 0.0;
            }
        }
    free(mask); free(top); free(mem);
}

/* brute-force version of the same specification (small inputs only; used to validate the sweep above) */
void orc_find_knns_brute(const int *labels, int lw, int lh, int nlabels, const uint8_t *lab, int h, int w, int samples,
                         int *knn_id, double *knn_w)
{
    const int n = h * w;
    (void)nlabels;
    uint32_t *mask = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)lw * lh);
    orc_cluster_masks(labels, lw, lh, mask);
    for (int p = 0; p < n; ++p) {
        top8_t t;
        t.n = 0;
        uint32_t mp = mask[(p / w / samples) * lw + (p % w) / samples];
        for (int q = 0; q < n; ++q) {
            if (q == p) continue;
            uint32_t mq = mask[(q / w / samples) * lw + (q % w) / samples];
            if (!(mp & mq)) continue;
            int dL = lab[(size_t)p * 3] - lab[(size_t)q * 3], da = lab[(size_t)p * 3 + 1] - lab[(size_t)q * 3 + 1],
                db = lab[(size_t)p * 3 + 2] - lab[(size_t)q * 3 + 2];
            uint64_t d2 = (uint64_t)(dL * dL + da * da + db * db);
            top8_insert(&t, (d2 << 32) | (uint32_t)q);
        }
        for (int k = 0; k < 8; ++k) {
            if (k < t.n) {
                double d = sqrt((double)(t.key[k] >> 32)) / 255.0;
                knn_id[(size_t)p * 8 + k] = (int)(t.key[k] & 0xffffffffu);
                knn_w[(size_t)p * 8 + k] = exp(1.0 - d / 3.0);
            } else {
                knn_id[(size_t)p * 8 + k] = -1;
                knn_w[(size_t)p * 8 + k] = 0.0;
            }
        }
    }
    free(mask);
}
