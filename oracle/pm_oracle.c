/*
 * oracle/pm_oracle.c -- CPU restatement of the reference's dense-correspondence stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (neural-color-transfer_b200/)
 * may link, import or execute this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Shorthand: NCT/ = /root/reference/code/windows/neural_color_transfer/source/
 *
 * Parity status: the reference has NO golden vectors / known-answer tests for this
 * stage, and its PatchMatch kernel is racy (all 10 iterations in one
 * launch, neighbours' NNF entries read while being written, both __syncthreads()
 * commented out, NCT/GeneralizedPatchMatch.cu:801,828), so its output is not
 * run-to-run reproducible.  => "parity unpinned" for the PatchMatch iteration order.
 * What IS pinned here:
 *   - the distance function in the reference's own summation order
 *     (orc_dist_ref_chw) is checked against the reference's verbatim
 *     dist_compute_single compiled from /root/reference into oracle/_ref (when
 *     that build exists), see tests/test_oracle_ref.py;
 *   - XORWOW is checked against cuRAND's device API on the GPU (tests -m gpu).
 *
 * Spec decisions (each is a place where the reference is undefined/racy):
 *   D1  PatchMatch schedule = jump flooding with double buffering: every neighbour
 *       read in step (iter, jump) sees the NNF as of the end of the previous step;
 *       the own entry is carried through the four directions L,R,U,D; random search
 *       reads only own state.  (The reference does the same reads in-place, racily.)
 *   D2  Canonical FP32 reduction order for the patch distance: 32 accumulator slots,
 *       slot s takes float4-vector (s + 32k) of every valid patch pixel (C >= 128) or
 *       vector (s mod V) of the patch pixels whose index == s / V (mod 32/V) (C < 128),
 *       fmaf chains in (patch pixel, vector, x/y/z/w) order, then an XOR butterfly
 *       16,8,4,2,1.  The reference sums sequentially over (dy,dx,c) in one accumulator
 *       (NCT/GeneralizedPatchMatch.cu:366-387); a warp-parallel kernel cannot keep
 *       that order, so the oracle fixes this one.
 *   D3  Candidates equal to the entry's value at step start or to an earlier
 *       candidate of the same step are not re-evaluated (provably the same result:
 *       acceptance needs d < dbest strictly).  Both counts are reported.
 *   D4  Unchanged-source skip: a propagation candidate whose source entry has not
 *       changed since the same (query, jump, direction) slot last judged it is not
 *       re-evaluated (it would be rejected again); details at orc_patchmatch_opts.
 *       Like D3 it cannot change the field, only the number of evaluations.
 *   D5  L2 normalise: a pixel with zero norm yields zeros (reference: 0/0 = NaN,
 *       NCT/GeneralizedPatchMatch.cu:276-277).
 *   D6  NNF upsample uses fmaf for `ax + (bxh-axh)*ratio` (what nvcc's default
 *       -fmad=true emits for the reference's device code, :571-572).
 */
#include <math.h>
#include <float.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- NNF packing: NCT/GeneralizedPatchMatch.cu:24-34 ---- */
static inline uint32_t xy_to_int(int x, int y) { return ((uint32_t)y << 12) | (uint32_t)x; }
static inline int int_to_x(uint32_t v) { return (int)(v & 0xFFFu); }
static inline int int_to_y(uint32_t v) { return (int)((v >> 12) & 0xFFFu); }

/* ---- XORWOW: CUDA curand_kernel.h:800-822 (init, subsequence=offset=0),
 *      :863-874 (curand), curand_uniform.h:69-72 (uniform);
 *      used at NCT/GeneralizedPatchMatch.cu:54-66 with seed = column index ---- */
typedef struct { uint32_t d, v[5]; } xorwow_t;

static void xorwow_init(xorwow_t *st, unsigned long long seed)
{
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    st->d = 6615241u + t1 + t0;
    st->v[0] = 123456789u + t0;
    st->v[1] = 362436069u ^ t0;
    st->v[2] = 521288629u + t1;
    st->v[3] = 88675123u ^ t1;
    st->v[4] = 5783321u + t0;
}

static uint32_t xorwow_next(xorwow_t *st)
{
    uint32_t t = st->v[0] ^ (st->v[0] >> 2);
    st->v[0] = st->v[1];
    st->v[1] = st->v[2];
    st->v[2] = st->v[3];
    st->v[3] = st->v[4];
    st->v[4] = (st->v[4] ^ (st->v[4] << 4)) ^ (t ^ (t << 1));
    st->d += 362437u;
    return st->v[4] + st->d;
}

static inline float xorwow_uniform(uint32_t x)
{
    /* 2.3283064e-10f == 2^-32 exactly, so the product is exact and FMA
       contraction cannot change the result. */
    return fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

/* out[col*ndraws + k] = k-th curand_uniform() of a generator seeded with `col`. */
void orc_xorwow_uniform_table(int ncols, int ndraws, float *out)
{
    for (int c = 0; c < ncols; ++c) {
        xorwow_t st;
        xorwow_init(&st, (unsigned long long)c);
        for (int k = 0; k < ndraws; ++k) out[(size_t)c * ndraws + k] = xorwow_uniform(xorwow_next(&st));
    }
}

void orc_xorwow_raw(unsigned long long seed, int ndraws, uint32_t *out)
{
    xorwow_t st;
    xorwow_init(&st, seed);
    for (int k = 0; k < ndraws; ++k) out[k] = xorwow_next(&st);
}

/* ---- init_Ann_kernel: NCT/GeneralizedPatchMatch.cu:527-544 ---- */
void orc_nnf_init(int ah, int aw, int bh, int bw, uint32_t *ann)
{
    for (int ay = 0; ay < ah; ++ay)
        for (int ax = 0; ax < aw; ++ax) {
            float fx = (float)ax / (float)(aw - 1) * (float)(bw - 1);
            float fy = (float)ay / (float)(ah - 1) * (float)(bh - 1);
            int bx = (int)fx, by = (int)fy;
            if (bx > bw - 1) bx = bw - 1;
            if (by > bh - 1) by = bh - 1;
            ann[ay * aw + ax] = xy_to_int(bx, by);
        }
}

static inline int clampi(int x, int hi, int lo) { return x > hi ? hi : (x < lo ? lo : x); }

/* ---- upSample_kernel: NCT/GeneralizedPatchMatch.cu:546-580 (decision D6) ---- */
void orc_nnf_upsample(const uint32_t *ann_half, int ah_half, int aw_half,
                      int ah, int aw, int bh, int bw, uint32_t *ann)
{
    float rx = (float)aw / (float)aw_half;
    float ry = (float)ah / (float)ah_half;
    for (int ay = 0; ay < ah; ++ay)
        for (int ax = 0; ax < aw; ++ax) {
            int axh = (int)(((double)ax + 0.5) / (double)rx);
            int ayh = (int)(((double)ay + 0.5) / (double)ry);
            axh = clampi(axh, aw_half - 1, 0);
            ayh = clampi(ayh, ah_half - 1, 0);
            uint32_t v = ann_half[ayh * aw_half + axh];
            int bxh = int_to_x(v), byh = int_to_y(v);
            int bx = (int)((double)fmaf((float)(bxh - axh), rx, (float)ax) + 0.5);
            int by = (int)((double)fmaf((float)(byh - ayh), ry, (float)ay) + 0.5);
            bx = clampi(bx, bw - 1, 0);
            by = clampi(by, bh - 1, 0);
            ann[ay * aw + ax] = xy_to_int(bx, by);
        }
}

/* ---- canonical 32-slot reduction (decision D2) ---- */
static inline float butterfly32(const float *acc_in)
{
    float acc[32], tmp[32];
    memcpy(acc, acc_in, sizeof(acc));
    for (int off = 16; off >= 1; off >>= 1) {
        for (int s = 0; s < 32; ++s) tmp[s] = acc[s] + acc[s ^ off];
        memcpy(acc, tmp, sizeof(acc));
    }
    return acc[0];
}

/* accumulate one pixel pair (C floats each) into the 32 slots; `pi` = patch pixel index 0..8 */
static inline void accum_pixel(float *acc, const float *a, const float *b, int C, int pi)
{
    const int V = C / 4;
    if (V >= 32) {
        for (int k = 0; k < V / 32; ++k)
            for (int s = 0; s < 32; ++s) {
                const float *pa = a + (size_t)(s + 32 * k) * 4, *pb = b + (size_t)(s + 32 * k) * 4;
                float t = acc[s];
                t = fmaf(pa[0], pb[0], t);
                t = fmaf(pa[1], pb[1], t);
                t = fmaf(pa[2], pb[2], t);
                t = fmaf(pa[3], pb[3], t);
                acc[s] = t;
            }
    } else {
        const int groups = 32 / V;
        const int g = pi % groups;
        for (int j = 0; j < V; ++j) {
            const float *pa = a + j * 4, *pb = b + j * 4;
            float t = acc[g * V + j];
            t = fmaf(pa[0], pb[0], t);
            t = fmaf(pa[1], pb[1], t);
            t = fmaf(pa[2], pb[2], t);
            t = fmaf(pa[3], pb[3], t);
            acc[g * V + j] = t;
        }
    }
}

/* dist in the canonical order; features are pixel-major ("HWC"): f[(y*w + x)*C + c].
 * Semantics: NCT/GeneralizedPatchMatch.cu:355-405 with weight=1, cutoff handled by caller. */
static float dist_canon(const float *a, const float *b, int C, int ah, int aw, int bh, int bw,
                        int ax, int ay, int bx, int by)
{
    float acc[32];
    for (int s = 0; s < 32; ++s) acc[s] = 0.f;
    int n = 0, pi = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx, ++pi) {
            int yy = ay + dy, xx = ax + dx, v = by + dy, u = bx + dx;
            if (yy < ah && yy >= 0 && xx < aw && xx >= 0 && v < bh && v >= 0 && u < bw && u >= 0) {
                accum_pixel(acc, a + ((size_t)yy * aw + xx) * C, b + ((size_t)v * bw + u) * C, C, pi);
                n++;
            }
        }
    if (n == 0) return 1.f;
    float s = butterfly32(acc);
    return (-s) / (float)n;
}

float orc_dist_canon(const float *a, const float *b, int C, int ah, int aw, int bh, int bw,
                     int ax, int ay, int bx, int by)
{
    return dist_canon(a, b, C, ah, aw, bh, bw, ax, ay, bx, by);
}

/* dist in the REFERENCE's order and layout (planar CHW, one sequential accumulator, the
 * fma nvcc contracts `sum -= a*b` into): NCT/GeneralizedPatchMatch.cu:355-405.
 * Used to pin this restatement against the verbatim reference build (oracle/_ref)
 * and to bound |canonical - reference| in tests. */
float orc_dist_ref_chw(const float *a1, const float *b1, int C, int a_rows, int a_cols,
                       int b_rows, int b_cols, int ax, int ay, int bx, int by, int patch_w,
                       float cutoff, int use_fma)
{
    float pixel_sum1 = 0, pixel_no = 0, pixel_dist;
    int a_slice = a_rows * a_cols, b_slice = b_rows * b_cols;
    for (int dy = -patch_w / 2; dy <= patch_w / 2; dy++)
        for (int dx = -patch_w / 2; dx <= patch_w / 2; dx++) {
            if ((ay + dy) < a_rows && (ay + dy) >= 0 && (ax + dx) < a_cols && (ax + dx) >= 0 &&
                (by + dy) < b_rows && (by + dy) >= 0 && (bx + dx) < b_cols && (bx + dx) >= 0) {
                for (int dc = 0; dc < C; dc++) {
                    float av = a1[(size_t)dc * a_slice + (ay + dy) * a_cols + (ax + dx)];
                    float bv = b1[(size_t)dc * b_slice + (by + dy) * b_cols + (bx + dx)];
                    if (use_fma) pixel_sum1 = fmaf(-av, bv, pixel_sum1);
                    else { float t = av * bv; pixel_sum1 -= t; }
                }
                pixel_no += 1;
            }
        }
    if (pixel_no == 0) pixel_dist = 1;
    else pixel_dist = (0 + 1.0f * pixel_sum1) / pixel_no;
    return pixel_dist >= cutoff ? cutoff : pixel_dist;
}

/* ---- L2 normalise across channels, pixel-major layout (decision D5).
 *      Semantics: norm(), NCT/GeneralizedPatchMatch.cu:237-283 (response output unused) ---- */
void orc_l2norm_hwc(const float *src, float *dst, int npix, int C)
{
    const int V = C / 4;
#pragma omp parallel for schedule(static)
    for (int p = 0; p < npix; ++p) {
        const float *x = src + (size_t)p * C;
        float acc[32];
        for (int s = 0; s < 32; ++s) acc[s] = 0.f;
        for (int v = 0; v < V; ++v) {
            int s = v % 32;
            float t = acc[s];
            t = fmaf(x[v * 4 + 0], x[v * 4 + 0], t);
            t = fmaf(x[v * 4 + 1], x[v * 4 + 1], t);
            t = fmaf(x[v * 4 + 2], x[v * 4 + 2], t);
            t = fmaf(x[v * 4 + 3], x[v * 4 + 3], t);
            acc[s] = t;
        }
        float ss = butterfly32(acc);
        float nrm = sqrtf(ss);
        float *y = dst + (size_t)p * C;
        if (ss > 0.f) for (int c = 0; c < C; ++c) y[c] = x[c] / nrm;
        else for (int c = 0; c < C; ++c) y[c] = 0.f;
    }
}

/* ---- layout helpers ---- */
void orc_chw_to_hwc(const float *src, float *dst, int C, int npix)
{
#pragma omp parallel for schedule(static)
    for (int p = 0; p < npix; ++p)
        for (int c = 0; c < C; ++c) dst[(size_t)p * C + c] = src[(size_t)c * npix + p];
}

/* ---- deterministic PatchMatch (decisions D1-D3); reference semantics
 *      NCT/GeneralizedPatchMatch.cu:677-831.
 * params[11] = {C, ah, aw, bh, bw, patch(3), iters, rs_max, flag_constraint(0), 10, 1}
 *   (contract of NCT/main.cu:204-214).
 * stats[0] = evaluations the reference semantics perform (neighbour in A, candidate in B,
 *            plus the initial one and every random-search candidate);
 * stats[1] = evaluations left after D3 de-duplication;
 * stats[2] = evaluations left after D3 + D4 (what the GPU kernel computes).
 * D4 (unchanged-source skip): step t = 4 * iter + jump index reads the neighbour q's entry as of the end of step
 *   t - 1; the same (query, jump, direction) slot read q's entry as of the end of step t - 5 one iteration earlier.
 *   If q's entry did not change during steps t-4 .. t-1 the candidate is the one already judged then, its distance
 *   was >= the query's best at that time >= the best now, and acceptance is strict (<): it would be rejected again,
 *   so it is not evaluated.  The resulting field is identical with and without D4 (orc_patchmatch_opts(.., d4 = 0)
 *   switches it off; tests/test_oracle_pm.py compares the two).  ---- */
int orc_patchmatch_opts(const float *a, const float *b, uint32_t *ann, float *annd,
                        const int *params, long long *stats, int d4);

int orc_patchmatch(const float *a, const float *b, uint32_t *ann, float *annd,
                   const int *params, long long *stats)
{
    return orc_patchmatch_opts(a, b, ann, annd, params, stats, 1);
}

int orc_patchmatch_opts(const float *a, const float *b, uint32_t *ann, float *annd,
                        const int *params, long long *stats, int d4)
{
    const int C = params[0], ah = params[1], aw = params[2], bh = params[3], bw = params[4];
    const int patch_w = params[5], iters = params[6], rs_max = params[7];
    if (patch_w != 3 || params[8] != 0) return -1;
    if (iters > 31) return -3; /* D4 step indices are kept in a signed char */
    if (C % 4 != 0) return -2;
    if (C >= 128 ? (C % 128 != 0) : (C != 64 && C != 32 && C != 16)) return -2;
    const int n = ah * aw;
    long long ev_ref = 0, ev_dedup = 0, ev_d4 = 0;

    int rs_start = rs_max;
    if (rs_start > (bw > bh ? bw : bh)) rs_start = (bw > bh ? bw : bh);
    int n_mag = 0;
    for (int mag = rs_start; mag >= 1; mag /= 2) n_mag++;
    const int ndraws = 2 * n_mag * iters;
    float *rng = (float *)malloc(sizeof(float) * (size_t)aw * (ndraws > 0 ? ndraws : 1));
    orc_xorwow_uniform_table(aw, ndraws, rng);

    uint32_t *prev = (uint32_t *)malloc(sizeof(uint32_t) * n);
    /* D4 bookkeeping: step of the last change of every entry (-1 = never), double buffered like the field */
    signed char *lc = (signed char *)malloc((size_t)n), *lc_prev = (signed char *)malloc((size_t)n);
    memset(lc, 0xff, (size_t)n);

    /* initial distance (:710-712) */
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ev_ref, ev_dedup, ev_d4)
    for (int p = 0; p < n; ++p) {
        int ax = p % aw, ay = p / aw;
        uint32_t v = ann[p];
        annd[p] = dist_canon(a, b, C, ah, aw, bh, bw, ax, ay, int_to_x(v), int_to_y(v));
        ev_ref++;
        ev_dedup++;
        ev_d4++;
    }

    for (int iter = 0; iter < iters; ++iter) {
        for (int jump = 8; jump > 0; jump /= 2) {
            memcpy(prev, ann, sizeof(uint32_t) * n);
            memcpy(lc_prev, lc, (size_t)n);
            const int t = 4 * iter + (jump == 8 ? 0 : jump == 4 ? 1 : jump == 2 ? 2 : 3);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ev_ref, ev_dedup, ev_d4)
            for (int p = 0; p < n; ++p) {
                int ax = p % aw, ay = p / aw;
                uint32_t v0 = prev[p];
                int xbest = int_to_x(v0), ybest = int_to_y(v0);
                float dbest = annd[p];
                uint32_t seen[5], seen4[5];
                int nseen = 0, nseen4 = 0;
                seen[nseen++] = v0;
                seen4[nseen4++] = v0;
                /* L, R, U, D (:725-798) */
                const int qx[4] = {ax - jump, ax + jump, ax, ax};
                const int qy[4] = {ay, ay, ay - jump, ay + jump};
                const int sx[4] = {jump, -jump, 0, 0};
                const int sy[4] = {0, 0, jump, -jump};
                for (int k = 0; k < 4; ++k) {
                    if (qx[k] < 0 || qx[k] >= aw || qy[k] < 0 || qy[k] >= ah) continue;
                    uint32_t vp = prev[qy[k] * aw + qx[k]];
                    int xp = int_to_x(vp) + sx[k], yp = int_to_y(vp) + sy[k];
                    if (!(yp >= 0 && yp < bh && xp >= 0 && xp < bw)) continue;
                    ev_ref++;
                    uint32_t cv = xy_to_int(xp, yp);
                    int dup = 0, dup4 = 0;
                    for (int i = 0; i < nseen; ++i) dup |= (seen[i] == cv);
                    for (int i = 0; i < nseen4; ++i) dup4 |= (seen4[i] == cv);
                    if (!dup) { seen[nseen++] = cv; ev_dedup++; }
                    if (d4) {   /* the GPU's rule: D3 against the candidates it evaluated, then D4 */
                        if (dup4) continue;
                        if (t >= 4 && lc_prev[qy[k] * aw + qx[k]] <= t - 5) continue;
                        seen4[nseen4++] = cv;
                    } else if (dup) continue;
                    ev_d4++;
                    float d = dist_canon(a, b, C, ah, aw, bh, bw, ax, ay, xp, yp);
                    if (d < dbest) { xbest = xp; ybest = yp; dbest = d; }
                }
                if (jump == 1) {
                    /* random search (:806-821) */
                    const float *u = rng + (size_t)ax * ndraws + (size_t)iter * 2 * n_mag;
                    int m = 0;
                    for (int mag = rs_start; mag >= 1; mag /= 2, ++m) {
                        int xmin = xbest - mag > 0 ? xbest - mag : 0;
                        int xmax = xbest + mag + 1 < bw ? xbest + mag + 1 : bw;
                        int ymin = ybest - mag > 0 ? ybest - mag : 0;
                        int ymax = ybest + mag + 1 < bh ? ybest + mag + 1 : bh;
                        int xp = xmin + (int)(u[2 * m] * (float)(xmax - xmin)) % (xmax - xmin);
                        int yp = ymin + (int)(u[2 * m + 1] * (float)(ymax - ymin)) % (ymax - ymin);
                        ev_ref++;
                        if (xp == xbest && yp == ybest) continue; /* D3: d == dbest, never accepted */
                        ev_dedup++;
                        ev_d4++;
                        float d = dist_canon(a, b, C, ah, aw, bh, bw, ax, ay, xp, yp);
                        if (d + FLT_MIN < dbest) { xbest = xp; ybest = yp; dbest = d; }
                    }
                }
                ann[p] = xy_to_int(xbest, ybest);
                annd[p] = dbest;
                if (ann[p] != v0) lc[p] = (signed char)t;
            }
        }
    }
    free(prev);
    free(lc);
    free(lc_prev);
    free(rng);
    if (stats) { stats[0] = ev_ref; stats[1] = ev_dedup; stats[2] = ev_d4; }
    return 0;
}

/* ---- reference-SEMANTICS PatchMatch on the CPU: same in-place reads as the reference
 *      kernel, pixels visited in raster order (one legal serialisation of the racy
 *      kernel), the reference's own summation order and planar layout.  This is the
 *      "port" CPU baseline of bench.py; it is NOT the bit-exact target. ---- */
int orc_patchmatch_ref_serial(const float *a1, const float *b1, uint32_t *ann, float *annd,
                              const int *params)
{
    const int ch = params[0], a_rows = params[1], a_cols = params[2], b_rows = params[3], b_cols = params[4];
    const int patch_w = params[5], pm_iters = params[6], rs_max = params[7];
    const int n = a_rows * a_cols;
    xorwow_t *states = (xorwow_t *)malloc(sizeof(xorwow_t) * a_cols);
    for (int x = 0; x < a_cols; ++x) xorwow_init(&states[x], (unsigned long long)x);
    /* every row of a column draws the same sequence: keep a per-pixel copy of the state */
    xorwow_t *pst = (xorwow_t *)malloc(sizeof(xorwow_t) * n);
    for (int p = 0; p < n; ++p) pst[p] = states[p % a_cols];
    free(states);

#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < n; ++p) {
        int ax = p % a_cols, ay = p / a_cols;
        uint32_t v = ann[p];
        annd[p] = orc_dist_ref_chw(a1, b1, ch, a_rows, a_cols, b_rows, b_cols, ax, ay, int_to_x(v), int_to_y(v),
                                   patch_w, (float)INT_MAX, 1);
    }
    for (int iter = 0; iter < pm_iters; ++iter) {
        /* rows are processed in parallel bands; within a band raster order. Reads of other
           bands' entries are racy exactly like the reference's. */
#pragma omp parallel for schedule(static)
        for (int ay = 0; ay < a_rows; ++ay)
            for (int ax = 0; ax < a_cols; ++ax) {
                int p = ay * a_cols + ax;
                uint32_t v = ann[p];
                int xbest = int_to_x(v), ybest = int_to_y(v);
                float dbest = annd[p];
                for (int jump = 8; jump > 0; jump /= 2) {
                    const int qx[4] = {ax - jump, ax + jump, ax, ax};
                    const int qy[4] = {ay, ay, ay - jump, ay + jump};
                    const int sx[4] = {jump, -jump, 0, 0};
                    const int sy[4] = {0, 0, jump, -jump};
                    for (int k = 0; k < 4; ++k) {
                        if (qx[k] < 0 || qx[k] >= a_cols || qy[k] < 0 || qy[k] >= a_rows) continue;
                        uint32_t vp = ann[qy[k] * a_cols + qx[k]];
                        int xp = int_to_x(vp) + sx[k], yp = int_to_y(vp) + sy[k];
                        if (!(yp >= 0 && yp < b_rows && xp >= 0 && xp < b_cols)) continue;
                        float d = orc_dist_ref_chw(a1, b1, ch, a_rows, a_cols, b_rows, b_cols, ax, ay, xp, yp,
                                                   patch_w, dbest, 1);
                        if (d < dbest) { xbest = xp; ybest = yp; dbest = d; }
                        ann[p] = xy_to_int(xbest, ybest);
                        annd[p] = dbest;
                    }
                }
                int rs_start = rs_max;
                if (rs_start > (b_cols > b_rows ? b_cols : b_rows)) rs_start = (b_cols > b_rows ? b_cols : b_rows);
                for (int mag = rs_start; mag >= 1; mag /= 2) {
                    int xmin = xbest - mag > 0 ? xbest - mag : 0;
                    int xmax = xbest + mag + 1 < b_cols ? xbest + mag + 1 : b_cols;
                    int ymin = ybest - mag > 0 ? ybest - mag : 0;
                    int ymax = ybest + mag + 1 < b_rows ? ybest + mag + 1 : b_rows;
                    float u1 = xorwow_uniform(xorwow_next(&pst[p]));
                    float u2 = xorwow_uniform(xorwow_next(&pst[p]));
                    int xp = xmin + (int)(u1 * (float)(xmax - xmin)) % (xmax - xmin);
                    int yp = ymin + (int)(u2 * (float)(ymax - ymin)) % (ymax - ymin);
                    float d = orc_dist_ref_chw(a1, b1, ch, a_rows, a_cols, b_rows, b_cols, ax, ay, xp, yp,
                                               patch_w, dbest, 1);
                    if (d + FLT_MIN < dbest) { xbest = xp; ybest = yp; dbest = d; }
                }
                ann[p] = xy_to_int(xbest, ybest);
                annd[p] = dbest;
            }
    }
    free(pst);
    return 0;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
