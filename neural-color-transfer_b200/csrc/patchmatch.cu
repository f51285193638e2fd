// Dense correspondence kernels: NNF init / upsample, XORWOW table, PatchMatch.
//
// Replaces init_Ann_kernel, upSample_kernel and patchmatch_single of
// NCT/GeneralizedPatchMatch.cu:527-580, 677-831 (launched from NCT/main.cu:230-284).
//
// Design:
//   * features are pixel-major (HWC) FP32, so one candidate patch pixel is one contiguous
//     C*4-byte row: a warp reads it with coalesced LDG.128 (32 lanes x float4 = 512 B / instr);
//   * one warp per query pixel; the query's own 3x3xC patch lives in registers (C <= 256) for
//     the whole step, candidates' patches are streamed from L2/HBM with all loads of a
//     candidate in flight at once; the distance is a 32-lane FMA chain + XOR-shuffle butterfly;
//   * the reference runs 10 iterations in ONE racy launch; here each (iter, jump) step is a
//     launch reading the previous step's NNF (double buffered) => deterministic and bit-exact
//     against oracle/pm_oracle.c.  Random search is fused into the jump==1 step and the initial
//     distance into the first step, both directions (A->B, B->A) share each launch.
//   * per-column XORWOW streams are materialised once per call into a small table.
#include "nct_internal.h"
#include <cfloat>
#include <cuda_fp16.h>

namespace {

__host__ __device__ __forceinline__ uint32_t xy_to_int(int x, int y) { return ((uint32_t)y << 12) | (uint32_t)x; }
__host__ __device__ __forceinline__ int int_to_x(uint32_t v) { return (int)(v & 0xFFFu); }
__host__ __device__ __forceinline__ int int_to_y(uint32_t v) { return (int)((v >> 12) & 0xFFFu); }

// ------------------------------------------------------------------ NNF init / upsample
__global__ void nnf_init_kernel(uint32_t *__restrict__ ann, int ah, int aw, int bh, int bw)
{
    int ax = blockIdx.x * blockDim.x + threadIdx.x;
    int ay = blockIdx.y * blockDim.y + threadIdx.y;
    if (ax < aw && ay < ah) {
        float fx = __fmul_rn(__fdiv_rn((float)ax, (float)(aw - 1)), (float)(bw - 1));
        float fy = __fmul_rn(__fdiv_rn((float)ay, (float)(ah - 1)), (float)(bh - 1));
        int bx = min((int)fx, bw - 1);
        int by = min((int)fy, bh - 1);
        ann[ay * aw + ax] = xy_to_int(bx, by);
    }
}

__device__ __forceinline__ int clampi(int x, int hi, int lo) { return x > hi ? hi : (x < lo ? lo : x); }

__global__ void nnf_upsample_kernel(const uint32_t *__restrict__ ann_half, int ah_half, int aw_half,
                                    uint32_t *__restrict__ ann, int ah, int aw, int bh, int bw)
{
    int ax = blockIdx.x * blockDim.x + threadIdx.x;
    int ay = blockIdx.y * blockDim.y + threadIdx.y;
    if (ax >= aw || ay >= ah) return;
    float rx = __fdiv_rn((float)aw, (float)aw_half);
    float ry = __fdiv_rn((float)ah, (float)ah_half);
    int axh = (int)(__ddiv_rn((double)ax + 0.5, (double)rx));
    int ayh = (int)(__ddiv_rn((double)ay + 0.5, (double)ry));
    axh = clampi(axh, aw_half - 1, 0);
    ayh = clampi(ayh, ah_half - 1, 0);
    uint32_t v = ann_half[ayh * aw_half + axh];
    int bxh = int_to_x(v), byh = int_to_y(v);
    int bx = (int)((double)__fmaf_rn((float)(bxh - axh), rx, (float)ax) + 0.5);
    int by = (int)((double)__fmaf_rn((float)(byh - ayh), ry, (float)ay) + 0.5);
    bx = clampi(bx, bw - 1, 0);
    by = clampi(by, bh - 1, 0);
    ann[ay * aw + ax] = xy_to_int(bx, by);
}

// ------------------------------------------------------------------ XORWOW table
// curand_init(seed = column, 0, 0) + curand_uniform, restated from CUDA's curand_kernel.h
// (no skip-ahead is needed for subsequence = offset = 0).
__global__ void xorwow_table_kernel(float *__restrict__ out, int ncols, int ndraws)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    unsigned long long seed = (unsigned long long)col;
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    uint32_t d = 6615241u + t1 + t0;
    uint32_t v0 = 123456789u + t0, v1 = 362436069u ^ t0, v2 = 521288629u + t1, v3 = 88675123u ^ t1,
             v4 = 5783321u + t0;
    for (int k = 0; k < ndraws; ++k) {
        uint32_t t = v0 ^ (v0 >> 2);
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        uint32_t x = v4 + d;
        out[(size_t)col * ndraws + k] = __fmaf_rn((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    }
}

// ------------------------------------------------------------------ PatchMatch step
// tuning knobs (overridable with -D for experiments; defaults = measured best, profiles/r1_pm_tuning.md)
#ifndef PM_A_REGS_MAXC
#define PM_A_REGS_MAXC 256   // keep the query patch in registers up to this channel count
#endif
#ifndef PM_BATCH_SMALL
#define PM_BATCH_SMALL 1     // propagation candidates whose row loads are issued together for C <= 64 (C == 128: half)
#endif
#ifndef PM_TPB
#define PM_TPB 128
#endif
#ifndef PM_MINB_SMALL
#define PM_MINB_SMALL 6      // __launch_bounds__ min CTAs/SM for C <= 128 (caps registers at 85)
#endif

struct PMDir {
    const void *a;           // query features   [ah][aw][C], FP32 or (FP16 feature store) FP16
    const void *b;           // target features  [bh][bw][C]
    const uint32_t *nnf_in;  // NNF at the end of the previous step
    uint32_t *nnf_out;
    float *nnd;              // own entry only: read + write in place
    const int8_t *lc_in;     // step of the last change of every entry at the end of the previous step (-1 = never)
    int8_t *lc_out;
    const float *rng;        // [aw][ndraws]
    int ah, aw, bh, bw;
    int rs_start, n_mag, ndraws;
};

struct PMStep {
    PMDir d[2];
    int nq0;        // number of queries of direction 0
    int nq_total;   // queries of both directions
    int jump;
    int iter;
    int t;          // step index 4 * iter + jump index (D4 bookkeeping)
    int first;      // compute the initial distance instead of reading nnd
    int do_random;  // jump == 1: fused random search
    unsigned long long *counters;  // nullptr or 2 x u64 {evaluated, reference-semantics}
};

template <int C>
struct PMTraits {
    static constexpr int V = C / 4;                        // float4 vectors per pixel
    static constexpr int VPL = (V >= 32) ? V / 32 : 1;     // vectors per lane per pixel
    static constexpr int GROUPS = (V >= 32) ? 1 : 32 / V;  // patch pixels processed side by side
    static constexpr int PPL = (9 + GROUPS - 1) / GROUPS;  // patch pixels per lane
    static constexpr bool A_IN_REGS = (C <= PM_A_REGS_MAXC);
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// FP16 feature store: four halves (8 bytes) -> four floats, exactly (every FP16 value is an FP32 value), so the products and
// sums below are those of the oracle run on the FP16-rounded volumes
__device__ __forceinline__ float4 ldg4(const __half *p)
{
    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&r.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float butterfly(float acc)
{
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 1));
    return acc;
}

__device__ __forceinline__ float fma4(float4 a, float4 b, float acc)
{
    acc = __fmaf_rn(a.x, b.x, acc);
    acc = __fmaf_rn(a.y, b.y, acc);
    acc = __fmaf_rn(a.z, b.z, acc);
    acc = __fmaf_rn(a.w, b.w, acc);
    return acc;
}

// 9-bit mask of patch pixels (dy outer, dx inner) that lie inside a h x w image around (x, y)
__device__ __forceinline__ unsigned patch_mask(int x, int y, int w, int h)
{
    unsigned mx = (x > 0 ? 1u : 0u) | 2u | (x + 1 < w ? 4u : 0u);
    unsigned m = 0;
    if (y > 0) m |= mx;
    m |= mx << 3;
    if (y + 1 < h) m |= mx << 6;
    return m;
}

template <int C>
struct QueryPatch {
    using T = PMTraits<C>;
    float4 a[T::A_IN_REGS ? T::PPL * T::VPL : 1];
    const float *a_base;  // address of the query pixel's own feature row (+ lane offset)
    int aw;
    unsigned amask;
};

// loads the query patch into registers (or records where to find it)
template <int C>
__device__ __forceinline__ void load_query(QueryPatch<C> &q, const float *__restrict__ a, int ax, int ay, int aw,
                                           int ah, int lane, bool fetch = true)
{
    using T = PMTraits<C>;
    q.aw = aw;
    q.amask = patch_mask(ax, ay, aw, ah);
    const int j = (T::GROUPS == 1) ? lane : (lane % T::V);
    q.a_base = a + ((size_t)ay * aw + ax) * C + j * 4;
    if (T::A_IN_REGS && fetch) {
        const int g = (T::GROUPS == 1) ? 0 : lane / T::V;
#pragma unroll
        for (int i = 0; i < T::PPL; ++i) {
            const int pi = i * T::GROUPS + g;
            const int dy = pi / 3 - 1, dx = pi % 3 - 1;
            const bool ok = pi < 9 && ((q.amask >> pi) & 1u);
#pragma unroll
            for (int k = 0; k < T::VPL; ++k) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v = ldg4(q.a_base + ((ptrdiff_t)dy * aw + dx) * C + k * 128);
                q.a[i * T::VPL + k] = v;
            }
        }
    }
}

// candidate patch rows -> registers (all loads of a candidate are issued back to back)
template <int C>
__device__ __forceinline__ void load_cand(const QueryPatch<C> &q, const float *__restrict__ b, int bx, int by, int bw, int bh,
                                          int lane, float4 (&bv)[PMTraits<C>::PPL * PMTraits<C>::VPL], unsigned &valid)
{
    using T = PMTraits<C>;
    valid = q.amask & patch_mask(bx, by, bw, bh);
    const int j = (T::GROUPS == 1) ? lane : (lane % T::V);
    const int g = (T::GROUPS == 1) ? 0 : lane / T::V;
    const float *b_base = b + ((size_t)by * bw + bx) * C + j * 4;
#pragma unroll
    for (int i = 0; i < T::PPL; ++i) {
        const int pi = i * T::GROUPS + g;
        const int dy = pi / 3 - 1, dx = pi % 3 - 1;
        const bool ok = pi < 9 && ((valid >> pi) & 1u);
#pragma unroll
        for (int k = 0; k < T::VPL; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) v = ldg4(b_base + ((ptrdiff_t)dy * bw + dx) * C + k * 128);
            bv[i * T::VPL + k] = v;
        }
    }
}

// canonical-order patch distance (oracle decision D2) from loaded rows; all lanes return the same value
template <int C>
__device__ __forceinline__ float reduce_cand(const QueryPatch<C> &q, const float4 (&bv)[PMTraits<C>::PPL * PMTraits<C>::VPL],
                                             unsigned valid, int lane)
{
    using T = PMTraits<C>;
    const int g = (T::GROUPS == 1) ? 0 : lane / T::V;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < T::PPL; ++i) {
        const int pi = i * T::GROUPS + g;
        const int dy = pi / 3 - 1, dx = pi % 3 - 1;
        const bool ok = pi < 9 && ((valid >> pi) & 1u);
        if (ok) {
#pragma unroll
            for (int k = 0; k < T::VPL; ++k) {
                float4 av;
                if (T::A_IN_REGS) av = q.a[i * T::VPL + k];
                else av = ldg4(q.a_base + ((ptrdiff_t)dy * q.aw + dx) * C + k * 128);
                acc = fma4(av, bv[i * T::VPL + k], acc);
            }
        }
    }
    acc = butterfly(acc);
    const int n = __popc(valid);
    return __fdiv_rn(-acc, (float)n);
}

template <int C>
__device__ __forceinline__ float eval_dist(const QueryPatch<C> &q, const float *__restrict__ b, int bx, int by,
                                           int bw, int bh, int lane)
{
    float4 bv[PMTraits<C>::PPL * PMTraits<C>::VPL];
    unsigned valid;
    load_cand<C>(q, b, bx, by, bw, bh, lane, bv, valid);
    return reduce_cand<C>(q, bv, valid, lane);
}

template <int C>
__global__ void __launch_bounds__(PM_TPB, (C <= 128) ? PM_MINB_SMALL : 1) pm_step_kernel(const PMStep s)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (warp_global >= s.nq_total) return;
    const int dsel = warp_global >= s.nq0 ? 1 : 0;
    const PMDir &D = s.d[dsel];
    const int p = warp_global - (dsel ? s.nq0 : 0);
    const int aw = D.aw, ah = D.ah, bw = D.bw, bh = D.bh;
    const int ax = p % aw, ay = p / aw;

    const uint32_t v0 = D.nnf_in[p];
    int xbest = int_to_x(v0), ybest = int_to_y(v0);
    float dbest;
    unsigned n_eval = 0, n_ref = 0;

    // ---- propagation: L, R, U, D candidates from the previous step's NNF
    const int jump = s.jump;
    uint32_t cand[4];
    bool use[4];
    {
        const int qx[4] = {ax - jump, ax + jump, ax, ax};
        const int qy[4] = {ay, ay, ay - jump, ay + jump};
        const int sx[4] = {jump, -jump, 0, 0};
        const int sy[4] = {0, 0, jump, -jump};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            use[k] = false;
            cand[k] = 0;
            if (qx[k] >= 0 && qx[k] < aw && qy[k] >= 0 && qy[k] < ah) {
                uint32_t vp = D.nnf_in[qy[k] * aw + qx[k]];
                int xp = int_to_x(vp) + sx[k], yp = int_to_y(vp) + sy[k];
                if (yp >= 0 && yp < bh && xp >= 0 && xp < bw) {
                    n_ref++;
                    cand[k] = xy_to_int(xp, yp);
                    bool dup = (cand[k] == v0);
#pragma unroll
                    for (int t = 0; t < k; ++t) dup = dup || (use[t] && cand[t] == cand[k]);
                    // D4 (unchanged-source skip, oracle/pm_oracle.c): the neighbour's entry has not changed since this
                    // slot last judged it, so the candidate would be rejected again
                    const bool stale = s.t >= 4 && (int)D.lc_in[qy[k] * aw + qx[k]] <= s.t - 5;
                    use[k] = !dup && !stale;
                }
            }
        }
    }
    // with D4 most queries of a converged region have nothing to evaluate in the jump 8/4/2 steps: the query patch is
    // only fetched when something will be compared against it
    QueryPatch<C> q;
    load_query<C>(q, (const float *)D.a, ax, ay, aw, ah, lane, s.first || s.do_random || use[0] || use[1] || use[2] || use[3]);
    if (s.first) {
        dbest = eval_dist<C>(q, (const float *)D.b, xbest, ybest, bw, bh, lane);
        n_eval++;
        n_ref++;
    } else {
        dbest = D.nnd[p];
    }
    // the distances do not depend on each other: issue the row loads of BATCH candidates before the first FMA
    // (memory-level parallelism; the kernel is latency-bound, profiles/r1_pm_step_ncu.md)
    constexpr int BATCH = (C <= 64) ? PM_BATCH_SMALL : (C == 128 ? (PM_BATCH_SMALL > 1 ? PM_BATCH_SMALL / 2 : 1) : 1);
    float dc[4];
#pragma unroll
    for (int k0 = 0; k0 < 4; k0 += BATCH) {
        float4 bv[BATCH][PMTraits<C>::PPL * PMTraits<C>::VPL];
        unsigned valid[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j)
            if (use[k0 + j]) load_cand<C>(q, (const float *)D.b, int_to_x(cand[k0 + j]), int_to_y(cand[k0 + j]), bw, bh, lane, bv[j], valid[j]);
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            dc[k0 + j] = 0.f;
            if (use[k0 + j]) {
                dc[k0 + j] = reduce_cand<C>(q, bv[j], valid[j], lane);
                n_eval++;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (use[k] && dc[k] < dbest) {
            dbest = dc[k];
            xbest = int_to_x(cand[k]);
            ybest = int_to_y(cand[k]);
        }
    }

    // ---- random search around the current best (fused into the jump == 1 step)
    if (s.do_random) {
        const float *u = D.rng + (size_t)ax * D.ndraws + (size_t)s.iter * 2 * D.n_mag;
        int m = 0;
        for (int mag = D.rs_start; mag >= 1; mag /= 2, ++m) {
            const int xmin = max(xbest - mag, 0), xmax = min(xbest + mag + 1, bw);
            const int ymin = max(ybest - mag, 0), ymax = min(ybest + mag + 1, bh);
            const float u1 = __ldg(u + 2 * m), u2 = __ldg(u + 2 * m + 1);
            // (int)(u * w) % w with u in (0, 1]: the product is in [0, w], so the modulo only maps w to 0
            const int wx = xmax - xmin, wy = ymax - ymin;
            const int tx = (int)(__fmul_rn(u1, (float)wx)), ty = (int)(__fmul_rn(u2, (float)wy));
            const int xp = xmin + (tx >= wx ? tx - wx : tx);
            const int yp = ymin + (ty >= wy ? ty - wy : ty);
            n_ref++;
            if (xp == xbest && yp == ybest) continue;
            n_eval++;
            const float d = eval_dist<C>(q, (const float *)D.b, xp, yp, bw, bh, lane);
            if (__fadd_rn(d, FLT_MIN) < dbest) {
                dbest = d;
                xbest = xp;
                ybest = yp;
            }
        }
    }

    if (lane == 0) {
        const uint32_t vnew = xy_to_int(xbest, ybest);
        D.nnf_out[p] = vnew;
        D.lc_out[p] = vnew != v0 ? (int8_t)s.t : D.lc_in[p];
        D.nnd[p] = dbest;
        if (s.counters) {
            atomicAdd(&s.counters[0], (unsigned long long)n_eval);
            atomicAdd(&s.counters[1], (unsigned long long)n_ref);
        }
    }
}

// ------------------------------------------------------------------ group-per-query distance (C = 64 ... 512)
// One candidate loop per query: [initial entry (first step only)] + compacted propagation candidates + random-search
// candidates, all through ONE inlined copy of the distance code; interior patches (all nine pixels valid on both sides
// -- almost all of them) take a predicate-free path with three row pointers and immediate offsets.  The FMA / reduction
// sequence per distance is the canonical one (oracle D2), whatever the lane mapping.
// HALF = true : 16 lanes per query (C = 64: 1 float4 per lane and pixel, C = 128: 2)
// HALF = false: 32 lanes per query (C = 256: 2 float4 per lane and pixel, C = 512: 4)
template <int C, bool HALF>
struct UTraits {
    static constexpr int LANES = HALF ? 16 : 32;
    static constexpr int NV = C / (4 * LANES);          // float4 per lane per pixel
    static constexpr int STRIDE = 4 * LANES;            // floats between a lane's consecutive vectors
    static constexpr bool A_IN_REGS = (C == 64) || (C == 256);
};

template <int C, bool HALF, typename E = float>
struct UQuery {
    using T = UTraits<C, HALF>;
    float4 a[T::A_IN_REGS ? 9 * T::NV : 1];
    const E *a_base;
    int aw;
    unsigned amask;
};

// canonical slot structure (oracle D2): 32 accumulator slots, slot = (vector index within the pixel row) mod 32 for
// C >= 128, and (pixel parity) * 16 + vector index for C = 64; butterfly xor 16, 8, 4, 2, 1.
//   HALF, C = 64 : a lane owns slots j (even pixels, acc0) and j + 16 (odd pixels, acc1)
//   HALF, C = 128: a lane owns slots j (vector j, acc0) and j + 16 (vector j + 16, acc1)
//   warp, C >= 256: a lane owns slot `lane` (vectors lane, lane + 32, ... in ascending order, one accumulator)
template <int C, bool HALF>
__device__ __forceinline__ void u_accumulate(float &acc0, float &acc1, int pi, const float4 *av, const float4 *bv)
{
    constexpr int NV = UTraits<C, HALF>::NV;
    if (HALF) {
        if (NV == 1) {
            if (pi & 1) acc1 = fma4(av[0], bv[0], acc1);
            else acc0 = fma4(av[0], bv[0], acc0);
        } else {
            acc0 = fma4(av[0], bv[0], acc0);
            acc1 = fma4(av[1], bv[1], acc1);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NV; ++k) acc0 = fma4(av[k], bv[k], acc0);
    }
}

template <int C, bool HALF>
__device__ __forceinline__ float u_finish(float acc0, float acc1, unsigned mask, int n)
{
    float acc;
    if (HALF) acc = __fadd_rn(acc0, acc1);  // == the xor-16 step of the 32-slot butterfly
    else acc = __fadd_rn(acc0, __shfl_xor_sync(mask, acc0, 16));
    acc = __fadd_rn(acc, __shfl_xor_sync(mask, acc, 8));
    acc = __fadd_rn(acc, __shfl_xor_sync(mask, acc, 4));
    acc = __fadd_rn(acc, __shfl_xor_sync(mask, acc, 2));
    acc = __fadd_rn(acc, __shfl_xor_sync(mask, acc, 1));
    return __fdiv_rn(-acc, (float)n);
}

template <int C, bool HALF, typename E>
__device__ __forceinline__ float u_eval(const UQuery<C, HALF, E> &q, const E *__restrict__ b, int bx, int by, int bw, int bh, int j,
                                        unsigned mask)
{
    using T = UTraits<C, HALF>;
    constexpr int NV = T::NV, ST = T::STRIDE;
    const unsigned valid = q.amask & patch_mask(bx, by, bw, bh);
    const E *b_base = b + ((size_t)by * bw + bx) * C + j * 4;
    float acc0 = 0.f, acc1 = 0.f;
    if (valid == 0x1FFu) {
        const E *rb[3] = {b_base - (ptrdiff_t)bw * C, b_base, b_base + (ptrdiff_t)bw * C};
        const E *ra[3] = {q.a_base - (ptrdiff_t)q.aw * C, q.a_base, q.a_base + (ptrdiff_t)q.aw * C};
        float4 bv[9 * NV];
#pragma unroll
        for (int pi = 0; pi < 9; ++pi)
#pragma unroll
            for (int k = 0; k < NV; ++k) bv[pi * NV + k] = ldg4(rb[pi / 3] + (pi % 3 - 1) * C + k * ST);
#pragma unroll
        for (int pi = 0; pi < 9; ++pi) {
            float4 av[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (T::A_IN_REGS) av[k] = q.a[pi * NV + k];
                else av[k] = ldg4(ra[pi / 3] + (pi % 3 - 1) * C + k * ST);
            }
            u_accumulate<C, HALF>(acc0, acc1, pi, av, &bv[pi * NV]);
        }
        return u_finish<C, HALF>(acc0, acc1, mask, 9);
    }
    float4 bv[9 * NV];
#pragma unroll
    for (int pi = 0; pi < 9; ++pi) {
        const int dy = pi / 3 - 1, dx = pi % 3 - 1;
        const bool ok = (valid >> pi) & 1u;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) v = ldg4(b_base + ((ptrdiff_t)dy * bw + dx) * C + k * ST);
            bv[pi * NV + k] = v;
        }
    }
#pragma unroll
    for (int pi = 0; pi < 9; ++pi) {
        const int dy = pi / 3 - 1, dx = pi % 3 - 1;
        if ((valid >> pi) & 1u) {
            float4 av[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (T::A_IN_REGS) av[k] = q.a[pi * NV + k];
                else av[k] = ldg4(q.a_base + ((ptrdiff_t)dy * q.aw + dx) * C + k * ST);
            }
            u_accumulate<C, HALF>(acc0, acc1, pi, av, &bv[pi * NV]);
        }
    }
    return u_finish<C, HALF>(acc0, acc1, mask, __popc(valid));
}

// ------------------------------------------------------------------ tiled step kernel (C = 64 ... 512)
// After D4 the late steps are sparse (about one candidate per query), and a kernel with one (half) warp per query is bound by the
// LIFETIME of its short warps: field entry -> four neighbour entries -> patch rows is a chain of dependent global
// loads (~2.5 us) paid by every (half) warp for one query (profiles/r1_pm_step_ncu.md).  Here a warp owns a TILE of
// up to 32 consecutive queries:
//   phase 1, lane = query: the field / last-change reads are coalesced, the candidate lists are built by 32 lanes at
//            once, queries with nothing to evaluate are finished right there (coalesced stores);
//   phase 2, 16 (or 32) lanes = one query: the queries that do have work are walked one after the other by the two
//            half warps (even / odd ranked work items), so the dependent-load chain is paid once per tile.
// Variants that keep more candidate rows of a query in flight (register prefetch, cp.async staging in shared memory, query
// patch in shared memory for more resident warps) were measured in rounds 1-2 and lost: profiles/r2_pm_tuning.md.
struct TQueryState {
    uint32_t v0, c0, c1, c2, c3;
    int qidx;
    int n;        // propagation candidates
    float dbest;
};

// MINB = resident CTAs per SM the register allocation aims at (half-warp kernels; 4 -> 127 registers, 5 -> 96, 6 -> 80)
template <int C, bool HALF, typename E, int MINB = 4>
__global__ void __launch_bounds__(128, HALF ? MINB : 2) pm_step_t_kernel(const PMStep s, const int tile)
{
    using T = UTraits<C, HALF>;
    __shared__ TQueryState st_all[4][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    TQueryState *st = st_all[wib];
    const int warp_global = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int q_first = warp_global * tile;
    if (q_first >= s.nq_total) return;  // whole warps leave together
    const int jump = s.jump;
    const int n_first = s.first ? 1 : 0;
    unsigned n_eval = 0, n_ref = 0;

    // ---- phase 1: one lane per query
    bool has_work = false;
    {
        const int qidx = q_first + lane;
        if (lane < tile && qidx < s.nq_total) {
            const int dsel = qidx >= s.nq0 ? 1 : 0;
            const PMDir &D = s.d[dsel];
            const int p = qidx - (dsel ? s.nq0 : 0);
            const int aw = D.aw, ah = D.ah, bw = D.bw, bh = D.bh;
            const int ax = p % aw, ay = p / aw;
            const uint32_t v0 = D.nnf_in[p];
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            int n = 0;
            const int qx[4] = {ax - jump, ax + jump, ax, ax};
            const int qy[4] = {ay, ay, ay - jump, ay + jump};
            const int sx[4] = {jump, -jump, 0, 0};
            const int sy[4] = {0, 0, jump, -jump};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (qx[k] >= 0 && qx[k] < aw && qy[k] >= 0 && qy[k] < ah) {
                    const int qi = qy[k] * aw + qx[k];
                    const uint32_t vp = D.nnf_in[qi];
                    const int xp = int_to_x(vp) + sx[k], yp = int_to_y(vp) + sy[k];
                    if (yp >= 0 && yp < bh && xp >= 0 && xp < bw) {
                        n_ref++;
                        const uint32_t cv = xy_to_int(xp, yp);
                        const bool dup = (cv == v0) || (n > 0 && cv == c0) || (n > 1 && cv == c1) || (n > 2 && cv == c2);
                        const bool stale = s.t >= 4 && (int)D.lc_in[qi] <= s.t - 5;  // D4
                        if (!dup && !stale) {
                            if (n == 0) c0 = cv;
                            else if (n == 1) c1 = cv;
                            else if (n == 2) c2 = cv;
                            else c3 = cv;
                            n++;
                        }
                    }
                }
            }
            const int total = n_first + n + (s.do_random ? D.n_mag : 0);
            if (total == 0) {  // nothing to compare: the entry stays
                D.nnf_out[p] = v0;
                D.lc_out[p] = D.lc_in[p];
            } else {
                has_work = true;
                st[lane].v0 = v0; st[lane].c0 = c0; st[lane].c1 = c1; st[lane].c2 = c2; st[lane].c3 = c3;
                st[lane].qidx = qidx;
                st[lane].n = n;
                st[lane].dbest = s.first ? 0.f : D.nnd[p];
            }
        }
    }
    const unsigned work = __ballot_sync(0xffffffffu, has_work);
    __syncwarp();

    // ---- phase 2: one group of 16 / 32 lanes per query with work
    const int grp = HALF ? (lane >> 4) : 0;
    const int j = HALF ? (lane & 15) : lane;
    const unsigned mask = HALF ? (grp ? 0xffff0000u : 0x0000ffffu) : 0xffffffffu;
    // even-ranked work items -> lanes 0-15, odd-ranked -> lanes 16-31: both groups run the SAME loop (converged), each on
    // its own item
    const int nwork = __popc(work);
#pragma unroll 1
    for (int r = grp; r < nwork; r += (HALF ? 2 : 1)) {
        const int src = (int)__fns(work, 0, r + 1);
        const TQueryState q0 = st[src];
        const int qidx = q0.qidx;
        const int dsel = qidx >= s.nq0 ? 1 : 0;
        const PMDir &D = s.d[dsel];
        const int p = qidx - (dsel ? s.nq0 : 0);
        const int aw = D.aw, ah = D.ah, bw = D.bw, bh = D.bh;
        const int ax = p % aw, ay = p / aw;
        const uint32_t v0 = q0.v0;
        int xbest = int_to_x(v0), ybest = int_to_y(v0);
        float dbest = q0.dbest;
        const int n_prop_end = n_first + q0.n;
        const int total = n_prop_end + (s.do_random ? D.n_mag : 0);

        UQuery<C, HALF, E> q;
        q.aw = aw;
        q.amask = patch_mask(ax, ay, aw, ah);
        q.a_base = (const E *)D.a + ((size_t)ay * aw + ax) * C + j * 4;
        if (T::A_IN_REGS) {
#pragma unroll
            for (int pi = 0; pi < 9; ++pi) {
                const int dy = pi / 3 - 1, dx = pi % 3 - 1;
#pragma unroll
                for (int k = 0; k < T::NV; ++k) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((q.amask >> pi) & 1u) v = ldg4(q.a_base + ((ptrdiff_t)dy * aw + dx) * C + k * T::STRIDE);
                    q.a[pi * T::NV + k] = v;
                }
            }
        }
        const float *u = D.rng + (size_t)ax * D.ndraws + (size_t)s.iter * 2 * D.n_mag;
#pragma unroll 1
        for (int i = 0; i < total; ++i) {
            int cx, cy;
            const bool is_rand = i >= n_prop_end;
            if (!is_rand) {
                const int k = i - n_first;
                const uint32_t cv = (i < n_first) ? v0 : (k == 0 ? q0.c0 : (k == 1 ? q0.c1 : (k == 2 ? q0.c2 : q0.c3)));
                cx = int_to_x(cv);
                cy = int_to_y(cv);
                if (i < n_first) n_ref += (j == 0);
            } else {
                const int m = i - n_prop_end;
                const int mag = D.rs_start >> m;
                const int xmin = max(xbest - mag, 0), xmax = min(xbest + mag + 1, bw);
                const int ymin = max(ybest - mag, 0), ymax = min(ybest + mag + 1, bh);
                const float u1 = __ldg(u + 2 * m), u2 = __ldg(u + 2 * m + 1);
                const int wx = xmax - xmin, wy = ymax - ymin;
                const int tx = (int)(__fmul_rn(u1, (float)wx)), ty = (int)(__fmul_rn(u2, (float)wy));
                cx = xmin + (tx >= wx ? tx - wx : tx);
                cy = ymin + (ty >= wy ? ty - wy : ty);
                n_ref += (j == 0);
                if (cx == xbest && cy == ybest) continue;  // D3
            }
            n_eval += (j == 0);
            const float d = u_eval<C, HALF, E>(q, (const E *)D.b, cx, cy, bw, bh, j, mask);
            const float dcmp = is_rand ? __fadd_rn(d, FLT_MIN) : d;
            if (i < n_first || dcmp < dbest) {
                dbest = d;
                xbest = cx;
                ybest = cy;
            }
        }
        if (j == 0) {
            const uint32_t vnew = xy_to_int(xbest, ybest);
            D.nnf_out[p] = vnew;
            D.lc_out[p] = vnew != v0 ? (int8_t)s.t : D.lc_in[p];
            D.nnd[p] = dbest;
        }
    }
    if (s.counters && (n_eval | n_ref)) {
        atomicAdd(&s.counters[0], (unsigned long long)n_eval);
        atomicAdd(&s.counters[1], (unsigned long long)n_ref);
    }
}

// queries per warp: 32 when that still leaves >= ~16 warps per SM, otherwise fewer (the coarse levels have few queries)
static int pm_tile_size(int nq_total, int num_sms)
{
    int tile = 32;
    while (tile > 2 && nq_total / tile < num_sms * 16) tile >>= 1;
    return tile;
}

template <int C>
struct UseHalfWarp { static constexpr bool value = (C == 64 || C == 128); };

// C >= 64: the tiled kernel (16 lanes per distance at C = 64 / 128, 32 at C = 256 / 512); C = 16, 32 (not used by the
// VGG levels; accepted by the ABI): the plain warp-per-query kernel
template <int C>
struct StepLauncher {
    static void go(const PMStep &s, cudaStream_t st, bool f16)
    {
        if constexpr (C >= 64) {
            const int tile = pm_tile_size(s.nq_total, 148);
            const int blocks = nct_div_up(nct_div_up(s.nq_total, tile), 4);  // 4 warps per block, `tile` queries per warp
            if constexpr (UseHalfWarp<C>::value) {
                // NCT_PM_MINB = 4 / 5 / 6 selects the register budget of the half-warp kernels (A/B knob).  Measured (B200, 700^2
                // level shapes, profiles/r2_pm_tuning.md): C = 64: 28.1 / 28.6 / 30.6 ms, C = 128: 15.3 / 14.7 / 15.9 ms
                static const int minb = getenv("NCT_PM_MINB") ? atoi(getenv("NCT_PM_MINB")) : (C == 128 ? 5 : 4);
                if (f16) {
                    if (minb >= 6) pm_step_t_kernel<C, true, __half, 6><<<blocks, 128, 0, st>>>(s, tile);
                    else if (minb == 5) pm_step_t_kernel<C, true, __half, 5><<<blocks, 128, 0, st>>>(s, tile);
                    else pm_step_t_kernel<C, true, __half, 4><<<blocks, 128, 0, st>>>(s, tile);
                } else {
                    if (minb >= 6) pm_step_t_kernel<C, true, float, 6><<<blocks, 128, 0, st>>>(s, tile);
                    else if (minb == 5) pm_step_t_kernel<C, true, float, 5><<<blocks, 128, 0, st>>>(s, tile);
                    else pm_step_t_kernel<C, true, float, 4><<<blocks, 128, 0, st>>>(s, tile);
                }
            } else {
                if (f16) pm_step_t_kernel<C, false, __half><<<blocks, 128, 0, st>>>(s, tile);
                else pm_step_t_kernel<C, false, float><<<blocks, 128, 0, st>>>(s, tile);
            }
        } else {
            const int blocks = nct_div_up(s.nq_total, PM_TPB / 32);
            pm_step_kernel<C><<<blocks, PM_TPB, 0, st>>>(s);
        }
    }
};

// iters == 0: only the initial distance (NCT/GeneralizedPatchMatch.cu:710-712)
template <int C>
__global__ void __launch_bounds__(PM_TPB) pm_init_dist_kernel(const PMStep s)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (warp_global >= s.nq_total) return;
    const int dsel = warp_global >= s.nq0 ? 1 : 0;
    const PMDir &D = s.d[dsel];
    const int p = warp_global - (dsel ? s.nq0 : 0);
    const int ax = p % D.aw, ay = p / D.aw;
    QueryPatch<C> q;
    load_query<C>(q, (const float *)D.a, ax, ay, D.aw, D.ah, lane);
    const uint32_t v0 = D.nnf_in[p];
    float d = eval_dist<C>(q, (const float *)D.b, int_to_x(v0), int_to_y(v0), D.bw, D.bh, lane);
    if (lane == 0) D.nnd[p] = d;
}

template <int C>
int launch_pm(nct_ctx *ctx, PMStep &s, int iters, uint32_t *tmp0, uint32_t *tmp1, int ndir, int8_t *lc, bool f16)
{
    if (f16 && (C < 64 || iters == 0))
        return nct_fail(ctx, NCT_ERR_ARG, "the FP16 feature store covers C >= 64 and iters >= 1 (got C = %d, iters = %d)", C, iters);
    const int warps_per_block = PM_TPB / 32;
    const int blocks = nct_div_up(s.nq_total, warps_per_block);
    s.counters = ctx->pm_count_evals ? ctx->pm_counters : nullptr;
    if (iters == 0) {
        pm_init_dist_kernel<C><<<blocks, PM_TPB, 0, ctx->stream>>>(s);
        NCT_CHECK_LAUNCH(ctx);
        return NCT_OK;
    }
    uint32_t *user[2] = {s.d[0].nnf_out, ndir > 1 ? s.d[1].nnf_out : nullptr};
    uint32_t *tmp[2] = {tmp0, tmp1};
    // last-change steps, double buffered like the field: [buffer][direction 0 | direction 1]
    const size_t nq = (size_t)s.nq_total;
    int8_t *lcbuf[2][2] = {{lc, lc + s.nq0}, {lc + nq, lc + nq + s.nq0}};
    auto steps = [&]() -> int {
        NCT_CUDA(ctx, cudaMemsetAsync(lc, 0xff, nq, ctx->stream));
        int step = 0;
        for (int iter = 0; iter < iters; ++iter)
            for (int jump = 8; jump > 0; jump /= 2, ++step) {
                s.iter = iter;
                s.jump = jump;
                s.t = step;
                s.first = (step == 0);
                s.do_random = (jump == 1);
                for (int d = 0; d < ndir; ++d) {
                    // even steps read the caller's buffer and write scratch; odd steps the reverse.
                    // 4*iters steps is even, so the final NNF lands in the caller's buffer.
                    s.d[d].nnf_in = (step & 1) ? tmp[d] : user[d];
                    s.d[d].nnf_out = (step & 1) ? user[d] : tmp[d];
                    s.d[d].lc_in = lcbuf[step & 1][d];
                    s.d[d].lc_out = lcbuf[(step & 1) ^ 1][d];
                }
                StepLauncher<C>::go(s, ctx->stream, f16);
                NCT_CHECK_LAUNCH(ctx);
            }
        return NCT_OK;
    };
    // The 4 * iters step launches of a level as ONE replayed graph per level shape (the pipeline calls with the same buffers
    // for every pair, so a pair re-captures nothing); NCT_PM_GRAPH=0: plain stream launches.
    const bool use_graph = !(getenv("NCT_PM_GRAPH") && atoi(getenv("NCT_PM_GRAPH")) == 0);
    if (!use_graph) return steps();
    auto P = [](const void *q) { return (unsigned long long)(uintptr_t)q; };
    std::vector<unsigned long long> key = {(unsigned long long)C, (unsigned long long)f16, (unsigned long long)iters, (unsigned long long)ndir,
                                           P(tmp0), P(tmp1), P(lc), P(s.counters), (unsigned long long)s.nq0, (unsigned long long)s.nq_total};
    for (int d = 0; d < ndir; ++d) {
        const PMDir &D = s.d[d];
        const unsigned long long part[] = {P(D.a), P(D.b), P(D.nnf_out), P(D.nnd), P(D.rng), (unsigned long long)D.ah, (unsigned long long)D.aw,
                                           (unsigned long long)D.bh, (unsigned long long)D.bw, (unsigned long long)D.rs_start,
                                           (unsigned long long)D.n_mag, (unsigned long long)D.ndraws};
        key.insert(key.end(), part, part + sizeof(part) / sizeof(part[0]));
    }
    char gname[64];
    snprintf(gname, sizeof(gname), "pm_steps_c%d_%dx%d_%d", C, s.d[0].ah, s.d[0].aw, ndir);
    if (!nct_graph_cached(ctx, gname, key)) {
        int rc = nct_graph_begin(ctx, gname, key, nullptr);
        if (rc) return rc;
        rc = steps();
        if (rc) return nct_graph_abort(ctx, gname, rc);
        rc = nct_graph_end(ctx, gname);
        if (rc) return rc;
    }
    return nct_graph_launch(ctx, gname);
}

int fill_dir(nct_ctx *ctx, PMDir &D, const void *a, const void *b, uint32_t *ann, float *annd, int ah, int aw,
             int bh, int bw, int iters, int rs_max, const char *rng_name)
{
    D.a = a;
    D.b = b;
    D.nnf_in = ann;
    D.nnf_out = ann;
    D.nnd = annd;
    D.ah = ah; D.aw = aw; D.bh = bh; D.bw = bw;
    int rs = rs_max;
    if (rs > (bw > bh ? bw : bh)) rs = (bw > bh ? bw : bh);
    D.rs_start = rs;
    D.n_mag = 0;
    for (int mag = rs; mag >= 1; mag /= 2) D.n_mag++;
    D.ndraws = 2 * D.n_mag * iters;
    D.rng = nullptr;
    if (D.ndraws > 0) {
        float *tab = (float *)nct_scratch(ctx, rng_name, sizeof(float) * (size_t)aw * D.ndraws);
        if (!tab) return NCT_ERR_NOMEM;
        xorwow_table_kernel<<<nct_div_up(aw, 128), 128, 0, ctx->stream>>>(tab, aw, D.ndraws);
        NCT_CHECK_LAUNCH(ctx);
        D.rng = tab;
    }
    return NCT_OK;
}

int check_params(nct_ctx *ctx, const int *p)
{
    NCT_REQUIRE(ctx, p != nullptr, "params is null");
    const int C = p[0];
    NCT_REQUIRE(ctx, C == 16 || C == 32 || C == 64 || (C >= 128 && C % 128 == 0 && C <= 512),
                "unsupported channel count %d (16, 32, 64, 128, 256, 384, 512)", C);
    NCT_REQUIRE(ctx, C != 384, "unsupported channel count 384");
    NCT_REQUIRE(ctx, p[1] > 0 && p[2] > 0 && p[3] > 0 && p[4] > 0 && p[1] <= 4096 && p[2] <= 4096 && p[3] <= 4096 &&
                         p[4] <= 4096,
                "image sides must be in [1, 4096] (12-bit NNF packing)");
    NCT_REQUIRE(ctx, p[5] == 3, "patch size %d unsupported (reference uses 3, CT/Config.h:70)", p[5]);
    NCT_REQUIRE(ctx, p[6] >= 0 && p[6] <= 31, "iters %d out of range [0, 31] (the reference uses 10, NCT/main.cu:65)", p[6]);
    NCT_REQUIRE(ctx, p[7] >= 0, "rs_max must be >= 0 (0 = no random search, as in the reference for sides below 64: NCT/main.cu:77-83)");
    NCT_REQUIRE(ctx, p[8] == 0, "flag_constraint must be 0 (the reference never enables it, NCT/main.cu:66)");
    return NCT_OK;
}

int dispatch_pm(nct_ctx *ctx, int C, PMStep &s, int iters, uint32_t *t0, uint32_t *t1, int ndir, bool f16 = false)
{
    int8_t *lc = (int8_t *)nct_scratch(ctx, "pm_last_change", 2 * (size_t)s.nq_total);
    if (!lc) return NCT_ERR_NOMEM;
    switch (C) {
    case 16: return launch_pm<16>(ctx, s, iters, t0, t1, ndir, lc, f16);
    case 32: return launch_pm<32>(ctx, s, iters, t0, t1, ndir, lc, f16);
    case 64: return launch_pm<64>(ctx, s, iters, t0, t1, ndir, lc, f16);
    case 128: return launch_pm<128>(ctx, s, iters, t0, t1, ndir, lc, f16);
    case 256: return launch_pm<256>(ctx, s, iters, t0, t1, ndir, lc, f16);
    case 512: return launch_pm<512>(ctx, s, iters, t0, t1, ndir, lc, f16);
    }
    return nct_fail(ctx, NCT_ERR_ARG, "unsupported channel count %d", C);
}

}  // namespace

extern "C" {

int nct_nnf_init(nct_ctx *ctx, uint32_t *ann_dev, int ah, int aw, int bh, int bw)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, ann_dev && ah > 0 && aw > 0 && bh > 0 && bw > 0, "bad arguments");
    dim3 block(32, 8), grid(nct_div_up(aw, 32), nct_div_up(ah, 8));
    nnf_init_kernel<<<grid, block, 0, ctx->stream>>>(ann_dev, ah, aw, bh, bw);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_nnf_upsample(nct_ctx *ctx, const uint32_t *ann_half_dev, int ah_half, int aw_half, uint32_t *ann_dev, int ah,
                     int aw, int bh, int bw)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, ann_half_dev && ann_dev && ann_half_dev != ann_dev, "bad / aliased NNF buffers");
    NCT_REQUIRE(ctx, ah_half > 0 && aw_half > 0 && ah > 0 && aw > 0 && bh > 0 && bw > 0, "bad sizes");
    dim3 block(32, 8), grid(nct_div_up(aw, 32), nct_div_up(ah, 8));
    nnf_upsample_kernel<<<grid, block, 0, ctx->stream>>>(ann_half_dev, ah_half, aw_half, ann_dev, ah, aw, bh, bw);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

int nct_xorwow_table(nct_ctx *ctx, float *out_dev, int ncols, int ndraws)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, out_dev && ncols > 0 && ndraws > 0, "bad arguments");
    xorwow_table_kernel<<<nct_div_up(ncols, 128), 128, 0, ctx->stream>>>(out_dev, ncols, ndraws);
    NCT_CHECK_LAUNCH(ctx);
    return NCT_OK;
}

static int patchmatch_single_impl(nct_ctx *ctx, const void *a, const void *b, uint32_t *ann, float *annd, const int params[11], bool f16)
{
    NCT_ENTER(ctx);
    int rc = check_params(ctx, params);
    if (rc) return rc;
    NCT_REQUIRE(ctx, a && b && ann && annd, "null device pointer");
    const int C = params[0], ah = params[1], aw = params[2], bh = params[3], bw = params[4];
    const int iters = params[6], rs_max = params[7];
    PMStep s{};
    rc = fill_dir(ctx, s.d[0], a, b, ann, annd, ah, aw, bh, bw, iters, rs_max, "pm_rng0");
    if (rc) return rc;
    s.d[1] = s.d[0];
    s.nq0 = ah * aw;
    s.nq_total = s.nq0;
    uint32_t *t0 = (uint32_t *)nct_scratch(ctx, "pm_nnf_tmp0", sizeof(uint32_t) * (size_t)ah * aw);
    if (!t0) return NCT_ERR_NOMEM;
    if (ctx->pm_count_evals) NCT_CUDA(ctx, cudaMemsetAsync(ctx->pm_counters, 0, 16, ctx->stream));
    return dispatch_pm(ctx, C, s, iters, t0, nullptr, 1, f16);
}

int nct_patchmatch(nct_ctx *ctx, const float *a, const float *b, uint32_t *ann, float *annd, const int params[11])
{
    return patchmatch_single_impl(ctx, a, b, ann, annd, params, false);
}

int nct_patchmatch_f16(nct_ctx *ctx, const uint16_t *a, const uint16_t *b, uint32_t *ann, float *annd, const int params[11])
{
    return patchmatch_single_impl(ctx, a, b, ann, annd, params, true);
}

static int patchmatch_bidir_impl(nct_ctx *ctx, const void *a, const void *b, uint32_t *ann, float *annd, uint32_t *bnn, float *bnnd,
                                 const int params[11], bool f16)
{
    NCT_ENTER(ctx);
    int rc = check_params(ctx, params);
    if (rc) return rc;
    NCT_REQUIRE(ctx, a && b && ann && annd && bnn && bnnd, "null device pointer");
    const int C = params[0], ah = params[1], aw = params[2], bh = params[3], bw = params[4];
    const int iters = params[6], rs_max = params[7];
    PMStep s{};
    rc = fill_dir(ctx, s.d[0], a, b, ann, annd, ah, aw, bh, bw, iters, rs_max, "pm_rng0");
    if (rc) return rc;
    rc = fill_dir(ctx, s.d[1], b, a, bnn, bnnd, bh, bw, ah, aw, iters, rs_max, "pm_rng1");
    if (rc) return rc;
    s.nq0 = ah * aw;
    s.nq_total = ah * aw + bh * bw;
    uint32_t *t0 = (uint32_t *)nct_scratch(ctx, "pm_nnf_tmp0", sizeof(uint32_t) * (size_t)ah * aw);
    uint32_t *t1 = (uint32_t *)nct_scratch(ctx, "pm_nnf_tmp1", sizeof(uint32_t) * (size_t)bh * bw);
    if (!t0 || !t1) return NCT_ERR_NOMEM;
    if (ctx->pm_count_evals) NCT_CUDA(ctx, cudaMemsetAsync(ctx->pm_counters, 0, 16, ctx->stream));
    return dispatch_pm(ctx, C, s, iters, t0, t1, 2, f16);
}

int nct_patchmatch_bidir(nct_ctx *ctx, const float *a, const float *b, uint32_t *ann, float *annd, uint32_t *bnn,
                         float *bnnd, const int params[11])
{
    return patchmatch_bidir_impl(ctx, a, b, ann, annd, bnn, bnnd, params, false);
}

int nct_patchmatch_bidir_f16(nct_ctx *ctx, const uint16_t *a, const uint16_t *b, uint32_t *ann, float *annd, uint32_t *bnn,
                             float *bnnd, const int params[11])
{
    return patchmatch_bidir_impl(ctx, a, b, ann, annd, bnn, bnnd, params, true);
}

int nct_patchmatch_count_evals(nct_ctx *ctx, int enable)
{
    NCT_ENTER(ctx);
    ctx->pm_count_evals = enable ? 1 : 0;
    return NCT_OK;
}

int nct_patchmatch_stats(nct_ctx *ctx, long long stats[2])
{
    if (!ctx || !stats) return NCT_ERR_ARG;
    cudaSetDevice(ctx->device);
    unsigned long long h[2];
    NCT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NCT_CUDA(ctx, cudaMemcpy(h, ctx->pm_counters, sizeof(h), cudaMemcpyDeviceToHost));
    stats[0] = (long long)h[0];
    stats[1] = (long long)h[1];
    return NCT_OK;
}

}  // extern "C"
