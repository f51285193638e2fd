// nct_solve_direct: the reference's CSR-level entry point of the WLS stage, for callers that keep its host assembly.
//
// Replaces solve_direct_cpu (CT/SparseSolver_CPU.h:35-43, CT/SparseSolver_CPU.cpp:104-286):
//     void solve_direct_cpu(Mat& aRes, Mat& bRes, int nonZeroNum, int eleNum, int oneBased, double* A, int* rowIndex,
//                           int* columns, double* Ba0, double* Xa0, ..., double* Bb2, double* Xb2)
// = MKL PARDISO, mtype 2 (real symmetric positive definite), the UPPER triangle of the matrix in CSR (one-based when
// oneBased != 0), six right-hand sides, all arrays on the HOST; the two Mat& receive copies of the solutions
// (copy_data) and are not part of the numerical interface.  PARDISO is a direct solver: the result is the exact solution
// to rounding.  Here: the symmetric matrix is expanded to full CSR on the host, uploaded, and solved for all six
// right-hand sides at once by Jacobi-preconditioned CG in FP64 to a relative residual `rel_tol` (default 1e-10), with
// deterministic two-stage reductions and an on-device stopping test.  It works for ANY SPD matrix in that format; for
// the WLS system itself nct_solve_wls (multigrid-preconditioned, no explicit matrix) is ~40x faster and is what the
// pipeline uses.
#include "nct_internal.h"
#include <cmath>
#include <vector>

namespace {

constexpr int TPB = 256;
constexpr int NR = 6;

struct DScalars {
    double rz[NR], alpha[NR], beta[NR], rr[NR], bb[NR];
    double tol2;
    int iters, max_iters, done, converged;
};

__device__ __forceinline__ void block_reduce6(double (&v)[NR], double *smem)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) smem[k * (TPB / 32) + w] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double s = 0.0;
            for (int i = 0; i < TPB / 32; ++i) s += smem[k * (TPB / 32) + i];
            v[k] = s;
        }
    }
    __syncthreads();
}

// fixed-order grid reduction: block partials in block order, summed by the block that finishes last
__device__ __forceinline__ bool grid_reduce6(double (&v)[NR], double *partials, unsigned *counter, double *smem)
{
    __shared__ bool last;
    block_reduce6(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) partials[(size_t)blockIdx.x * NR + k] = v[k];
        __threadfence();
        last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += TPB) {
#pragma unroll
        for (int k = 0; k < NR; ++k) acc[k] += __ldcg(&partials[(size_t)b * NR + k]);
    }
    block_reduce6(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) v[k] = acc[k];
        *counter = 0;
        return true;
    }
    return false;
}

__device__ __forceinline__ bool all_converged(const DScalars *sc)
{
    bool ok = true;
    for (int k = 0; k < NR; ++k) ok = ok && (sc->bb[k] > 0.0 ? sc->rr[k] <= sc->tol2 * sc->bb[k] : sc->rr[k] == 0.0);
    return ok;
}

// vectors: [NR][n] planar.  x = 0, r = b, z = D^-1 r, p = z; bb = rr = b.b, rz = r.z
__global__ void __launch_bounds__(TPB) d_init_kernel(int n, const double *__restrict__ b, const double *__restrict__ invd, double *__restrict__ x,
                                                     double *__restrict__ r, double *__restrict__ p, DScalars *sc, double *partials,
                                                     unsigned *counter, double tol2, int max_iters)
{
    __shared__ double smem[NR * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double d0[NR], d1[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) d0[k] = d1[k] = 0.0;
    if (i < n) {
        const double id = invd[i];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const double bi = b[(size_t)k * n + i];
            x[(size_t)k * n + i] = 0.0;
            r[(size_t)k * n + i] = bi;
            p[(size_t)k * n + i] = bi * id;
            d0[k] = bi * bi;
            d1[k] = bi * bi * id;
        }
    }
    const bool fin0 = grid_reduce6(d0, partials, counter, smem);
    if (fin0) {
        for (int k = 0; k < NR; ++k) { sc->bb[k] = d0[k]; sc->rr[k] = d0[k]; sc->alpha[k] = 0.0; sc->beta[k] = 0.0; }
        sc->iters = 0;
        sc->tol2 = tol2;
        sc->max_iters = max_iters;
    }
    __syncthreads();
    const bool fin1 = grid_reduce6(d1, partials + (size_t)gridDim.x * NR, counter + 1, smem);
    if (fin1) {
        for (int k = 0; k < NR; ++k) sc->rz[k] = d1[k];
    }
}

// done flag after the init reductions (both "last blocks" may differ, so a tiny follow-up kernel sets it)
__global__ void d_flag_kernel(DScalars *sc)
{
    sc->converged = all_converged(sc) ? 1 : 0;
    sc->done = (sc->converged || sc->max_iters <= 0) ? 1 : 0;
}

// Ap = A p (full symmetric CSR, zero-based, row-parallel, entries in storage order) ; alpha = rz / p.Ap
__global__ void __launch_bounds__(TPB) d_spmv_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ cols,
                                                     const double *__restrict__ vals, const double *__restrict__ p, double *__restrict__ Ap,
                                                     DScalars *sc, double *partials, unsigned *counter)
{
    __shared__ double smem[NR * TPB / 32];
    if (sc->done) return;
    const int i = blockIdx.x * TPB + threadIdx.x;
    double dots[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) dots[k] = 0.0;
    if (i < n) {
        double s[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) s[k] = 0.0;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const double a = vals[e];
            const int c = cols[e];
#pragma unroll
            for (int k = 0; k < NR; ++k) s[k] += a * p[(size_t)k * n + c];
        }
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            Ap[(size_t)k * n + i] = s[k];
            dots[k] = p[(size_t)k * n + i] * s[k];
        }
    }
    if (grid_reduce6(dots, partials, counter, smem)) {
        for (int k = 0; k < NR; ++k) sc->alpha[k] = dots[k] > 0.0 ? sc->rz[k] / dots[k] : 0.0;
    }
}

// x += alpha p ; r -= alpha Ap ; rr = r.r ; rz_new = r.D^-1 r ; beta ; stopping test
__global__ void __launch_bounds__(TPB) d_update_kernel(int n, const double *__restrict__ invd, double *__restrict__ x, double *__restrict__ r,
                                                       const double *__restrict__ p, const double *__restrict__ Ap, DScalars *sc,
                                                       double *partials, unsigned *counter)
{
    __shared__ double smem[NR * TPB / 32];
    if (sc->done) return;
    const int i = blockIdx.x * TPB + threadIdx.x;
    double d0[NR], d1[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) d0[k] = d1[k] = 0.0;
    if (i < n) {
        const double id = invd[i];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const size_t q = (size_t)k * n + i;
            const double a = sc->alpha[k];
            x[q] += a * p[q];
            const double ri = r[q] - a * Ap[q];
            r[q] = ri;
            d0[k] = ri * ri;
            d1[k] = ri * ri * id;
        }
    }
    // one reduction of 12 values in two halves (the helper reduces six at a time)
    const bool fin0 = grid_reduce6(d0, partials, counter, smem);
    if (fin0) {
        for (int k = 0; k < NR; ++k) sc->rr[k] = d0[k];
    }
    __syncthreads();
    const bool fin1 = grid_reduce6(d1, partials + (size_t)gridDim.x * NR, counter + 1, smem);
    if (fin1) {
        for (int k = 0; k < NR; ++k) {
            sc->beta[k] = sc->rz[k] > 0.0 ? d1[k] / sc->rz[k] : 0.0;
            sc->rz[k] = d1[k];
        }
    }
}

// p = D^-1 r + beta p ; iteration count + stopping flag (single thread of block 0, after both reductions of d_update)
__global__ void __launch_bounds__(TPB) d_pupdate_kernel(int n, const double *__restrict__ invd, const double *__restrict__ r, double *__restrict__ p,
                                                        DScalars *sc, int *flag_scratch)
{
    if (sc->done) return;
    const int i = blockIdx.x * TPB + threadIdx.x;
    if (i < n) {
        const double id = invd[i];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const size_t q = (size_t)k * n + i;
            p[q] = r[q] * id + sc->beta[k] * p[q];
        }
    }
    (void)flag_scratch;
}

__global__ void d_step_kernel(DScalars *sc)
{
    if (sc->done) return;
    sc->iters += 1;
    sc->converged = all_converged(sc) ? 1 : 0;
    sc->done = (sc->converged || sc->iters >= sc->max_iters) ? 1 : 0;
}

}  // namespace

extern "C" int nct_solve_direct(nct_ctx *ctx, int nnz, int n, int one_based, const double *A, const int *row_index, const int *columns,
                                const double *const B[6], double *const X[6], double rel_tol, int max_iters, int *iters_out,
                                double *rel_res_out)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, nnz > 0 && n > 0 && A && row_index && columns && B && X, "bad arguments");
    for (int k = 0; k < NR; ++k) NCT_REQUIRE(ctx, B[k] && X[k], "null right-hand side / solution pointer %d", k);
    if (rel_tol <= 0) rel_tol = 1e-10;
    if (max_iters <= 0) max_iters = 100000;
    const int base = one_based ? 1 : 0;
    NCT_REQUIRE(ctx, row_index[n] - base == nnz, "rowIndex[n] (%d) does not match nonZeroNum (%d)", row_index[n] - base, nnz);
    // ---- host: upper triangle -> full symmetric CSR (zero-based); every row keeps ascending source order: its mirrored
    //      entries (columns < row, in ascending row order of their origin) first, then its own stored entries
    std::vector<int> cnt((size_t)n + 1, 0);
    std::vector<double> diag((size_t)n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int e = row_index[i] - base; e < row_index[i + 1] - base; ++e) {
            const int j = columns[e] - base;
            NCT_REQUIRE(ctx, j >= i && j < n, "entry (%d, %d) is not in the upper triangle of an %d x %d matrix", i, j, n, n);
            cnt[(size_t)i + 1]++;
            if (j != i) cnt[(size_t)j + 1]++;
            else diag[(size_t)i] += A[e];
        }
    for (int i = 0; i < n; ++i) {
        NCT_REQUIRE(ctx, diag[(size_t)i] > 0.0, "non-positive diagonal at row %d: the matrix is not SPD", i);
        cnt[(size_t)i + 1] += cnt[(size_t)i];
    }
    const size_t nz_full = (size_t)cnt[(size_t)n];
    std::vector<int> fcols(nz_full), fill(cnt.begin(), cnt.end() - 1);
    std::vector<double> fvals(nz_full), invd((size_t)n);
    for (int i = 0; i < n; ++i) {
        invd[(size_t)i] = 1.0 / diag[(size_t)i];
        for (int e = row_index[i] - base; e < row_index[i + 1] - base; ++e) {
            const int j = columns[e] - base;
            if (j != i) {   // mirrored entry lands in row j before row j's own entries are appended (i < j)
                fcols[(size_t)fill[(size_t)j]] = i;
                fvals[(size_t)fill[(size_t)j]++] = A[e];
            }
        }
        for (int e = row_index[i] - base; e < row_index[i + 1] - base; ++e) {
            fcols[(size_t)fill[(size_t)i]] = columns[e] - base;
            fvals[(size_t)fill[(size_t)i]++] = A[e];
        }
    }
    // ---- device
    const int blocks = nct_div_up(n, TPB);
    int *d_rowptr = (int *)nct_scratch(ctx, "dir_rowptr", sizeof(int) * ((size_t)n + 1));
    int *d_cols = (int *)nct_scratch(ctx, "dir_cols", sizeof(int) * nz_full);
    double *d_vals = (double *)nct_scratch(ctx, "dir_vals", sizeof(double) * nz_full);
    double *d_invd = (double *)nct_scratch(ctx, "dir_invd", sizeof(double) * (size_t)n);
    double *vec = (double *)nct_scratch(ctx, "dir_vec", sizeof(double) * (size_t)n * NR * 5);   // b, x, r, p, Ap
    double *partials = (double *)nct_scratch(ctx, "solver_partials", sizeof(double) * 18 * (size_t)(blocks + 1));
    char *misc = (char *)nct_scratch(ctx, "solver_misc", 1024);
    if (!d_rowptr || !d_cols || !d_vals || !d_invd || !vec || !partials || !misc) return NCT_ERR_NOMEM;
    DScalars *sc = (DScalars *)misc;
    unsigned *counter = (unsigned *)(misc + 512);
    static_assert(sizeof(DScalars) <= 512, "scalar block too large");
    double *b = vec, *x = vec + (size_t)n * NR, *r = vec + (size_t)n * NR * 2, *p = vec + (size_t)n * NR * 3, *Ap = vec + (size_t)n * NR * 4;
    NCT_CUDA(ctx, cudaMemcpyAsync(d_rowptr, cnt.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_cols, fcols.data(), sizeof(int) * nz_full, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_vals, fvals.data(), sizeof(double) * nz_full, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_invd, invd.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    for (int k = 0; k < NR; ++k)
        NCT_CUDA(ctx, cudaMemcpyAsync(b + (size_t)k * n, B[k], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    NCT_CUDA(ctx, cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned), ctx->stream));
    d_init_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, b, d_invd, x, r, p, sc, partials, counter, rel_tol * rel_tol, max_iters);
    NCT_CHECK_LAUNCH(ctx);
    d_flag_kernel<<<1, 1, 0, ctx->stream>>>(sc);
    NCT_CHECK_LAUNCH(ctx);
    DScalars hs;
    while (true) {
        for (int it = 0; it < 64; ++it) {   // queued iterations past the stopping point return at once (sc->done)
            d_spmv_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, d_rowptr, d_cols, d_vals, p, Ap, sc, partials, counter);
            NCT_CHECK_LAUNCH(ctx);
            d_update_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, d_invd, x, r, p, Ap, sc, partials, counter);
            NCT_CHECK_LAUNCH(ctx);
            d_pupdate_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, d_invd, r, p, sc, nullptr);
            NCT_CHECK_LAUNCH(ctx);
            d_step_kernel<<<1, 1, 0, ctx->stream>>>(sc);
            NCT_CHECK_LAUNCH(ctx);
        }
        NCT_CUDA(ctx, cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        NCT_CUDA(ctx, nct_stream_wait(ctx));
        if (hs.done) break;
    }
    for (int k = 0; k < NR; ++k)
        NCT_CUDA(ctx, cudaMemcpyAsync(X[k], x + (size_t)k * n, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    NCT_CUDA(ctx, nct_stream_wait(ctx));   // the host vectors above must outlive the uploads, the caller reads X next
    double worst = 0.0;
    for (int k = 0; k < NR; ++k) {
        const double rel = hs.bb[k] > 0.0 ? sqrt(hs.rr[k] / hs.bb[k]) : (hs.rr[k] > 0.0 ? 1.0 : 0.0);
        if (rel > worst) worst = rel;
    }
    if (iters_out) *iters_out = hs.iters;
    if (rel_res_out) *rel_res_out = worst;
    if (!hs.converged) return nct_fail(ctx, NCT_ERR_STATE, "nct_solve_direct: %d iterations reached at relative residual %.3e (target %.1e)", hs.iters, worst, rel_tol);
    return NCT_OK;
}
