// nct_solve_ls_cg: the reference's explicit-CSR least-squares entry point, for callers that keep its host assembly.
//
// Replaces solve_ls_cg_gpu (CT/SparseSolver_GPU.cu:3-198; declared CT/SparseSolver_GPU.cuh:12):
//     void solve_ls_cg_gpu(int size, int constraints, double* A, int* columns, int* rowindex, double* x, double* b,
//                          int nonzeros, double tolerance, int maxitrs)
// A is `constraints` x `size` in CSR with ONE-based rowindex / columns, all arrays on the HOST; x holds the start vector and
// receives the result.  The reference forms A^T A with cusparseDcsrgemm and runs un-preconditioned CG on
// A^T A x = A^T b: `k = 1; while (r1 > tol*tol && k <= maxit) {...; k++}` (:119-159).
//
// Here A^T A is never formed: each iteration applies q = A^T (A p) with two CSR products (A and a host-built transpose,
// both row-parallel with a sequential, hence deterministic, sum per row), the dot products are two-stage block
// reductions in a fixed order, and the loop control lives on the device (no host round trip per iteration -- the
// reference synchronises after every dot product).  Mathematically the same Krylov iteration; the rounding differs from
// cuSPARSE's (unspecified) summation order exactly as discussed in DESIGN.md section 6.  The pipeline itself uses the
// matrix-free nct_solve_nonlocal; this entry point exists for drop-in use of the reference's assembly code.
#include "nct_internal.h"
#include <vector>

namespace {

constexpr int TPB = 256;

struct LsScalars {
    double r1, r0, dot;
    int active, iters;
};

// y[row] = sum_k vals[k] * x[cols[k]] over the row's entries in storage order (zero-based arrays)
__global__ void csr_spmv_kernel(int rows, const int *__restrict__ rowptr, const int *__restrict__ cols, const double *__restrict__ vals,
                                const double *__restrict__ x, double *__restrict__ y, const LsScalars *sc)
{
    if (sc && !sc->active) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double s = 0.0;
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) s = __dadd_rn(s, __dmul_rn(vals[k], x[cols[k]]));
    y[r] = s;
}

__device__ __forceinline__ double block_sum(double v, double *smem)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) smem[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < TPB / 32; ++i) s += smem[i];
    __syncthreads();
    return s;  // valid in thread 0
}

// partials[block] = sum over the block of a[i] * b[i]
__global__ void dot_partial_kernel(int n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ partials,
                                   const LsScalars *sc)
{
    __shared__ double smem[TPB / 32];
    if (sc && !sc->active) return;
    const int i = blockIdx.x * TPB + threadIdx.x;
    const double v = i < n ? a[i] * b[i] : 0.0;
    const double s = block_sum(v, smem);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// one block: total of the partials in a fixed order, then the scalar update selected by `what`
//   0: r1 = total (initial residual), active = r1 > tol2 && maxit >= 1
//   1: dot = total (p . A^T A p)
//   2: r0 = r1, r1 = total, iters++, active = r1 > tol2 && iters < maxit
__global__ void dot_final_kernel(int nblocks, const double *__restrict__ partials, LsScalars *sc, int what, double tol2, int maxit)
{
    __shared__ double smem[TPB / 32];
    if (what != 0 && !sc->active) return;
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += TPB) acc += partials[b];
    const double total = block_sum(acc, smem);
    if (threadIdx.x == 0) {
        if (what == 0) {
            sc->r1 = total;
            sc->r0 = 0.0;
            sc->iters = 0;
            sc->active = (total > tol2 && maxit >= 1) ? 1 : 0;
        } else if (what == 1) {
            sc->dot = total;
        } else {
            sc->r0 = sc->r1;
            sc->r1 = total;
            sc->iters += 1;
            sc->active = (total > tol2 && sc->iters < maxit) ? 1 : 0;
        }
    }
}

// r = atb - q (initial residual)
__global__ void residual_kernel(int n, const double *__restrict__ atb, const double *__restrict__ q, double *__restrict__ r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = atb[i] - q[i];
}

// p = r (first iteration) or p = (r1 / r0) p + r
__global__ void p_update_kernel(int n, const double *__restrict__ r, double *__restrict__ p, const LsScalars *sc)
{
    if (!sc->active) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sc->iters == 0) p[i] = r[i];
    else p[i] = (sc->r1 / sc->r0) * p[i] + r[i];
}

// x += va p ; r -= va q, va = r1 / dot
__global__ void xr_update_kernel(int n, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ p,
                                 const double *__restrict__ q, const LsScalars *sc)
{
    if (!sc->active) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double va = sc->r1 / sc->dot;
    x[i] += va * p[i];
    r[i] -= va * q[i];
}

}  // namespace

extern "C" {

int nct_solve_ls_cg(nct_ctx *ctx, int size, int constraints, const double *A, const int *columns, const int *rowindex, double *x,
                    const double *b, int nonzeros, double tolerance, int maxitrs, int *iters_out)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, A && columns && rowindex && x && b, "null pointer");
    NCT_REQUIRE(ctx, size > 0 && constraints > 0 && nonzeros > 0, "empty system");
    NCT_REQUIRE(ctx, rowindex[0] == 1 && rowindex[constraints] == nonzeros + 1,
                "rowindex must be one-based with rowindex[constraints] == nonzeros + 1 (CUSPARSE_INDEX_BASE_ONE, CT/SparseSolver_GPU.cu:27)");
    // ---- zero-based copies, the transpose in CSR (stable counting sort: ascending constraint index inside a column)
    std::vector<int> rp((size_t)constraints + 1), ci((size_t)nonzeros);
    for (int r = 0; r <= constraints; ++r) {
        rp[r] = rowindex[r] - 1;
        if (r > 0 && rp[r] < rp[r - 1]) return nct_fail(ctx, NCT_ERR_ARG, "rowindex is not non-decreasing at row %d", r);
    }
    std::vector<int> trp((size_t)size + 1, 0);
    for (int k = 0; k < nonzeros; ++k) {
        const int c = columns[k] - 1;
        if (c < 0 || c >= size) return nct_fail(ctx, NCT_ERR_ARG, "column index %d out of range [1, %d] at entry %d", columns[k], size, k);
        ci[k] = c;
        trp[c + 1]++;
    }
    for (int c = 0; c < size; ++c) trp[c + 1] += trp[c];
    std::vector<int> tci((size_t)nonzeros), cursor(trp.begin(), trp.end() - 1);
    std::vector<double> tv((size_t)nonzeros);
    for (int r = 0; r < constraints; ++r)
        for (int k = rp[r]; k < rp[r + 1]; ++k) {
            const int pos = cursor[ci[k]]++;
            tci[pos] = r;
            tv[pos] = A[k];
        }

    // ---- device buffers
    const size_t nz = (size_t)nonzeros;
    int *d_rp = (int *)nct_scratch(ctx, "ls_rp", sizeof(int) * ((size_t)constraints + 1));
    int *d_ci = (int *)nct_scratch(ctx, "ls_ci", sizeof(int) * nz);
    double *d_v = (double *)nct_scratch(ctx, "ls_v", sizeof(double) * nz);
    int *d_trp = (int *)nct_scratch(ctx, "ls_trp", sizeof(int) * ((size_t)size + 1));
    int *d_tci = (int *)nct_scratch(ctx, "ls_tci", sizeof(int) * nz);
    double *d_tv = (double *)nct_scratch(ctx, "ls_tv", sizeof(double) * nz);
    double *d_b = (double *)nct_scratch(ctx, "ls_b", sizeof(double) * (size_t)constraints);
    double *d_t = (double *)nct_scratch(ctx, "ls_t", sizeof(double) * (size_t)constraints);
    double *d_vec = (double *)nct_scratch(ctx, "ls_vec", sizeof(double) * (size_t)size * 5);
    const int blocks = nct_div_up(size, TPB), cblocks = nct_div_up(constraints, TPB);
    double *d_part = (double *)nct_scratch(ctx, "ls_partials", sizeof(double) * (size_t)blocks);
    LsScalars *sc = (LsScalars *)nct_scratch(ctx, "ls_scalars", sizeof(LsScalars));
    if (!d_rp || !d_ci || !d_v || !d_trp || !d_tci || !d_tv || !d_b || !d_t || !d_vec || !d_part || !sc) return NCT_ERR_NOMEM;
    double *d_x = d_vec, *d_r = d_vec + size, *d_p = d_vec + 2 * (size_t)size, *d_q = d_vec + 3 * (size_t)size, *d_atb = d_vec + 4 * (size_t)size;
    cudaStream_t st = ctx->stream;
    NCT_CUDA(ctx, cudaMemcpyAsync(d_rp, rp.data(), sizeof(int) * rp.size(), cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_ci, ci.data(), sizeof(int) * nz, cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_v, A, sizeof(double) * nz, cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_trp, trp.data(), sizeof(int) * trp.size(), cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_tci, tci.data(), sizeof(int) * nz, cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_tv, tv.data(), sizeof(double) * nz, cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_b, b, sizeof(double) * (size_t)constraints, cudaMemcpyHostToDevice, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)size, cudaMemcpyHostToDevice, st));

    const double tol2 = tolerance * tolerance;
    // A^T b ; r = A^T b - A^T (A x0) ; r1 = r.r
    csr_spmv_kernel<<<blocks, TPB, 0, st>>>(size, d_trp, d_tci, d_tv, d_b, d_atb, nullptr);
    NCT_CHECK_LAUNCH(ctx);
    csr_spmv_kernel<<<cblocks, TPB, 0, st>>>(constraints, d_rp, d_ci, d_v, d_x, d_t, nullptr);
    NCT_CHECK_LAUNCH(ctx);
    csr_spmv_kernel<<<blocks, TPB, 0, st>>>(size, d_trp, d_tci, d_tv, d_t, d_q, nullptr);
    NCT_CHECK_LAUNCH(ctx);
    residual_kernel<<<blocks, TPB, 0, st>>>(size, d_atb, d_q, d_r);
    NCT_CHECK_LAUNCH(ctx);
    dot_partial_kernel<<<blocks, TPB, 0, st>>>(size, d_r, d_r, d_part, nullptr);
    NCT_CHECK_LAUNCH(ctx);
    dot_final_kernel<<<1, TPB, 0, st>>>(blocks, d_part, sc, 0, tol2, maxitrs);
    NCT_CHECK_LAUNCH(ctx);
    for (int k = 1; k <= maxitrs; ++k) {
        p_update_kernel<<<blocks, TPB, 0, st>>>(size, d_r, d_p, sc);
        NCT_CHECK_LAUNCH(ctx);
        csr_spmv_kernel<<<cblocks, TPB, 0, st>>>(constraints, d_rp, d_ci, d_v, d_p, d_t, sc);
        NCT_CHECK_LAUNCH(ctx);
        csr_spmv_kernel<<<blocks, TPB, 0, st>>>(size, d_trp, d_tci, d_tv, d_t, d_q, sc);
        NCT_CHECK_LAUNCH(ctx);
        dot_partial_kernel<<<blocks, TPB, 0, st>>>(size, d_p, d_q, d_part, sc);
        NCT_CHECK_LAUNCH(ctx);
        dot_final_kernel<<<1, TPB, 0, st>>>(blocks, d_part, sc, 1, tol2, maxitrs);
        NCT_CHECK_LAUNCH(ctx);
        xr_update_kernel<<<blocks, TPB, 0, st>>>(size, d_x, d_r, d_p, d_q, sc);
        NCT_CHECK_LAUNCH(ctx);
        dot_partial_kernel<<<blocks, TPB, 0, st>>>(size, d_r, d_r, d_part, sc);
        NCT_CHECK_LAUNCH(ctx);
        dot_final_kernel<<<1, TPB, 0, st>>>(blocks, d_part, sc, 2, tol2, maxitrs);
        NCT_CHECK_LAUNCH(ctx);
    }
    LsScalars hs;
    NCT_CUDA(ctx, cudaMemcpyAsync(x, d_x, sizeof(double) * (size_t)size, cudaMemcpyDeviceToHost, st));
    NCT_CUDA(ctx, cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, st));
    NCT_CUDA(ctx, nct_stream_wait(ctx));  // the host vectors above must outlive the copies, and x is a host result
    if (iters_out) *iters_out = hs.iters;
    return NCT_OK;
}

}  // extern "C"
