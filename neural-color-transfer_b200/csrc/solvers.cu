// FP64 sparse solves of the colour stage, matrix-free on the GPU.
//
// (1) nct_solve_nonlocal replaces solve_nonlocal_downsample_gpu_gradient + solve_ls_cg_gpu
//     (CT/ColorTransfer.cpp:548-949, CT/SparseSolver_GPU.cu:3-198).  The reference assembles an explicit
//     constraint matrix A on the host (serial, ~50 nnz per pixel, three value arrays), uploads it three times,
//     forms A^T A with cuSPARSE csrgemm and runs un-preconditioned CG with one cuBLAS/cuSPARSE call per
//     vector operation.  Here nothing is assembled: A^T A p is evaluated per pixel from the 3x4-neighbour
//     Laplacian weights, the 8 forward + the reverse non-local links and the data term; the three channels'
//     CG recurrences run side by side (they share everything but the data coefficient), two kernels per
//     iteration, scalars (alpha, beta, r.r) stay on the device.  Same start vector, same iteration budget,
//     same stopping rule `while (r1 > tol^2 && k <= maxit)`, FP64.
// (2) nct_solve_wls replaces solve_WLS_roughness_cpu + solve_direct_cpu (MKL PARDISO)
//     (CT/ColorTransfer.cpp:951-1125, CT/SparseSolver_CPU.cpp:104-286): the full-resolution SPD system
//     (W + L_g) x = W x0 for the six coefficient maps, solved by preconditioned CG to a relative residual
//     that makes it indistinguishable from the direct solve at the parity tolerance.
//
// Both are bandwidth-bound (SpMV-like, ~11 vector passes of 6 doubles per pixel per iteration).
#include "device_utils.cuh"

namespace {

constexpr int TPB = 256;

// deterministic block reduction of NV doubles per thread -> thread 0 holds the sums
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *smem /* NV * TPB/32 */)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) smem[k * (TPB / 32) + w] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int i = 0; i < TPB / 32; ++i) s += smem[k * (TPB / 32) + i];
            v[k] = s;
        }
    }
    __syncthreads();
}

// "last block done" pattern: every block stores its partial sums, the last one to arrive adds all of them
// in a fixed order (deterministic) and returns true on its thread 0 with the totals in `v`.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials, unsigned *counter, double *smem)
{
    __shared__ bool last;
    block_reduce<NV>(v, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[(size_t)blockIdx.x * NV + k] = v[k];
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += TPB) {
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += __ldcg(&partials[(size_t)b * NV + k]);
    }
    block_reduce<NV>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = acc[k];
        *counter = 0;
        return true;
    }
    return false;
}

// ====================================================================== non-local least squares (CG)
struct NlScalars {
    double r1[3], r0[3], alpha[3], beta[3];
    int active[3];
    int iters[3];
};

struct NlSystem {
    int n, h, w;
    const uint8_t *src;   // level content Lab, 8 bit (s = u8 / 255)
    const uint8_t *ref;   // level BDS-reconstructed style Lab, 8 bit
    const double *d2;     // data weight^2 per pixel
    const double *wx2;    // 2 g^2 of edge (p, p+1)
    const double *wy2;    // 2 g^2 of edge (p, p+w)
    const int *knn_id;    // [n][8], -1 = none
    const double *knn_w2; // [n][8] link weight^2
    const int *rev_start; // [n+1]
    const int *rev_src;   // reverse links: source pixel
    const double *rev_w2;
};

__global__ void nl_setup_kernel(const double *__restrict__ weight, const uint8_t *__restrict__ src, int h, int w, double lam,
                                const double *__restrict__ ptab, double sqrt_dw, double *__restrict__ d2, double *__restrict__ wx2,
                                double *__restrict__ wy2)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int x = p % w, y = p / w;
    const double dw = __dmul_rn(__dsqrt_rn(weight[p]), sqrt_dw);
    d2[p] = __dmul_rn(dw, dw);
    const double L = __dmul_rn((double)src[(size_t)p * 3], 1.0 / 255.0);
    double vx = 0.0, vy = 0.0;
    if (x + 1 < w) {
        const double g = __dsqrt_rn(__ddiv_rn(lam, __dadd_rn(ptab[(int)src[(size_t)p * 3] * 256 + (int)src[(size_t)(p + 1) * 3]], 1e-4)));
        const double gg = __dmul_rn(g, g);
        vx = __dadd_rn(gg, gg);
    }
    if (y + 1 < h) {
        const double g = __dsqrt_rn(__ddiv_rn(lam, __dadd_rn(ptab[(int)src[(size_t)p * 3] * 256 + (int)src[(size_t)(p + w) * 3]], 1e-4)));
        const double gg = __dmul_rn(g, g);
        vy = __dadd_rn(gg, gg);
    }
    wx2[p] = vx;
    wy2[p] = vy;
}

// forward link weights: iw = sqrt(w) * sqrt(nl / k); w2 = iw * iw; count targets for the reverse lists
__global__ void nl_links_kernel(const int *__restrict__ knn_id, const double *__restrict__ knn_w, int n, int k, double nlw,
                                double *__restrict__ w2, int *__restrict__ count)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int id = knn_id[t];
    double v = 0.0;
    if (id >= 0 && id < n) {
        const double iw = __dmul_rn(__dsqrt_rn(knn_w[t]), nlw);
        v = __dmul_rn(iw, iw);
        atomicAdd(&count[id], 1);
    }
    w2[t] = v;
}

__global__ void nl_rev_keys_kernel(const int *__restrict__ knn_id, int n, int k, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const int id = knn_id[t];
    keys[t] = (id >= 0 && id < n) ? (uint32_t)id : (uint32_t)n;  // invalid links sort to the end
    vals[t] = (uint32_t)t;
}

__global__ void nl_rev_fill_kernel(const uint32_t *__restrict__ sorted_vals, const double *__restrict__ w2, int total, int k,
                                   int *__restrict__ rev_src, double *__restrict__ rev_w2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const uint32_t flat = sorted_vals[t];
    rev_src[t] = (int)(flat / (uint32_t)k);
    rev_w2[t] = w2[flat];
}

// CG vectors are [pixel][6] records (a0 a1 a2 b0 b1 b2, 48 bytes): a neighbour gather touches two 32-byte sectors
// (the first version kept the a and b blocks apart and recomputed p = r + beta p_old per neighbour: 8 sectors)
__device__ __forceinline__ void load6(const double *__restrict__ v, int n, int i, double (&o)[6])
{
    (void)n;
    const double2 *q = reinterpret_cast<const double2 *>(v + (size_t)i * 6);
    const double2 t0 = q[0], t1 = q[1], t2 = q[2];
    o[0] = t0.x; o[1] = t0.y; o[2] = t1.x; o[3] = t1.y; o[4] = t2.x; o[5] = t2.y;
}
__device__ __forceinline__ void store6(double *__restrict__ v, int n, int i, const double (&o)[6])
{
    (void)n;
    double2 *q = reinterpret_cast<double2 *>(v + (size_t)i * 6);
    q[0] = make_double2(o[0], o[1]);
    q[1] = make_double2(o[2], o[3]);
    q[2] = make_double2(o[4], o[5]);
}

// y = (A^T A) x at pixel i, x given through a functor returning the 6 values of a pixel
template <class GetX>
__device__ __forceinline__ void nl_apply(const NlSystem &S, int i, const double (&xi)[6], GetX getx, double (&out)[6])
{
    const int w = S.w, h = S.h, n = S.n;
    const int x = i % w, y = i / w;
    const double d2 = S.d2[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double s = __dmul_rn((double)S.src[(size_t)i * 3 + c], 1.0 / 255.0);
        const double t = __dadd_rn(__dmul_rn(s, xi[c]), xi[3 + c]);
        const double dt = __dmul_rn(d2, t);
        out[c] = __dmul_rn(dt, s);
        out[3 + c] = dt;
    }
    auto link = [&](int j, double w2) {
        double xj[6];
        getx(j, xj);
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k] = __dadd_rn(out[k], __dmul_rn(w2, __dsub_rn(xi[k], xj[k])));
    };
    // Non-local links, four at a time: ids and weights first, then the four neighbour records, then the accumulation in
    // the canonical order (k ascending, then ascending reverse-list position).  The arithmetic is that of one link()
    // per neighbour; only the loads are issued together -- one link at a time is a chain of two dependent gathers per
    // link (id -> record), 16-20 round trips per pixel, which was all of this kernel's time at the coarse levels.
    auto link4 = [&](const int (&j)[4], const double (&w2)[4], const bool (&ok)[4]) {
        double xj[4][6];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) getx(j[u], xj[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) {
#pragma unroll
                for (int k = 0; k < 6; ++k) out[k] = __dadd_rn(out[k], __dmul_rn(w2[u], __dsub_rn(xi[k], xj[u][k])));
            }
    };
    {   // the 4-neighbour links x+1, x-1, y+1, y-1 (in this order)
        const bool ok[4] = {x + 1 < w, x > 0, y + 1 < h, y > 0};
        const int j[4] = {i + 1, i - 1, i + w, i - w};
        const double w2[4] = {ok[0] ? S.wx2[i] : 0.0, ok[1] ? S.wx2[i - 1] : 0.0, ok[2] ? S.wy2[i] : 0.0, ok[3] ? S.wy2[i - w] : 0.0};
        link4(j, w2, ok);
    }
#pragma unroll 1
    for (int k0 = 0; k0 < 8; k0 += 4) {
        const int4 jv = *reinterpret_cast<const int4 *>(S.knn_id + (size_t)i * 8 + k0);
        const double2 wa = *reinterpret_cast<const double2 *>(S.knn_w2 + (size_t)i * 8 + k0);
        const double2 wb = *reinterpret_cast<const double2 *>(S.knn_w2 + (size_t)i * 8 + k0 + 2);
        const int j[4] = {jv.x, jv.y, jv.z, jv.w};
        const double w2[4] = {wa.x, wa.y, wb.x, wb.y};
        const bool ok[4] = {j[0] >= 0 && j[0] < n, j[1] >= 0 && j[1] < n, j[2] >= 0 && j[2] < n, j[3] >= 0 && j[3] < n};
        link4(j, w2, ok);
    }
    const int e = S.rev_start[i + 1];
#pragma unroll 1
    for (int t = S.rev_start[i]; t < e; t += 4) {
        int j[4];
        double w2[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = t + u < e;
            j[u] = ok[u] ? S.rev_src[t + u] : 0;
            w2[u] = ok[u] ? S.rev_w2[t + u] : 0.0;
        }
        link4(j, w2, ok);
    }
}

// r = A^T b - A^T A x0 ; r1 = r.r per channel ; p_old = 0
__global__ void __launch_bounds__(TPB) nl_init_kernel(NlSystem S, const double *__restrict__ x, double *__restrict__ r,
                                                      double *__restrict__ pold, NlScalars *sc, double tol2,
                                                      double *partials, unsigned *counter)
{
    __shared__ double smem[3 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double dots[3] = {0.0, 0.0, 0.0};
    if (i < S.n) {
        double xi[6], ax[6], ri[6];
        load6(x, S.n, i, xi);
        nl_apply(S, i, xi, [&](int j, double (&o)[6]) { load6(x, S.n, j, o); }, ax);
        const double d2 = S.d2[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double s = __dmul_rn((double)S.src[(size_t)i * 3 + c], 1.0 / 255.0);
            const double rr = __dmul_rn((double)S.ref[(size_t)i * 3 + c], 1.0 / 255.0);
            const double db = __dmul_rn(d2, rr);
            ri[c] = __dsub_rn(__dmul_rn(db, s), ax[c]);
            ri[3 + c] = __dsub_rn(db, ax[3 + c]);
            dots[c] = __dadd_rn(__dmul_rn(ri[c], ri[c]), __dmul_rn(ri[3 + c], ri[3 + c]));
        }
        store6(r, S.n, i, ri);
        const double z[6] = {0, 0, 0, 0, 0, 0};
        store6(pold, S.n, i, z);
    }
    if (grid_reduce<3>(dots, partials, counter, smem)) {
        for (int c = 0; c < 3; ++c) {
            sc->r1[c] = dots[c];
            sc->r0[c] = 0.0;
            sc->alpha[c] = 0.0;
            sc->beta[c] = 0.0;
            sc->active[c] = dots[c] > tol2 ? 1 : 0;
            sc->iters[c] = 0;
        }
    }
}

// p = beta p + r (in place, element-wise)
__global__ void __launch_bounds__(TPB) nl_pupdate_kernel(int n, const double *__restrict__ r, double *__restrict__ p, const NlScalars *sc)
{
    const int i = blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const double beta[3] = {sc->beta[0], sc->beta[1], sc->beta[2]};
    double ri[6], pi[6];
    load6(r, n, i, ri);
    load6(p, n, i, pi);
#pragma unroll
    for (int k = 0; k < 6; ++k) pi[k] = __dadd_rn(__dmul_rn(beta[k % 3], pi[k]), ri[k]);
    store6(p, n, i, pi);
}

// Ap = (A^T A) p ; alpha = r1 / (p.Ap)
__global__ void __launch_bounds__(TPB) nl_spmv_kernel(NlSystem S, const double *__restrict__ p, double *__restrict__ Ap, NlScalars *sc,
                                                      double *partials, unsigned *counter)
{
    __shared__ double smem[3 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double dots[3] = {0.0, 0.0, 0.0};
    if (i < S.n) {
        double pi[6], api[6];
        load6(p, S.n, i, pi);
        nl_apply(S, i, pi, [&](int j, double (&o)[6]) { load6(p, S.n, j, o); }, api);
        store6(Ap, S.n, i, api);
#pragma unroll
        for (int c = 0; c < 3; ++c) dots[c] = __dadd_rn(__dmul_rn(pi[c], api[c]), __dmul_rn(pi[3 + c], api[3 + c]));
    }
    if (grid_reduce<3>(dots, partials, counter, smem)) {
        for (int c = 0; c < 3; ++c) sc->alpha[c] = sc->active[c] ? sc->r1[c] / dots[c] : 0.0;
    }
}

// x += alpha p ; r -= alpha Ap ; r0 = r1 ; r1 = r.r ; beta = r1 / r0
__global__ void __launch_bounds__(TPB) nl_update_kernel(int n, double *__restrict__ x, double *__restrict__ r,
                                                        const double *__restrict__ p, const double *__restrict__ Ap,
                                                        NlScalars *sc, double tol2, double *partials, unsigned *counter)
{
    __shared__ double smem[3 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    const double alpha[3] = {sc->alpha[0], sc->alpha[1], sc->alpha[2]};
    double dots[3] = {0.0, 0.0, 0.0};
    if (i < n) {
        double xi[6], ri[6], pi[6], api[6];
        load6(x, n, i, xi);
        load6(r, n, i, ri);
        load6(p, n, i, pi);
        load6(Ap, n, i, api);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            xi[k] = __dadd_rn(xi[k], __dmul_rn(alpha[k % 3], pi[k]));
            ri[k] = __dsub_rn(ri[k], __dmul_rn(alpha[k % 3], api[k]));
        }
        store6(x, n, i, xi);
        store6(r, n, i, ri);
#pragma unroll
        for (int c = 0; c < 3; ++c) dots[c] = __dadd_rn(__dmul_rn(ri[c], ri[c]), __dmul_rn(ri[3 + c], ri[3 + c]));
    }
    if (grid_reduce<3>(dots, partials, counter, smem)) {
        for (int c = 0; c < 3; ++c) {
            if (sc->active[c]) {
                sc->r0[c] = sc->r1[c];
                sc->r1[c] = dots[c];
                sc->beta[c] = dots[c] / sc->r0[c];
                sc->iters[c] += 1;
                sc->active[c] = dots[c] > tol2 ? 1 : 0;
            }
        }
    }
}

__global__ void pack_ab_kernel(const double *__restrict__ a, const double *__restrict__ b, int n3, double *__restrict__ x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = pixel * 3 + channel
    if (i < n3) {
        const int px = i / 3, c = i % 3;
        x[(size_t)px * 6 + c] = a[i];
        x[(size_t)px * 6 + 3 + c] = b[i];
    }
}
__global__ void unpack_ab_kernel(const double *__restrict__ x, int n3, double *__restrict__ a, double *__restrict__ b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3) {
        const int px = i / 3, c = i % 3;
        a[i] = x[(size_t)px * 6 + c];
        b[i] = x[(size_t)px * 6 + 3 + c];
    }
}

// ====================================================================== WLS (Jacobi-preconditioned CG, 6 RHS)
struct WlsScalars {
    double rz[6], rz_old[6], alpha[6], beta[6], rr[6], bb[6];
    int iters;
};

__global__ void wls_setup_kernel(const uint8_t *__restrict__ lab, int H, int W, double lam, const double *__restrict__ ptab,
                                 double *__restrict__ wx, double *__restrict__ wy)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int x = p % W, y = p / W;
    const double L = __dmul_rn((double)lab[(size_t)p * 3], 1.0 / 255.0);
    double vx = 0.0, vy = 0.0;
    if (x + 1 < W) {
        const double g = __dsqrt_rn(__ddiv_rn(lam, __dadd_rn(ptab[(int)lab[(size_t)p * 3] * 256 + (int)lab[(size_t)(p + 1) * 3]], 1e-4)));
        vx = __dmul_rn(g, g);
    }
    if (y + 1 < H) {
        const double g = __dsqrt_rn(__ddiv_rn(lam, __dadd_rn(ptab[(int)lab[(size_t)p * 3] * 256 + (int)lab[(size_t)(p + W) * 3]], 1e-4)));
        vy = __dmul_rn(g, g);
    }
    wx[p] = vx;
    wy[p] = vy;
}

struct WlsSystem {
    int n, H, W;
    const double *rough, *wx, *wy;
    double *inv_diag;
};

template <class GetX>
__device__ __forceinline__ void wls_apply(const WlsSystem &S, int i, const double (&xi)[6], GetX getx, double (&out)[6],
                                          double *diag_out)
{
    const int W = S.W, H = S.H;
    const int x = i % W, y = i / W;
    const double rg = S.rough[i];
    double diag = rg;
#pragma unroll
    for (int k = 0; k < 6; ++k) out[k] = rg * xi[k];
    auto link = [&](int j, double w) {
        double xj[6];
        getx(j, xj);
        diag += w;
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k] += w * (xi[k] - xj[k]);
    };
    if (x + 1 < W) link(i + 1, S.wx[i]);
    if (x > 0) link(i - 1, S.wx[i - 1]);
    if (y + 1 < H) link(i + W, S.wy[i]);
    if (y > 0) link(i - W, S.wy[i - W]);
    if (diag_out) *diag_out = diag;
}

// x[p][6] interleaved here (a0,a1,a2,b0,b1,b2): one 48-byte record per pixel
__device__ __forceinline__ void ld6(const double *__restrict__ v, int i, double (&o)[6])
{
    const double2 *q = reinterpret_cast<const double2 *>(v + (size_t)i * 6);
    const double2 t0 = q[0], t1 = q[1], t2 = q[2];
    o[0] = t0.x; o[1] = t0.y; o[2] = t1.x; o[3] = t1.y; o[4] = t2.x; o[5] = t2.y;
}
__device__ __forceinline__ void st6(double *__restrict__ v, int i, const double (&o)[6])
{
    double2 *q = reinterpret_cast<double2 *>(v + (size_t)i * 6);
    q[0] = make_double2(o[0], o[1]);
    q[1] = make_double2(o[2], o[3]);
    q[2] = make_double2(o[4], o[5]);
}

__global__ void wls_pack_kernel(const double *__restrict__ a, const double *__restrict__ b, int n, double *__restrict__ x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double o[6] = {a[(size_t)i * 3], a[(size_t)i * 3 + 1], a[(size_t)i * 3 + 2], b[(size_t)i * 3], b[(size_t)i * 3 + 1], b[(size_t)i * 3 + 2]};
    st6(x, i, o);
}
__global__ void wls_unpack_kernel(const double *__restrict__ x, int n, double *__restrict__ a, double *__restrict__ b)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[6];
    ld6(x, i, o);
    a[(size_t)i * 3] = o[0]; a[(size_t)i * 3 + 1] = o[1]; a[(size_t)i * 3 + 2] = o[2];
    b[(size_t)i * 3] = o[3]; b[(size_t)i * 3 + 1] = o[4]; b[(size_t)i * 3 + 2] = o[5];
}

// rhs = W x0 ; r = rhs - M x0 ; z = r / diag ; rz = r.z ; rr = r.r ; bb = rhs.rhs ; p_old = 0
__global__ void __launch_bounds__(TPB) wls_init_kernel(WlsSystem S, const double *__restrict__ x, double *__restrict__ r,
                                                       double *__restrict__ z, double *__restrict__ pold, WlsScalars *sc,
                                                       double *partials, unsigned *counter)
{
    __shared__ double smem[18 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double dots[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) dots[k] = 0.0;
    if (i < S.n) {
        double xi[6], mx[6], ri[6], zi[6], diag;
        ld6(x, i, xi);
        wls_apply(S, i, xi, [&](int j, double (&o)[6]) { ld6(x, j, o); }, mx, &diag);
        const double inv = 1.0 / diag;
        S.inv_diag[i] = inv;
        const double rg = S.rough[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double rhs = rg * xi[k];
            ri[k] = rhs - mx[k];
            zi[k] = ri[k] * inv;
            dots[k] = ri[k] * zi[k];
            dots[6 + k] = ri[k] * ri[k];
            dots[12 + k] = rhs * rhs;
        }
        st6(r, i, ri);
        st6(z, i, zi);
        const double zero[6] = {0, 0, 0, 0, 0, 0};
        st6(pold, i, zero);
    }
    if (grid_reduce<18>(dots, partials, counter, smem)) {
        for (int k = 0; k < 6; ++k) {
            sc->rz[k] = dots[k];
            sc->rz_old[k] = 0.0;
            sc->rr[k] = dots[6 + k];
            sc->bb[k] = dots[12 + k];
            sc->alpha[k] = 0.0;
            sc->beta[k] = 0.0;
        }
        sc->iters = 0;
    }
}

// p = z + beta p_old ; Ap = M p ; alpha = rz / p.Ap
__global__ void __launch_bounds__(TPB) wls_spmv_kernel(WlsSystem S, const double *__restrict__ z, const double *__restrict__ pold,
                                                       double *__restrict__ pnew, double *__restrict__ Ap, WlsScalars *sc,
                                                       double *partials, unsigned *counter)
{
    __shared__ double smem[6 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double beta[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) beta[k] = sc->beta[k];
    double dots[6] = {0, 0, 0, 0, 0, 0};
    auto getp = [&](int j, double (&o)[6]) {
        double zj[6], pj[6];
        ld6(z, j, zj);
        ld6(pold, j, pj);
#pragma unroll
        for (int k = 0; k < 6; ++k) o[k] = zj[k] + beta[k] * pj[k];
    };
    if (i < S.n) {
        double pi[6], api[6];
        getp(i, pi);
        st6(pnew, i, pi);
        wls_apply(S, i, pi, getp, api, nullptr);
        st6(Ap, i, api);
#pragma unroll
        for (int k = 0; k < 6; ++k) dots[k] = pi[k] * api[k];
    }
    if (grid_reduce<6>(dots, partials, counter, smem)) {
        for (int k = 0; k < 6; ++k) sc->alpha[k] = dots[k] > 0.0 ? sc->rz[k] / dots[k] : 0.0;
    }
}

// x += alpha p ; r -= alpha Ap ; z = r / diag ; rz, rr ; beta = rz_new / rz
__global__ void __launch_bounds__(TPB) wls_update_kernel(int n, const double *__restrict__ inv_diag, double *__restrict__ x,
                                                         double *__restrict__ r, double *__restrict__ z,
                                                         const double *__restrict__ p, const double *__restrict__ Ap,
                                                         WlsScalars *sc, double *partials, unsigned *counter)
{
    __shared__ double smem[12 * TPB / 32];
    const int i = blockIdx.x * TPB + threadIdx.x;
    double alpha[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) alpha[k] = sc->alpha[k];
    double dots[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) dots[k] = 0.0;
    if (i < n) {
        double xi[6], ri[6], pi[6], api[6], zi[6];
        ld6(x, i, xi);
        ld6(r, i, ri);
        ld6(p, i, pi);
        ld6(Ap, i, api);
        const double inv = inv_diag[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            xi[k] += alpha[k] * pi[k];
            ri[k] -= alpha[k] * api[k];
            zi[k] = ri[k] * inv;
            dots[k] = ri[k] * zi[k];
            dots[6 + k] = ri[k] * ri[k];
        }
        st6(x, i, xi);
        st6(r, i, ri);
        st6(z, i, zi);
    }
    if (grid_reduce<12>(dots, partials, counter, smem)) {
        for (int k = 0; k < 6; ++k) {
            sc->rz_old[k] = sc->rz[k];
            sc->rz[k] = dots[k];
            sc->beta[k] = sc->rz_old[k] > 0.0 ? dots[k] / sc->rz_old[k] : 0.0;
            sc->rr[k] = dots[6 + k];
        }
        sc->iters += 1;
    }
}

}  // namespace

int nct_sort_pairs_u32(nct_ctx *ctx, const uint32_t *keys_in, uint32_t *keys_out, const uint32_t *vals_in, uint32_t *vals_out,
                       int n, int end_bit);

extern "C" {

int nct_solve_nonlocal(nct_ctx *ctx, double *a_dev, double *b_dev, const double *weight_dev, const uint8_t *cnt_lab_dev,
                       const uint8_t *stl_lab_dev, const int *knn_id_dev, const double *knn_w_dev, int h, int w, int layer,
                       double local_weight, double alpha, double nonlocal_weight, int knum, double d_weight, int iters_out[3])
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, a_dev && b_dev && weight_dev && cnt_lab_dev && stl_lab_dev && knn_id_dev && knn_w_dev, "null pointer");
    NCT_REQUIRE(ctx, h > 0 && w > 0 && knum == 8, "bad size / only k = 8 neighbours supported (CT/Config.h:69)");
    const int n = h * w, n3 = 3 * n;
    const int blocks = nct_div_up(n, TPB);
    // float parameters of the reference's signature (CT/ColorTransfer.cpp:548-550)
    const double lam = (double)(float)local_weight, alpha_f = (double)(float)alpha;
    const double sqrt_dw = (double)sqrtf((float)d_weight);
    const double nlw = sqrt(nonlocal_weight / (double)knum);

    double *d2 = (double *)nct_scratch(ctx, "nl_d2", sizeof(double) * n);
    double *wx2 = (double *)nct_scratch(ctx, "nl_wx2", sizeof(double) * n);
    double *wy2 = (double *)nct_scratch(ctx, "nl_wy2", sizeof(double) * n);
    double *kw2 = (double *)nct_scratch(ctx, "nl_kw2", sizeof(double) * (size_t)n * 8);
    int *count = (int *)nct_scratch(ctx, "nl_count", sizeof(int) * ((size_t)n + 1));
    int *rstart = (int *)nct_scratch(ctx, "nl_rstart", sizeof(int) * ((size_t)n + 1));
    uint32_t *keys = (uint32_t *)nct_scratch(ctx, "nl_keys", sizeof(uint32_t) * (size_t)n * 8 * 4);
    int *rsrc = (int *)nct_scratch(ctx, "nl_rsrc", sizeof(int) * (size_t)n * 8);
    double *rw2 = (double *)nct_scratch(ctx, "nl_rw2", sizeof(double) * (size_t)n * 8);
    double *vec = (double *)nct_scratch(ctx, "nl_vec", sizeof(double) * (size_t)n * 6 * 5);
    double *partials = (double *)nct_scratch(ctx, "solver_partials", sizeof(double) * 18 * (size_t)(blocks + 1));
    char *misc = (char *)nct_scratch(ctx, "solver_misc", 1024);
    if (!d2 || !wx2 || !wy2 || !kw2 || !count || !rstart || !keys || !rsrc || !rw2 || !vec || !partials || !misc) return NCT_ERR_NOMEM;
    NlScalars *sc = (NlScalars *)misc;
    unsigned *counter = (unsigned *)(misc + 512);
    uint32_t *keys_out = keys + (size_t)n * 8, *vals = keys + (size_t)n * 16, *vals_out = keys + (size_t)n * 24;
    double *x = vec, *r = vec + (size_t)n * 6, *p0 = vec + (size_t)n * 12, *p1 = vec + (size_t)n * 18, *Ap = vec + (size_t)n * 24;

    NCT_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    NCT_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
    const double *ptab = nct_pow_table(ctx, alpha_f);
    if (!ptab) return NCT_ERR_NOMEM;
    nl_setup_kernel<<<blocks, TPB, 0, ctx->stream>>>(weight_dev, cnt_lab_dev, h, w, lam, ptab, sqrt_dw, d2, wx2, wy2);
    NCT_CHECK_LAUNCH(ctx);
    nl_links_kernel<<<nct_div_up(n * 8, TPB), TPB, 0, ctx->stream>>>(knn_id_dev, knn_w_dev, n, 8, nlw, kw2, count);
    NCT_CHECK_LAUNCH(ctx);
    int rc = nct_exclusive_scan_i32(ctx, count, rstart, n);
    if (rc) return rc;
    nl_rev_keys_kernel<<<nct_div_up(n * 8, TPB), TPB, 0, ctx->stream>>>(knn_id_dev, n, 8, keys, vals);
    NCT_CHECK_LAUNCH(ctx);
    int bits = 1;
    while ((1 << bits) <= n) bits++;
    rc = nct_sort_pairs_u32(ctx, keys, keys_out, vals, vals_out, n * 8, bits);
    if (rc) return rc;
    nl_rev_fill_kernel<<<nct_div_up(n * 8, TPB), TPB, 0, ctx->stream>>>(vals_out, kw2, n * 8, 8, rsrc, rw2);
    NCT_CHECK_LAUNCH(ctx);

    NlSystem S{n, h, w, cnt_lab_dev, stl_lab_dev, d2, wx2, wy2, knn_id_dev, kw2, rstart, rsrc, rw2};
    pack_ab_kernel<<<nct_div_up(n3, TPB), TPB, 0, ctx->stream>>>(a_dev, b_dev, n3, x);
    NCT_CHECK_LAUNCH(ctx);
    const double tol = 1e-6, tol2 = tol * tol;
    const int maxit = layer == 4 ? 50 : 100;  // CT/ColorTransfer.cpp:917
    nl_init_kernel<<<blocks, TPB, 0, ctx->stream>>>(S, x, r, p0, sc, tol2, partials, counter);
    NCT_CHECK_LAUNCH(ctx);
    (void)p1;
    // maxit iterations of 3 launches each, all with device-side `active` flags (no host check): queued as replays of one
    // captured block of 10 iterations (NCT_NL_GRAPH=0: plain stream launches, for ncu launch lists)
    auto iteration = [&]() -> int {
        nl_pupdate_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, r, p0, sc);
        NCT_CHECK_LAUNCH(ctx);
        nl_spmv_kernel<<<blocks, TPB, 0, ctx->stream>>>(S, p0, Ap, sc, partials, counter);
        NCT_CHECK_LAUNCH(ctx);
        nl_update_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, x, r, p0, Ap, sc, tol2, partials, counter);
        NCT_CHECK_LAUNCH(ctx);
        return NCT_OK;
    };
    const bool use_graph = !(getenv("NCT_NL_GRAPH") && atoi(getenv("NCT_NL_GRAPH")) == 0);
    constexpr int kBlock = 10;
    int k = 1;
    if (use_graph && maxit >= kBlock) {
        char gname[32];
        snprintf(gname, sizeof(gname), "nl_cg_block_l%d", layer);   // one graph per pyramid level: a pair never re-captures
        auto P = [](const void *q) { return (unsigned long long)(uintptr_t)q; };
        std::vector<unsigned long long> key = {(unsigned long long)n, (unsigned long long)h, (unsigned long long)w, P(cnt_lab_dev), P(stl_lab_dev),
                                               P(d2), P(wx2), P(wy2), P(knn_id_dev), P(kw2), P(rstart), P(rsrc), P(rw2), P(vec), P(partials), P(misc)};
        if (!nct_graph_cached(ctx, gname, key)) {
            rc = nct_graph_begin(ctx, gname, key, nullptr);
            if (rc) return rc;
            for (int q = 0; q < kBlock; ++q) {
                rc = iteration();
                if (rc) return nct_graph_abort(ctx, gname, rc);
            }
            rc = nct_graph_end(ctx, gname);
            if (rc) return rc;
        }
        for (; k + kBlock - 1 <= maxit; k += kBlock) {
            rc = nct_graph_launch(ctx, gname);
            if (rc) return rc;
        }
    }
    for (; k <= maxit; ++k) {
        rc = iteration();
        if (rc) return rc;
    }
    unpack_ab_kernel<<<nct_div_up(n3, TPB), TPB, 0, ctx->stream>>>(x, n3, a_dev, b_dev);
    NCT_CHECK_LAUNCH(ctx);
    if (iters_out) {
        NlScalars hs;
        NCT_CUDA(ctx, cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        NCT_CUDA(ctx, nct_stream_wait(ctx));
        for (int c = 0; c < 3; ++c) iters_out[c] = hs.iters[c];
    }
    return NCT_OK;
}

int nct_solve_wls_jacobi(nct_ctx *ctx, double *a_dev, double *b_dev, const double *rough_dev, const uint8_t *cnt_lab_full_dev, int H,
                  int W, double lam, double alpha, double rel_tol, int max_iters, int *iters_out, double *rel_res_out)
{
    NCT_ENTER(ctx);
    NCT_REQUIRE(ctx, a_dev && b_dev && rough_dev && cnt_lab_full_dev && H > 0 && W > 0, "bad arguments");
    if (rel_tol <= 0) rel_tol = 1e-10;
    if (max_iters <= 0) max_iters = 50000;
    const int n = H * W;
    const int blocks = nct_div_up(n, TPB);
    double *wx = (double *)nct_scratch(ctx, "wls_wx", sizeof(double) * n);
    double *wy = (double *)nct_scratch(ctx, "wls_wy", sizeof(double) * n);
    double *invd = (double *)nct_scratch(ctx, "wls_invd", sizeof(double) * n);
    double *vec = (double *)nct_scratch(ctx, "wlsj_vec", sizeof(double) * (size_t)n * 6 * 6);
    double *partials = (double *)nct_scratch(ctx, "solver_partials", sizeof(double) * 18 * (size_t)(blocks + 1));
    char *misc = (char *)nct_scratch(ctx, "solver_misc", 1024);
    if (!wx || !wy || !invd || !vec || !partials || !misc) return NCT_ERR_NOMEM;
    WlsScalars *sc = (WlsScalars *)misc;
    unsigned *counter = (unsigned *)(misc + 512);
    double *x = vec, *r = vec + (size_t)n * 6, *z = vec + (size_t)n * 12, *p0 = vec + (size_t)n * 18, *p1 = vec + (size_t)n * 24,
           *Ap = vec + (size_t)n * 30;

    NCT_CUDA(ctx, cudaMemsetAsync(counter, 0, sizeof(unsigned), ctx->stream));
    const double *ptab = nct_pow_table(ctx, alpha);
    if (!ptab) return NCT_ERR_NOMEM;
    wls_setup_kernel<<<blocks, TPB, 0, ctx->stream>>>(cnt_lab_full_dev, H, W, lam, ptab, wx, wy);
    NCT_CHECK_LAUNCH(ctx);
    wls_pack_kernel<<<blocks, TPB, 0, ctx->stream>>>(a_dev, b_dev, n, x);
    NCT_CHECK_LAUNCH(ctx);
    WlsSystem S{n, H, W, rough_dev, wx, wy, invd};
    wls_init_kernel<<<blocks, TPB, 0, ctx->stream>>>(S, x, r, z, p0, sc, partials, counter);
    NCT_CHECK_LAUNCH(ctx);
    double *pold = p0, *pnew = p1;
    const int check_every = 32;
    WlsScalars hs;
    int done = 0;
    double worst = 0.0;
    while (!done) {
        NCT_CUDA(ctx, cudaMemcpyAsync(&hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
        NCT_CUDA(ctx, nct_stream_wait(ctx));
        worst = 0.0;
        for (int k = 0; k < 6; ++k) {
            const double rel = hs.bb[k] > 0.0 ? sqrt(hs.rr[k] / hs.bb[k]) : (hs.rr[k] > 0.0 ? 1.0 : 0.0);
            if (rel > worst) worst = rel;
        }
        if (worst <= rel_tol || hs.iters >= max_iters) break;
        for (int it = 0; it < check_every; ++it) {
            wls_spmv_kernel<<<blocks, TPB, 0, ctx->stream>>>(S, z, pold, pnew, Ap, sc, partials, counter);
            NCT_CHECK_LAUNCH(ctx);
            wls_update_kernel<<<blocks, TPB, 0, ctx->stream>>>(n, invd, x, r, z, pnew, Ap, sc, partials, counter);
            NCT_CHECK_LAUNCH(ctx);
            double *t = pold; pold = pnew; pnew = t;
        }
    }
    wls_unpack_kernel<<<blocks, TPB, 0, ctx->stream>>>(x, n, a_dev, b_dev);
    NCT_CHECK_LAUNCH(ctx);
    if (iters_out) *iters_out = hs.iters;
    if (rel_res_out) *rel_res_out = worst;
    if (worst > rel_tol) return nct_fail(ctx, NCT_ERR_STATE, "WLS PCG did not reach %.1e in %d iterations (at %.3e)", rel_tol, hs.iters, worst);
    return NCT_OK;
}

}  // extern "C"
