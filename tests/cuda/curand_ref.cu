// Test helper (NOT part of the product): draws from cuRAND's own device API exactly the way
// the reference does (curand_init(seed = column, 0, 0); curand_uniform), so the XORWOW
// restatements in libnct.so and oracle/pm_oracle.c can be pinned against the real generator.
#include <curand_kernel.h>
#include <cuda_runtime.h>

__global__ void curand_ref_kernel(float *out, int ncols, int ndraws)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    curandState st;
    curand_init(col, 0, 0, &st);
    for (int k = 0; k < ndraws; ++k) out[(size_t)col * ndraws + k] = curand_uniform(&st);
}

extern "C" int curand_ref_table(float *out_dev, int ncols, int ndraws)
{
    curand_ref_kernel<<<(ncols + 127) / 128, 128>>>(out_dev, ncols, ndraws);
    return (int)cudaDeviceSynchronize();
}
