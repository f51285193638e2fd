#!/bin/bash
# Builds oracle/_ref/libref_dist.so from the REFERENCE'S OWN SOURCE where it lies under /root/reference:
# the __host__ __device__ distance code of NCT/GeneralizedPatchMatch.cu (lines 9-52 helpers, 355-405
# dist_compute_single, 461-488 dist_constraint + dist_single) compiled verbatim for the host with g++.
# Nothing is copied into the repository: the extracted text lives in a temporary directory, only the .so is kept
# (oracle/_ref/ is git-ignored but travels to the GPU box).  The rest of the reference cannot be built here
# (Windows-only includes, Caffe, OpenCV 2.4.10, MKL, cuDNN <= 5, legacy cuSPARSE).
set -euo pipefail
SRC=/root/reference/code/windows/neural_color_transfer/source/GeneralizedPatchMatch.cu
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$SRC" ] || { echo "reference source not present; keeping any prebuilt oracle/_ref" >&2; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
{
  echo '#include <climits>'
  echo '#include <cfloat>'
  echo '#define __host__'
  echo '#define __device__'
  sed -n '9,52p' "$SRC"
  sed -n '355,405p' "$SRC"
  sed -n '461,488p' "$SRC"
  cat <<'EOT'
extern "C" float ref_dist_single(float *a1, float *b1, int channels, int a_rows, int a_cols, int b_rows, int b_cols,
                                 int ax, int ay, int xp, int yp, int patch_w, float cutoff)
{
    return dist_single(a1, b1, (float *)0, 1.0f, channels, a_rows, a_cols, b_rows, b_cols, ax, ay, xp, yp, 0, patch_w, 10.0f, cutoff);
}
extern "C" unsigned int ref_xy_to_int(int x, int y) { return XY_TO_INT(x, y); }
extern "C" int ref_int_to_x(unsigned int v) { return INT_TO_X(v); }
extern "C" int ref_int_to_y(unsigned int v) { return INT_TO_Y(v); }
EOT
} > "$TMP/ref_dist.cpp"
mkdir -p "$HERE/_ref"
# host flavour: separate multiply and subtract (what a host compiler without contraction does)
/usr/bin/g++ -O2 -fPIC -shared -ffp-contract=off -o "$HERE/_ref/libref_dist.so" "$TMP/ref_dist.cpp"
# device flavour: the same text compiled by nvcc for sm_100a with its defaults (-fmad=true), wrapped in a kernel that
# evaluates dist_single for a list of (ax, ay, bx, by) -- what the reference's patchmatch_single computes per candidate
{
  echo '#include <climits>'
  echo '#include <cfloat>'
  echo '#include <cuda_runtime.h>'
  sed -n '9,52p' "$SRC"
  sed -n '355,405p' "$SRC"
  sed -n '461,488p' "$SRC"
  cat <<'EOT'
__global__ void ref_dist_kernel(float *a1, float *b1, int channels, int a_rows, int a_cols, int b_rows, int b_cols,
                                const int *q, int nq, float *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) out[i] = dist_single(a1, b1, (float *)0, 1.0f, channels, a_rows, a_cols, b_rows, b_cols, q[4 * i], q[4 * i + 1],
                                     q[4 * i + 2], q[4 * i + 3], 0, 3, 10.0f);
}
extern "C" int ref_dist_device(float *a1, float *b1, int channels, int a_rows, int a_cols, int b_rows, int b_cols,
                               const int *q_dev, int nq, float *out_dev)
{
    ref_dist_kernel<<<(nq + 127) / 128, 128>>>(a1, b1, channels, a_rows, a_cols, b_rows, b_cols, q_dev, nq, out_dev);
    return (int)cudaDeviceSynchronize();
}
EOT
} > "$TMP/ref_dist_dev.cu"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o "$HERE/_ref/libref_dist_dev.so" "$TMP/ref_dist_dev.cu" -lcudart_static -lpthread -ldl -lrt
# the reference's PatchMatch kernel itself (patchmatch_single, :677-831, with its helpers :9-66, :355-405, :461-488,
# :505-515), verbatim, launched with the reference's geometry (24 x 24 blocks, NCT/main.cu:196-201, CT/Config.h:4) on
# planar CHW volumes.  It is racy by construction (both __syncthreads are commented out), so it is used for
# STATISTICAL agreement with the deterministic restatement (tests/test_oracle_ref.py), not for bit-level parity.
{
  echo '#include <climits>'
  echo '#include <cfloat>'
  echo '#include <cuda_runtime.h>'
  echo '#include <curand_kernel.h>'
  sed -n '9,66p' "$SRC"
  sed -n '355,405p' "$SRC"
  sed -n '461,488p' "$SRC"
  sed -n '505,515p' "$SRC"
  sed -n '677,831p' "$SRC"
  cat <<'EOT'
extern "C" int ref_patchmatch_device(float *a1_chw_dev, float *b1_chw_dev, unsigned int *ann_dev, float *annd_dev, const int *params_host)
{
    int *dparams = 0;
    if (cudaMalloc(&dparams, 11 * sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpy(dparams, params_host, 11 * sizeof(int), cudaMemcpyHostToDevice);
    dim3 block(24, 24), grid(params_host[2] / 24 + 1, params_host[1] / 24 + 1);
    patchmatch_single<<<grid, block>>>(a1_chw_dev, b1_chw_dev, (float *)0, ann_dev, annd_dev, dparams);
    int rc = (int)cudaDeviceSynchronize();
    cudaFree(dparams);
    return rc;
}
EOT
} > "$TMP/ref_pm_dev.cu"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -o "$HERE/_ref/libref_pm_dev.so" "$TMP/ref_pm_dev.cu" -lcudart_static -lpthread -ldl -lrt
echo "built $HERE/_ref/libref_dist.so libref_dist_dev.so libref_pm_dev.so from $SRC"
