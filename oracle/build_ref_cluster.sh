#!/bin/bash
# Builds oracle/_ref/libref_cluster.so from the REFERENCE'S OWN SOURCES where they lie under /root/reference: the cvflann
# k-means (CT/Flann/kmeans_index.h and the headers it includes), nanoflann (CT/Flann/nanoflann.hpp) and the clustering /
# k-NN member functions of CT/ColorTransfer.cpp (PointColor + cmpDist :12-44, sortMergeComputeWeight :60-110, findSubKNNs
# :136-220, insertClusterPixel + getClusters :255-353, clusterFeastures :355-395, findKnns :397-423), extracted verbatim by
# line number into a temporary directory and compiled with g++ behind a shim that supplies only what those lines need
# from outside the tree: a stub logger.h (absent from the reference), minimal stand-ins for cv::Mat / Vec3d, the subset
# of the ColorTransfer class the functions touch, and the MSVC C runtime's rand / srand / random_shuffle (the reference is
# built with Visual Studio 2013; restated from public knowledge of that CRT -- the same restatement as decision K1 of
# oracle/cluster_oracle.c).  Nothing is copied into the repository; only the .so is kept (oracle/_ref/ is git-ignored).
# Used by tests/test_oracle_ref_cluster.py to pin oracle/cluster_oracle.c against reference-built code.
set -euo pipefail
CT=/root/reference/code/windows/neural_color_transfer/source/ColorTransfer
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$CT/ColorTransfer.cpp" ] || { echo "reference source not present; keeping any prebuilt oracle/_ref" >&2; exit 0; }
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cat > "$TMP/logger.h" <<'EOT'
#pragma once
#include <cstdio>
namespace cvflann { struct Logger { template <class... A> static int info(const char *, A...) { return 0; }
                                    template <class... A> static int error(const char *, A...) { return 0; } }; }
EOT
{
  cat <<'EOT'
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>
typedef unsigned char uchar;
// ---- MSVC CRT (VS2013): rand() LCG, RAND_MAX 0x7fff, std::random_shuffle drawing 15 bits at a time
namespace std {
static unsigned int g_msvc_seed = 1;
inline void msvc_srand(unsigned int s) { g_msvc_seed = s; }
inline int msvc_rand() { g_msvc_seed = g_msvc_seed * 214013u + 2531011u; return (int)((g_msvc_seed >> 16) & 0x7fff); }
template <class It> inline void msvc_random_shuffle(It first, It last)
{
    const unsigned long RBITS = 15, RMAX = (1UL << 15) - 1;
    It next = first;
    for (unsigned long index = 2; ++next != last; ++index) {
        unsigned long rm = RMAX;
        unsigned long rn = (unsigned long)msvc_rand() & RMAX;
        for (; rm < index && rm != ~0UL; rm = rm << RBITS | RMAX) rn = rn << RBITS | ((unsigned long)msvc_rand() & RMAX);
        std::iter_swap(next, first + (long)(rn % index));
    }
}
}
using std::msvc_rand;
using std::msvc_srand;
#undef RAND_MAX
#define RAND_MAX 0x7fff
#define rand msvc_rand
#define srand msvc_srand
#define random_shuffle msvc_random_shuffle
#include "kmeans_index.h"
#include "nanoflann.hpp"
namespace cvflann {
EOT
  sed -n '279,290p' "$CT/Flann/flann_base.hpp"
  cat <<'EOT'
}
using namespace std;
// ---- minimal stand-ins for the OpenCV types the extracted lines use
struct Vec3d { double v[3]; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; } };
enum { CV_8UC1 = 1, CV_64FC3 = 24 };
struct Mat {
    int rows = 0, cols = 0, esz = 0;
    std::vector<unsigned char> buf;
    static Mat zeros(int h, int w, int type) { Mat m; m.rows = h; m.cols = w; m.esz = type; m.buf.assign((size_t)h * w * type, 0); return m; }
    template <class T> T &at(int y, int x) { return *reinterpret_cast<T *>(&buf[((size_t)y * cols + x) * esz]); }
    template <class T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(&buf[((size_t)y * cols + x) * esz]); }
};
#define MAX_VAL 1e8
struct Config { int m_clusterNum = 10, m_kNum = 8; };
EOT
  sed -n '11,31p' "$CT/ColorTransfer.h"
  cat <<'EOT'
class ColorTransfer {
public:
    ColorTransfer(const Config &c) : m_config(c) {}
    void sortMergeComputeWeight(const vector<vector<vector<int>>>& nns, const vector<vector<vector<double>>>& nnds);
    void findSubKNNs(vector<vector<int>>& nns, vector<vector<double>>& nnds, const vector<ClusterPixel>& subCluster, int width, int height);
    void insertClusterPixel(vector<ClusterPixel>& subClusterPixels, int& count, const Mat& cntLab, int x, int y, int samples);
    void getClusters(Mat& visMat, vector<vector<ClusterPixel>>& subClusterPixels, const Mat& cntLab, int samples);
    void clusterFeastures(Mat& dvisMat, float* features, int width, int height, int channel);
    void findKnns(Mat& visMat, const Mat& cntRgb, int samples);
    vector<vector<int>> m_knn;
    vector<vector<double>> m_knnd;
    vector<vector<NN>> m_knnid;
    vector<int> m_labels;
    int m_labelNum = 0, m_labelWidth = 0, m_labelHeight = 0;
    const Config &m_config;
};
using namespace nanoflann;
EOT
  sed -n '12,44p' "$CT/ColorTransfer.cpp"
  sed -n '60,110p' "$CT/ColorTransfer.cpp"
  sed -n '136,220p' "$CT/ColorTransfer.cpp"
  sed -n '255,353p' "$CT/ColorTransfer.cpp"
  sed -n '355,395p' "$CT/ColorTransfer.cpp"
  sed -n '397,423p' "$CT/ColorTransfer.cpp"
  cat <<'EOT'
// features: [width*height][channel] unit-norm rows (NCT/main.cu:139-165) -> labels[width*height]; returns m_labelNum
extern "C" int ref_cluster_features(float *features, int width, int height, int channel, int *labels_out)
{
    Config cfg;
    ColorTransfer ct(cfg);
    Mat vis;
    ct.clusterFeastures(vis, features, width, height, channel);
    for (int i = 0; i < width * height; ++i) labels_out[i] = ct.m_labels[i];
    return ct.m_labelNum;
}
// labels: [lh*lw] from ref_cluster_features; lab_d: [H*W*3] doubles (8-bit Lab / 255, CT/ColorTransfer.h:59) -> ids / w [H*W][8]
extern "C" int ref_find_knns(const int *labels, int lw, int lh, int label_num, const double *lab_d, int H, int W, int samples,
                             int *ids_out, double *w_out)
{
    Config cfg;
    ColorTransfer ct(cfg);
    ct.m_labels.assign(labels, labels + (size_t)lw * lh);
    ct.m_labelNum = label_num;
    ct.m_labelWidth = lw;
    ct.m_labelHeight = lh;
    Mat lab = Mat::zeros(H, W, CV_64FC3), vis;
    memcpy(lab.buf.data(), lab_d, sizeof(double) * 3 * (size_t)H * W);
    ct.findKnns(vis, lab, samples);
    for (size_t i = 0; i < (size_t)H * W; ++i)
        for (int k = 0; k < 8; ++k) {
            ids_out[i * 8 + k] = ct.m_knnid[i][k].id;
            w_out[i * 8 + k] = ct.m_knnid[i][k].w;
        }
    return 0;
}
extern "C" void ref_msvc_shuffle(int n, int *out)
{
    srand(1);
    std::vector<int> v(n);
    for (int i = 0; i < n; ++i) v[i] = i;
    std::random_shuffle(v.begin(), v.end());
    for (int i = 0; i < n; ++i) out[i] = v[i];
}
EOT
} > "$TMP/ref_cluster.cpp"
mkdir -p "$HERE/_ref"
# -DNDEBUG: the reference's release build compiles its assert()s out (sortMergeComputeWeight pads short lists instead)
/usr/bin/g++ -O2 -std=c++14 -fpermissive -w -DNDEBUG -ffp-contract=off -fPIC -shared -I"$TMP" -I"$CT/Flann" -o "$HERE/_ref/libref_cluster.so" "$TMP/ref_cluster.cpp"
echo "built $HERE/_ref/libref_cluster.so from $CT/{Flann/kmeans_index.h, Flann/nanoflann.hpp, ColorTransfer.cpp}"
