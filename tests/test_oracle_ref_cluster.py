"""Pins oracle/cluster_oracle.c (decisions K1-K5) against REFERENCE-BUILT code: the reference's own cvflann k-means
(CT/Flann/kmeans_index.h), nanoflann (CT/Flann/nanoflann.hpp) and the clustering / k-NN member functions of
CT/ColorTransfer.cpp, compiled verbatim from /root/reference into oracle/_ref/libref_cluster.so by
oracle/build_ref_cluster.sh (behind an emulated MSVC rand / random_shuffle).  CPU only; skipped when the library has not
been built (it is built by __graft_entry__.build() wherever /root/reference exists and travels to the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle
from oracle import color, synth

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_cluster.so")
pytestmark = pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libref_cluster.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(SO)
    L.ref_cluster_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ref_cluster_features.restype = C.c_int
    L.ref_find_knns.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_msvc_shuffle.argtypes = [C.c_int, C.c_void_p]
    return L


def ref_kmeans(L, feats, w, h):
    f = np.ascontiguousarray(feats, np.float32).copy()
    labels = np.full(w * h, -1, np.int32)
    n = L.ref_cluster_features(f.ctypes.data, w, h, f.shape[1], labels.ctypes.data)
    return labels, n


def ref_knn(L, labels, lw, lh, nl, lab_u8, samples):
    H, W, _ = lab_u8.shape
    lab_d = np.ascontiguousarray(lab_u8.astype(np.float64) * (1.0 / 255.0))  # Mat::convertTo(CV_64F, 1/255), CT/ColorTransfer.h:59
    ids = np.empty((H * W, 8), np.int32)
    wts = np.empty((H * W, 8), np.float64)
    lb = np.ascontiguousarray(labels, np.int32)
    L.ref_find_knns(lb.ctypes.data, lw, lh, nl, lab_d.ctypes.data, H, W, samples, ids.ctypes.data, wts.ctypes.data)
    return ids, wts


def test_random_shuffle_restatements_agree(ref):
    for n in (2, 10, 1936, 40000):   # 40000 > 2^15: exercises the widening of the 15-bit draws
        out = np.empty(n, np.int32)
        ref.ref_msvc_shuffle(n, out.ctypes.data)
        assert np.array_equal(out, oracle.msvc_shuffle(n))


@pytest.mark.parametrize("seed,h,w,c", [(1, 44, 44, 512), (2, 16, 16, 512), (3, 32, 28, 512), (4, 20, 24, 64)])
def test_kmeans_labels_equal_the_reference_flann_kmeans(ref, seed, h, w, c):
    """The 10-way root split of the reference's KMeansIndex (random distinct initial centres under srand(1), <= 11 Lloyd
    iterations, float L2 in groups of four, getMinVarianceClusters) on unit-norm conv5_1-like rows: labels identical."""
    feats = oracle.l2norm_hwc(synth.feature_volume(seed, h, w, c)).reshape(h * w, c)
    want, n_ref = ref_kmeans(ref, feats, w, h)
    got, n = oracle.kmeans_labels(feats, 10, 11)
    assert n == n_ref
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} labels differ from the reference-built k-means"


def test_kmeans_on_real_pipeline_features_equals_reference(ref):
    """The same on the features the pipeline really clusters: the fixed-point VGG conv5_1 map of a synthetic image."""
    from oracle import vgg

    wts = synth.vgg19_weights(19)
    img, _ = synth.pair(3, 160, 176)
    f = vgg.features_fixedpoint(img, wts, 0)[0]
    h, w, c = f.shape
    feats = oracle.l2norm_hwc(f).reshape(h * w, c)
    want, n_ref = ref_kmeans(ref, feats, w, h)
    got, n = oracle.kmeans_labels(feats, 10, 11)
    assert n == n_ref and np.array_equal(got, want)


@pytest.mark.parametrize("samples,h,w,quant", [(1, 12, 12, 1), (2, 24, 22, 1), (4, 47, 48, 1), (2, 24, 24, 16)])
def test_knn_equals_the_reference_nanoflann_search_up_to_ties(ref, samples, h, w, quant):
    """findKnns of the reference (cluster dilation, sample blocks, per-cluster KD-tree 9-NN, merge, sort, unique, weights)
    against the oracle's exact search.  Decision K4 breaks distance ties by the smaller pixel id, the reference by KD-tree
    traversal order of a shuffled point set; everything that does not depend on tie-breaking must be IDENTICAL:
    the multiset of the 8 neighbour distances (hence all 8 weights, bit for bit), and every neighbour that is not tied
    with the 8th / 9th distance."""
    lw, lh = (w + samples - 1) // samples, (h + samples - 1) // samples
    rng = np.random.default_rng(samples * 100 + h)
    # blocky label map with 10 labels, like a k-means root split of a smooth image
    labels = (rng.integers(0, 10, ((lh + 3) // 4, (lw + 3) // 4)).repeat(4, 0).repeat(4, 1)[:lh, :lw]).astype(np.int32).ravel()
    # k-means never returns an empty cluster (m_labelNum counts the clusters found; an empty one would crash the
    # reference's KD-tree build): make the labels contiguous
    _, labels = np.unique(labels, return_inverse=True)
    labels = labels.astype(np.int32)
    nl = int(labels.max()) + 1
    cnt, _ = synth.pair(5 + samples, h, w)
    lab = color.bgr2lab_u8(cnt)
    lab = (lab // quant) * quant          # quant > 1: many exactly tied distances
    r_ids, r_w = ref_knn(ref, labels, lw, lh, nl, lab, samples)
    o_ids, o_w = oracle.find_knns(labels, lw, lh, lab, samples, nlabels=nl)
    o_ids, o_w = o_ids.reshape(-1, 8), o_w.reshape(-1, 8)
    full = (o_ids >= 0).all(1) & (r_ids >= 0).all(1)     # pixels with at least 8 candidates on both sides
    assert full.mean() > 0.9
    # (1) weights: the reference computes exp(1 - d/3) from the double Euclidean distance of u8/255 values, the oracle from
    #     sqrt(integer D2)/255 -- the same real number, rounded differently in the last place at most
    rel = np.abs(r_w[full] - o_w[full]) / o_w[full]
    assert rel.max() < 1e-14, rel.max()
    # (2) ids: where several candidates share a distance the two searches may pick different ones (K4: smaller pixel id;
    #     reference: KD-tree traversal order of the shuffled points).  Every neighbour the reference picked must be a VALID
    #     alternative: a different pixel, not the query itself, at exactly the distance the oracle has at that rank.
    flat = lab.reshape(-1, 3).astype(np.int64)
    q = np.repeat(np.arange(h * w), 8).reshape(-1, 8)
    d2 = ((flat[q[full]] - flat[r_ids[full]]) ** 2).sum(-1)
    w_of_ref_ids = np.exp(1.0 - (np.sqrt(d2.astype(np.float64)) / 255.0) / 3.0)
    assert (np.abs(w_of_ref_ids - o_w[full]) / o_w[full]).max() < 1e-14
    assert (r_ids[full] != q[full]).all()
    srt = np.sort(r_ids[full], axis=1)
    assert (np.diff(srt, axis=1) != 0).all(), "the reference returned a duplicate neighbour"
    same = (r_ids[full] == o_ids[full])
    frac = same.mean()
    # rows without any tie (8 strictly decreasing weights) can differ only in the last element (tie with the first cut-off candidate)
    strict = (np.diff(o_w[full], axis=1) < 0).all(1)
    rows_equal = same.all(1)
    bad = strict & ~rows_equal
    assert (same[bad][:, :7]).all(), "a non-tied neighbour differs from the reference"
    print(f"k-NN vs reference-built nanoflann (samples {samples}, {h}x{w}, quant {quant}): weights max rel diff {rel.max():.1e}, "
          f"{100 * frac:.2f} % of ids identical, {100 * rows_equal.mean():.2f} % of rows identical "
          f"({100 * strict.mean():.1f} % of rows are tie-free), {int(bad.sum())} tie-free rows differ in the cut-off neighbour only")
    # (3) short rows: fewer than 8 candidates -> the reference pads with NN() = (id -1, w 0) (K5)
    short = ~(r_ids >= 0).all(1)
    assert np.array_equal(short, ~(o_ids >= 0).all(1))
    assert np.array_equal((r_ids[short] >= 0).sum(1), (o_ids[short] >= 0).sum(1))
