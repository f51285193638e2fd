"""CPU test: the composed oracle pipeline (oracle/pipeline.py) runs BASELINE config 1's plumbing at a small size."""
import numpy as np

from oracle import pipeline, synth


def test_oracle_pipeline_small_pair_runs_and_is_deterministic():
    w = synth.vgg19_weights(19)
    cnt, stl = synth.pair(0, 64, 64)
    seen = []
    t = {}
    out1 = pipeline.transfer_pair(cnt, stl, w, on_level=lambda l, d: seen.append((l, d["cg_iters"], d["result"].shape)), timings=t)
    out2 = pipeline.transfer_pair(cnt, stl, w)
    assert out1.shape == cnt.shape and out1.dtype == np.uint8
    assert np.array_equal(out1, out2)
    assert [s[0] for s in seen] == [0, 1, 2, 3, 4]
    assert all(max(s[1]) <= (50 if s[0] == 4 else 100) for s in seen)
    assert set(t) >= {"vgg", "patchmatch", "bds", "knn", "nonlocal", "wls", "kmeans"}
    # the transfer moves the content colours towards the style's statistics
    assert abs(out1.astype(float).mean() - stl.astype(float).mean()) < abs(cnt.astype(float).mean() - stl.astype(float).mean()) + 5


def test_level_dims_match_reference_geometry():
    assert [d[1] for d in pipeline.level_dims(700, 700)] == [44, 88, 175, 350, 700]
    assert [d[2] for d in pipeline.level_dims(600, 960)] == [60, 120, 240, 480, 960]
