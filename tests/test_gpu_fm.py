"""GPU (-m gpu): the per-frame fundamental-matrix RANSAC kernels through the C ABI against
(1) the committed outputs of the real cv2.findFundamentalMat and (2) the CPU oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from urmvo_b200 import synth
import urmvo_b200 as U
from make_golden_fm_cases import CASES

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(GOLDEN, "golden_fm_r01.npz"))


@pytest.mark.parametrize("k", range(len(CASES)))
def test_gpu_reproduces_opencv_golden(ctx, k):
    p0, p1 = G[f"p0_{k}"], G[f"p1_{k}"]
    g = ctx.fm_ransac(p0, p1, 3.0, 0.99, 1000)
    assert g["found"] == 1
    assert np.array_equal(g["mask"], G[f"mask_{k}"]), f"{(g['mask'] != G[f'mask_{k}']).sum()} flags differ from OpenCV"
    F = G[f"F_{k}"] / G[f"F_{k}"][2, 2]
    assert np.abs(g["F"] - F).max() < 1e-9 * max(1.0, np.abs(F).max())


@pytest.mark.parametrize("seed,n,inl", [(3001, 15, 0.9), (3002, 33, 0.6), (3003, 300, 0.4), (3004, 1000, 0.7),
                                        (3005, 1000, 0.25), (3006, 4000, 0.6)])
def test_gpu_matches_oracle(ctx, oracle, seed, n, inl):
    p0, p1 = synth.make_fm(seed, n, inl)
    g = ctx.fm_ransac(p0, p1)
    o = oracle.fm_ransac(p0, p1)
    assert g["found"] == o["found"]
    assert np.array_equal(g["mask"], o["mask"])
    assert (g["iters"], g["n_inliers"], g["models"]) == (o["iters"], o["n_inliers"], o["models"])
    assert np.abs(g["F"] - o["F"]).max() <= 1e-11 * max(1.0, np.abs(o["F"]).max())


def test_gpu_batch_equals_single_calls(ctx, oracle):
    pairs = [synth.make_fm(3100 + b, 100 + 37 * b, 0.4 + 0.05 * (b % 8)) for b in range(12)]
    masks, stats = ctx.fm_ransac_batch(pairs)
    for (p0, p1), m, st in zip(pairs, masks, stats):
        o = oracle.fm_ransac(p0, p1)
        assert np.array_equal(m, o["mask"]) and st.iters == o["iters"] and st.n_inliers == o["n_inliers"]


def test_gpu_plan_rerun_is_deterministic(ctx):
    pairs = [synth.make_fm(3200 + b, 500, 0.6) for b in range(4)]
    plan = U.FMPlan(ctx, pairs)
    plan.run()
    m1, s1 = plan.finish()
    # adaptive rounds: at least the iterations OpenCV's loop needs, at most the whole budget
    assert sum(s.iters for s in s1) <= plan.hypotheses <= 4 * 1000
    plan.run()
    m2, s2 = plan.finish()
    for a, b in zip(m1, m2):
        assert np.array_equal(a, b)
    assert [s.iters for s in s1] == [s.iters for s in s2]
    plan.close()


def test_gpu_custom_threshold_confidence_and_budget(ctx, oracle):
    p0, p1 = synth.make_fm(3300, 400, 0.5)
    for thr, conf, its in [(1.0, 0.99, 1000), (3.0, 0.999, 2000), (5.0, 0.9, 50), (3.0, 0.99, 1)]:
        g = ctx.fm_ransac(p0, p1, thr, conf, its)
        o = oracle.fm_ransac(p0, p1, thr, conf, its)
        assert np.array_equal(g["mask"], o["mask"]) and g["iters"] == o["iters"], (thr, conf, its)


def test_gpu_collinear_heavy_input(ctx, oracle):
    """Half of the keypoints on one image row: checkSubset re-draws must follow OpenCV's sequence."""
    p0, p1 = synth.make_fm(3400, 60, 0.8)
    p0[:30, 1] = 120.0
    g = ctx.fm_ransac(p0, p1)
    o = oracle.fm_ransac(p0, p1)
    assert np.array_equal(g["mask"], o["mask"]) and g["iters"] == o["iters"]


def test_gpu_degenerate_all_points_identical(ctx, oracle):
    """No valid subset exists: OpenCV returns no model; every flag stays 0."""
    p0 = np.full((20, 2), 50.0, dtype=np.float32)
    p1 = np.full((20, 2), 60.0, dtype=np.float32)
    g = ctx.fm_ransac(p0, p1, max_iters=3)
    o = oracle.fm_ransac(p0, p1, max_iters=3)
    assert g["found"] == 0 and o["found"] == 0 and g["mask"].sum() == 0


def test_gpu_fewer_than_15_points_is_refused(ctx):
    p0, p1 = synth.make_fm(3500, 14, 0.9)
    with pytest.raises(U.UrmvoError, match="fewer than 15"):
        ctx.fm_ransac(p0, p1)


def test_gpu_triangulate_batch_matches_oracle(ctx, oracle):
    """Mapping::TriangulateMappoint batched: same accept / reject flags as the
    CPU restatement, positions to 1e-9, rejected mappoints keep their previous position."""
    t = synth.make_triangulation(13, n_pts=800, degenerate_frac=0.1)
    prev = np.full((800, 3), 7.0)
    pts, ok = ctx.triangulate_batch(t["obs_off"], t["obs_pose"], t["obs_uv"], t["poses_Rp"], t["intr"], pts=prev)
    n_rej = 0
    for l in range(800):
        s = slice(t["obs_off"][l], t["obs_off"][l + 1])
        o_ok, X = oracle.triangulate(t["poses_Rp"][t["obs_pose"][s]], t["obs_uv"][s], t["intr"])
        assert bool(ok[l]) == o_ok, l
        if o_ok:
            assert np.abs(pts[l] - X).max() <= 1e-9 * max(1.0, np.abs(X).max())
        else:
            n_rej += 1
            assert np.array_equal(pts[l], prev[l])
    assert 20 < n_rej < 200


def test_gpu_triangulate_batch_rejects_bad_indices(ctx):
    t = synth.make_triangulation(14, n_pts=10)
    bad = t["obs_pose"].copy(); bad[0] = 99
    with pytest.raises(U.UrmvoError, match="pose index"):
        ctx.triangulate_batch(t["obs_off"], bad, t["obs_uv"], t["poses_Rp"], t["intr"])
