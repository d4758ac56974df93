"""CPU: the oracle reproduces the committed golden vectors (tests/golden/make_golden.py)."""
import os

import numpy as np

from conftest import GOLDEN
from urmvo_b200 import synth

G = np.load(os.path.join(GOLDEN, "golden_r01.npz"))


def test_generator_is_stable():
    p = synth.small_ba(seed=7)
    for k in ("poses", "fixed", "pts", "uv", "obs_cam", "obs_pt", "intr"):
        assert np.array_equal(p[k], G["ba_in_" + k]), k
    assert np.array_equal(synth.draw_sets(1000, 16, 0), G["sets_1000x16"])


def test_ba_oracle_matches_golden(oracle):
    p = {k: G["ba_in_" + k] for k in ("poses", "fixed", "pts", "uv", "obs_cam", "obs_pt", "intr")}
    poses, pts, inl, st = oracle.local_ba(p)
    assert np.array_equal(inl, G["ba_inlier"])
    assert np.allclose(poses, G["ba_poses"], rtol=0, atol=1e-10)
    assert np.allclose(pts, G["ba_pts"], rtol=0, atol=1e-9)
    tr = np.array(st.rows())
    assert tr.shape == G["ba_trace"].shape
    assert np.array_equal(tr[:, 3:], G["ba_trace"][:, 3:])  # trials / accepted pattern
    assert np.allclose(tr[:, :3], G["ba_trace"][:, :3], rtol=1e-9)


def test_pose_only_oracle_matches_golden(oracle):
    b = {k: G["po_in_" + k] for k in ("poses", "obs_offset", "uv", "Xw", "intr")}
    poses, inl, n = oracle.pose_only_batch(b)
    assert np.array_equal(inl, G["po_inlier"]) and np.array_equal(n, G["po_n_inlier"])
    assert np.allclose(poses, G["po_poses"], rtol=0, atol=1e-10)


def test_two_view_oracle_matches_golden_bit_for_bit(oracle):
    tv = {k: G["tv_in_" + k] for k in ("keys1", "keys2", "matches12", "K", "sets")}
    tv["sigma"] = 1.0
    r = oracle.two_view(tv)
    assert bool(G["tv_ok"]) == r["ok"]
    for k, v in (("tv_T21", r["T21"]), ("tv_P3D", r["P3D"])):
        assert np.array_equal(G[k].view(np.uint32), v.view(np.uint32)), k
    assert np.array_equal(G["tv_tri"], r["triangulated"])
    assert np.array_equal(G["tv_mask_F"], r["mask_F"]) and np.array_equal(G["tv_mask_H"], r["mask_H"])
    for model, tag in ((0, "F"), (1, "H")):
        s, m, M = oracle.score_all(tv, model)
        assert np.array_equal(s.view(np.uint32), G["tv_scores_" + tag].view(np.uint32))
        assert np.array_equal(m, G["tv_masks_" + tag])
        assert np.array_equal(M.view(np.uint32), G["tv_models_" + tag].view(np.uint32))
