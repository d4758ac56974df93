"""CPU: the fundamental-matrix RANSAC oracle (oracle/fm_oracle.cpp) against the committed outputs of
the REAL cv2.findFundamentalMat (tests/golden/make_golden_fm.py) — the OpenCV call the reference
makes at src/point_matching.cc:53.  This pins the oracle of SURVEY.md §8f row 1."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from urmvo_b200 import synth
from make_golden_fm_cases import CASES

G = np.load(os.path.join(GOLDEN, "golden_fm_r01.npz"))


def test_generator_is_stable():
    assert int(G["n_cases"]) == len(CASES)
    for k, (seed, n, inl, sig, rot) in enumerate(CASES):
        p0, p1 = synth.make_fm(seed, n, inl, sig, rot)
        assert np.array_equal(p0, G[f"p0_{k}"]) and np.array_equal(p1, G[f"p1_{k}"]), k


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_reproduces_opencv_mask_and_model(oracle, k):
    p0, p1 = G[f"p0_{k}"], G[f"p1_{k}"]
    o = oracle.fm_ransac(p0, p1, 3.0, 0.99, 1000)
    assert o["found"] == 1
    assert np.array_equal(o["mask"], G[f"mask_{k}"]), f"case {k}: {(o['mask'] != G[f'mask_{k}']).sum()} flags differ"
    F = G[f"F_{k}"]
    assert np.abs(o["F"] - F / F[2, 2]).max() < 1e-9 * max(1.0, np.abs(F / F[2, 2]).max())


def test_cv_rng_known_answers(oracle):
    """cv::RNG(-1): the multiply-with-carry sequence that drives getSubset.  First outputs computed
    from the definition (state = lo32 * 4164903690 + hi32)."""
    state, want = 0xFFFFFFFFFFFFFFFF, []
    for _ in range(7 * 3):
        state = ((state & 0xFFFFFFFF) * 4164903690 + (state >> 32)) & 0xFFFFFFFFFFFFFFFF
        want.append(state & 0xFFFFFFFF)
    # a problem without collinear triples and with N > 2^20 never re-draws: subsets = next() % N
    rng = np.random.default_rng(5)
    N = 4099
    p0 = rng.uniform(0, 1000, (N, 2)).astype(np.float32)
    p1 = rng.uniform(0, 1000, (N, 2)).astype(np.float32)
    idx = oracle.fm_subsets(p0, p1, 3)
    assert idx.shape == (3, 7)
    assert [int(v) for v in idx.ravel()] == [w % N for w in want]


def test_subsets_are_distinct_and_not_collinear(oracle):
    p0, p1 = synth.make_fm(2020, 40, 0.7)
    p0[:20, 1] = 100.0  # half of the points on one image row: collinear triples must be re-drawn
    idx = oracle.fm_subsets(p0, p1, 200)
    assert len(idx) == 200
    for s in idx:
        assert len(set(s.tolist())) == 7
        on_row = [i for i in s if i < 20]
        # the collinearity test only involves the LAST point of the subset (OpenCV's rule)
        assert not (s[6] < 20 and len(on_row) >= 3)


def test_seven_point_models_satisfy_epipolar_constraint(oracle):
    p0, p1 = synth.make_fm(2021, 7, 1.0, px_sigma=0.0)
    Fs = oracle.fm_run7(p0, p1)
    assert 1 <= len(Fs) <= 3
    h0 = np.c_[p0.astype(np.float64), np.ones(7)]
    h1 = np.c_[p1.astype(np.float64), np.ones(7)]
    for F in Fs:
        assert abs(np.linalg.det(F)) < 1e-10 * np.abs(F).max() ** 3 + 1e-18
        r = np.einsum("ni,ij,nj->n", h1, F, h0)
        assert np.abs(r).max() < 1e-8 * np.abs(F).max() * 1e3


def test_error_is_max_of_squared_point_line_distances(oracle):
    p0, p1 = synth.make_fm(2022, 50, 0.8)
    F = oracle.fm_ransac(p0, p1)["F"]
    err = oracle.fm_errors(p0, p1, F)
    h0 = np.c_[p0.astype(np.float64), np.ones(50)]
    h1 = np.c_[p1.astype(np.float64), np.ones(50)]
    l1 = h0 @ F.T  # epipolar lines in image 1 (of points 0)
    l0 = h1 @ F
    d1 = (np.sum(l1 * h1, 1) ** 2) / (l1[:, 0] ** 2 + l1[:, 1] ** 2)
    d0 = (np.sum(l0 * h0, 1) ** 2) / (l0[:, 0] ** 2 + l0[:, 1] ** 2)
    assert np.allclose(err, np.maximum(d0, d1).astype(np.float32), rtol=1e-5)


def test_below_15_points_is_not_this_path(oracle):
    p0, p1 = synth.make_fm(2023, 14, 0.9)
    assert oracle.fm_ransac(p0, p1)["found"] == -1


# ------------------------------------------------------------------ fewer than 15 matches (7-point / LMedS branches)

GS = np.load(os.path.join(GOLDEN, "golden_fm_small_r02.npz"))


@pytest.mark.parametrize("k", range(int(GS["n_cases"])))
def test_oracle_reproduces_opencv_below_15_matches(oracle, k):
    """N == 7: cv::findFundamentalMat solves the 7 points directly and flags every match; N == 14: LMedS, where
    the median is the smallest error outside the 7-point sample.  Both reproduce the real cv2 bit for bit."""
    p0, p1 = GS[f"p0_{k}"], GS[f"p1_{k}"]
    o = oracle.find_fundamental(p0, p1, 3.0, 0.99, 1000)
    assert o["found"] == int(GS[f"found_{k}"])
    assert np.array_equal(o["mask"], GS[f"mask_{k}"])
    Fs = [f for f in GS[f"F_{k}"].reshape(3, 3, 3) if f[2, 2] != 0]
    if len(p0) == 7:  # the same set of <= 3 solutions (cv2 orders them by its own null-space basis)
        mine = oracle.fm_run7(p0, p1)
        assert len(mine) == len(Fs) and o["models"] == len(Fs)
        for f in Fs:
            assert min(np.abs(m - f).max() for m in mine) < 1e-7 * max(1.0, np.abs(f).max())
        assert min(np.abs(o["F"] - f).max() for f in Fs) < 1e-7 * max(1.0, np.abs(o["F"]).max())
    else:
        assert np.abs(o["F"] - Fs[0]).max() < 1e-7 * max(1.0, np.abs(Fs[0]).max())


def test_lmeds_between_8_and_13_matches_is_the_same_algorithm_but_not_reproducible(oracle):
    """8 <= N <= 13: the count/2-th smallest error belongs to an exactly-fitted sample point (~1e-27), so the winner
    is rounding noise (the generator recorded 1-28 % agreement with cv2).  What is checkable: the result is a valid
    LMedS answer — the seven points of some sample are inliers of the returned model and sigma follows the rule."""
    agree = GS["agree_8_13"]
    assert (agree[:, 0] < agree[:, 1]).all()  # cv2 itself is not reproduced there — documented, not hidden
    for n in range(8, 14):
        p0, p1 = synth.make_fm(7000 + n, n, 0.8, 0.5, 4.0)
        o = oracle.find_fundamental(p0, p1)
        assert o["iters"] == 300 and o["mask"].sum() >= 7 and o["found"] == 1
        err = oracle.fm_errors(p0, p1, o["F"])
        assert np.sort(err)[6] < 1e-12  # seven exactly-fitted points


def test_find_fundamental_dispatch(oracle):
    p0, p1 = synth.make_fm(7100, 40, 0.7)
    a, b = oracle.find_fundamental(p0, p1), oracle.fm_ransac(p0, p1)
    assert np.array_equal(a["mask"], b["mask"]) and a["iters"] == b["iters"]
    assert oracle.find_fundamental(p0[:6], p1[:6])["found"] == -1
