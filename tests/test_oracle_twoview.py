"""CPU: the two-view oracle against numpy / OpenCV and known answers."""
import cv2
import numpy as np

from urmvo_b200 import synth


def test_jacobi_svd_matches_lapack(oracle):
    rng = np.random.default_rng(0)
    for (m, n) in [(8, 9), (16, 9), (4, 4), (3, 3)]:
        A = rng.standard_normal((m, n)).astype(np.float32)
        U, s, V = oracle.svd(A)
        ref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
        k = min(m, n)
        assert np.allclose(s[:k], ref, rtol=2e-5, atol=2e-6)
        assert np.allclose(V.T @ V, np.eye(n), atol=1e-5)
        assert np.allclose((U[:, :k] * s[:k]) @ V[:, :k].T, A, atol=1e-5)
        if m < n:  # the last right singular vector spans the null space (what _compute_F21 uses)
            assert np.abs(A @ V[:, n - 1]).max() < 1e-5


def _fit_from_oracle(oracle, tv, idx, model):
    return oracle.score_all(tv, model, sets=np.asarray([idx], dtype=np.int32))[2][0].reshape(3, 3).astype(np.float64)


def test_fundamental_8pt_matches_opencv(oracle):
    """Same algorithm family: normalised 8-point + rank-2 enforcement (cv2.FM_8POINT)."""
    tv = synth.make_two_view(7, n_keys=300, inlier_frac=1.0, px_sigma=0.0)
    idx = [3, 40, 77, 120, 160, 199, 250, 290]
    F = _fit_from_oracle(oracle, tv, idx, 0)
    Fcv, _ = cv2.findFundamentalMat(tv["keys1"][idx].astype(np.float64), tv["keys2"][idx].astype(np.float64), cv2.FM_8POINT)
    F /= np.linalg.norm(F); Fcv /= np.linalg.norm(Fcv)
    if np.sum(F * Fcv) < 0:
        Fcv = -Fcv
    # rounded integer pixels: both are least-squares fits, compare through the epipolar residuals
    x1 = np.c_[tv["keys1"], np.ones(300)].astype(np.float64); x2 = np.c_[tv["keys2"], np.ones(300)].astype(np.float64)
    r = np.abs(np.einsum("ni,ij,nj->n", x2[idx], F, x1[idx]))
    rcv = np.abs(np.einsum("ni,ij,nj->n", x2[idx], Fcv, x1[idx]))
    assert r.max() < 2 * rcv.max() + 1e-3  # rank-2 enforcement leaves the same order of residual
    assert abs(np.linalg.det(F)) < 1e-12  # rank 2
    # OpenCV normalises isotropically, the reference per axis by mean absolute deviation
    # (src/epipolar_geometry.cc:735-780): close, not identical
    assert np.abs(F - Fcv).max() < 1e-3


def test_homography_dlt_matches_opencv(oracle):
    """Exact (un-rounded) correspondences of a plane: the DLT null vector is the true homography, so
    the oracle and cv2.findHomography(method=0) must agree over the whole image."""
    tv = synth.make_two_view(8, n_keys=300, inlier_frac=1.0, px_sigma=0.0, planar=True)
    rng = np.random.default_rng(2)
    u = rng.uniform(20, 620, 300); v = rng.uniform(20, 490, 300)
    d = 4.0 + 0.2 * ((u - synth.CX) / synth.FX) + 0.1 * ((v - synth.CY) / synth.FY)
    X1 = np.c_[(u - synth.CX) / synth.FX * d, (v - synth.CY) / synth.FY * d, d]
    T = tv["gt_T21"]
    X2 = X1 @ T[:3, :3].T + T[:3, 3]
    tv["keys1"] = np.c_[u, v].astype(np.float32)
    tv["keys2"] = np.c_[X2[:, 0] / X2[:, 2] * synth.FX + synth.CX, X2[:, 1] / X2[:, 2] * synth.FY + synth.CY].astype(np.float32)
    idx = [5, 50, 90, 130, 170, 210, 260, 295]
    H = _fit_from_oracle(oracle, tv, idx, 1)
    Hcv, _ = cv2.findHomography(tv["keys1"][idx].astype(np.float64), tv["keys2"][idx].astype(np.float64), 0)
    p = np.c_[tv["keys1"], np.ones(300)].astype(np.float64) @ H.T
    pcv = np.c_[tv["keys1"], np.ones(300)].astype(np.float64) @ Hcv.T
    assert np.abs(p[:, :2] / p[:, 2:] - pcv[:, :2] / pcv[:, 2:]).max() < 0.05  # pixels
    assert np.abs(p[:, :2] / p[:, 2:] - tv["keys2"]).max() < 0.05


def test_noise_free_scene_recovers_motion_with_full_inlier_mask(oracle):
    tv = synth.make_two_view(9, n_keys=500, inlier_frac=1.0, px_sigma=0.0)
    # un-rounded exact correspondences
    rng = np.random.default_rng(1)
    u = rng.uniform(20, 620, 500); v = rng.uniform(20, 490, 500); d = rng.uniform(2, 8, 500)
    X1 = np.c_[(u - synth.CX) / synth.FX * d, (v - synth.CY) / synth.FY * d, d]
    T = tv["gt_T21"]
    X2 = X1 @ T[:3, :3].T + T[:3, 3]
    tv["keys1"] = np.c_[u, v].astype(np.float32)
    tv["keys2"] = np.c_[X2[:, 0] / X2[:, 2] * synth.FX + synth.CX, X2[:, 1] / X2[:, 2] * synth.FY + synth.CY].astype(np.float32)
    tv["sets"] = synth.draw_sets(500, 50, 0)
    r = oracle.two_view(tv)
    assert r["ok"] and r["stats"].used_H == 0
    assert r["mask_F"].mean() > 0.99
    R, t = r["T21"][:3, :3].astype(np.float64), r["T21"][:3, 3].astype(np.float64)
    assert np.abs(R - T[:3, :3]).max() < 2e-3
    tn = T[:3, 3] / np.linalg.norm(T[:3, 3])
    assert np.abs(t - tn).max() < 2e-2
    # triangulated points equal the true ones up to the monocular scale |t|
    tri = r["triangulated"].astype(bool)
    assert tri.sum() > 450
    scale = np.linalg.norm(T[:3, 3])
    assert np.median(np.abs(r["P3D"][tri] * scale - X1[tri])) < 0.05
    # cross-check the decomposition with OpenCV
    E, _ = cv2.findEssentialMat(tv["keys1"].astype(np.float64), tv["keys2"].astype(np.float64), tv["K"].astype(np.float64), cv2.LMEDS)
    _, Rcv, tcv, _ = cv2.recoverPose(E, tv["keys1"].astype(np.float64), tv["keys2"].astype(np.float64), tv["K"].astype(np.float64))
    assert np.abs(Rcv - R).max() < 5e-3 and np.abs(tcv.ravel() - t).max() < 5e-2


def test_triangulation_matches_opencv(oracle):
    tv = synth.cfg3(n_hyp=100)
    r = oracle.two_view(tv)
    assert r["ok"]
    K = tv["K"].astype(np.float64)
    P1 = K @ np.c_[np.eye(3), np.zeros(3)]
    P2 = K @ r["T21"][:3, :].astype(np.float64)
    tri = np.where(r["triangulated"])[0]
    X = cv2.triangulatePoints(P1, P2, tv["keys1"][tri].T.astype(np.float64), tv["keys2"][tri].T.astype(np.float64))
    X = (X[:3] / X[3]).T
    rel = np.linalg.norm(X - r["P3D"][tri], axis=1) / np.linalg.norm(X, axis=1)
    assert np.median(rel) < 1e-3


def test_planar_scene_selects_homography_score(oracle):
    tv = synth.make_two_view(5, planar=True, inlier_frac=0.9)
    tv["sets"] = synth.draw_sets(1000, 200, 0)
    r = oracle.two_view(tv)
    st = r["stats"]
    assert st.SH > 0 and st.SF > 0 and st.SH / (st.SH + st.SF) > 0.40  # near-planar: H competes with F


def test_best_hypothesis_is_first_maximum(oracle):
    tv = synth.cfg3(n_hyp=300)
    r = oracle.two_view(tv)
    for model, best, S in ((0, r["stats"].best_F, r["stats"].SF), (1, r["stats"].best_H, r["stats"].SH)):
        scores, masks, models = oracle.score_all(tv, model)
        assert best == int(np.argmax(scores)) and scores[best] == S  # np.argmax returns the earliest maximum
        bits = np.unpackbits(masks[best].view(np.uint8), bitorder="little")[:1000]
        assert np.array_equal(bits, r["mask_F"] if model == 0 else r["mask_H"])


def test_duplicate_sets_give_identical_results_and_unmatched_keys_are_skipped(oracle):
    tv = synth.make_two_view(12, n_keys=300, n_unmatched=37)
    N = int((tv["matches12"] >= 0).sum())
    assert N == 263
    sets = synth.draw_sets(N, 20, 0)
    sets[7] = sets[2]
    s, m, M = oracle.score_all(tv, 0, sets=sets)
    assert s[7] == s[2] and np.array_equal(m[7], m[2]) and np.array_equal(M[7], M[2])
    assert m.shape[1] == (N + 31) // 32
    assert (m[:, -1] >> (N % 32)).max() == 0  # no bits beyond N


def test_sampson_mode_matches_float64_sampson_distance(oracle):
    po = oracle
    """The extra scoring mode (Sampson error, BASELINE.json north_star (4)): the oracle's fp32 inlier
    decisions agree with a float64 evaluation of the same formula except within 1e-3 of the threshold,
    and the mode only changes the fundamental model's scores."""
    tv = synth.make_two_view(1003, n_keys=500)
    tv["sets"] = synth.draw_sets(500, 40, 0)
    s0, m0, M0 = po.score_all(tv, 0)
    s1, m1, M1 = po.score_all(tv, 0, score_mode=1)
    assert np.array_equal(M0, M1) and not np.array_equal(s0, s1)
    h0 = po.score_all(tv, 1)
    h1 = po.score_all(tv, 1, score_mode=1)
    assert np.array_equal(h0[0], h1[0]) and np.array_equal(h0[1], h1[1])
    m = np.asarray(tv["matches12"])
    i1 = np.nonzero(m >= 0)[0]
    x1 = np.c_[np.asarray(tv["keys1"], dtype=np.float64)[i1], np.ones(len(i1))]
    x2 = np.c_[np.asarray(tv["keys2"], dtype=np.float64)[m[i1]], np.ones(len(i1))]
    n_checked = 0
    for h in range(40):
        F = M1[h].astype(np.float64).reshape(3, 3)
        Fx1 = x1 @ F.T
        Ftx2 = x2 @ F
        num = (x2 * Fx1).sum(1)
        chi = num ** 2 / (Fx1[:, 0] ** 2 + Fx1[:, 1] ** 2 + Ftx2[:, 0] ** 2 + Ftx2[:, 1] ** 2)
        ref = chi <= 3.841
        bits = ((m1[h][np.arange(len(i1)) >> 5] >> (np.arange(len(i1)) & 31)) & 1).astype(bool)
        clear = np.abs(chi - 3.841) > 1e-3 * np.maximum(1.0, chi)
        assert np.array_equal(bits[clear], ref[clear])
        assert abs(float(s1[h]) - float((5.991 - chi[ref]).sum())) <= 1e-3 * max(1.0, float(s1[h]))
        n_checked += int(clear.sum())
    assert n_checked > 15000
    full = po.two_view(tv, score_mode=1)
    assert full["stats"].best_F >= 0
