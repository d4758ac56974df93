"""CPU: the BA / pose-only oracle against known answers and scipy."""
import pytest
import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from urmvo_b200 import synth


def _noise_free(seed=21):
    p = synth.small_ba(seed=seed, outlier_frac=0.0, rot_sigma_deg=0.3, trans_sigma=0.01, pt_sigma=0.02)
    uv, _ = synth.project(p["gt_poses"], p["obs_cam"], p["gt_pts"][p["obs_pt"]])
    p["uv"] = np.ascontiguousarray(uv)  # exact, un-rounded projections of the ground truth
    return p


def test_noise_free_scene_returns_ground_truth(oracle):
    p = _noise_free()
    poses, pts, inl, st = oracle.local_ba(p, it0=15, it1=15)
    assert inl.all()
    assert st.chi2_final[1] < 1e-16
    assert np.abs(pts - p["gt_pts"]).max() < 1e-7
    # quaternion sign convention: compare rotation matrices
    assert np.abs(synth.quat_to_R(poses[:, :4]) - synth.quat_to_R(p["gt_poses"][:, :4])).max() < 1e-8
    assert np.abs(poses[:, 4:] - p["gt_poses"][:, 4:]).max() < 1e-8


def test_fixed_cameras_do_not_move(oracle):
    p = synth.small_ba(seed=5)
    poses, _, _, _ = oracle.local_ba(p)
    fx = p["fixed"] == 1
    assert np.abs(synth.quat_to_R(poses[fx, :4]) - synth.quat_to_R(p["poses"][fx, :4])).max() < 1e-12
    assert np.abs(poses[fx, 4:] - p["poses"][fx, 4:]).max() < 1e-12
    assert np.abs(poses[~fx] - p["poses"][~fx]).max() > 1e-4


def test_lm_trace_is_monotone_and_matches_g2o_lambda_rule(oracle):
    p = synth.small_ba(seed=11, rot_sigma_deg=4.0, trans_sigma=0.3, pt_sigma=0.5)
    _, _, _, st = oracle.local_ba(p)
    rows = st.rows()
    assert len(rows) == st.iters[0] + st.iters[1]
    for before, after, lam, trials, accepted in rows:
        assert after <= before + 1e-9
        assert 1 <= trials <= 10
        assert lam > 0
    # accepted first-try steps shrink lambda by at most 3x (alpha clamp 1/3 .. 2/3)
    first = rows[: st.iters[0]]
    for (b0, a0, l0, t0, acc0), (b1, a1, l1, t1, acc1) in zip(first, first[1:]):
        if t1 == 1 and acc1:
            assert l0 / 3 - 1e-12 <= l1 <= l0 * 2 / 3 + 1e-12


def test_second_pass_minimum_matches_scipy_least_squares(oracle):
    """it0 = 0 leaves every edge at level 0 with no kernel, so the second optimize() is plain
    non-linear least squares; scipy's trust-region solver must find the same minimum."""
    p = synth.small_ba(seed=9, n_cams=4, n_pts=40, outlier_frac=0.0)
    poses, pts, inl, st = oracle.local_ba(p, it0=0, it1=40)
    Nc, Np = p["poses"].shape[0], p["pts"].shape[0]
    free = np.where(p["fixed"] == 0)[0]

    def unpack(x):
        P = p["poses"].copy()
        for k, c in enumerate(free):
            r = Rotation.from_rotvec(x[k * 6:k * 6 + 3]) * Rotation.from_quat(p["poses"][c, :4])
            P[c, :4] = r.as_quat()
            P[c, 4:] = p["poses"][c, 4:] + x[k * 6 + 3:k * 6 + 6]
        X = p["pts"] + x[len(free) * 6:].reshape(Np, 3)
        return P, X

    def resid(x):
        P, X = unpack(x)
        uv, _ = synth.project(P, p["obs_cam"], X[p["obs_pt"]])
        return (p["uv"] - uv).ravel()

    sol = least_squares(resid, np.zeros(len(free) * 6 + Np * 3), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=400)
    assert abs(st.chi2_final[1] - 2 * sol.cost) / (2 * sol.cost) < 1e-6
    _, X = unpack(sol.x)
    assert np.abs(X - pts).max() < 1e-4


def test_robust_first_pass_minimum_matches_scipy_huber(oracle):
    """The FIRST optimize() runs with RobustKernelHuber on every edge (src/g2o_optimization.cc:74-77, delta =
    (float)sqrt(cfg.mono_point)): g2o's robustified cost is sum rho(|e|^2) with rho(s) = s for s <= delta^2, else
    2 delta sqrt(s) - delta^2.  scipy's loss='huber' is the same rho applied to the square of each residual, so with
    ONE residual |e| per edge and f_scale = delta its cost is exactly half of g2o's: an independent minimiser
    (trust region, numerical Jacobian) must find the same robust minimum, outliers included."""
    p = synth.small_ba(seed=12, n_cams=4, n_pts=40, outlier_frac=0.08)
    assert p["is_outlier"].sum() >= 5
    poses, pts, inl, st = oracle.local_ba(p, it0=60, it1=0)
    Np = p["pts"].shape[0]
    free = np.where(p["fixed"] == 0)[0]
    delta = float(np.float32(np.sqrt(10.0)))

    def unpack(x):
        P = p["poses"].copy()
        for k, c in enumerate(free):
            r = Rotation.from_rotvec(x[k * 6:k * 6 + 3]) * Rotation.from_quat(p["poses"][c, :4])
            P[c, :4] = r.as_quat()
            P[c, 4:] = p["poses"][c, 4:] + x[k * 6 + 3:k * 6 + 6]
        return P, p["pts"] + x[len(free) * 6:].reshape(Np, 3)

    def resid(x):
        P, X = unpack(x)
        uv, _ = synth.project(P, p["obs_cam"], X[p["obs_pt"]])
        return np.linalg.norm(p["uv"] - uv, axis=1)

    # the oracle's minimiser in scipy's parametrisation, perturbed: |e| is not smooth at 0, so a cold start with a
    # numerical Jacobian crawls; from a nearby point the trust region must come back to the same robust cost
    x = np.zeros(len(free) * 6 + Np * 3)
    for k, c in enumerate(free):
        r = Rotation.from_quat(poses[c, :4]) * Rotation.from_quat(p["poses"][c, :4]).inv()
        x[k * 6:k * 6 + 3] = r.as_rotvec()
        x[k * 6 + 3:k * 6 + 6] = poses[c, 4:] - p["poses"][c, 4:]
    x[len(free) * 6:] = (pts - p["pts"]).ravel()
    r0 = resid(x)
    mine = np.where(r0 * r0 <= delta * delta, r0 * r0, 2 * delta * r0 - delta * delta).sum()
    assert abs(mine - st.chi2_final[0]) <= 1e-9 * mine  # the reported robust chi2 is sum rho(|e|^2)
    x0 = x + 1e-3 * np.random.default_rng(0).standard_normal(x.shape)
    assert 2 * least_squares(resid, x0, method="trf", loss="huber", f_scale=delta, max_nfev=1).cost > mine * (1 + 1e-4)
    sol = least_squares(resid, x0, method="trf", loss="huber", f_scale=delta, xtol=1e-15, ftol=1e-15, gtol=1e-12,
                        max_nfev=400)
    assert (resid(sol.x) > delta).sum() >= 3  # the robust branch of rho is exercised at the minimum
    assert abs(st.chi2_final[0] - 2 * sol.cost) / (2 * sol.cost) < 1e-6


def test_stereo_second_pass_minimum_matches_scipy_least_squares(oracle):
    """EdgeStereoSE3ProjectXYZ next to EdgeSE3ProjectXYZ (src/g2o_optimization.cc:96-118): with it0 = 0 the graph is
    plain least squares over 2-row mono and 3-row stereo residuals (third row u_right - (u - bf / z)); scipy finds the
    same minimum."""
    p = synth.add_stereo(synth.small_ba(seed=9, n_cams=4, n_pts=40, outlier_frac=0.0), 3, stereo_frac=0.5)
    assert 10 < p["kind"].sum() < len(p["kind"]) - 10
    poses, pts, inl, st = oracle.local_ba_stereo(p, 10.0, 75.0, it0=0, it1=40)
    Np = p["pts"].shape[0]
    free = np.where(p["fixed"] == 0)[0]
    fx, bf = p["intr5"][0], p["intr5"][4]
    stereo = p["kind"] != 0

    def unpack(x):
        P = p["poses"].copy()
        for k, c in enumerate(free):
            r = Rotation.from_rotvec(x[k * 6:k * 6 + 3]) * Rotation.from_quat(p["poses"][c, :4])
            P[c, :4] = r.as_quat()
            P[c, 4:] = p["poses"][c, 4:] + x[k * 6 + 3:k * 6 + 6]
        return P, p["pts"] + x[len(free) * 6:].reshape(Np, 3)

    def resid(x):
        P, X = unpack(x)
        uv, z = synth.project(P, p["obs_cam"], X[p["obs_pt"]])
        e = (p["uv3"][:, :2] - uv).ravel()
        er = (p["uv3"][:, 2] - (uv[:, 0] - bf / z))[stereo]
        return np.r_[e, er]

    assert fx > 0
    sol = least_squares(resid, np.zeros(len(free) * 6 + Np * 3), method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-12, max_nfev=400)
    assert abs(st.chi2_final[1] - 2 * sol.cost) / (2 * sol.cost) < 1e-6
    assert np.abs(unpack(sol.x)[1] - pts).max() < 1e-4


def test_outliers_are_flagged(oracle):
    p = synth.cfg1()
    _, _, inl, st = oracle.local_ba(p)
    agree = ((inl == 0) == p["is_outlier"]).mean()
    assert agree > 0.99
    assert st.chi2_final[1] < st.chi2_final[0]


def test_pose_only_recovers_pose_and_counts_inliers(oracle):
    b = synth.make_pose_batch(3, B=6, n_obs=400)
    poses, inl, n_inl = oracle.pose_only_batch(b)
    for f in range(6):
        s = slice(b["obs_offset"][f], b["obs_offset"][f + 1])
        assert n_inl[f] == inl[s].sum()
        assert ((inl[s] == 0) == b["is_outlier"][s]).mean() > 0.98
    assert np.abs(synth.quat_to_R(poses[:, :4]) - synth.quat_to_R(b["gt_poses"][:, :4])).max() < 2e-3
    assert np.abs(poses[:, 4:] - b["gt_poses"][:, 4:]).max() < 2e-2


def test_pose_only_less_than_ten_edges_stops_after_first_round(oracle):
    """src/g2o_optimization.cc:310-311."""
    b = synth.make_pose_batch(4, B=1, n_obs=9, outlier_frac=0.0)
    pose, inl, n, st = oracle.pose_only(b["poses"][0], b["uv"], b["Xw"], b["intr"])
    assert st.iters[0] > 0 and st.iters[1] == 0 and st.iters[2] == 0 and st.iters[3] == 0


def test_pose_only_threads_match_serial(oracle):
    b = synth.make_pose_batch(8, B=12, n_obs=150)
    a = oracle.pose_only_batch(b, n_threads=1)
    c = oracle.pose_only_batch(b, n_threads=4)
    assert all(np.array_equal(x, y) for x, y in zip(a, c))


# ------------------------------------------------------------------------ per-constraint camera models

def test_multicam_with_one_model_is_the_stereo_call(oracle):
    """camera_list[mpc->id_camera] with a single camera (every reference configuration): the per-edge model
    lookup must reproduce the single-intrinsics restatement bit for bit."""
    p = synth.add_stereo(synth.small_ba(seed=13, n_pts=200), 5)
    q = dict(p, kind_model=p["kind"].copy(), intr5_tab=p["intr5"].reshape(1, 5))
    a = oracle.local_ba_stereo(p, 10.0, 75.0)
    b = oracle.local_ba_multicam(q, 10.0, 75.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3].chi2_final[1] == b[3].chi2_final[1]


def test_multicam_residual_uses_the_model_of_the_edge(oracle):
    """One edge per model: the cost of a zero-iteration solve equals the sum of the robustified errors computed by
    the single-edge hook with that model's intrinsics (src/g2o_optimization.cc:86-89, :106-113)."""
    p = synth.add_camera_models(synth.add_stereo(synth.small_ba(seed=3, n_pts=80), 9), 4, n_models=4)
    assert len(set((p["kind_model"] >> 1).tolist())) == 4
    _, _, _, st = oracle.local_ba_multicam(p, 10.0, 75.0, it0=1, it1=0)
    chi0 = st.rows()[0][0] if st.rows() else None
    # the initial robust cost, recomputed edge by edge
    tot = 0.0
    for o in range(len(p["obs_cam"])):
        c, l = p["obs_cam"][o], p["obs_pt"][o]
        q, t = p["poses"][c, :4], p["poses"][c, 4:]
        R = synth.quat_to_R(q[None])[0]
        Tcw = np.r_[-q[:3], q[3], -(R.T @ t)]
        km = int(p["kind_model"][o])
        row = p["intr5_tab"][km >> 1]
        if km & 1:
            e = oracle.edge_stereo(Tcw, p["pts"][l], p["uv3"][o], row)[0]
            d = float(np.float32(np.sqrt(75.0)))
        else:
            e = np.r_[oracle.edge(Tcw, p["pts"][l], p["uv3"][o, :2], row[:4])[0], 0.0]
            d = float(np.float32(np.sqrt(10.0)))
        e2 = float(e @ e)
        tot += e2 if e2 <= d * d else 2 * np.sqrt(e2) * d - d * d
    assert chi0 is not None and abs(chi0 - tot) <= 1e-9 * tot


def test_multicam_noise_free_scene_returns_ground_truth(oracle):
    p = synth.add_stereo(_noise_free(seed=8), 2, stereo_frac=0.5, px_sigma=0.0)
    # exact right-image columns of the ground truth
    uv, z = synth.project(p["gt_poses"], p["obs_cam"], p["gt_pts"][p["obs_pt"]])
    p["uv3"] = np.ascontiguousarray(np.c_[uv, np.where(p["kind"] != 0, uv[:, 0] - synth.BF / z, 0.0)])
    p = synth.add_camera_models(p, 6, n_models=3)
    poses, pts, inl, st = oracle.local_ba_multicam(p, 10.0, 75.0, it0=15, it1=15)
    assert inl.all() and st.chi2_final[1] < 1e-14
    assert np.abs(pts - p["gt_pts"]).max() < 1e-6
    assert np.abs(poses[:, 4:] - p["gt_poses"][:, 4:]).max() < 1e-7


def test_multicam_pose_only_matches_single_model_and_rejects_bad_index(oracle):
    b = synth.make_pose_batch_stereo(4, B=3, n_obs=150)
    one = dict(b, kind_model=b["kind"].copy(), intr5_tab=b["intr5"].reshape(1, 5))
    a = oracle.pose_only_batch_stereo(b, 10.0, 75.0)
    c = oracle.pose_only_batch_multicam(one, 10.0, 75.0)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[2], c[2])
    many = synth.add_camera_models(b, 12, n_models=5)
    gp, gi, gn = oracle.pose_only_batch_multicam(many, 10.0, 75.0)
    # the same scene seen through other pinhole models: the recovered poses stay near the ground truth
    assert np.abs(gp[:, 4:] - b["gt_poses"][:, 4:]).max() < 0.05 and (gn > 0.7 * 150).all()
    bad = dict(many, kind_model=(many["kind_model"] | 0xF0).astype(np.uint8))
    assert (oracle.pose_only_batch_multicam(bad, 10.0, 75.0)[2] < 0).all()


def test_pose_only_final_pose_is_opencvs_least_squares_pose_over_the_inliers(oracle):
    """An independent pin of FrameOptimization's result with the REAL OpenCV: rounds 3 and 4 run without the robust
    kernel over the edges that survived the re-classification (src/g2o_optimization.cc:287-288), so the returned pose
    is the plain least-squares minimum over the final inlier set.  cv2.solvePnP(ITERATIVE) + solvePnPRefineLM, started
    from the same INPUT pose on that set, must reach the same pose."""
    cv2 = pytest.importorskip("cv2")
    b = synth.make_pose_batch(3, B=6, n_obs=400)
    K = np.array([[b["intr"][0], 0, b["intr"][2]], [0, b["intr"][1], b["intr"][3]], [0, 0, 1.0]])
    for f in range(6):
        s = slice(b["obs_offset"][f], b["obs_offset"][f + 1])
        pose, inl, n, _ = oracle.pose_only(b["poses"][f], b["uv"][s], b["Xw"][s], b["intr"])
        m = inl.astype(bool)
        R = synth.quat_to_R(pose[None, :4])[0]
        Rcw, tcw = R.T, -R.T @ pose[4:]
        R0 = synth.quat_to_R(b["poses"][f][None, :4])[0]
        rvec, _ = cv2.Rodrigues(R0.T)
        tvec = (-R0.T @ b["poses"][f][4:]).reshape(3, 1)
        ok, rv, tv = cv2.solvePnP(b["Xw"][s][m], b["uv"][s][m], K, None, rvec.copy(), tvec.copy(), useExtrinsicGuess=True,
                                  flags=cv2.SOLVEPNP_ITERATIVE)
        rv, tv = cv2.solvePnPRefineLM(b["Xw"][s][m], b["uv"][s][m], K, None, rv, tv,
                                      criteria=(cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 100, 1e-12))
        R2, _ = cv2.Rodrigues(rv)
        assert ok and np.abs(R0.T - Rcw).max() > 1e-3      # the input pose is far from the answer
        assert np.abs(R2 - Rcw).max() < 1e-6 and np.abs(tv.ravel() - tcw).max() < 1e-6
