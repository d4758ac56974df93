"""CPU: the TriangulateMappoint oracle (oracle/tri_oracle.cpp, reference src/mapping.cc:151-205)
against numpy on the same normal equations and against known answers."""
import numpy as np

from urmvo_b200 import synth


def _normal_equations(Rp, uv, intr):
    b = np.stack([(uv[:, 0] - intr[2]) / intr[0], (uv[:, 1] - intr[3]) / intr[1], np.ones(len(uv))], axis=-1)
    b = np.einsum("kij,kj->ki", Rp[:, :9].reshape(-1, 3, 3), b)
    p = Rp[:, 9:]
    inv = 1.0 / (b * b).sum(1)
    A = len(uv) * np.eye(3) - np.einsum("ki,k,kj->ij", b, inv, b)
    rhs = p.sum(0) - np.einsum("ki,k,k->i", b, inv, (b * p).sum(1))
    return A, rhs


def test_noise_free_points_are_recovered(oracle):
    t = synth.make_triangulation(11, n_pts=50, px_sigma=0.0, degenerate_frac=0.0)
    for l in range(50):
        s = slice(t["obs_off"][l], t["obs_off"][l + 1])
        ok, X = oracle.triangulate(t["poses_Rp"][t["obs_pose"][s]], t["obs_uv"][s], t["intr"])
        assert ok and np.abs(X - t["gt"][l]).max() < 1e-9


def test_matches_numpy_solution_and_rank_rule(oracle):
    t = synth.make_triangulation(12, n_pts=300, degenerate_frac=0.2)
    n_fail = 0
    for l in range(300):
        s = slice(t["obs_off"][l], t["obs_off"][l + 1])
        Rp, uv = t["poses_Rp"][t["obs_pose"][s]], t["obs_uv"][s]
        ok, X = oracle.triangulate(Rp, uv, t["intr"])
        if len(uv) < 2:
            assert not ok
            n_fail += 1
            continue
        A, rhs = _normal_equations(Rp, uv, t["intr"])
        # Eigen's rule: rank = #{|R_ii| > 1e-5 max|R_jj|} of the column-pivoted QR; for a symmetric PSD
        # 3x3 the |R_ii| are within a small factor of the singular values
        sv = np.linalg.svd(A, compute_uv=False)
        if sv[2] > 1e-4 * sv[0]:
            assert ok and np.allclose(X, np.linalg.solve(A, rhs), rtol=1e-9, atol=1e-9)
        elif sv[2] < 1e-7 * sv[0]:
            assert not ok
            n_fail += 1
    assert n_fail > 10


def test_two_observers_with_parallel_bearings_are_rejected(oracle):
    intr = np.array([400.0, 400.0, 320.0, 256.0])
    Rp = np.array([np.r_[np.eye(3).ravel(), [0.0, 0.0, 0.0]], np.r_[np.eye(3).ravel(), [0.0, 0.0, 1.0]]])
    uv = np.array([[320.0, 256.0], [320.0, 256.0]])  # both look down the baseline: depth unobservable
    ok, _ = oracle.triangulate(Rp, uv, intr)
    assert not ok
