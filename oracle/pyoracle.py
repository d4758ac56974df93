"""ctypes wrapper of the CPU parity oracle (oracle/liburmvo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (ur-mvo_b200/) never imports this.
PARITY UNPINNED — see oracle/oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
TRACE_MAX = 64


class TraceRow(C.Structure):
    _fields_ = [("chi2_before", C.c_double), ("chi2_after", C.c_double), ("lambda_after", C.c_double),
                ("trials", C.c_int32), ("accepted", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("iters", C.c_int32 * 4), ("chi2_final", C.c_double * 4),
                ("lambda_final", C.c_double * 4), ("trace", TraceRow * TRACE_MAX)]

    def rows(self):
        return [(r.chi2_before, r.chi2_after, r.lambda_after, r.trials, r.accepted)
                for r in self.trace[:self.n_rows]]


class TVStats(C.Structure):
    _fields_ = [("SH", C.c_float), ("SF", C.c_float), ("best_H", C.c_int32), ("best_F", C.c_int32),
                ("H21", C.c_float * 9), ("F21", C.c_float * 9), ("used_H", C.c_int32),
                ("n_good", C.c_int32 * 8), ("parallax", C.c_float * 8), ("best_motion", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "liburmvo_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.urmvo_oracle_local_ba.restype = C.c_int
        _LIB.urmvo_oracle_pose_only.restype = C.c_int
        _LIB.urmvo_oracle_two_view.restype = C.c_int
        _LIB.urmvo_oracle_score_all.restype = C.c_int
        _LIB.urmvo_oracle_two_view_mode.restype = C.c_int
        _LIB.urmvo_oracle_score_all_mode.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def local_ba(prob, chi2_thr=10.0, it0=10, it1=5):
    """Returns (poses, pts, inlier, Stats). Inputs are not modified."""
    poses = np.ascontiguousarray(prob["poses"], dtype=np.float64).copy()
    pts = np.ascontiguousarray(prob["pts"], dtype=np.float64).copy()
    fixed = np.ascontiguousarray(prob["fixed"], dtype=np.uint8)
    uv = np.ascontiguousarray(prob["uv"], dtype=np.float64)
    cam = np.ascontiguousarray(prob["obs_cam"], dtype=np.int32)
    pt = np.ascontiguousarray(prob["obs_pt"], dtype=np.int32)
    intr = np.ascontiguousarray(prob["intr"], dtype=np.float64)
    inl = np.zeros(uv.shape[0], dtype=np.uint8)
    st = Stats()
    lib().urmvo_oracle_local_ba(C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]), _p(pts),
                                C.c_int(uv.shape[0]), _p(uv), _p(cam), _p(pt), _p(intr), C.c_double(chi2_thr),
                                C.c_int(it0), C.c_int(it1), _p(inl), C.byref(st))
    return poses, pts, inl, st


def local_ba_stereo(prob, chi2_thr=10.0, chi2_thr_stereo=75.0, it0=10, it1=5):
    """Stereo camera: prob has uv3 (No, 3), kind (No,), intr5 = fx fy cx cy bf. Returns (poses, pts, inlier, Stats)."""
    poses = np.ascontiguousarray(prob["poses"], dtype=np.float64).copy()
    pts = np.ascontiguousarray(prob["pts"], dtype=np.float64).copy()
    fixed = np.ascontiguousarray(prob["fixed"], dtype=np.uint8)
    uv3 = np.ascontiguousarray(prob["uv3"], dtype=np.float64)
    kind = np.ascontiguousarray(prob["kind"], dtype=np.uint8)
    cam = np.ascontiguousarray(prob["obs_cam"], dtype=np.int32)
    pt = np.ascontiguousarray(prob["obs_pt"], dtype=np.int32)
    intr5 = np.ascontiguousarray(prob["intr5"], dtype=np.float64)
    inl = np.zeros(uv3.shape[0], dtype=np.uint8)
    st = Stats()
    lib().urmvo_oracle_local_ba_stereo(C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]), _p(pts),
                                       C.c_int(uv3.shape[0]), _p(uv3), _p(kind), _p(cam), _p(pt), _p(intr5),
                                       C.c_double(chi2_thr), C.c_double(chi2_thr_stereo), C.c_int(it0), C.c_int(it1),
                                       _p(inl), C.byref(st))
    return poses, pts, inl, st


def pose_only_batch_stereo(batch, chi2_thr=10.0, chi2_thr_stereo=75.0, rounds=4, its=10):
    poses = np.ascontiguousarray(batch["poses"], dtype=np.float64).copy()
    off = np.ascontiguousarray(batch["obs_offset"], dtype=np.int32)
    uv3 = np.ascontiguousarray(batch["uv3"], dtype=np.float64)
    kind = np.ascontiguousarray(batch["kind"], dtype=np.uint8)
    Xw = np.ascontiguousarray(batch["Xw"], dtype=np.float64)
    intr5 = np.ascontiguousarray(batch["intr5"], dtype=np.float64)
    inl = np.ones(uv3.shape[0], dtype=np.uint8)
    n_inl = np.zeros(poses.shape[0], dtype=np.int32)
    lib().urmvo_oracle_pose_only_stereo.restype = C.c_int
    for f in range(poses.shape[0]):
        o0, o1 = int(off[f]), int(off[f + 1])
        pose = poses[f].copy(); i_f = inl[o0:o1].copy()
        u = np.ascontiguousarray(uv3[o0:o1]); k = np.ascontiguousarray(kind[o0:o1]); X = np.ascontiguousarray(Xw[o0:o1])
        n_inl[f] = lib().urmvo_oracle_pose_only_stereo(_p(pose), C.c_int(o1 - o0), _p(u), _p(k), _p(X), _p(intr5),
                                                       C.c_double(chi2_thr), C.c_double(chi2_thr_stereo), C.c_int(rounds),
                                                       C.c_int(its), _p(i_f), None)
        poses[f] = pose; inl[o0:o1] = i_f
    return poses, inl, n_inl


def local_ba_multicam(prob, chi2_thr=10.0, chi2_thr_stereo=75.0, it0=10, it1=5):
    """Per-constraint camera models: prob has uv3 (No, 3), kind_model (No,) = stereo bit | model << 1 and
    intr5_tab (n_models, 5).  Returns (poses, pts, inlier, Stats)."""
    poses = np.ascontiguousarray(prob["poses"], dtype=np.float64).copy()
    pts = np.ascontiguousarray(prob["pts"], dtype=np.float64).copy()
    fixed = np.ascontiguousarray(prob["fixed"], dtype=np.uint8)
    uv3 = np.ascontiguousarray(prob["uv3"], dtype=np.float64)
    km = np.ascontiguousarray(prob["kind_model"], dtype=np.uint8)
    cam = np.ascontiguousarray(prob["obs_cam"], dtype=np.int32)
    pt = np.ascontiguousarray(prob["obs_pt"], dtype=np.int32)
    tab = np.ascontiguousarray(prob["intr5_tab"], dtype=np.float64).reshape(-1, 5)
    inl = np.zeros(uv3.shape[0], dtype=np.uint8)
    st = Stats()
    lib().urmvo_oracle_local_ba_multicam.restype = C.c_int
    rc = lib().urmvo_oracle_local_ba_multicam(C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]),
                                              _p(pts), C.c_int(uv3.shape[0]), _p(uv3), _p(km), _p(cam), _p(pt),
                                              C.c_int(tab.shape[0]), _p(tab), C.c_double(chi2_thr),
                                              C.c_double(chi2_thr_stereo), C.c_int(it0), C.c_int(it1), _p(inl),
                                              C.byref(st))
    if rc < 0:
        raise ValueError("bad camera model index")
    return poses, pts, inl, st


def pose_only_batch_multicam(batch, chi2_thr=10.0, chi2_thr_stereo=75.0, rounds=4, its=10):
    poses = np.ascontiguousarray(batch["poses"], dtype=np.float64).copy()
    off = np.ascontiguousarray(batch["obs_offset"], dtype=np.int32)
    uv3 = np.ascontiguousarray(batch["uv3"], dtype=np.float64)
    km = np.ascontiguousarray(batch["kind_model"], dtype=np.uint8)
    Xw = np.ascontiguousarray(batch["Xw"], dtype=np.float64)
    tab = np.ascontiguousarray(batch["intr5_tab"], dtype=np.float64).reshape(-1, 5)
    inl = np.ones(uv3.shape[0], dtype=np.uint8)
    n_inl = np.zeros(poses.shape[0], dtype=np.int32)
    lib().urmvo_oracle_pose_only_multicam.restype = C.c_int
    for f in range(poses.shape[0]):
        o0, o1 = int(off[f]), int(off[f + 1])
        pose = poses[f].copy(); i_f = inl[o0:o1].copy()
        u = np.ascontiguousarray(uv3[o0:o1]); k = np.ascontiguousarray(km[o0:o1]); X = np.ascontiguousarray(Xw[o0:o1])
        n_inl[f] = lib().urmvo_oracle_pose_only_multicam(_p(pose), C.c_int(o1 - o0), _p(u), _p(k), _p(X),
                                                         C.c_int(tab.shape[0]), _p(tab), C.c_double(chi2_thr),
                                                         C.c_double(chi2_thr_stereo), C.c_int(rounds), C.c_int(its),
                                                         _p(i_f), None)
        poses[f] = pose; inl[o0:o1] = i_f
    return poses, inl, n_inl


def edge_stereo(Tcw, X, uv3, intr5):
    """EdgeStereoSE3ProjectXYZ: returns (e[3], Jpose[3,6], Jpoint[3,3], depth_positive)."""
    Tcw = np.ascontiguousarray(Tcw, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
    uv3 = np.ascontiguousarray(uv3, dtype=np.float64); intr5 = np.ascontiguousarray(intr5, dtype=np.float64)
    e = np.zeros(3); Jp = np.zeros((3, 6)); Jx = np.zeros((3, 3))
    lib().urmvo_oracle_edge_stereo.restype = C.c_int
    dp = lib().urmvo_oracle_edge_stereo(_p(Tcw), _p(X), _p(uv3), _p(intr5), _p(e), _p(Jp), _p(Jx))
    return e, Jp, Jx, bool(dp)


def pose_only(pose, uv, Xw, intr, chi2_thr=10.0, rounds=4, its=10, inlier=None):
    pose = np.ascontiguousarray(pose, dtype=np.float64).copy()
    uv = np.ascontiguousarray(uv, dtype=np.float64)
    Xw = np.ascontiguousarray(Xw, dtype=np.float64)
    intr = np.ascontiguousarray(intr, dtype=np.float64)
    inl = np.ones(uv.shape[0], dtype=np.uint8) if inlier is None else np.ascontiguousarray(inlier, dtype=np.uint8).copy()
    st = Stats()
    n = lib().urmvo_oracle_pose_only(_p(pose), C.c_int(uv.shape[0]), _p(uv), _p(Xw), _p(intr), C.c_double(chi2_thr),
                                     C.c_int(rounds), C.c_int(its), _p(inl), C.byref(st))
    return pose, inl, n, st


def pose_only_batch(batch, chi2_thr=10.0, rounds=4, its=10, n_threads=1):
    poses = np.ascontiguousarray(batch["poses"], dtype=np.float64).copy()
    off = np.ascontiguousarray(batch["obs_offset"], dtype=np.int32)
    uv = np.ascontiguousarray(batch["uv"], dtype=np.float64)
    Xw = np.ascontiguousarray(batch["Xw"], dtype=np.float64)
    intr = np.ascontiguousarray(batch["intr"], dtype=np.float64)
    inl = np.ones(uv.shape[0], dtype=np.uint8)
    B = poses.shape[0]
    n_inl = np.zeros(B, dtype=np.int32)
    lib().urmvo_oracle_pose_only_batch(C.c_int(B), _p(off), _p(poses), _p(uv), _p(Xw), _p(intr), C.c_double(chi2_thr),
                                       C.c_int(rounds), C.c_int(its), _p(inl), _p(n_inl), C.c_int(n_threads))
    return poses, inl, n_inl


def two_view(tv, sets=None, score_mode=0):
    k1 = np.ascontiguousarray(tv["keys1"], dtype=np.float32)
    k2 = np.ascontiguousarray(tv["keys2"], dtype=np.float32)
    m = np.ascontiguousarray(tv["matches12"], dtype=np.int32)
    K = np.ascontiguousarray(tv["K"], dtype=np.float32)
    sets = np.ascontiguousarray(tv["sets"] if sets is None else sets, dtype=np.int32)
    N = int((m >= 0).sum())
    T21 = np.zeros((4, 4), dtype=np.float32)
    P3D = np.zeros((k1.shape[0], 3), dtype=np.float32)
    tri = np.zeros(k1.shape[0], dtype=np.uint8)
    mH = np.zeros(N, dtype=np.uint8)
    mF = np.zeros(N, dtype=np.uint8)
    st = TVStats()
    ok = lib().urmvo_oracle_two_view_mode(C.c_int(k1.shape[0]), _p(k1), C.c_int(k2.shape[0]), _p(k2), _p(m), _p(K),
                                          C.c_float(tv.get("sigma", 1.0)), C.c_int(sets.shape[0]), _p(sets),
                                          C.c_int(score_mode), _p(T21), _p(P3D), _p(tri), _p(mH), _p(mF), C.byref(st))
    return dict(ok=bool(ok), T21=T21, P3D=P3D, triangulated=tri, mask_H=mH, mask_F=mF, stats=st)


def score_all(tv, model, sets=None, score_mode=0):
    """model 0 = F, 1 = H. Returns scores[n_hyp], masks[n_hyp, words], models[n_hyp, 9]."""
    k1 = np.ascontiguousarray(tv["keys1"], dtype=np.float32)
    k2 = np.ascontiguousarray(tv["keys2"], dtype=np.float32)
    m = np.ascontiguousarray(tv["matches12"], dtype=np.int32)
    sets = np.ascontiguousarray(tv["sets"] if sets is None else sets, dtype=np.int32)
    N = int((m >= 0).sum())
    words = (N + 31) // 32
    scores = np.zeros(sets.shape[0], dtype=np.float32)
    masks = np.zeros((sets.shape[0], words), dtype=np.uint32)
    models = np.zeros((sets.shape[0], 9), dtype=np.float32)
    lib().urmvo_oracle_score_all_mode(C.c_int(k1.shape[0]), _p(k1), C.c_int(k2.shape[0]), _p(k2), _p(m),
                                      C.c_float(tv.get("sigma", 1.0)), C.c_int(sets.shape[0]), _p(sets), C.c_int(model),
                                      C.c_int(score_mode), _p(scores), _p(masks), _p(models))
    return scores, masks, models


def draw_sets(N, n_hyp, seed=0):
    sets = np.zeros((n_hyp, 8), dtype=np.int32)
    lib().urmvo_oracle_draw_sets(C.c_int(N), C.c_int(n_hyp), C.c_int(1), C.c_int(seed), _p(sets))
    return sets


def svd(A):
    A = np.ascontiguousarray(A, dtype=np.float32)
    m, n = A.shape
    s = np.zeros(n, dtype=np.float32)
    U = np.zeros((m, n), dtype=np.float32)
    V = np.zeros((n, n), dtype=np.float32)
    lib().urmvo_oracle_svd(C.c_int(m), C.c_int(n), _p(A), _p(s), _p(U), _p(V))
    return U, s, V


def edge(Tcw, X, uv, intr):
    e = np.zeros(2); Jp = np.zeros((2, 6)); Jx = np.zeros((2, 3))
    Tcw = np.ascontiguousarray(Tcw, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
    uv = np.ascontiguousarray(uv, dtype=np.float64); intr = np.ascontiguousarray(intr, dtype=np.float64)
    pos = lib().urmvo_oracle_edge(_p(Tcw), _p(X), _p(uv), _p(intr), _p(e), _p(Jp), _p(Jx))
    return e, Jp, Jx, bool(pos)


def huber(e2, delta):
    rho = np.zeros(3)
    lib().urmvo_oracle_huber(C.c_double(e2), C.c_double(delta), _p(rho))
    return rho


def se3_oplus(Tcw, upd):
    T = np.ascontiguousarray(Tcw, dtype=np.float64).copy()
    upd = np.ascontiguousarray(upd, dtype=np.float64)
    lib().urmvo_oracle_se3_oplus(_p(T), _p(upd))
    return T


def se3_inverse(T):
    T = np.ascontiguousarray(T, dtype=np.float64)
    out = np.zeros(7)
    lib().urmvo_oracle_se3_inverse(_p(T), _p(out))
    return out


# ---- per-frame fundamental-matrix RANSAC (cv::findFundamentalMat FM_RANSAC restatement, fm_oracle.cpp)

def fm_ransac(p0, p1, thresh=3.0, confidence=0.99, max_iters=1000):
    """Returns dict(found, mask[N] u8, F[3,3], iters, n_inliers, models)."""
    p0 = np.ascontiguousarray(p0, dtype=np.float32); p1 = np.ascontiguousarray(p1, dtype=np.float32)
    N = len(p0)
    mask = np.zeros(N, dtype=np.uint8); F = np.zeros(9); st = np.zeros(3, dtype=np.int32)
    rc = lib().urmvo_oracle_fm_ransac(C.c_int(N), _p(p0), _p(p1), C.c_double(thresh), C.c_double(confidence),
                                      C.c_int(max_iters), _p(mask), _p(F), _p(st))
    return dict(found=rc, mask=mask, F=F.reshape(3, 3), iters=int(st[0]), n_inliers=int(st[1]), models=int(st[2]))


def find_fundamental(p0, p1, thresh=3.0, confidence=0.99, max_iters=1000):
    """cv::findFundamentalMat(FM_RANSAC) for any N >= 7 (direct 7-point / LMedS / RANSAC). Same dict as fm_ransac."""
    p0 = np.ascontiguousarray(p0, dtype=np.float32); p1 = np.ascontiguousarray(p1, dtype=np.float32)
    N = len(p0)
    mask = np.zeros(N, dtype=np.uint8); F = np.zeros(9); st = np.zeros(3, dtype=np.int32)
    lib().urmvo_oracle_find_fundamental.restype = C.c_int
    rc = lib().urmvo_oracle_find_fundamental(C.c_int(N), _p(p0), _p(p1), C.c_double(thresh), C.c_double(confidence),
                                             C.c_int(max_iters), _p(mask), _p(F), _p(st))
    return dict(found=rc, mask=mask, F=F.reshape(3, 3), iters=int(st[0]), n_inliers=int(st[1]), models=int(st[2]))


def fm_subsets(p0, p1, max_iters=1000):
    p0 = np.ascontiguousarray(p0, dtype=np.float32); p1 = np.ascontiguousarray(p1, dtype=np.float32)
    idx = np.zeros((max_iters, 7), dtype=np.int32)
    n = lib().urmvo_oracle_fm_subsets(C.c_int(len(p0)), _p(p0), _p(p1), C.c_int(max_iters), _p(idx))
    return idx[:n]


def fm_run7(p0, p1):
    p0 = np.ascontiguousarray(p0, dtype=np.float32); p1 = np.ascontiguousarray(p1, dtype=np.float32)
    F = np.zeros(27)
    n = lib().urmvo_oracle_fm_run7(_p(p0), _p(p1), _p(F))
    return F.reshape(3, 3, 3)[:max(n, 0)]


def fm_errors(p0, p1, F):
    p0 = np.ascontiguousarray(p0, dtype=np.float32); p1 = np.ascontiguousarray(p1, dtype=np.float32)
    F = np.ascontiguousarray(F, dtype=np.float64)
    err = np.zeros(len(p0), dtype=np.float32)
    lib().urmvo_oracle_fm_errors(C.c_int(len(p0)), _p(p0), _p(p1), _p(F), _p(err))
    return err


# ---- Mapping::TriangulateMappoint (tri_oracle.cpp)

def triangulate(Rp, uv, intr):
    """Rp: (n,12) [R row-major | p] of the observing keyframes' T_wc, uv: (n,2). Returns (ok, X[3])."""
    Rp = np.ascontiguousarray(Rp, dtype=np.float64); uv = np.ascontiguousarray(uv, dtype=np.float64)
    intr = np.ascontiguousarray(intr, dtype=np.float64)
    X = np.zeros(3)
    lib().urmvo_oracle_triangulate.restype = C.c_int
    ok = lib().urmvo_oracle_triangulate(C.c_int(len(uv)), _p(Rp), _p(uv), _p(intr), _p(X))
    return bool(ok), X


# ---- SolvePnPWithCV (cv::solvePnPRansac restatement, pnp_oracle.cpp)

def pnp_ransac(obj, img, K4, max_iters=100, reproj=20.0, confidence=0.99):
    """Returns dict(found, R[3,3], t[3] (T_cw after the refinement over the inliers), mask[N] u8, iters, n_inliers,
    models, counts[iters])."""
    obj = np.ascontiguousarray(obj, dtype=np.float32); img = np.ascontiguousarray(img, dtype=np.float32)
    K4 = np.ascontiguousarray(K4, dtype=np.float64)
    N = len(obj)
    R = np.zeros(9); t = np.zeros(3); mask = np.zeros(N, dtype=np.uint8); st = np.zeros(4, dtype=np.int32)
    counts = np.zeros(max(max_iters, 1), dtype=np.int32)
    rc = lib().urmvo_oracle_pnp_ransac(C.c_int(N), _p(obj), _p(img), _p(K4), C.c_int(max_iters), C.c_double(reproj),
                                       C.c_double(confidence), _p(R), _p(t), _p(mask), _p(st), _p(counts))
    return dict(found=rc, R=R.reshape(3, 3), t=t, mask=mask, iters=int(st[0]), n_inliers=int(st[1]), models=int(st[2]),
                counts=counts[:int(st[0])])


def pnp_subsets(N, max_iters=100):
    idx = np.zeros((max_iters, 5), dtype=np.int32)
    lib().urmvo_oracle_pnp_subsets(C.c_int(N), C.c_int(max_iters), _p(idx))
    return idx


def pnp_epnp5(obj, img, K4, idx5):
    obj = np.ascontiguousarray(obj, dtype=np.float32); img = np.ascontiguousarray(img, dtype=np.float32)
    K4 = np.ascontiguousarray(K4, dtype=np.float64); idx5 = np.ascontiguousarray(idx5, dtype=np.int32)
    R = np.zeros(9); t = np.zeros(3)
    ok = lib().urmvo_oracle_pnp_epnp5(_p(obj), _p(img), _p(K4), _p(idx5), _p(R), _p(t))
    return ok, R.reshape(3, 3), t
