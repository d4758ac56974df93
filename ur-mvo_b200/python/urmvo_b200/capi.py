import ctypes as C
import os
import subprocess
import numpy as np

_PKG_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))  # ur-mvo_b200/
_LIB = None

EXPORTED_SYMBOLS = [
    "urmvo_version", "urmvo_last_error", "urmvo_create", "urmvo_destroy", "urmvo_stream", "urmvo_sync",
    "urmvo_launch_count", "urmvo_local_ba", "urmvo_local_ba_batch", "urmvo_local_ba_stereo", "urmvo_local_ba_batch_stereo",
    "urmvo_pose_only_batch_stereo", "urmvo_local_ba_multicam", "urmvo_local_ba_batch_multicam", "urmvo_pose_only_batch_multicam",
    "urmvo_ba_plan_create",
    "urmvo_ba_plan_run", "urmvo_ba_plan_download", "urmvo_ba_plan_destroy", "urmvo_ba_plan_phase_info", "urmvo_debug_ba_timing", "urmvo_debug_lg_timing", "urmvo_nccl_unique_id", "urmvo_comm_init", "urmvo_ba_covisibility",
    "urmvo_sharded_ba_create", "urmvo_sharded_ba_run", "urmvo_pose_only_batch",
    "urmvo_pose_plan_create", "urmvo_pose_plan_run", "urmvo_pose_plan_download", "urmvo_pose_plan_destroy",
    "urmvo_two_view", "urmvo_two_view_scored", "urmvo_tv_plan_set_score_mode", "urmvo_tv_plan_create", "urmvo_tv_plan_run_ransac", "urmvo_tv_plan_download_hyps",
    "urmvo_tv_plan_reconstruct", "urmvo_tv_plan_destroy",
    "urmvo_fm_ransac", "urmvo_fm_ransac_batch", "urmvo_fm_plan_create", "urmvo_fm_plan_run", "urmvo_fm_plan_finish",
    "urmvo_fm_plan_hypotheses", "urmvo_fm_plan_destroy", "urmvo_triangulate_batch",
    "urmvo_pnp_ransac", "urmvo_pnp_ransac_batch",
    "urmvo_map_create", "urmvo_map_destroy", "urmvo_map_set_keyframes", "urmvo_map_set_points", "urmvo_map_add_observations",
    "urmvo_map_remove_observations", "urmvo_map_get_keyframes", "urmvo_map_get_points", "urmvo_map_local_ba",
]


class UrmvoError(RuntimeError):
    pass


class FMStats(C.Structure):
    _fields_ = [("found", C.c_int32), ("iters", C.c_int32), ("n_inliers", C.c_int32), ("n_models", C.c_int32),
                ("F", C.c_double * 9)]


class PnPStats(C.Structure):
    _fields_ = [("found", C.c_int32), ("iters", C.c_int32), ("n_inliers", C.c_int32), ("n_models", C.c_int32),
                ("R", C.c_double * 9), ("t", C.c_double * 3)]


class BAOptions(C.Structure):
    _fields_ = [("pcg_tol", C.c_double), ("pcg_max_iter", C.c_int32), ("cluster_size", C.c_int32),
                ("threads", C.c_int32), ("force_atomic", C.c_int32), ("dense_solver", C.c_int32),
                ("large_mode", C.c_int32), ("band_solver", C.c_int32)]


class BAStats(C.Structure):
    _fields_ = [("iters", C.c_int32 * 2), ("trials", C.c_int32 * 2), ("pcg_iters", C.c_int32 * 2),
                ("n_level1", C.c_int32), ("chi2_initial", C.c_double), ("chi2_final", C.c_double * 2),
                ("lambda_final", C.c_double * 2)]


class TVStats(C.Structure):
    _fields_ = [("SH", C.c_float), ("SF", C.c_float), ("best_H", C.c_int32), ("best_F", C.c_int32),
                ("H21", C.c_float * 9), ("F21", C.c_float * 9), ("used_H", C.c_int32),
                ("n_good", C.c_int32 * 8), ("parallax", C.c_float * 8), ("best_motion", C.c_int32)]


def lib_path():
    # URMVO_B200_LIB: development override for A/B runs of two builds of the same library
    return os.environ.get("URMVO_B200_LIB") or os.path.join(_PKG_ROOT, "lib", "liburmvo_b200.so")


def build_library():
    subprocess.check_call(["make", "-C", _PKG_ROOT, "-s"])
    return lib_path()


def load_library():
    """Loads liburmvo_b200.so. Raises (never falls back) when the CUDA extension is missing."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise UrmvoError(f"{path} is missing: build it with `make -C {_PKG_ROOT}` "
                             "(python __graft_entry__.py build). There is no CPU fallback.")
        L = C.CDLL(path)
        L.urmvo_last_error.restype = C.c_char_p
        L.urmvo_stream.restype = C.c_void_p
        L.urmvo_launch_count.restype = C.c_int64
        for name in EXPORTED_SYMBOLS:
            f = getattr(L, name)
            if name not in ("urmvo_last_error", "urmvo_stream", "urmvo_launch_count", "urmvo_destroy",
                            "urmvo_ba_plan_destroy", "urmvo_pose_plan_destroy", "urmvo_tv_plan_destroy",
                            "urmvo_fm_plan_destroy", "urmvo_map_destroy"):
                f.restype = C.c_int
        for name in ("urmvo_destroy", "urmvo_ba_plan_destroy", "urmvo_pose_plan_destroy", "urmvo_tv_plan_destroy",
                     "urmvo_fm_plan_destroy", "urmvo_map_destroy"):
            getattr(L, name).restype = None
        _LIB = L
    return _LIB


def _check(rc, what):
    if rc != 0:
        raise UrmvoError(f"{what} failed ({rc}): {load_library().urmvo_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Context:
    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.urmvo_create(C.byref(self._h), C.c_int(device)), "urmvo_create")

    def close(self):
        if self._h:
            self._L.urmvo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self._L.urmvo_stream(self._h)

    def sync(self):
        _check(self._L.urmvo_sync(self._h), "urmvo_sync")

    def comm_init(self, rank, world, unique_id):
        """Joins the NCCL communicator used by the point-sharded BA (one process per GPU)."""
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        _check(self._L.urmvo_comm_init(self._h, C.c_int(rank), C.c_int(world), _p(uid)), "urmvo_comm_init")
        self.rank, self.world = rank, world

    def lg_timing(self, reset=True):
        """SM cycles per segment of the tile-mode band solve since the last reset (development aid)."""
        t = (C.c_uint64 * 8)()
        _check(self._L.urmvo_debug_lg_timing(t, C.c_int(1 if reset else 0)), "urmvo_debug_lg_timing")
        return list(t)

    def ba_timing(self, reset=True):
        """SM cycles per BA phase of window 0 since the last reset (development aid)."""
        out = (C.c_uint64 * 8)()
        _check(self._L.urmvo_debug_ba_timing(out, C.c_int(1 if reset else 0)), "urmvo_debug_ba_timing")
        return list(out)

    @property
    def launches(self):
        return int(self._L.urmvo_launch_count(self._h))

    # ---- SolvePnPWithCV (cv::solvePnPRansac)
    def pnp_ransac_batch(self, problems, intr, max_iters=100, reproj=20.0, confidence=0.99):
        """problems: list of (obj[N,3] f32, img[N,2] f32).  Returns a list of dict(found, R, t, mask, iters, n_inliers, models)."""
        off = np.zeros(len(problems) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(o) for o, _ in problems])
        obj = _f32(np.concatenate([np.asarray(o, dtype=np.float32).reshape(-1, 3) for o, _ in problems]))
        img = _f32(np.concatenate([np.asarray(i, dtype=np.float32).reshape(-1, 2) for _, i in problems]))
        K4 = _f64(intr)
        inl = np.zeros(int(off[-1]), dtype=np.uint8)
        st = (PnPStats * len(problems))()
        _check(self._L.urmvo_pnp_ransac_batch(self._h, C.c_int(len(problems)), _p(off), _p(obj), _p(img), _p(K4),
                                              C.c_int(max_iters), C.c_double(reproj), C.c_double(confidence), _p(inl), st),
               "urmvo_pnp_ransac_batch")
        return [dict(found=int(s.found), R=np.array(s.R).reshape(3, 3), t=np.array(s.t), mask=inl[off[b]:off[b + 1]].copy(),
                     iters=int(s.iters), n_inliers=int(s.n_inliers), models=int(s.n_models)) for b, s in enumerate(st)]

    def pnp_ransac(self, obj, img, intr, max_iters=100, reproj=20.0, confidence=0.99):
        return self.pnp_ransac_batch([(obj, img)], intr, max_iters, reproj, confidence)[0]

    # ---- one-shot, host buffers in / out (the reference-facing calls)
    def local_ba(self, prob, chi2_thr=10.0, it0=10, it1=5, opts=None):
        poses = _f64(prob["poses"]).copy(); pts = _f64(prob["pts"]).copy()
        fixed = _u8(prob["fixed"]); uv = _f64(prob["uv"]); cam = _i32(prob["obs_cam"]); pt = _i32(prob["obs_pt"])
        intr = _f64(prob["intr"])
        inl = np.zeros(uv.shape[0], dtype=np.uint8)
        st = BAStats()
        _check(self._L.urmvo_local_ba(self._h, C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]),
                                      _p(pts), C.c_int(uv.shape[0]), _p(uv), _p(cam), _p(pt), _p(intr),
                                      C.c_double(chi2_thr), C.c_int(it0), C.c_int(it1), _p(inl), C.byref(st),
                                      C.byref(opts) if opts is not None else None), "urmvo_local_ba")
        return poses, pts, inl, st

    def local_ba_stereo(self, prob, chi2_thr=10.0, chi2_thr_stereo=75.0, it0=10, it1=5, opts=None):
        """prob: dict like synth.make_ba_stereo: uv3 (No, 3), kind (No,), intr5 = fx fy cx cy bf."""
        poses = _f64(prob["poses"]).copy(); pts = _f64(prob["pts"]).copy()
        fixed = _u8(prob["fixed"]); uv3 = _f64(prob["uv3"]); kind = _u8(prob["kind"])
        cam = _i32(prob["obs_cam"]); pt = _i32(prob["obs_pt"]); intr5 = _f64(prob["intr5"])
        inl = np.zeros(uv3.shape[0], dtype=np.uint8)
        st = BAStats()
        _check(self._L.urmvo_local_ba_stereo(self._h, C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]),
                                             _p(pts), C.c_int(uv3.shape[0]), _p(uv3), _p(kind), _p(cam), _p(pt), _p(intr5),
                                             C.c_double(chi2_thr), C.c_double(chi2_thr_stereo), C.c_int(it0), C.c_int(it1),
                                             _p(inl), C.byref(st), C.byref(opts) if opts is not None else None),
               "urmvo_local_ba_stereo")
        return poses, pts, inl, st

    def pose_only_batch_stereo(self, batch, chi2_thr=10.0, chi2_thr_stereo=75.0, rounds=4, its=10, inlier=None):
        """batch: uv3 (No, 3), kind (No,), Xw, obs_offset, poses, intr5."""
        poses = _f64(batch["poses"]).copy()
        off = _i32(batch["obs_offset"]); uv3 = _f64(batch["uv3"]); kind = _u8(batch["kind"]); Xw = _f64(batch["Xw"])
        intr5 = _f64(batch["intr5"])
        inl = np.ones(uv3.shape[0], dtype=np.uint8) if inlier is None else _u8(inlier).copy()
        n_inl = np.zeros(poses.shape[0], dtype=np.int32)
        _check(self._L.urmvo_pose_only_batch_stereo(self._h, C.c_int(poses.shape[0]), _p(off), _p(poses), _p(uv3), _p(kind),
                                                    _p(Xw), _p(intr5), C.c_double(chi2_thr), C.c_double(chi2_thr_stereo),
                                                    C.c_int(rounds), C.c_int(its), _p(inl), _p(n_inl)),
               "urmvo_pose_only_batch_stereo")
        return poses, inl, n_inl

    def local_ba_multicam(self, prob, chi2_thr=10.0, chi2_thr_stereo=75.0, it0=10, it1=5, opts=None):
        """Several camera models in one window (camera_list[mpc->id_camera]): prob has uv3 (No, 3),
        kind_model (No,) = stereo bit | model << 1, intr5_tab (n_models, 5)."""
        poses = _f64(prob["poses"]).copy(); pts = _f64(prob["pts"]).copy()
        fixed = _u8(prob["fixed"]); uv3 = _f64(prob["uv3"]); km = _u8(prob["kind_model"])
        cam = _i32(prob["obs_cam"]); pt = _i32(prob["obs_pt"]); tab = _f64(prob["intr5_tab"]).reshape(-1, 5)
        inl = np.zeros(uv3.shape[0], dtype=np.uint8)
        st = BAStats()
        _check(self._L.urmvo_local_ba_multicam(self._h, C.c_int(poses.shape[0]), _p(poses), _p(fixed), C.c_int(pts.shape[0]),
                                               _p(pts), C.c_int(uv3.shape[0]), _p(uv3), _p(km), _p(cam), _p(pt),
                                               C.c_int(tab.shape[0]), _p(tab), C.c_double(chi2_thr),
                                               C.c_double(chi2_thr_stereo), C.c_int(it0), C.c_int(it1), _p(inl),
                                               C.byref(st), C.byref(opts) if opts is not None else None),
               "urmvo_local_ba_multicam")
        return poses, pts, inl, st

    def pose_only_batch_multicam(self, batch, chi2_thr=10.0, chi2_thr_stereo=75.0, rounds=4, its=10, inlier=None):
        """batch: uv3 (No, 3), kind_model (No,), Xw, obs_offset, poses, intr5_tab (n_models, 5)."""
        poses = _f64(batch["poses"]).copy()
        off = _i32(batch["obs_offset"]); uv3 = _f64(batch["uv3"]); km = _u8(batch["kind_model"]); Xw = _f64(batch["Xw"])
        tab = _f64(batch["intr5_tab"]).reshape(-1, 5)
        inl = np.ones(uv3.shape[0], dtype=np.uint8) if inlier is None else _u8(inlier).copy()
        n_inl = np.zeros(poses.shape[0], dtype=np.int32)
        _check(self._L.urmvo_pose_only_batch_multicam(self._h, C.c_int(poses.shape[0]), _p(off), _p(poses), _p(uv3), _p(km),
                                                      _p(Xw), C.c_int(tab.shape[0]), _p(tab), C.c_double(chi2_thr),
                                                      C.c_double(chi2_thr_stereo), C.c_int(rounds), C.c_int(its),
                                                      _p(inl), _p(n_inl)),
               "urmvo_pose_only_batch_multicam")
        return poses, inl, n_inl

    def local_ba_batch(self, batch, chi2_thr=10.0, it0=10, it1=5, opts=None, out=None):
        """batch: dict from pack_ba_batch(). Returns (poses, pts, inlier, [BAStats])."""
        B = len(batch["cam_off"]) - 1
        poses = batch["poses"].copy() if out is None else out["poses"]
        pts = batch["pts"].copy() if out is None else out["pts"]
        if out is not None:
            poses[...] = batch["poses"]; pts[...] = batch["pts"]
        inl = np.zeros(batch["uv"].shape[0], dtype=np.uint8) if out is None else out["inlier"]
        st = (BAStats * B)()
        _check(self._L.urmvo_local_ba_batch(self._h, C.c_int(B), _p(batch["cam_off"]), _p(batch["pt_off"]),
                                            _p(batch["obs_off"]), _p(poses), _p(batch["fixed"]), _p(pts),
                                            _p(batch["uv"]), _p(batch["obs_cam"]), _p(batch["obs_pt"]),
                                            _p(batch["intr"]), C.c_double(chi2_thr), C.c_int(it0), C.c_int(it1),
                                            _p(inl), st, C.byref(opts) if opts is not None else None),
               "urmvo_local_ba_batch")
        return poses, pts, inl, list(st)

    def pose_only_batch(self, batch, chi2_thr=10.0, rounds=4, its=10, inlier=None):
        poses = _f64(batch["poses"]).copy()
        off = _i32(batch["obs_offset"]); uv = _f64(batch["uv"]); Xw = _f64(batch["Xw"]); intr = _f64(batch["intr"])
        inl = np.ones(uv.shape[0], dtype=np.uint8) if inlier is None else _u8(inlier).copy()
        n_inl = np.zeros(poses.shape[0], dtype=np.int32)
        _check(self._L.urmvo_pose_only_batch(self._h, C.c_int(poses.shape[0]), _p(off), _p(poses), _p(uv), _p(Xw),
                                             _p(intr), C.c_double(chi2_thr), C.c_int(rounds), C.c_int(its),
                                             _p(inl), _p(n_inl)), "urmvo_pose_only_batch")
        return poses, inl, n_inl

    def fm_ransac(self, p0, p1, thresh=3.0, confidence=0.99, max_iters=1000):
        """cv::findFundamentalMat(p0, p1, FM_RANSAC, thresh, confidence, mask) for one frame pair."""
        p0 = _f32(p0); p1 = _f32(p1)
        mask = np.zeros(len(p0), dtype=np.uint8); st = FMStats()
        _check(self._L.urmvo_fm_ransac(self._h, C.c_int(len(p0)), _p(p0), _p(p1), C.c_double(thresh),
                                       C.c_double(confidence), C.c_int(max_iters), _p(mask), C.byref(st)),
               "urmvo_fm_ransac")
        return dict(found=int(st.found), mask=mask, F=np.array(st.F).reshape(3, 3), iters=int(st.iters),
                    n_inliers=int(st.n_inliers), models=int(st.n_models))

    def fm_ransac_batch(self, pairs, thresh=3.0, confidence=0.99, max_iters=1000):
        off, p0, p1 = pack_fm_batch(pairs)
        mask = np.zeros(int(off[-1]), dtype=np.uint8); st = (FMStats * len(pairs))()
        _check(self._L.urmvo_fm_ransac_batch(self._h, C.c_int(len(pairs)), _p(off), _p(p0), _p(p1), C.c_double(thresh),
                                             C.c_double(confidence), C.c_int(max_iters), _p(mask), st),
               "urmvo_fm_ransac_batch")
        return [mask[off[b]:off[b + 1]] for b in range(len(pairs))], list(st)

    def triangulate_batch(self, obs_off, obs_pose, obs_uv, poses_Rp, intr, pts=None):
        """Mapping::TriangulateMappoint for a batch of mappoints. Returns (pts[n,3], ok[n])."""
        obs_off = _i32(obs_off); obs_pose = _i32(obs_pose); obs_uv = _f64(obs_uv); poses_Rp = _f64(poses_Rp)
        n = len(obs_off) - 1
        out = np.zeros((n, 3)) if pts is None else _f64(pts).copy()
        ok = np.zeros(n, dtype=np.uint8)
        _check(self._L.urmvo_triangulate_batch(self._h, C.c_int(n), _p(obs_off), _p(obs_pose), _p(obs_uv),
                                               C.c_int(len(poses_Rp)), _p(poses_Rp), _p(_f64(intr)), _p(out), _p(ok)),
               "urmvo_triangulate_batch")
        return out, ok

    def two_view(self, tv, sets=None, score_mode=0):
        k1 = _f32(tv["keys1"]); k2 = _f32(tv["keys2"]); m = _i32(tv["matches12"]); K = _f32(tv["K"])
        sets = _i32(tv["sets"] if sets is None else sets)
        N = int((m >= 0).sum())
        T21 = np.zeros((4, 4), dtype=np.float32); P3D = np.zeros((k1.shape[0], 3), dtype=np.float32)
        tri = np.zeros(k1.shape[0], dtype=np.uint8)
        mH = np.zeros(N, dtype=np.uint8); mF = np.zeros(N, dtype=np.uint8)
        st = TVStats(); ok = C.c_int(0)
        _check(self._L.urmvo_two_view_scored(self._h, C.c_int(k1.shape[0]), _p(k1), C.c_int(k2.shape[0]), _p(k2), _p(m),
                                             _p(K), C.c_float(tv.get("sigma", 1.0)), C.c_int(sets.shape[0]), _p(sets),
                                             C.c_int(score_mode), _p(T21), _p(P3D), _p(tri), _p(mH), _p(mF), C.byref(st),
                                             C.byref(ok)), "urmvo_two_view_scored")
        return dict(ok=bool(ok.value), T21=T21, P3D=P3D, triangulated=tri, mask_H=mH, mask_F=mF, stats=st)


def pack_fm_batch(pairs):
    """[(p0, p1), ...] -> (off int32[B+1], pts0 float32[T,2], pts1 float32[T,2])."""
    off = np.zeros(len(pairs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(a) for a, _ in pairs])
    pts0 = np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float32).reshape(-1, 2) for a, _ in pairs]))
    pts1 = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.float32).reshape(-1, 2) for _, b in pairs]))
    return off, pts0, pts1


class FMPlan:
    """Device-resident batch of per-frame fundamental-matrix RANSAC problems (urmvo_fm_plan_*)."""

    def __init__(self, ctx, pairs, thresh=3.0, confidence=0.99, max_iters=1000):
        self._L = ctx._L
        self.ctx = ctx
        self.off, p0, p1 = pack_fm_batch(pairs)
        self.B = len(pairs)
        self._h = C.c_void_p()
        _check(self._L.urmvo_fm_plan_create(ctx._h, C.byref(self._h), C.c_int(self.B), _p(self.off), _p(p0), _p(p1),
                                            C.c_double(thresh), C.c_double(confidence), C.c_int(max_iters)),
               "urmvo_fm_plan_create")
        self.hypotheses = 0  # RANSAC iterations evaluated by the last run()

    def run(self):
        """All RANSAC rounds of every problem (kernels + the host replay of OpenCV's budget logic)."""
        _check(self._L.urmvo_fm_plan_run(self._h), "urmvo_fm_plan_run")
        self.hypotheses = int(self._L.urmvo_fm_plan_hypotheses(self._h))

    def finish(self):
        mask = np.zeros(int(self.off[-1]), dtype=np.uint8)
        st = (FMStats * self.B)()
        _check(self._L.urmvo_fm_plan_finish(self._h, _p(mask), st), "urmvo_fm_plan_finish")
        return [mask[self.off[b]:self.off[b + 1]] for b in range(self.B)], list(st)

    def close(self):
        if self._h:
            self._L.urmvo_fm_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id():
    uid = np.zeros(128, dtype=np.uint8)
    _check(load_library().urmvo_nccl_unique_id(_p(uid)), "urmvo_nccl_unique_id")
    return uid


def ba_covisibility(prob):
    """Upper-triangular co-visibility of the free cameras through the observations of `prob`."""
    fixed = _u8(prob["fixed"]); cam = _i32(prob["obs_cam"]); pt = _i32(prob["obs_pt"])
    ncf = int((fixed == 0).sum())
    upper = np.zeros((ncf, ncf), dtype=np.uint8)
    rc = load_library().urmvo_ba_covisibility(C.c_int(fixed.shape[0]), _p(fixed), C.c_int(prob["pts"].shape[0]),
                                              C.c_int(cam.shape[0]), _p(cam), _p(pt), _p(upper))
    if rc < 0:
        _check(rc, "urmvo_ba_covisibility")
    return upper


def shard_points(prob, rank, world):
    """This rank's share of a BA problem sharded by point: a contiguous range of points balanced by
    observation count (observations are point-major sorted), all cameras replicated."""
    obs_pt = np.asarray(prob["obs_pt"])
    Np, No = prob["pts"].shape[0], obs_pt.shape[0]
    start = np.searchsorted(obs_pt, np.arange(Np + 1), side="left")  # CSR over points
    cuts = [int(np.searchsorted(start, No * r / world, side="left")) for r in range(world)] + [Np]
    cuts[0] = 0
    p0, p1 = cuts[rank], max(cuts[rank + 1], cuts[rank])
    o0, o1 = int(start[p0]), int(start[p1])
    loc = dict(prob)
    loc.update(pts=np.ascontiguousarray(prob["pts"][p0:p1]), uv=np.ascontiguousarray(prob["uv"][o0:o1]),
               obs_cam=np.ascontiguousarray(prob["obs_cam"][o0:o1]),
               obs_pt=np.ascontiguousarray(obs_pt[o0:o1] - p0, dtype=np.int32), point_range=(p0, p1), obs_range=(o0, o1))
    return loc


def _phase_info(L, h):
    """Phase times (ms) and counters of the last run of a large / sharded plan (tile mode)."""
    ms = (C.c_float * 4)()
    info = (C.c_int32 * 5)()
    _check(L.urmvo_ba_plan_phase_info(h, ms, info), "urmvo_ba_plan_phase_info")
    solver = "none"
    if info[0] == 1:
        solver = "sequential block-banded Cholesky"
    elif info[0] & 2:
        solver = f"block cyclic reduction ({info[0] >> 12} super-blocks, {(info[0] >> 4) & 255} levels)"
    return {"tile_mode": bool(info[0]), "band_solver": solver, "half_bandwidth_blocks": int(info[1]), "trials_enqueued": int(info[2]),
            "host_syncs": int(info[3]), "allreduce_doubles_per_trial": int(info[4]),
            "phase_ms_first_trial_of_each_batch": {"lin": float(ms[0]), "allreduce": float(ms[1]), "solve": float(ms[2]),
                                                   "backsub_decide": float(ms[3])}}


class ShardedBAPlan:
    """One large BA sharded by point over the ranks of ctx's NCCL communicator (or alone)."""

    def __init__(self, ctx, local_prob, covis=None, chi2_thr=10.0, it0=10, it1=5, opts=None):
        self._L = ctx._L
        self.ctx = ctx
        self._keep = [_f64(local_prob["poses"]), _u8(local_prob["fixed"]), _f64(local_prob["pts"]), _f64(local_prob["uv"]),
                      _i32(local_prob["obs_cam"]), _i32(local_prob["obs_pt"]), _f64(local_prob["intr"])]
        k = self._keep
        cv = None if covis is None else _u8(covis)
        self.shapes = (k[0].shape, k[2].shape, k[3].shape[0])
        self._h = C.c_void_p()
        _check(self._L.urmvo_sharded_ba_create(ctx._h, C.byref(self._h), C.c_int(k[0].shape[0]), _p(k[0]), _p(k[1]),
                                               C.c_int(k[2].shape[0]), _p(k[2]), C.c_int(k[3].shape[0]), _p(k[3]),
                                               _p(k[4]), _p(k[5]), _p(k[6]), C.c_double(chi2_thr), C.c_int(it0), C.c_int(it1),
                                               _p(cv), C.byref(opts) if opts is not None else None), "urmvo_sharded_ba_create")

    def run(self):
        _check(self._L.urmvo_sharded_ba_run(self._h), "urmvo_sharded_ba_run")

    def phase_info(self):
        return _phase_info(self._L, self._h)

    def download(self):
        poses = np.zeros(self.shapes[0]); pts = np.zeros(self.shapes[1]); inl = np.zeros(self.shapes[2], dtype=np.uint8)
        st = (BAStats * 1)()
        _check(self._L.urmvo_ba_plan_download(self._h, _p(poses), _p(pts), _p(inl), st), "urmvo_ba_plan_download")
        return poses, pts, inl, st[0]

    def close(self):
        if self._h:
            self._L.urmvo_ba_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_ba_batch(probs):
    """Concatenates BA problems (dicts like synth.make_ba) into the batch layout of the C ABI."""
    cam_off = np.zeros(len(probs) + 1, dtype=np.int32)
    pt_off = np.zeros(len(probs) + 1, dtype=np.int32)
    obs_off = np.zeros(len(probs) + 1, dtype=np.int32)
    for i, p in enumerate(probs):
        cam_off[i + 1] = cam_off[i] + p["poses"].shape[0]
        pt_off[i + 1] = pt_off[i] + p["pts"].shape[0]
        obs_off[i + 1] = obs_off[i] + p["uv"].shape[0]
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate([np.asarray(p[k]) for p in probs], axis=0), dtype=dt)
    return dict(cam_off=cam_off, pt_off=pt_off, obs_off=obs_off, poses=cat("poses", np.float64),
                fixed=cat("fixed", np.uint8), pts=cat("pts", np.float64), uv=cat("uv", np.float64),
                obs_cam=cat("obs_cam", np.int32), obs_pt=cat("obs_pt", np.int32), intr=_f64(probs[0]["intr"]))


class BAPlan:
    """Device-resident batch of BA windows: upload once, run() many times, download()."""

    def __init__(self, ctx, batch, chi2_thr=10.0, it0=10, it1=5, opts=None):
        self._L = ctx._L
        self.ctx = ctx
        self.B = len(batch["cam_off"]) - 1
        self.shapes = (batch["poses"].shape, batch["pts"].shape, batch["uv"].shape[0])
        self._h = C.c_void_p()
        _check(self._L.urmvo_ba_plan_create(ctx._h, C.byref(self._h), C.c_int(self.B), _p(batch["cam_off"]),
                                            _p(batch["pt_off"]), _p(batch["obs_off"]), _p(batch["poses"]),
                                            _p(batch["fixed"]), _p(batch["pts"]), _p(batch["uv"]),
                                            _p(batch["obs_cam"]), _p(batch["obs_pt"]), _p(batch["intr"]),
                                            C.c_double(chi2_thr), C.c_int(it0), C.c_int(it1),
                                            C.byref(opts) if opts is not None else None), "urmvo_ba_plan_create")

    def run(self):
        _check(self._L.urmvo_ba_plan_run(self._h), "urmvo_ba_plan_run")

    def phase_info(self):
        return _phase_info(self._L, self._h)

    def download(self):
        poses = np.zeros(self.shapes[0]); pts = np.zeros(self.shapes[1]); inl = np.zeros(self.shapes[2], dtype=np.uint8)
        st = (BAStats * self.B)()
        _check(self._L.urmvo_ba_plan_download(self._h, _p(poses), _p(pts), _p(inl), st), "urmvo_ba_plan_download")
        return poses, pts, inl, list(st)

    def close(self):
        if self._h:
            self._L.urmvo_ba_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PosePlan:
    def __init__(self, ctx, batch, chi2_thr=10.0, rounds=4, its=10, inlier=None):
        self._L = ctx._L
        self.ctx = ctx
        self.B = batch["poses"].shape[0]
        self.No = batch["uv"].shape[0]
        self._keep = [_i32(batch["obs_offset"]), _f64(batch["poses"]), _f64(batch["uv"]), _f64(batch["Xw"]), _f64(batch["intr"])]
        inl = None if inlier is None else _u8(inlier)
        self._h = C.c_void_p()
        k = self._keep
        _check(self._L.urmvo_pose_plan_create(ctx._h, C.byref(self._h), C.c_int(self.B), _p(k[0]), _p(k[1]), _p(k[2]),
                                              _p(k[3]), _p(k[4]), C.c_double(chi2_thr), C.c_int(rounds), C.c_int(its),
                                              _p(inl)), "urmvo_pose_plan_create")

    def run(self):
        _check(self._L.urmvo_pose_plan_run(self._h), "urmvo_pose_plan_run")

    def download(self):
        poses = np.zeros((self.B, 7)); inl = np.zeros(self.No, dtype=np.uint8)
        n_inl = np.zeros(self.B, dtype=np.int32); iters = np.zeros(self.B, dtype=np.int32)
        _check(self._L.urmvo_pose_plan_download(self._h, _p(poses), _p(inl), _p(n_inl), _p(iters)), "urmvo_pose_plan_download")
        return poses, inl, n_inl, iters

    def close(self):
        if self._h:
            self._L.urmvo_pose_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TVPlan:
    def __init__(self, ctx, tv, sets=None):
        self._L = ctx._L
        self.ctx = ctx
        k1 = _f32(tv["keys1"]); k2 = _f32(tv["keys2"]); m = _i32(tv["matches12"]); K = _f32(tv["K"])
        sets = _i32(tv["sets"] if sets is None else sets)
        self.n1 = k1.shape[0]; self.N = int((m >= 0).sum()); self.n_hyp = sets.shape[0]
        self.words = (self.N + 31) // 32
        self._h = C.c_void_p()
        _check(self._L.urmvo_tv_plan_create(ctx._h, C.byref(self._h), C.c_int(k1.shape[0]), _p(k1), C.c_int(k2.shape[0]),
                                            _p(k2), _p(m), _p(K), C.c_float(tv.get("sigma", 1.0)), C.c_int(self.n_hyp),
                                            _p(sets)), "urmvo_tv_plan_create")

    def set_score_mode(self, score_mode):
        """0: the reference's symmetric point-line chi2, 1: Sampson error (fundamental hypotheses)."""
        _check(self._L.urmvo_tv_plan_set_score_mode(self._h, C.c_int(score_mode)), "urmvo_tv_plan_set_score_mode")

    def run_ransac(self):
        _check(self._L.urmvo_tv_plan_run_ransac(self._h), "urmvo_tv_plan_run_ransac")

    def download_hyps(self, model):
        scores = np.zeros(self.n_hyp, dtype=np.float32)
        masks = np.zeros((self.n_hyp, self.words), dtype=np.uint32)
        models = np.zeros((self.n_hyp, 9), dtype=np.float32)
        _check(self._L.urmvo_tv_plan_download_hyps(self._h, C.c_int(model), _p(scores), _p(masks), _p(models)),
               "urmvo_tv_plan_download_hyps")
        return scores, masks, models

    def reconstruct(self):
        T21 = np.zeros((4, 4), dtype=np.float32); P3D = np.zeros((self.n1, 3), dtype=np.float32)
        tri = np.zeros(self.n1, dtype=np.uint8)
        mH = np.zeros(self.N, dtype=np.uint8); mF = np.zeros(self.N, dtype=np.uint8)
        st = TVStats(); ok = C.c_int(0)
        _check(self._L.urmvo_tv_plan_reconstruct(self._h, _p(T21), _p(P3D), _p(tri), _p(mH), _p(mF), C.byref(st),
                                                 C.byref(ok)), "urmvo_tv_plan_reconstruct")
        return dict(ok=bool(ok.value), T21=T21, P3D=P3D, triangulated=tri, mask_H=mH, mask_F=mF, stats=st)

    def close(self):
        if self._h:
            self._L.urmvo_tv_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMap:
    """Device-resident map (urmvo_map_*): keyframe poses, mappoint positions and observations stay in HBM across
    keyframes; windows are selected by id lists."""

    def __init__(self, ctx, intr):
        self._L = ctx._L
        self.ctx = ctx
        self._h = C.c_void_p()
        _check(self._L.urmvo_map_create(ctx._h, C.byref(self._h), _p(_f64(intr))), "urmvo_map_create")

    def set_keyframes(self, ids, poses):
        ids = _i32(ids); poses = _f64(poses)
        _check(self._L.urmvo_map_set_keyframes(self._h, C.c_int(len(ids)), _p(ids), _p(poses)), "urmvo_map_set_keyframes")

    def set_points(self, ids, xyz):
        ids = _i32(ids); xyz = _f64(xyz)
        _check(self._L.urmvo_map_set_points(self._h, C.c_int(len(ids)), _p(ids), _p(xyz)), "urmvo_map_set_points")

    def add_observations(self, kf_ids, pt_ids, uv):
        kf_ids = _i32(kf_ids); pt_ids = _i32(pt_ids); uv = _f64(uv)
        _check(self._L.urmvo_map_add_observations(self._h, C.c_int(len(kf_ids)), _p(kf_ids), _p(pt_ids), _p(uv)),
               "urmvo_map_add_observations")

    def remove_observations(self, kf_ids, pt_ids):
        kf_ids = _i32(kf_ids); pt_ids = _i32(pt_ids)
        _check(self._L.urmvo_map_remove_observations(self._h, C.c_int(len(kf_ids)), _p(kf_ids), _p(pt_ids)),
               "urmvo_map_remove_observations")

    def get_keyframes(self, ids):
        ids = _i32(ids); out = np.zeros((len(ids), 7))
        _check(self._L.urmvo_map_get_keyframes(self._h, C.c_int(len(ids)), _p(ids), _p(out)), "urmvo_map_get_keyframes")
        return out

    def get_points(self, ids):
        ids = _i32(ids); out = np.zeros((len(ids), 3))
        _check(self._L.urmvo_map_get_points(self._h, C.c_int(len(ids)), _p(ids), _p(out)), "urmvo_map_get_points")
        return out

    def local_ba(self, kf_ids, kf_fixed, pt_ids, max_obs, chi2_thr=10.0, it0=10, it1=5, opts=None):
        """Returns (obs_kf, obs_pt, inlier, stats) of the observations used; the map holds the optimised values."""
        kf_ids = _i32(kf_ids); kf_fixed = _u8(kf_fixed); pt_ids = _i32(pt_ids)
        n = C.c_int32(0)
        okf = np.zeros(max_obs, dtype=np.int32); opt = np.zeros(max_obs, dtype=np.int32); inl = np.zeros(max_obs, dtype=np.uint8)
        st = BAStats()
        _check(self._L.urmvo_map_local_ba(self._h, C.c_int(len(kf_ids)), _p(kf_ids), _p(kf_fixed), C.c_int(len(pt_ids)), _p(pt_ids),
                                          C.c_double(chi2_thr), C.c_int(it0), C.c_int(it1),
                                          C.byref(opts) if opts is not None else None, C.c_int32(max_obs), C.byref(n),
                                          _p(okf), _p(opt), _p(inl), C.byref(st)), "urmvo_map_local_ba")
        k = n.value
        return okf[:k], opt[:k], inl[:k], st

    def close(self):
        if self._h:
            self._L.urmvo_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
