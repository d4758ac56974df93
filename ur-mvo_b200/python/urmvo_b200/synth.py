"""Deterministic Aqualoc-shaped synthetic inputs for the five BASELINE.json configs.

Camera: 640x512 pinhole from the reference's configs/camera_settings/aqua.yaml:37.
Keypoints are integer pixels inside a 4-px border (SuperPoint emits integer heat-map coordinates,
reference src/super_point.cpp:196-226).  Poses are T_wc as (qx,qy,qz,qw,px,py,pz) like the
reference's Pose3d (include/types.h:18-31); observations are ordered by point then frame id like
Mapping::LocalMapOptimization emits them (src/mapping.cc:406-469).
"""
import ctypes
import numpy as np

FX, FY, CX, CY = 413.32595366566017, 413.70198739483686, 305.9507483284928, 259.4439948946375
INTR = np.array([FX, FY, CX, CY], dtype=np.float64)
K33 = np.array([[FX, 0, CX], [0, FY, CY], [0, 0, 1]], dtype=np.float32)
W_IMG, H_IMG, BORDER = 640, 512, 4
MONO_POINT = 10.0  # configs/configs_aqua.yaml:40-48


def quat_to_R(q):
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotvec_to_quat(rv):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv, axis=-1, keepdims=True)
    small = th < 1e-12
    k = np.where(small, 0.5, np.sin(th / 2) / np.where(small, 1.0, th))
    return np.concatenate([rv * k, np.cos(th / 2)], axis=-1)


def quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def project(poses_wc, cam_idx, X):
    """Project world points X[i] into camera poses_wc[cam_idx[i]] (T_wc). Returns uv, depth."""
    R = quat_to_R(poses_wc[cam_idx, :4])
    p = poses_wc[cam_idx, 4:]
    pc = np.einsum('nji,nj->ni', R, X - p)  # R^T (X - p)
    z = pc[:, 2]
    uv = np.stack([pc[:, 0] / z * FX + CX, pc[:, 1] / z * FY + CY], axis=-1)
    return uv, z


def _trajectory(rng, n):
    """Smooth forward-lateral trajectory, 0.15 per keyframe, yaw <= 3 deg per keyframe."""
    yaw_step = np.deg2rad(rng.uniform(-3.0, 3.0, size=n)) * 0.5
    yaw = np.cumsum(yaw_step)
    # keep the heading bounded so that long trajectories stay roughly straight
    yaw = yaw - np.linspace(0, yaw[-1], n) if n > 1 else yaw
    pos = np.zeros((n, 3))
    step = np.stack([0.15 * np.cos(yaw) * 0.8 + 0.0, 0.01 * rng.standard_normal(n), 0.15 * np.sin(yaw) * 0.3 + 0.05], axis=-1)
    pos[1:] = np.cumsum(step[1:], axis=0)
    q = rotvec_to_quat(np.stack([np.zeros(n), yaw, np.zeros(n)], axis=-1))
    return np.concatenate([q, pos], axis=-1)


def make_ba(seed, n_cams, n_pts, obs_per_pt, span, n_fixed, outlier_frac,
            rot_sigma_deg=0.5, trans_sigma=0.02, pt_sigma=0.05, px_sigma=0.5):
    """Local-BA problem. Returns dict of arrays (ground truth + perturbed initial estimate)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gt_poses = _trajectory(rng, n_cams)
    span = min(span, n_cams)
    c0 = rng.integers(0, n_cams, size=n_pts)
    # sample points in the frustum of their centre camera, depth 2..8
    u = rng.uniform(40, W_IMG - 40, size=n_pts)
    v = rng.uniform(40, H_IMG - 40, size=n_pts)
    d = rng.uniform(2.0, 8.0, size=n_pts)
    pc = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], axis=-1)
    Rc = quat_to_R(gt_poses[c0, :4])
    gt_pts = np.einsum('nij,nj->ni', Rc, pc) + gt_poses[c0, 4:]
    # candidate observers: a window of `span` cameras around c0
    lo = np.clip(c0 - span // 2, 0, n_cams - span)
    cand_cam = (lo[:, None] + np.arange(span)[None, :]).reshape(-1)
    cand_pt = np.repeat(np.arange(n_pts), span)
    keep = rng.uniform(size=cand_cam.size) < min(1.0, obs_per_pt / span)
    keep |= (cand_cam == np.repeat(c0, span))  # the centre camera always sees the point
    cand_cam, cand_pt = cand_cam[keep], cand_pt[keep]
    uv, z = project(gt_poses, cand_cam, gt_pts[cand_pt])
    uv = uv + px_sigma * rng.standard_normal(uv.shape)
    is_out = rng.uniform(size=uv.shape[0]) < outlier_frac
    uv_out = np.stack([rng.uniform(BORDER, W_IMG - BORDER, size=uv.shape[0]),
                       rng.uniform(BORDER, H_IMG - BORDER, size=uv.shape[0])], axis=-1)
    uv = np.where(is_out[:, None], uv_out, uv)
    uv = np.rint(uv)
    vis = (z > 0.5) & (uv[:, 0] >= BORDER) & (uv[:, 0] < W_IMG - BORDER) & \
          (uv[:, 1] >= BORDER) & (uv[:, 1] < H_IMG - BORDER)
    cand_cam, cand_pt, uv, is_out = cand_cam[vis], cand_pt[vis], uv[vis], is_out[vis]
    # src/mapping.cc:459-460: points with <= 1 mono observation are dropped
    cnt = np.bincount(cand_pt, minlength=n_pts)
    ok_pt = cnt > 1
    remap = np.cumsum(ok_pt) - 1
    sel = ok_pt[cand_pt]
    obs_cam = cand_cam[sel].astype(np.int32)
    obs_pt = remap[cand_pt[sel]].astype(np.int32)
    uv = np.ascontiguousarray(uv[sel])
    is_out = is_out[sel]
    gt_pts = gt_pts[ok_pt]
    order = np.lexsort((obs_cam, obs_pt))  # by point, then frame id
    obs_cam, obs_pt, uv, is_out = obs_cam[order], obs_pt[order], uv[order], is_out[order]
    fixed = np.zeros(n_cams, dtype=np.uint8)
    fixed[:n_fixed] = 1
    # initial estimate = GT (+) noise on the free cameras and on all points
    dq = rotvec_to_quat(np.deg2rad(rot_sigma_deg) * rng.standard_normal((n_cams, 3)))
    dp = trans_sigma * rng.standard_normal((n_cams, 3))
    free = (fixed == 0)[:, None]
    poses = gt_poses.copy()
    poses[:, :4] = np.where(free, quat_mul(dq, gt_poses[:, :4]), gt_poses[:, :4])
    poses[:, 4:] = np.where(free, gt_poses[:, 4:] + dp, gt_poses[:, 4:])
    pts = gt_pts + pt_sigma * rng.standard_normal(gt_pts.shape)
    return dict(poses=np.ascontiguousarray(poses), fixed=fixed, pts=np.ascontiguousarray(pts),
                uv=uv, obs_cam=obs_cam, obs_pt=obs_pt, intr=INTR.copy(), is_outlier=is_out,
                gt_poses=gt_poses, gt_pts=gt_pts)


BF = 0.12 * FX  # stereo baseline 12 cm (a stereo rig is not part of the Aqualoc configuration: test value)
STEREO_POINT = 75.0  # the reference's thHuberStereo scale: cfg.stereo_point is not set in configs_aqua.yaml


def add_stereo(prob, seed, stereo_frac=0.6, px_sigma=0.5):
    """Turns a mono BA problem into a STEREO-camera one (reference src/g2o_optimization.cc:96-118): a
    fraction of the observations gets a right-image column u_right = u - bf / z (+ noise, rounded like a
    keypoint) and becomes a stereo edge; the others stay mono edges in the same graph."""
    rng = np.random.Generator(np.random.PCG64(seed))
    uv, z = project(prob["gt_poses"], prob["obs_cam"], prob["gt_pts"][prob["obs_pt"]])
    ur = np.rint(uv[:, 0] - BF / z + px_sigma * rng.standard_normal(z.shape))
    kind = (rng.uniform(size=z.shape) < stereo_frac) & (ur >= BORDER) & (~prob["is_outlier"])
    out = dict(prob)
    out["uv3"] = np.ascontiguousarray(np.c_[prob["uv"], np.where(kind, ur, 0.0)])
    out["kind"] = kind.astype(np.uint8)
    out["intr5"] = np.r_[INTR, BF]
    return out


def make_pose_batch_stereo(seed, B=8, n_obs=300, stereo_frac=0.6, outlier_frac=0.1):
    """Pose-only frames of a STEREO camera: uv3 / kind / intr5 beside the mono arrays of make_pose_batch."""
    b = make_pose_batch(seed, B=B, n_obs=n_obs, outlier_frac=outlier_frac)
    rng = np.random.Generator(np.random.PCG64(seed + 17))
    frame = np.repeat(np.arange(B), np.diff(b["obs_offset"]))
    _, z = project(b["gt_poses"], frame, b["Xw"])
    u_gt, _ = project(b["gt_poses"], frame, b["Xw"])
    ur = np.rint(u_gt[:, 0] - BF / z + 0.5 * rng.standard_normal(z.shape))
    kind = (rng.uniform(size=z.shape) < stereo_frac) & (ur >= BORDER)
    out = dict(b)
    out["uv3"] = np.ascontiguousarray(np.c_[b["uv"], np.where(kind, ur, 0.0)])
    out["kind"] = kind.astype(np.uint8)
    out["intr5"] = np.r_[INTR, BF]
    return out


def add_camera_models(prob, seed, n_models=3):
    """Gives every observation of a stereo-capable problem (uv3 / kind / intr5) one of n_models camera models, the
    way the reference reads camera_list[mpc->id_camera] per constraint (src/g2o_optimization.cc:86-89): model 0 is
    the base camera, the others have focal lengths within 15 %, shifted principal points and their own baseline;
    the measurements are re-expressed in their model (sub-pixel, not re-rounded).  Adds intr5_tab (n_models, 5) and
    kind_model = stereo bit | model << 1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = np.asarray(prob["intr5"], dtype=np.float64)
    tab = np.tile(base, (n_models, 1))
    for m in range(1, n_models):
        s = 1.0 + 0.15 * rng.uniform(-1, 1)
        tab[m] = [base[0] * s, base[1] * s * (1.0 + 0.01 * rng.uniform(-1, 1)), base[2] + rng.uniform(-12, 12),
                  base[3] + rng.uniform(-12, 12), base[4] * rng.uniform(0.7, 1.4)]
    uv3 = np.asarray(prob["uv3"], dtype=np.float64)
    kind = np.asarray(prob["kind"], dtype=np.uint8)
    model = rng.integers(0, n_models, size=uv3.shape[0])
    K = tab[model]
    xn = (uv3[:, 0] - base[2]) / base[0]
    yn = (uv3[:, 1] - base[3]) / base[1]
    disp = (uv3[:, 0] - uv3[:, 2]) / base[4]  # 1 / z as the base camera measured it
    u = K[:, 0] * xn + K[:, 2]
    v = K[:, 1] * yn + K[:, 3]
    ur = np.where(kind != 0, u - K[:, 4] * disp, 0.0)
    out = dict(prob)
    out["uv3"] = np.ascontiguousarray(np.c_[u, v, ur])
    out["intr5_tab"] = tab
    out["kind_model"] = (kind | (model << 1)).astype(np.uint8)
    return out


def cfg1(seed=1001):
    """10 keyframes (ids 0,1,2 fixed by src/mapping.cc:355-356), 2000 points, ~15k observations."""
    return make_ba(seed, 10, 2000, 7.7, 10, 3, 0.05)


def cfg4(seed=1004):
    """50 keyframes (first 2 fixed), 50k points, ~400k observations."""
    return make_ba(seed, 50, 50000, 8.2, 14, 2, 0.02)


def cfg5(seed=1005, n_cams=1000, n_pts=200000):
    """1000 cameras (first 2 fixed), 200k points, ~2M observations, banded co-visibility."""
    return make_ba(seed, n_cams, n_pts, 10.3, 16, 2, 0.01)


def sort_points_by_first_camera(prob):
    """Renumbers the points of a BA problem by the first camera that observes them (the order a SLAM map creates
    its mappoints in: the generator above draws them at random along the trajectory).  The problem is the same; a
    contiguous range of point ids is then local to a stretch of the trajectory, which is what a point-sharded solve
    wants.  Observations stay point-major."""
    obs_pt, obs_cam = prob["obs_pt"], prob["obs_cam"]
    n_pts = prob["pts"].shape[0]
    first = np.full(n_pts, np.iinfo(np.int32).max, dtype=np.int64)
    np.minimum.at(first, obs_pt, obs_cam)
    order = np.argsort(first, kind="stable")          # new index -> old index
    inv = np.empty(n_pts, dtype=np.int64); inv[order] = np.arange(n_pts)
    new_pt = inv[obs_pt]
    o = np.lexsort((obs_cam, new_pt))
    out = dict(prob)
    out["pts"] = np.ascontiguousarray(prob["pts"][order])
    if "gt_pts" in prob:
        out["gt_pts"] = np.ascontiguousarray(prob["gt_pts"][order])
    out["obs_pt"] = np.ascontiguousarray(new_pt[o].astype(np.int32))
    out["obs_cam"] = np.ascontiguousarray(obs_cam[o])
    out["uv"] = np.ascontiguousarray(prob["uv"][o])
    if "is_outlier" in prob:
        out["is_outlier"] = prob["is_outlier"][o]
    return out


def small_ba(seed=7, n_cams=6, n_pts=120, n_fixed=2, outlier_frac=0.05, **kw):
    return make_ba(seed, n_cams, n_pts, 4.5, n_cams, n_fixed, outlier_frac, **kw)


def make_pose_batch(seed=1002, B=256, n_obs=1000, outlier_frac=0.10, rot_deg=2.0, trans=0.1, px_sigma=0.5):
    """cfg2: B independent pose-only problems with n_obs 3D-2D matches each."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yaw = np.deg2rad(rng.uniform(-20, 20, size=B))
    q = rotvec_to_quat(np.stack([np.deg2rad(rng.uniform(-5, 5, size=B)), yaw, np.zeros(B)], axis=-1))
    p = rng.uniform(-1, 1, size=(B, 3))
    gt = np.concatenate([q, p], axis=-1)
    No = B * n_obs
    fr = np.repeat(np.arange(B), n_obs)
    u = rng.uniform(BORDER + 2, W_IMG - BORDER - 2, size=No)
    v = rng.uniform(BORDER + 2, H_IMG - BORDER - 2, size=No)
    d = rng.uniform(2.0, 8.0, size=No)
    pc = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], axis=-1)
    Xw = np.einsum('nij,nj->ni', quat_to_R(gt[fr, :4]), pc) + gt[fr, 4:]
    uv = np.stack([u, v], axis=-1) + px_sigma * rng.standard_normal((No, 2))
    is_out = rng.uniform(size=No) < outlier_frac
    uv_out = np.stack([rng.uniform(BORDER, W_IMG - BORDER, size=No), rng.uniform(BORDER, H_IMG - BORDER, size=No)], axis=-1)
    uv = np.rint(np.where(is_out[:, None], uv_out, uv))
    uv[:, 0] = np.clip(uv[:, 0], BORDER, W_IMG - BORDER - 1)
    uv[:, 1] = np.clip(uv[:, 1], BORDER, H_IMG - BORDER - 1)
    dq = rotvec_to_quat(np.deg2rad(rot_deg) * rng.standard_normal((B, 3)) / np.sqrt(3))
    poses = gt.copy()
    poses[:, :4] = quat_mul(dq, gt[:, :4])
    poses[:, 4:] += trans * rng.standard_normal((B, 3)) / np.sqrt(3)
    obs_offset = (np.arange(B + 1) * n_obs).astype(np.int32)
    return dict(poses=np.ascontiguousarray(poses), obs_offset=obs_offset, uv=np.ascontiguousarray(uv),
                Xw=np.ascontiguousarray(Xw), intr=INTR.copy(), gt_poses=gt, is_outlier=is_out)


def cfg2(seed=1002):
    return make_pose_batch(seed, 256, 1000)


def draw_sets(N, n_hyp, seed=0):
    """8-point index sets exactly as EpipolarGeometry::reconstruct draws them
    (src/epipolar_geometry.cc:53-71,100-112): glibc srand(seed), RandomInt, swap-with-back."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    sets = np.empty((n_hyp, 8), dtype=np.int32)
    base = list(range(N))
    for it in range(n_hyp):
        avail = base.copy()
        for j in range(8):
            d = len(avail)
            r = int((libc.rand() / (2147483647 + 1.0)) * d)
            sets[it, j] = avail[r]
            avail[r] = avail[-1]
            avail.pop()
    return sets


def make_two_view(seed=1003, n_keys=1000, inlier_frac=0.70, px_sigma=0.7, planar=False,
                  rot_deg=5.0, t=(0.3, 0.02, 0.05), n_unmatched=0):
    """cfg3: two images with n_keys keypoints each and an identity match vector."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q21 = rotvec_to_quat(np.array([0.0, np.deg2rad(rot_deg), 0.0]))
    R21 = quat_to_R(q21)
    t21 = np.asarray(t, dtype=np.float64)
    u = rng.uniform(BORDER + 8, W_IMG - BORDER - 8, size=n_keys)
    v = rng.uniform(BORDER + 8, H_IMG - BORDER - 8, size=n_keys)
    if planar:
        d = 4.0 + 0.2 * ((u - CX) / FX) + 0.1 * ((v - CY) / FY)
    else:
        d = rng.uniform(2.0, 8.0, size=n_keys)
    X1 = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], axis=-1)
    X2 = X1 @ R21.T + t21
    uv2 = np.stack([X2[:, 0] / X2[:, 2] * FX + CX, X2[:, 1] / X2[:, 2] * FY + CY], axis=-1)
    k1 = np.stack([u, v], axis=-1) + px_sigma * rng.standard_normal((n_keys, 2))
    k2 = uv2 + px_sigma * rng.standard_normal((n_keys, 2))
    is_out = rng.uniform(size=n_keys) >= inlier_frac
    k2_out = np.stack([rng.uniform(BORDER, W_IMG - BORDER, size=n_keys), rng.uniform(BORDER, H_IMG - BORDER, size=n_keys)], axis=-1)
    k2 = np.where(is_out[:, None], k2_out, k2)
    k1, k2 = np.rint(k1), np.rint(k2)
    for k in (k1, k2):
        k[:, 0] = np.clip(k[:, 0], BORDER, W_IMG - BORDER - 1)
        k[:, 1] = np.clip(k[:, 1], BORDER, H_IMG - BORDER - 1)
    matches = np.arange(n_keys, dtype=np.int32)
    if n_unmatched:
        drop = rng.choice(n_keys, size=n_unmatched, replace=False)
        matches[drop] = -1
    T21 = np.eye(4)
    T21[:3, :3] = R21
    T21[:3, 3] = t21
    return dict(keys1=np.ascontiguousarray(k1, dtype=np.float32), keys2=np.ascontiguousarray(k2, dtype=np.float32),
                matches12=matches, K=K33.copy(), sigma=1.0, gt_T21=T21, is_outlier=is_out)


def cfg3(seed=1003, n_hyp=8192):
    tv = make_two_view(seed)
    N = int((tv["matches12"] >= 0).sum())
    tv["sets"] = draw_sets(N, n_hyp, 0)
    return tv


def make_fm(seed=1006, n=1000, inlier_frac=0.7, px_sigma=0.7, rot_deg=5.0, t=(0.3, 0.02, 0.05)):
    """Matched keypoints of one frame pair for the per-frame fundamental-matrix RANSAC
    (reference src/point_matching.cc:44-58): two (n, 2) float32 arrays of integer pixel positions."""
    tv = make_two_view(seed, n_keys=n, inlier_frac=inlier_frac, px_sigma=px_sigma, rot_deg=rot_deg, t=t)
    return tv["keys1"], tv["keys2"]


def make_fm_batch(seed=1006, B=256, n=1000, inlier_frac=0.7):
    """B independent frame pairs (stereo + temporal matches of consecutive keyframes): list of (p0, p1)."""
    return [make_fm(seed + 7919 * b, n, inlier_frac) for b in range(B)]


def make_triangulation(seed=1007, n_pts=500, n_poses=12, max_obs=10, px_sigma=0.5, degenerate_frac=0.05):
    """New mappoints of a keyframe for Mapping::TriangulateMappoint (reference src/mapping.cc:151-205):
    poses_Rp (n_poses, 12) = keyframe poses T_wc as [R row-major | p], CSR observer lists, keypoints.
    A few mappoints are degenerate on purpose (one observer, or all observers at the same centre)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gt = _trajectory(rng, n_poses)
    Rs = quat_to_R(gt[:, :4])
    poses_Rp = np.concatenate([Rs.reshape(n_poses, 9), gt[:, 4:]], axis=1)
    X = np.stack([rng.uniform(-2, 2, n_pts), rng.uniform(-1.5, 1.5, n_pts), rng.uniform(3, 9, n_pts)], axis=-1)
    X = X @ Rs[0].T + gt[0, 4:]
    off, idx, uv = [0], [], []
    kind = rng.uniform(size=n_pts)
    for l in range(n_pts):
        k = int(rng.integers(2, max_obs + 1))
        cams = np.sort(rng.choice(n_poses, size=min(k, n_poses), replace=False))
        if kind[l] < degenerate_frac / 2:
            cams = cams[:1]                      # fewer than 2 observers
        for c in cams:
            xc = Rs[c].T @ (X[l] - gt[c, 4:])
            p = np.array([xc[0] / xc[2] * FX + CX, xc[1] / xc[2] * FY + CY]) + px_sigma * rng.standard_normal(2)
            if degenerate_frac / 2 <= kind[l] < degenerate_frac:
                c = cams[0]                      # the same keyframe, same pixel: rank-deficient system
                xc = Rs[c].T @ (X[l] - gt[c, 4:])
                p = np.array([xc[0] / xc[2] * FX + CX, xc[1] / xc[2] * FY + CY])
            idx.append(int(c)); uv.append(p)
        off.append(len(idx))
    return dict(obs_off=np.array(off, dtype=np.int32), obs_pose=np.array(idx, dtype=np.int32),
                obs_uv=np.array(uv, dtype=np.float64).reshape(-1, 2), poses_Rp=poses_Rp,
                intr=np.array([FX, FY, CX, CY]), gt=X)


def make_pnp(seed=1008, n=300, outlier_frac=0.2, px_sigma=0.7, rot_deg=10.0, trans=0.5):
    """2D-3D matches of one tracked frame for SolvePnPWithCV (reference src/g2o_optimization.cc:323-377):
    obj (n, 3) float32 mappoint positions, img (n, 2) float32 keypoints, intr = fx fy cx cy, and the true T_cw
    (R, t).  Outliers are keypoints at random pixel positions."""
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.uniform(BORDER + 2, W_IMG - BORDER - 2, size=n)
    v = rng.uniform(BORDER + 2, H_IMG - BORDER - 2, size=n)
    d = rng.uniform(2.0, 8.0, size=n)
    pc = np.stack([(u - CX) / FX * d, (v - CY) / FY * d, d], axis=-1)
    rv = np.deg2rad(rot_deg) * rng.standard_normal(3) / np.sqrt(3.0)
    R = quat_to_R(rotvec_to_quat(rv[None, :]))[0]  # T_cw rotation
    t = trans * rng.standard_normal(3)
    X = (pc - t) @ R  # X_w = R^T (X_c - t)
    uv = np.stack([u, v], axis=-1) + px_sigma * rng.standard_normal((n, 2))
    is_out = rng.uniform(size=n) < outlier_frac
    uv[is_out] = np.stack([rng.uniform(BORDER, W_IMG - BORDER, size=int(is_out.sum())),
                           rng.uniform(BORDER, H_IMG - BORDER, size=int(is_out.sum()))], axis=-1)
    return dict(obj=X.astype(np.float32), img=uv.astype(np.float32), intr=np.array([FX, FY, CX, CY]), R=R, t=t,
                is_outlier=is_out)
