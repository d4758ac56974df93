"""GPU (-m gpu): the device-resident map (urmvo_map_*).  A sliding window of local bundle
adjustments over a trajectory, run (a) through the map — values stay in HBM, every keyframe uploads only what is new,
windows are selected by id lists — and (b) as one-shot urmvo_local_ba calls on host arrays that the test keeps in sync
by hand (what the reference does with its std::map containers every keyframe, src/mapping.cc:335-535).  Both must give
the same poses, points and inlier flags, window after window."""
import numpy as np
import pytest

import urmvo_b200 as U
from urmvo_b200 import synth

pytestmark = pytest.mark.gpu


def _window(prob, kfs, fixed_ids, dead):
    """Host-side construction of the window (the reference's way): observations of the listed keyframes, point-major,
    points with fewer than two of them left out.  Returns a one-shot problem + the (kf, pt) list of its observations."""
    local = {k: i for i, k in enumerate(kfs)}
    by_pt = {}
    for o in range(len(prob["obs_cam"])):
        c, p = int(prob["obs_cam"][o]), int(prob["obs_pt"][o])
        if c in local and (c, p) not in dead:
            by_pt.setdefault(p, []).append(o)
    pts_used, cam, pt, uv, pairs = [], [], [], [], []
    for p in sorted(by_pt):
        if len(by_pt[p]) < 2:
            continue
        for o in by_pt[p]:
            cam.append(local[int(prob["obs_cam"][o])]); pt.append(len(pts_used)); uv.append(prob["uv"][o])
            pairs.append((int(prob["obs_cam"][o]), p))
        pts_used.append(p)
    return dict(kfs=kfs, pts_used=pts_used, obs_cam=np.array(cam, dtype=np.int32), obs_pt=np.array(pt, dtype=np.int32),
                uv=np.array(uv), fixed=np.array([1 if k in fixed_ids else 0 for k in kfs], dtype=np.uint8), pairs=pairs)


def test_sliding_windows_through_the_map_equal_one_shot_calls(ctx):
    prob = synth.make_ba(61, 16, 2400, 6.0, 9, 2, 0.04)
    n_cams = prob["poses"].shape[0]
    poses = prob["poses"].copy(); pts = prob["pts"].copy()   # the host copy the one-shot path keeps in sync
    m = U.DeviceMap(ctx, prob["intr"])
    # ids as the reference has them: sparse frame ids, mappoint ids
    kf_id = lambda c: 10 + 3 * c
    pt_id = lambda p: 1000 + 7 * p
    m.set_points([pt_id(p) for p in range(len(pts))], pts)
    dead = set()
    uploaded = 0
    for step, last in enumerate(range(9, n_cams)):
        # a new keyframe arrives: upload ITS pose and ITS observations only
        new_kfs = range(uploaded, last + 1)
        m.set_keyframes([kf_id(c) for c in new_kfs], poses[list(new_kfs)])
        sel = np.flatnonzero((prob["obs_cam"] >= uploaded) & (prob["obs_cam"] <= last))
        m.add_observations([kf_id(int(c)) for c in prob["obs_cam"][sel]], [pt_id(int(p)) for p in prob["obs_pt"][sel]], prob["uv"][sel])
        uploaded = last + 1
        kfs = list(range(last - 9, last + 1))
        fixed_ids = set(kfs[:3])
        w = _window(prob, kfs, fixed_ids, dead)
        # (b) one-shot call on host arrays
        one = dict(poses=poses[kfs], fixed=w["fixed"], pts=pts[w["pts_used"]], uv=w["uv"], obs_cam=w["obs_cam"], obs_pt=w["obs_pt"],
                   intr=prob["intr"])
        gp, gx, gi, gs = ctx.local_ba(one)
        # (a) the same window through the map: every mappoint seen so far is offered, the map drops the under-observed ones
        okf, opt, inl, st = m.local_ba([kf_id(c) for c in kfs], w["fixed"], [pt_id(p) for p in range(len(pts))], max_obs=len(prob["uv"]))
        assert len(okf) == len(w["pairs"])
        assert [(int(a), int(b)) for a, b in zip(okf, opt)] == [(kf_id(c), pt_id(p)) for c, p in w["pairs"]]
        assert np.array_equal(inl, gi), f"window {step}: {(inl != gi).sum()} inlier flags differ"
        assert list(st.iters) == list(gs.iters) and list(st.trials) == list(gs.trials)
        assert abs(st.chi2_final[1] - gs.chi2_final[1]) <= 1e-12 * abs(gs.chi2_final[1])
        mp = m.get_keyframes([kf_id(c) for c in kfs])
        mx = m.get_points([pt_id(p) for p in w["pts_used"]])
        free = w["fixed"] == 0
        assert np.abs(mp[free] - gp[free]).max() < 1e-12 and np.abs(mx - gx).max() < 1e-12
        assert np.array_equal(mp[~free], poses[kfs][~free])  # fixed keyframes are not written
        # the host copy follows (what mapping.cc does after LocalmapOptimization) and the outliers are erased
        poses[np.array(kfs)[free]] = gp[free]
        pts[w["pts_used"]] = gx
        out = [w["pairs"][o] for o in np.flatnonzero(gi == 0)]
        dead.update(out)
        m.remove_observations([kf_id(c) for c, _ in out], [pt_id(p) for _, p in out])
    m.close()


def test_map_growth_overwrite_and_errors(ctx):
    intr = np.array([400.0, 400.0, 320.0, 256.0])
    m = U.DeviceMap(ctx, intr)
    rng = np.random.default_rng(5)
    ids = np.arange(5000, dtype=np.int32) * 2 + 1   # more than the initial capacity of the slot arrays
    xyz = rng.normal(size=(5000, 3))
    m.set_points(ids[:3000], xyz[:3000])
    m.set_points(ids[3000:], xyz[3000:])             # grows, keeps the old contents
    assert np.array_equal(m.get_points(ids), xyz)
    m.set_points(ids[10:20], xyz[10:20] + 1.0)        # overwrite
    assert np.array_equal(m.get_points(ids[10:20]), xyz[10:20] + 1.0)
    with pytest.raises(U.UrmvoError):
        m.get_points([4])                             # unknown id
    with pytest.raises(U.UrmvoError):
        m.add_observations([7], [1], np.zeros((1, 2)))  # unknown keyframe
    pose = np.array([[0, 0, 0, 1.0, 0, 0, 0]])
    m.set_keyframes([7], pose)
    okf, opt, inl, st = m.local_ba([7], [0], ids[:10], max_obs=10)
    assert len(okf) == 0                              # no observations: empty graph, silent no-op
    m.close()
