"""The C++ drop-in adapters (ur-mvo_b200/adapter/*.cc) keep the reference's signatures
(include/g2o_optimization.h:13-19, include/epipolar_geometry.h:20-48).  CPU: they compile and link
against the C ABI using stand-in headers (tests/shim — Eigen / OpenCV / g2o are not installed).
GPU: driven through LocalmapOptimization / FrameOptimization / EpipolarGeometry::reconstruct they
give the oracle's results, including the id -> dense index mapping and the in-place conventions."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from urmvo_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "shim", "adapter_driver")


def build_driver():
    import urmvo_b200
    urmvo_b200.load_library()
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "shim"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "shim", "adapter_driver.cc"), os.path.join(ROOT, "ur-mvo_b200", "adapter", "g2o_optimization.cc"),
           os.path.join(ROOT, "ur-mvo_b200", "adapter", "epipolar_geometry.cc"),
           os.path.join(ROOT, "ur-mvo_b200", "adapter", "point_matching_outliers.cc"), "-I", os.path.join(ROOT, "ur-mvo_b200", "adapter"), "-L", os.path.join(ROOT, "ur-mvo_b200", "lib"),
           "-lurmvo_b200", "-Wl,-rpath," + os.path.join(ROOT, "ur-mvo_b200", "lib"), "-o", DRIVER]
    subprocess.check_call(cmd)


def test_adapters_compile_and_link_against_the_c_abi():
    build_driver()
    assert os.path.exists(DRIVER)
    syms = subprocess.run(["nm", "-C", DRIVER], capture_output=True, text=True).stdout
    for s in ("LocalmapOptimization(", "FrameOptimization(", "SolvePnPWithCV(", "EpipolarGeometry::reconstruct(", "EpipolarGeometry::Random::RandomInt(",
              "FindFundamentalInliersGPU("):
        assert s in syms, s


def _run(mode, payload):
    if not os.path.exists(DRIVER):
        build_driver()
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        open(fi, "wb").write(payload)
        subprocess.check_call([DRIVER, mode, fi, fo])
        return open(fo, "rb").read()


@pytest.mark.gpu
def test_localmap_optimization_through_the_adapter(oracle):
    p = synth.small_ba(seed=7)
    Nc, Np, No = p["poses"].shape[0], p["pts"].shape[0], p["uv"].shape[0]
    frame_ids = (np.arange(Nc) * 3 + 11).astype(np.int32)      # sparse, ascending ids like frame ids
    point_ids = (np.arange(Np) * 7 + 5).astype(np.int32)
    buf = struct.pack("3i", Nc, Np, No) + p["intr"].tobytes() + frame_ids.tobytes() + p["poses"].tobytes() + p["fixed"].tobytes() \
        + point_ids.tobytes() + p["pts"].tobytes() + p["uv"].tobytes() + p["obs_cam"].tobytes() + p["obs_pt"].tobytes()
    out = _run("ba", buf)
    poses = np.frombuffer(out[:Nc * 56], dtype=np.float64).reshape(Nc, 7)
    pts = np.frombuffer(out[Nc * 56:Nc * 56 + Np * 24], dtype=np.float64).reshape(Np, 3)
    inl = np.frombuffer(out[Nc * 56 + Np * 24:], dtype=np.uint8)
    op, ox, oi, _ = oracle.local_ba(p)
    assert np.abs(poses - op).max() < 1e-5 and np.abs(pts - ox).max() < 1e-4 and np.array_equal(inl, oi)


@pytest.mark.gpu
def test_frame_optimization_through_the_adapter(oracle):
    b = synth.make_pose_batch(5, B=1, n_obs=250)
    buf = struct.pack("i", 250) + b["intr"].tobytes() + b["poses"][0].tobytes() + b["uv"].tobytes() + b["Xw"].tobytes()
    out = _run("pose", buf)
    pose = np.frombuffer(out[:56], dtype=np.float64)
    inl = np.frombuffer(out[56:56 + 250], dtype=np.uint8)
    n = struct.unpack("i", out[56 + 250:])[0]
    op, oi, on, _ = oracle.pose_only(b["poses"][0], b["uv"], b["Xw"], b["intr"])
    assert np.abs(pose - op).max() < 1e-5 and np.array_equal(inl, oi) and n == on


@pytest.mark.gpu
def test_stereo_camera_through_the_adapters(oracle):
    """Camera type STEREO (reference src/g2o_optimization.cc:96-118, 235-258): the adapters add one 3-row
    edge per StereoPointConstraint next to the mono edges; results and the per-vector inlier flags match the oracle."""
    p = synth.add_stereo(synth.small_ba(seed=9), 21)
    Nc, Np, No = p["poses"].shape[0], p["pts"].shape[0], p["uv3"].shape[0]
    frame_ids = (np.arange(Nc) * 2 + 3).astype(np.int32)
    point_ids = (np.arange(Np) * 5 + 1).astype(np.int32)
    buf = struct.pack("3i", Nc, Np, No) + p["intr5"].tobytes() + frame_ids.tobytes() + p["poses"].tobytes() + p["fixed"].tobytes() \
        + point_ids.tobytes() + p["pts"].tobytes() + p["uv3"].tobytes() + p["kind"].tobytes() + p["obs_cam"].tobytes() + p["obs_pt"].tobytes()
    out = _run("ba_stereo", buf)
    poses = np.frombuffer(out[:Nc * 56], dtype=np.float64).reshape(Nc, 7)
    pts = np.frombuffer(out[Nc * 56:Nc * 56 + Np * 24], dtype=np.float64).reshape(Np, 3)
    inl = np.frombuffer(out[Nc * 56 + Np * 24:Nc * 56 + Np * 24 + No], dtype=np.uint8)
    status = struct.unpack("i", out[Nc * 56 + Np * 24 + No:])[0]
    # the adapter concatenates [mono constraints | stereo constraints]; the oracle gets the same order
    order = np.r_[np.nonzero(p["kind"] == 0)[0], np.nonzero(p["kind"] == 1)[0]]
    q = dict(p, uv3=p["uv3"][order], kind=p["kind"][order], obs_cam=p["obs_cam"][order], obs_pt=p["obs_pt"][order])
    op, ox, oi, _ = oracle.local_ba_stereo(q, 10.0, 75.0)
    back = np.empty_like(oi); back[order] = oi
    assert status == 0
    assert np.abs(poses - op).max() < 1e-5 and np.abs(pts - ox).max() < 1e-4 and np.array_equal(inl, back)
    b = synth.make_pose_batch_stereo(6, B=1, n_obs=220)
    buf = struct.pack("i", 220) + b["intr5"].tobytes() + b["poses"][0].tobytes() + b["uv3"].tobytes() + b["kind"].tobytes() + b["Xw"].tobytes()
    out = _run("pose_stereo", buf)
    pose = np.frombuffer(out[:56], dtype=np.float64)
    inl = np.frombuffer(out[56:56 + 220], dtype=np.uint8)
    n = struct.unpack("i", out[56 + 220:])[0]
    order = np.r_[np.nonzero(b["kind"] == 0)[0], np.nonzero(b["kind"] == 1)[0]]
    q = dict(b, uv3=b["uv3"][order], kind=b["kind"][order], Xw=b["Xw"][order])
    op, oi, on = oracle.pose_only_batch_stereo(q, 10.0, 75.0)
    back = np.empty_like(oi); back[order] = oi
    assert np.abs(pose - op[0]).max() < 1e-5 and np.array_equal(inl, back) and n == on[0]


@pytest.mark.gpu
def test_several_cameras_in_camera_list_through_the_adapters(oracle):
    """camera_list with several different cameras: the reference reads the intrinsics of every edge from
    camera_list[mpc->id_camera] (src/g2o_optimization.cc:86-89, :106-113, :221-224, :243-250); the adapters send the
    camera-model table and per-edge indices through the multicam entry points."""
    p = synth.add_camera_models(synth.add_stereo(synth.small_ba(seed=15, n_pts=200), 4), 8, n_models=3)
    Nc, Np, No = p["poses"].shape[0], p["pts"].shape[0], p["uv3"].shape[0]
    frame_ids = (np.arange(Nc) * 2 + 3).astype(np.int32)
    point_ids = (np.arange(Np) * 5 + 1).astype(np.int32)
    km = p["kind_model"]
    buf = struct.pack("4i", Nc, Np, No, 3) + p["intr5_tab"].tobytes() + frame_ids.tobytes() + p["poses"].tobytes() + p["fixed"].tobytes() \
        + point_ids.tobytes() + p["pts"].tobytes() + p["uv3"].tobytes() + km.tobytes() + p["obs_cam"].tobytes() + p["obs_pt"].tobytes()
    out = _run("ba_multicam", buf)
    poses = np.frombuffer(out[:Nc * 56], dtype=np.float64).reshape(Nc, 7)
    pts = np.frombuffer(out[Nc * 56:Nc * 56 + Np * 24], dtype=np.float64).reshape(Np, 3)
    inl = np.frombuffer(out[Nc * 56 + Np * 24:Nc * 56 + Np * 24 + No], dtype=np.uint8)
    status = struct.unpack("i", out[Nc * 56 + Np * 24 + No:])[0]
    order = np.r_[np.nonzero((km & 1) == 0)[0], np.nonzero((km & 1) == 1)[0]]
    # the adapter numbers the camera models in order of first use; the oracle takes the same table order
    first = []
    for m in (km[order] >> 1):
        if m not in first:
            first.append(int(m))
    remap = np.zeros(3, dtype=np.uint8); remap[first] = np.arange(len(first))
    q = dict(p, uv3=p["uv3"][order], kind_model=((km[order] & 1) | (remap[km[order] >> 1] << 1)).astype(np.uint8),
             intr5_tab=p["intr5_tab"][first], obs_cam=p["obs_cam"][order], obs_pt=p["obs_pt"][order])
    op, ox, oi, _ = oracle.local_ba_multicam(q, 10.0, 75.0)
    back = np.empty_like(oi); back[order] = oi
    assert status == 0
    assert np.abs(poses - op).max() < 1e-5 and np.abs(pts - ox).max() < 1e-4 and np.array_equal(inl, back)
    b = synth.add_camera_models(synth.make_pose_batch_stereo(16, B=1, n_obs=220), 3, n_models=4)
    km = b["kind_model"]
    buf = struct.pack("2i", 220, 4) + b["intr5_tab"].tobytes() + b["poses"][0].tobytes() + b["uv3"].tobytes() + km.tobytes() + b["Xw"].tobytes()
    out = _run("pose_multicam", buf)
    pose = np.frombuffer(out[:56], dtype=np.float64)
    inl = np.frombuffer(out[56:56 + 220], dtype=np.uint8)
    n, status = struct.unpack("2i", out[56 + 220:])
    order = np.r_[np.nonzero((km & 1) == 0)[0], np.nonzero((km & 1) == 1)[0]]
    q = dict(b, uv3=b["uv3"][order], kind_model=km[order], Xw=b["Xw"][order])
    op, oi, on = oracle.pose_only_batch_multicam(q, 10.0, 75.0)
    back = np.empty_like(oi); back[order] = oi
    assert status == 0
    assert np.abs(pose - op[0]).max() < 1e-5 and np.array_equal(inl, back) and n == on[0]


@pytest.mark.gpu
def test_reconstruct_through_the_adapter_uses_glibc_rand_sets(oracle):
    tv = synth.make_two_view(1003, n_keys=400)
    its = 64
    buf = struct.pack("3i", 400, 400, its) + tv["K"].tobytes() + tv["keys1"].tobytes() + tv["keys2"].tobytes() + tv["matches12"].tobytes()
    out = _run("tv", buf)
    ok = struct.unpack("i", out[:4])[0]
    T = np.frombuffer(out[4:68], dtype=np.float32).reshape(4, 4)
    P = np.frombuffer(out[68:68 + 400 * 12], dtype=np.float32).reshape(400, 3)
    tri = np.frombuffer(out[68 + 400 * 12:], dtype=np.uint8)
    tv["sets"] = synth.draw_sets(400, its, 0)   # a fresh process seeds rand() once with 0 (:56)
    o = oracle.two_view(tv)
    assert bool(ok) == o["ok"]
    assert np.array_equal(T.view(np.uint32), o["T21"].view(np.uint32))
    assert np.array_equal(P.view(np.uint32), o["P3D"].view(np.uint32)) and np.array_equal(tri, o["triangulated"])


@pytest.mark.gpu
def test_point_matching_outlier_rejection_through_the_adapter():
    """FindFundamentalInliersGPU (the call of src/point_matching.cc:53) against the committed output
    of the real cv2.findFundamentalMat (RANSAC, and the LMedS / 7-point branches below 15 matches); fewer than 7
    matches are handed back to the caller."""
    G = np.load(os.path.join(ROOT, "tests", "golden", "golden_fm_r01.npz"))
    for k in (3, 11):
        p0, p1 = G[f"p0_{k}"], G[f"p1_{k}"]
        out = _run("fm", struct.pack("i", len(p0)) + p0.tobytes() + p1.tobytes())
        assert struct.unpack("i", out[:4])[0] == 1
        assert np.array_equal(np.frombuffer(out[4:], dtype=np.uint8), G[f"mask_{k}"])
    GS = np.load(os.path.join(ROOT, "tests", "golden", "golden_fm_small_r02.npz"))
    for k in (1, 9):  # N == 7 (every match flagged) and N == 14 (LMedS)
        p0, p1 = GS[f"p0_{k}"], GS[f"p1_{k}"]
        out = _run("fm", struct.pack("i", len(p0)) + p0.tobytes() + p1.tobytes())
        assert struct.unpack("i", out[:4])[0] == 1
        assert np.array_equal(np.frombuffer(out[4:], dtype=np.uint8), GS[f"mask_{k}"])
    p0, p1 = synth.make_fm(3500, 6, 0.9)
    out = _run("fm", struct.pack("i", 6) + p0.tobytes() + p1.tobytes())
    assert struct.unpack("i", out[:4])[0] == 0 and set(out[4:]) == {7}  # untouched


@pytest.mark.gpu
def test_solve_pnp_with_cv_through_the_adapter(oracle):
    """SolvePnPWithCV (reference src/g2o_optimization.cc:323-377): null / invalid mappoints are skipped, the pose
    comes back as T_wc, inliers carry mappoint ids at the FRAME's slot indices."""
    p = synth.make_pnp(4400, 400, 0.25)
    n = 400
    valid = np.ones(n, dtype=np.uint8)
    valid[::17] = 0   # no mappoint in this slot
    valid[5::23] = 2  # mappoint exists but is invalid
    buf = struct.pack("i", n) + p["intr"].tobytes() + p["obj"].astype(np.float64).tobytes() + p["img"].astype(np.float64).tobytes() + valid.tobytes()
    out = _run("pnp", buf)
    cnt, status = struct.unpack("2i", out[:8])
    Twc = np.frombuffer(out[8:8 + 128], dtype=np.float64).reshape(4, 4)
    inl = np.frombuffer(out[8 + 128:], dtype=np.int32)
    use = valid == 1
    o = oracle.pnp_ransac(p["obj"][use], p["img"][use], p["intr"])
    assert status == 0 and cnt == o["n_inliers"]
    want = np.full(n, -1, dtype=np.int32)
    want[np.flatnonzero(use)[o["mask"] == 1]] = 5000 + 3 * np.flatnonzero(use)[o["mask"] == 1]
    assert np.array_equal(inl, want)
    Rwc, twc = o["R"].T, -o["R"].T @ o["t"]
    assert np.abs(Twc[:3, :3] - Rwc).max() < 1e-9 and np.abs(Twc[:3, 3] - twc).max() < 1e-9 and np.array_equal(Twc[3], [0, 0, 0, 1])


@pytest.mark.gpu
def test_solve_pnp_with_cv_needs_eight_points():
    p = synth.make_pnp(4401, 7, 0.0)
    buf = struct.pack("i", 7) + p["intr"].tobytes() + p["obj"].astype(np.float64).tobytes() + p["img"].astype(np.float64).tobytes() + np.ones(7, dtype=np.uint8).tobytes()
    out = _run("pnp", buf)
    assert struct.unpack("2i", out[:8]) == (0, 0)  # reference :349-350: fewer than 8 correspondences -> 0
