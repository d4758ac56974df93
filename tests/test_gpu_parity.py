"""GPU (B200): the CUDA path through the C ABI against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star / SURVEY.md §8d):
  BA       final robust cost within 1e-6 relative, poses within 1e-5 per element (fp64), inlier flags
           identical, LM accept/reject pattern identical.  The reduced camera system is solved by
           block-Jacobi PCG (tolerance 1e-10) on the GPU and by Cholesky in the oracle.
  RANSAC   scores, inlier masks and models bit-exact for identical 8-point sets.
"""
import os

import numpy as np
import pytest

import urmvo_b200 as U
from conftest import GOLDEN
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch

pytestmark = pytest.mark.gpu

REL_COST = 1e-6
POSE_TOL = 1e-5


def _check_ba(oracle, ctx, prob, opts=None, it0=10, it1=5):
    gp, gx, gi, gs = ctx.local_ba(prob, it0=it0, it1=it1, opts=opts)
    op, ox, oi, os_ = oracle.local_ba(prob, it0=it0, it1=it1)
    for k in range(2):
        if abs(os_.chi2_final[k]) > 0:
            assert abs(gs.chi2_final[k] - os_.chi2_final[k]) <= REL_COST * abs(os_.chi2_final[k]), (k, gs.chi2_final[k], os_.chi2_final[k])
        assert gs.iters[k] == os_.iters[k]
    rows = os_.rows()
    assert gs.trials[0] == sum(r[3] for r in rows[: os_.iters[0]])
    assert gs.trials[1] == sum(r[3] for r in rows[os_.iters[0]:])
    assert np.abs(gp - op).max() <= POSE_TOL
    assert np.abs(gx - ox).max() <= 1e-4 * max(1.0, np.abs(ox).max())
    # flags identical except observations within 1e-6 of the chi2 threshold
    assert (gi != oi).sum() <= 0
    return gs, os_


@pytest.mark.parametrize("seed", [3, 7, 19])
def test_ba_small_windows(oracle, ctx, seed):
    _check_ba(oracle, ctx, synth.small_ba(seed=seed))


def test_ba_noisy_start_with_rejected_steps(oracle, ctx):
    p = synth.small_ba(seed=31, rot_sigma_deg=6.0, trans_sigma=0.5, pt_sigma=0.8, outlier_frac=0.1)
    gs, os_ = _check_ba(oracle, ctx, p)
    assert gs.trials[0] >= gs.iters[0]


@pytest.mark.parametrize("atomic", [0, 1, 2])
def test_ba_cfg1_window(oracle, ctx, atomic):
    gs, _ = _check_ba(oracle, ctx, synth.cfg1(), opts=U.BAOptions(0, 0, 0, 0, atomic))
    assert gs.iters[0] == 10 and gs.iters[1] == 5


def test_ba_points_without_observations(oracle, ctx):
    p = synth.small_ba(seed=6)
    q = dict(p, pts=np.vstack([p["pts"], [[9.0, 9.0, 9.0], [1.0, 2.0, 3.0]]]))  # two points nobody observes
    gp, gx, gi, gs = ctx.local_ba(q)
    op, ox, oi, os_ = oracle.local_ba(q)
    assert np.abs(gp - op).max() < 1e-5 and np.array_equal(gi, oi)
    assert np.array_equal(gx[-2:], q["pts"][-2:])


def test_ba_smem_path_is_bit_reproducible(ctx):
    """The small-window paths have no atomics: two runs are bit-identical."""
    p = synth.cfg1()
    a = ctx.local_ba(p)
    b = ctx.local_ba(p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3].chi2_final[1] == b[3].chi2_final[1]


def test_ba_sixteen_free_cameras_and_duplicate_camera_fallback(oracle, ctx):
    _check_ba(oracle, ctx, synth.make_ba(27, 12, 400, 7.0, 12, 2, 0.03))   # Ncf = 10: 55 blocks, 2 per lane
    _check_ba(oracle, ctx, synth.make_ba(28, 13, 400, 7.0, 13, 2, 0.03))   # Ncf = 11: 66 blocks -> smem RMW
    _check_ba(oracle, ctx, synth.make_ba(29, 18, 400, 7.0, 18, 2, 0.03))   # Ncf = 16: largest smem system
    _check_ba(oracle, ctx, synth.make_ba(30, 24, 400, 7.0, 24, 2, 0.03))   # Ncf = 22: atomic path
    p = synth.small_ba(seed=8)
    dup = dict(p, uv=np.vstack([p["uv"], p["uv"][:1] + 1.0]), obs_cam=np.r_[p["obs_cam"], p["obs_cam"][:1]],
               obs_pt=np.r_[p["obs_pt"], p["obs_pt"][:1]])
    _check_ba(oracle, ctx, dup)  # same camera sees a point twice -> atomic path, still correct


@pytest.mark.parametrize("atomic", [0, 1, 2])
@pytest.mark.parametrize("cs", [1, 2, 4, 8, 16])
def test_ba_cluster_sizes_and_accumulation_modes_agree(oracle, ctx, cs, atomic):
    """force_atomic 0: packed groups + register-resident Schur blocks (default for small windows);
    1: global fp64 atomics + BSR PCG; 2: shared-memory read-modify-write copies + dense PCG."""
    _check_ba(oracle, ctx, synth.small_ba(seed=5, n_pts=300), opts=U.BAOptions(0, 0, cs, 128, atomic))


def test_ba_unsorted_observations_and_flag_order(oracle, ctx):
    p = synth.small_ba(seed=13)
    rng = np.random.default_rng(0)
    perm = rng.permutation(p["uv"].shape[0])
    q = dict(p, uv=np.ascontiguousarray(p["uv"][perm]), obs_cam=np.ascontiguousarray(p["obs_cam"][perm]),
             obs_pt=np.ascontiguousarray(p["obs_pt"][perm]))
    gp, gx, gi, gs = ctx.local_ba(q)
    sp, sx, si, ss = ctx.local_ba(p)
    assert np.abs(gp - sp).max() < 1e-9 and np.abs(gx - sx).max() < 1e-8
    assert np.array_equal(gi, si[perm])  # flags come back in the caller's order


def test_ba_all_cameras_fixed_only_moves_points(oracle, ctx):
    p = synth.small_ba(seed=17)
    p["fixed"][:] = 1
    gs, _ = _check_ba(oracle, ctx, p)
    gp = ctx.local_ba(p)[0]
    assert np.abs(synth.quat_to_R(gp[:, :4]) - synth.quat_to_R(p["poses"][:, :4])).max() < 1e-12


def test_ba_point_seen_by_more_than_32_cameras(oracle, ctx):
    p = synth.make_ba(23, 40, 60, 36, 40, 2, 0.02)
    assert np.bincount(p["obs_pt"]).max() > 32
    _check_ba(oracle, ctx, p)


def test_ba_zero_iterations_is_identity_on_state(oracle, ctx):
    p = synth.small_ba(seed=2)
    gp, gx, gi, gs = ctx.local_ba(p, it0=0, it1=0)
    op, ox, oi, _ = oracle.local_ba(p, it0=0, it1=0)
    assert np.abs(gp - op).max() < 1e-12 and np.abs(gx - p["pts"]).max() == 0
    assert np.array_equal(gi, oi)


def test_ba_batch_plan_equals_single_windows_and_is_rerunnable(oracle, ctx):
    probs = [synth.small_ba(seed=40 + i, n_cams=5 + i % 3, n_pts=100 + 10 * i) for i in range(6)]
    batch = pack_ba_batch(probs)
    plan = U.BAPlan(ctx, batch)
    plan.run()
    a = plan.download()
    plan.run()  # restarts from the uploaded estimate
    b = plan.download()
    assert np.abs(a[0] - b[0]).max() < 1e-9 and np.array_equal(a[2], b[2])
    for w, p in enumerate(probs):
        op, ox, oi, os_ = oracle.local_ba(p)
        c = slice(batch["cam_off"][w], batch["cam_off"][w + 1])
        o = slice(batch["obs_off"][w], batch["obs_off"][w + 1])
        assert np.abs(a[0][c] - op).max() <= POSE_TOL
        assert np.array_equal(a[2][o], oi)
        assert abs(a[3][w].chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    plan.close()


def test_ba_golden_fixture(ctx):
    G = np.load(os.path.join(GOLDEN, "golden_r01.npz"))
    p = {k: G["ba_in_" + k] for k in ("poses", "fixed", "pts", "uv", "obs_cam", "obs_pt", "intr")}
    gp, gx, gi, gs = ctx.local_ba(p)
    assert np.array_equal(gi, G["ba_inlier"])
    assert np.abs(gp - G["ba_poses"]).max() <= POSE_TOL
    tr = G["ba_trace"]
    assert gs.trials[0] + gs.trials[1] == int(tr[:, 3].sum())
    assert abs(gs.chi2_final[1] - tr[-1, 1]) <= REL_COST * abs(tr[-1, 1])


@pytest.mark.parametrize("large_mode", [0, 1])
def test_ba_large_window_cfg4_grid_kernel(oracle, ctx, large_mode):
    """400k observations: tile mode (phase kernels, S as a block band, direct banded Cholesky;
    large_mode 0) and the round-1 cooperative whole-grid kernel (global atomics + PCG; large_mode 1).
    The oracle takes a few seconds."""
    p = synth.cfg4()
    assert p["uv"].shape[0] > 300000
    _check_ba(oracle, ctx, p, opts=U.BAOptions(0, 0, 0, 0, 0, 0, large_mode))


def test_ba_large_tile_mode_is_selected_and_bit_reproducible(ctx):
    """cfg4 runs in tile mode by default; two runs of the same plan agree to the last bit in the LM
    decisions (iterations, trials) and to 1e-12 in the cost (the chunk flush uses fp64 atomics)."""
    plan = U.BAPlan(ctx, pack_ba_batch([synth.cfg4()]))
    plan.run()
    a = plan.download()
    info = plan.phase_info()
    plan.run()
    b = plan.download()
    plan.close()
    assert info["tile_mode"] and 1 <= info["half_bandwidth_blocks"] <= 16
    assert info["host_syncs"] <= 4  # whole LM iterations are enqueued without synchronising
    assert list(a[3][0].trials) == list(b[3][0].trials) and np.array_equal(a[2], b[2])
    assert abs(a[3][0].chi2_final[1] - b[3][0].chi2_final[1]) <= 1e-12 * abs(b[3][0].chi2_final[1])


def test_ba_bal_scale_cfg5_single_rank_matches_oracle(oracle, ctx):
    """BASELINE configs[4] at full size (1000 cameras, 200k points, ~2M observations) on one rank of
    the point-sharded path: cost, poses, flags and the LM trace against the CPU oracle."""
    p = synth.cfg5()
    assert p["poses"].shape[0] == 1000 and p["uv"].shape[0] > 1500000
    loc = U.shard_points(p, 0, 1)
    plan = U.ShardedBAPlan(ctx, loc, covis=U.ba_covisibility(loc))
    plan.run()
    gp, gx, gi, gs = plan.download()
    assert plan.phase_info()["tile_mode"]
    plan.close()
    op, ox, oi, os_ = oracle.local_ba(p)
    assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    assert np.abs(gp - op).max() <= POSE_TOL and np.abs(gx - ox).max() <= 1e-4
    assert np.array_equal(gi, oi)
    assert list(gs.iters) == list(os_.iters)[:2]
    assert gs.trials[0] + gs.trials[1] == sum(r[3] for r in os_.rows())


@pytest.mark.parametrize("n_cams,n_pts,span", [(40, 1500, 14), (47, 1800, 6), (130, 6000, 16), (333, 12000, 9)])
def test_ba_cyclic_reduction_solver_matches_oracle_and_band_solver(oracle, ctx, n_cams, n_pts, span):
    """Block cyclic reduction of the banded reduced camera system (csrc/ba_bcr.cu, forced with
    band_solver = 2) against the oracle's skyline Cholesky and against the sequential band solve:
    2 ... 6 levels, a partial last super-block, odd and even numbers of super-blocks."""
    p = synth.make_ba(500 + n_cams, n_cams, n_pts, 0.6 * span, span, 2, 0.02)
    res = []
    for solver in (2, 1):
        plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p), opts=U.BAOptions(0, 0, 0, 0, 0, 0, 0, solver))
        plan.run()
        info = plan.phase_info()
        assert info["tile_mode"] and ("cyclic" in info["band_solver"]) == (solver == 2), info
        res.append(plan.download())
        plan.close()
    op, ox, oi, os_ = oracle.local_ba(p)
    for gp, gx, gi, gs in res:
        assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
        assert np.abs(gp - op).max() <= POSE_TOL and np.array_equal(gi, oi)
        assert list(gs.iters) == list(os_.iters)[:2]
        assert gs.trials[0] + gs.trials[1] == sum(r[3] for r in os_.rows())
    assert np.abs(res[0][0] - res[1][0]).max() <= 1e-9


def test_ba_cyclic_reduction_has_no_camera_count_limit(oracle, ctx):
    """3 200 cameras: the sequential band solve keeps the whole right-hand side in shared memory (a limit of about
    2 900 free cameras at half-bandwidth 15), the cyclic reduction does not; parity with the oracle at that size."""
    p = synth.make_ba(4242, 3200, 40000, 5.4, 9, 2, 0.02)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p))
    plan.run()
    info = plan.phase_info()
    assert info["tile_mode"] and "cyclic" in info["band_solver"], info
    gp, gx, gi, gs = plan.download()
    plan.close()
    op, ox, oi, os_ = oracle.local_ba(p)
    assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    assert np.abs(gp - op).max() <= POSE_TOL and np.array_equal(gi, oi)
    assert list(gs.iters) == list(os_.iters)[:2]


# ------------------------------------------------------------------------------- stereo edges (S1)

@pytest.mark.parametrize("force_atomic", [0, 1])
def test_ba_stereo_edges_match_oracle(oracle, ctx, force_atomic):
    """Mono + stereo edges in one window (reference src/g2o_optimization.cc:96-118): accumulation modes 5
    (shared-memory copies) and 6 (global atomics) against the oracle's 3-row edges."""
    for seed, frac in ((7, 0.6), (19, 1.0), (23, 0.15)):
        p = synth.add_stereo(synth.small_ba(seed=seed, n_pts=300), seed + 100, stereo_frac=frac)
        assert p["kind"].sum() > 0
        gp, gx, gi, gs = ctx.local_ba_stereo(p, 10.0, 75.0, opts=U.BAOptions(0, 0, 0, 0, force_atomic))
        op, ox, oi, os_ = oracle.local_ba_stereo(p, 10.0, 75.0)
        assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
        assert list(gs.iters) == list(os_.iters)[:2]
        assert np.abs(gp - op).max() <= POSE_TOL and np.abs(gx - ox).max() <= 1e-4
        assert np.array_equal(gi, oi)


def test_ba_stereo_with_no_stereo_edge_equals_the_mono_call(oracle, ctx):
    p = synth.add_stereo(synth.small_ba(seed=5, n_pts=200), 1, stereo_frac=0.0)
    assert p["kind"].sum() == 0
    a = ctx.local_ba_stereo(p, 10.0, 75.0)
    b = ctx.local_ba(p, opts=U.BAOptions(0, 0, 0, 0, 2))  # the same one-point-per-warp accumulation
    assert np.abs(a[0] - b[0]).max() < 1e-11 and np.array_equal(a[2], b[2])
    assert abs(a[3].chi2_final[1] - b[3].chi2_final[1]) <= 1e-12 * abs(b[3].chi2_final[1])


def test_ba_stereo_medium_window_cfg1_shape(oracle, ctx):
    """A cfg1-sized window of a stereo camera (10 keyframes, 2000 points, 60 % stereo observations)."""
    p = synth.add_stereo(synth.cfg1(), 77)
    gp, gx, gi, gs = ctx.local_ba_stereo(p, 10.0, 75.0)
    op, ox, oi, os_ = oracle.local_ba_stereo(p, 10.0, 75.0)
    assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    assert np.abs(gp - op).max() <= POSE_TOL and np.array_equal(gi, oi)


def test_pose_only_stereo_edges_match_oracle(oracle, ctx):
    """FrameOptimization with EdgeStereoSE3ProjectXYZOnlyPose edges (:235-258): poses, flags, counts."""
    for B in (5, 200):  # both kernel instantiations (one / two CTAs per SM)
        b = synth.make_pose_batch_stereo(31 + B, B=B, n_obs=120)
        gp, gi, gn = ctx.pose_only_batch_stereo(b, 10.0, 75.0)
        op, oi, on = oracle.pose_only_batch_stereo(b, 10.0, 75.0)
        assert np.abs(gp - op).max() <= POSE_TOL
        assert np.array_equal(gi, oi) and np.array_equal(gn, on)


@pytest.mark.parametrize("force_atomic", [0, 1])
def test_ba_several_camera_models_match_oracle(oracle, ctx, force_atomic):
    """Per-constraint intrinsics, camera_list[mpc->id_camera] (reference src/g2o_optimization.cc:86-89, :106-113):
    three / five camera models mixed inside one window, mono and stereo edges, both accumulation modes."""
    for seed, frac, nm in ((7, 0.6, 3), (23, 0.0, 5), (31, 1.0, 2)):
        p = synth.add_camera_models(synth.add_stereo(synth.small_ba(seed=seed, n_pts=300), seed + 100, stereo_frac=frac),
                                    seed + 200, n_models=nm)
        gp, gx, gi, gs = ctx.local_ba_multicam(p, 10.0, 75.0, opts=U.BAOptions(0, 0, 0, 0, force_atomic))
        op, ox, oi, os_ = oracle.local_ba_multicam(p, 10.0, 75.0)
        assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
        assert list(gs.iters) == list(os_.iters)[:2]
        assert np.abs(gp - op).max() <= POSE_TOL and np.abs(gx - ox).max() <= 1e-4
        assert np.array_equal(gi, oi)


def test_ba_one_camera_model_equals_the_stereo_call_and_cfg1_shape(oracle, ctx):
    p = synth.add_stereo(synth.small_ba(seed=5, n_pts=200), 1)
    q = dict(p, kind_model=p["kind"].copy(), intr5_tab=p["intr5"].reshape(1, 5))
    a = ctx.local_ba_stereo(p, 10.0, 75.0)
    b = ctx.local_ba_multicam(q, 10.0, 75.0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    big = synth.add_camera_models(synth.add_stereo(synth.cfg1(), 77), 78, n_models=4)  # 10 keyframes, 2000 points
    gp, gx, gi, gs = ctx.local_ba_multicam(big, 10.0, 75.0)
    op, ox, oi, os_ = oracle.local_ba_multicam(big, 10.0, 75.0)
    assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    assert np.abs(gp - op).max() <= POSE_TOL and np.array_equal(gi, oi)
    bad = dict(big, kind_model=(big["kind_model"] | 0x80).astype(np.uint8))
    with pytest.raises(U.UrmvoError):
        ctx.local_ba_multicam(bad, 10.0, 75.0)


def test_pose_only_several_camera_models_match_oracle(oracle, ctx):
    """FrameOptimization with per-constraint camera models (:221-224, :243-250): poses, flags, counts."""
    for B in (5, 200):
        b = synth.add_camera_models(synth.make_pose_batch_stereo(31 + B, B=B, n_obs=120), B, n_models=4)
        gp, gi, gn = ctx.pose_only_batch_multicam(b, 10.0, 75.0)
        op, oi, on = oracle.pose_only_batch_multicam(b, 10.0, 75.0)
        assert np.abs(gp - op).max() <= POSE_TOL
        assert np.array_equal(gi, oi) and np.array_equal(gn, on)


def test_ba_rejects_bad_input(ctx):
    p = synth.small_ba(seed=2)
    bad = dict(p, obs_cam=p["obs_cam"].copy())
    bad["obs_cam"][5] = 99
    with pytest.raises(U.UrmvoError, match="out of range"):
        ctx.local_ba(bad)


# ------------------------------------------------------------------------------- pose only

def _check_pose(oracle, ctx, b, **kw):
    gp, gi, gn = ctx.pose_only_batch(b, **kw)
    op, oi, on = oracle.pose_only_batch(b)
    assert np.abs(gp - op).max() <= POSE_TOL
    assert np.array_equal(gi, oi) and np.array_equal(gn, on)


def test_pose_only_cfg2_batch(oracle, ctx):
    _check_pose(oracle, ctx, synth.cfg2())


def test_pose_only_ragged_and_tiny_frames(oracle, ctx):
    b = synth.make_pose_batch(6, B=5, n_obs=300)
    # ragged: frame sizes 300, 7 (< 10 edges: one round only), 0 (empty), 150, 300
    keep = np.r_[np.arange(0, 300), np.arange(300, 307), np.arange(900, 1050), np.arange(1200, 1500)]
    b2 = dict(b, uv=np.ascontiguousarray(b["uv"][keep]), Xw=np.ascontiguousarray(b["Xw"][keep]),
              obs_offset=np.array([0, 300, 307, 307, 457, 757], dtype=np.int32))
    _check_pose(oracle, ctx, b2)


def test_pose_only_all_outliers_and_golden(oracle, ctx):
    b = synth.make_pose_batch(9, B=3, n_obs=120, outlier_frac=1.0)
    _check_pose(oracle, ctx, b)
    G = np.load(os.path.join(GOLDEN, "golden_r01.npz"))
    g = {k: G["po_in_" + k] for k in ("poses", "obs_offset", "uv", "Xw", "intr")}
    gp, gi, gn = ctx.pose_only_batch(g)
    assert np.array_equal(gi, G["po_inlier"]) and np.array_equal(gn, G["po_n_inlier"])
    assert np.abs(gp - G["po_poses"]).max() <= POSE_TOL


# ------------------------------------------------------------------------------- two view

def _check_hyps(oracle, plan, tv, score_mode=0):
    for model in (0, 1):
        gs, gm, gM = plan.download_hyps(model)
        os_, om, oM = oracle.score_all(tv, model, score_mode=score_mode)
        assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32)), f"scores differ (model {model})"
        assert np.array_equal(gm, om), f"masks differ (model {model})"
        assert np.array_equal(gM.view(np.uint32), oM.view(np.uint32)), f"models differ (model {model})"


def _check_reconstruct(g, o):
    assert g["ok"] == o["ok"]
    gs, os_ = g["stats"], o["stats"]
    assert (gs.best_F, gs.best_H, gs.used_H, gs.best_motion) == (os_.best_F, os_.best_H, os_.used_H, os_.best_motion)
    assert gs.SF == os_.SF and gs.SH == os_.SH
    assert list(gs.n_good) == list(os_.n_good)
    assert np.array_equal(g["mask_F"], o["mask_F"]) and np.array_equal(g["mask_H"], o["mask_H"])
    assert np.array_equal(g["T21"].view(np.uint32), o["T21"].view(np.uint32))
    assert np.array_equal(g["P3D"].view(np.uint32), o["P3D"].view(np.uint32))
    assert np.array_equal(g["triangulated"], o["triangulated"])


def test_two_view_cfg3_bit_exact(oracle, ctx):
    tv = synth.cfg3(n_hyp=2048)
    plan = U.TVPlan(ctx, tv)
    plan.run_ransac()
    _check_hyps(oracle, plan, tv)
    _check_reconstruct(plan.reconstruct(), oracle.two_view(tv))
    plan.close()


def test_two_view_sampson_mode_bit_exact(oracle, ctx):
    """BASELINE.json north_star (4): fundamental hypotheses scored with the Sampson error (extra mode;
    the reference's rule stays the default).  Scores, masks, winner and the reconstruction are bit-exact
    against the oracle's restatement of the same rule; switching back restores the reference rule."""
    tv = synth.cfg3(n_hyp=1024)
    plan = U.TVPlan(ctx, tv)
    plan.set_score_mode(1)
    plan.run_ransac()
    _check_hyps(oracle, plan, tv, score_mode=1)
    _check_reconstruct(plan.reconstruct(), oracle.two_view(tv, score_mode=1))
    s1 = plan.download_hyps(0)[0]
    plan.set_score_mode(0)
    plan.run_ransac()
    _check_hyps(oracle, plan, tv, score_mode=0)
    assert not np.array_equal(s1, plan.download_hyps(0)[0])
    plan.close()
    small = synth.make_two_view(11, n_keys=300)
    small["sets"] = synth.draw_sets(300, 200, 0)
    _check_reconstruct(ctx.two_view(small, score_mode=1), oracle.two_view(small, score_mode=1))
    with pytest.raises(U.UrmvoError, match="unknown mode"):
        ctx.two_view(small, score_mode=7)


def test_two_view_full_8192_argmax_properties(oracle, ctx):
    """BASELINE size: arg-max is the first maximum, masks have no bits beyond N, and a sample of
    hypotheses is bit-exact against the oracle."""
    tv = synth.cfg3()
    plan = U.TVPlan(ctx, tv)
    plan.run_ransac()
    r = plan.reconstruct()
    for model, best in ((0, r["stats"].best_F), (1, r["stats"].best_H)):
        s, m, M = plan.download_hyps(model)
        assert best == int(np.argmax(s))
        assert (m[:, -1] >> (1000 % 32)).max() == 0
        sub = np.r_[0:64, 4000:4064, 8128:8192]
        os_, om, oM = oracle.score_all(tv, model, sets=tv["sets"][sub])
        assert np.array_equal(s[sub].view(np.uint32), os_.view(np.uint32)) and np.array_equal(m[sub], om)
    plan.close()


@pytest.mark.parametrize("kw", [dict(seed=5, planar=True), dict(seed=12, n_keys=300, n_unmatched=37),
                                dict(seed=14, n_keys=77), dict(seed=15, inlier_frac=0.05)])
def test_two_view_variants(oracle, ctx, kw):
    tv = synth.make_two_view(**kw)
    N = int((tv["matches12"] >= 0).sum())
    tv["sets"] = synth.draw_sets(N, 200, 0)
    plan = U.TVPlan(ctx, tv)
    plan.run_ransac()
    _check_hyps(oracle, plan, tv)
    _check_reconstruct(plan.reconstruct(), oracle.two_view(tv))
    plan.close()


def test_two_view_golden_fixture(ctx):
    G = np.load(os.path.join(GOLDEN, "golden_r01.npz"))
    tv = {k: G["tv_in_" + k] for k in ("keys1", "keys2", "matches12", "K", "sets")}
    tv["sigma"] = 1.0
    g = ctx.two_view(tv)
    assert g["ok"] == bool(G["tv_ok"])
    assert np.array_equal(g["T21"].view(np.uint32), G["tv_T21"].view(np.uint32))
    assert np.array_equal(g["mask_F"], G["tv_mask_F"]) and np.array_equal(g["mask_H"], G["tv_mask_H"])
    assert np.array_equal(g["triangulated"], G["tv_tri"])


def test_two_view_rejects_bad_input(ctx):
    tv = synth.make_two_view(3, n_keys=50)
    tv["sets"] = synth.draw_sets(50, 4, 0)
    tv["sets"][1, 3] = 50
    with pytest.raises(U.UrmvoError, match="out of range"):
        ctx.two_view(tv)
    few = synth.make_two_view(3, n_keys=50, n_unmatched=45)
    few["sets"] = np.zeros((1, 8), dtype=np.int32)
    with pytest.raises(U.UrmvoError, match="fewer than 8"):
        ctx.two_view(few)


# ------------------------------------------------------------------------------- point-sharded BA

@pytest.mark.parametrize("large_mode", [0, 1])
def test_sharded_ba_single_rank_matches_oracle(oracle, ctx, large_mode):
    """The NCCL-sharded phase-kernel path with a world of one rank (all-reduces are identities)."""
    p = synth.make_ba(77, 40, 1500, 8.0, 14, 2, 0.02)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p), opts=U.BAOptions(0, 0, 0, 0, 0, 0, large_mode))
    plan.run()
    assert plan.phase_info()["tile_mode"] == (large_mode == 0)
    gp, gx, gi, gs = plan.download()
    op, ox, oi, os_ = oracle.local_ba(p)
    assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= REL_COST * abs(os_.chi2_final[1])
    assert np.abs(gp - op).max() <= POSE_TOL and np.array_equal(gi, oi)
    assert list(gs.iters) == list(os_.iters)[:2]
    plan.close()


@pytest.mark.parametrize("large_mode", [0, 1])
def test_sharded_ba_two_ranks_over_nccl(oracle, large_mode):
    """Two GPUs, one process each (torchrun): skipped on a single-GPU box."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "sharded_ba.py"), "small", "--check",
                          f"--mode={large_mode}"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if "rel cost diff" in l][0]
    rel = float(line.split("rel cost diff")[1].split(";")[0])
    assert rel < REL_COST and "inlier mismatches 0" in line
    assert ("tile=True" in out.stdout) == (large_mode == 0)


def test_hypothesis_sharded_ransac_two_ranks(oracle):
    """cfg3 hypotheses split over two GPUs, winners merged with the earliest-index tie rule."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29534", os.path.join(root, "scripts", "sharded_ransac.py"), "2048", "--check"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "match" in out.stdout and "MISMATCH" not in out.stdout


# ------------------------------------------------------------------------------- more LM edge cases

def test_ba_far_start_exercises_rejections_and_termination(oracle, ctx):
    """A bad initial estimate: rejected trials (lambda growth), possibly early termination; the
    trial / iteration pattern and the stale-error classification must follow the oracle."""
    for seed, rot, tr, pt in ((51, 12.0, 1.0, 1.5), (52, 20.0, 2.0, 2.0), (53, 0.0, 0.0, 0.0)):
        p = synth.small_ba(seed=seed, rot_sigma_deg=rot, trans_sigma=tr, pt_sigma=pt, outlier_frac=0.15)
        gp, gx, gi, gs = ctx.local_ba(p)
        op, ox, oi, os_ = oracle.local_ba(p)
        rows = os_.rows()
        assert list(gs.iters) == list(os_.iters)[:2]
        assert gs.trials[0] == sum(r[3] for r in rows[: os_.iters[0]])
        if abs(os_.chi2_final[1]) > 0:
            assert abs(gs.chi2_final[1] - os_.chi2_final[1]) <= 1e-6 * abs(os_.chi2_final[1])
        assert (gi != oi).sum() == 0


def test_ba_batch_of_full_size_windows_matches_single_solves(oracle, ctx):
    """BASELINE size: a batch of cfg1 windows (one window per SM) gives, window by window, what the
    single-window (16-CTA cluster) solve and the oracle give."""
    probs = [synth.make_ba(3001 + i, 10, 2000, 7.7, 10, 3, 0.05) for i in range(6)]
    batch = pack_ba_batch(probs * 30)  # 180 windows > 148 SMs
    plan = U.BAPlan(ctx, batch)
    plan.run()
    poses, pts, inl, st = plan.download()
    plan.close()
    for w in (0, 5, 77, 179):
        p = probs[w % 6]
        c = slice(batch["cam_off"][w], batch["cam_off"][w + 1]); o = slice(batch["obs_off"][w], batch["obs_off"][w + 1])
        sp, sx, si, ss = ctx.local_ba(p)
        assert np.abs(poses[c] - sp).max() < 1e-9 and np.array_equal(inl[o], si)
    op, ox, oi, os_ = oracle.local_ba(probs[0])
    assert abs(st[0].chi2_final[1] - os_.chi2_final[1]) <= 1e-6 * abs(os_.chi2_final[1]) and np.array_equal(inl[:oi.size], oi)
    # identical windows give bit-identical results wherever they run
    c0 = slice(batch["cam_off"][0], batch["cam_off"][1]); c6 = slice(batch["cam_off"][6], batch["cam_off"][7])
    assert np.array_equal(poses[c0], poses[c6])


def test_ba_dense_solver_ldlt_and_pcg_agree(ctx, oracle):
    """The in-shared-memory reduced camera system: tiled Cholesky (default), the one-barrier-per-pivot LDL^T
    (dense_solver = 1) and block-Jacobi PCG (dense_solver = 2) all meet the parity bar; the direct solves report
    no PCG iterations."""
    prob = synth.cfg1()
    gs0, _ = _check_ba(oracle, ctx, prob, opts=U.BAOptions(0, 0, 0, 0, 0, 0))
    gs1, _ = _check_ba(oracle, ctx, prob, opts=U.BAOptions(0, 0, 0, 0, 0, 1))
    gs2, _ = _check_ba(oracle, ctx, prob, opts=U.BAOptions(0, 0, 0, 0, 0, 2))
    assert gs0.pcg_iters[0] == 0 and gs0.pcg_iters[1] == 0 and gs1.pcg_iters[0] == 0
    assert gs2.pcg_iters[0] > 0
    assert list(gs0.trials) == list(gs1.trials) == list(gs2.trials)
    # 16 free cameras (n = 96: 528 + 32 tiles do not fit 256 threads -> falls back to the LDL^T) and 2 free cameras
    for n_cams, n_fixed in ((18, 2), (4, 2)):
        p = synth.make_ba(900 + n_cams, n_cams, 60 * n_cams, 4.0, n_cams, n_fixed, 0.03)
        _check_ba(oracle, ctx, p)


def _thin(prob, keep):
    """Keep the observations selected by the boolean mask (arrays stay point-major)."""
    q = dict(prob)
    for k in ("uv", "obs_cam", "obs_pt"):
        q[k] = np.ascontiguousarray(prob[k][keep])
    return q


def test_ba_packed_groups_of_32_single_observation_points(oracle, ctx):
    """Packed mode edge: groups of 32 points with one observation each (the slot table is full, the
    slot prefetch of point q+1 clamps at row 31) mixed with ordinary points."""
    p = synth.make_ba(41, 8, 600, 6.0, 8, 2, 0.02)
    first = np.r_[True, p["obs_pt"][1:] != p["obs_pt"][:-1]]
    keep = first | (p["obs_pt"] >= 300)          # points 0..299 keep only their first observation
    q = _thin(p, keep)
    counts = np.bincount(q["obs_pt"], minlength=600)
    assert (counts[:300] == 1).all() and counts[300:].max() > 3
    _check_ba(oracle, ctx, q)


def test_ba_packed_point_with_32_observations(oracle, ctx):
    """Packed mode edge: kmax = 32 (one point fills a whole group) with 7 free of 40 cameras."""
    p = synth.make_ba(43, 40, 300, 30.0, 40, 33, 0.02)
    counts = np.bincount(p["obs_pt"], minlength=300)
    order = np.argsort(-counts)
    # cap every point at 32 observations (drop the surplus of the few that have more)
    rank_in_pt = np.arange(len(p["obs_pt"])) - np.r_[0, np.cumsum(counts)][p["obs_pt"]]
    q = _thin(p, rank_in_pt < 32)
    assert np.bincount(q["obs_pt"]).max() == 32 and int((q["fixed"] == 0).sum()) == 7
    _check_ba(oracle, ctx, q)
