"""Development check on a GPU box: GPU path vs the CPU oracle, verbose. Not part of the product."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
import pyoracle as po

ctx = U.Context(0)
which = sys.argv[1:] or ["ba_small", "ba_cfg1", "pose", "tv"]

def cmp_ba(name, prob, opts=None):
    t = time.time(); gp, gx, gi, gs = ctx.local_ba(prob, opts=opts); tg = time.time() - t
    t = time.time(); op, ox, oi, os_ = po.local_ba(prob); to = time.time() - t
    print(f"[{name}] Nc={prob['poses'].shape[0]} Np={prob['pts'].shape[0]} No={prob['uv'].shape[0]} gpu {tg*1e3:.2f} ms oracle {to*1e3:.2f} ms")
    print("   gpu iters", list(gs.iters), "trials", list(gs.trials), "pcg", list(gs.pcg_iters), "lvl1", gs.n_level1,
          "chi0", gs.chi2_initial, "chi", list(gs.chi2_final), "lam", list(gs.lambda_final))
    print("   orc iters", list(os_.iters)[:2], "chi", list(os_.chi2_final)[:2], "lam", list(os_.lambda_final)[:2], "chi0", os_.trace[0].chi2_before)
    rel = [abs(a - b) / max(abs(b), 1e-300) for a, b in zip(gs.chi2_final, list(os_.chi2_final)[:2])]
    print("   rel chi2 diff", rel, "pose maxdiff", np.abs(gp - op).max(), "pts maxdiff", np.abs(gx - ox).max(),
          "inlier mismatches", int((gi != oi).sum()), "/", gi.size)

if "ba_small" in which:
    cmp_ba("small", synth.small_ba())
    cmp_ba("small_cs1", synth.small_ba(seed=3), U.BAOptions(0, 0, 1, 64, 0))
    cmp_ba("small_noisy", synth.small_ba(seed=11, rot_sigma_deg=4.0, trans_sigma=0.3, pt_sigma=0.5))
if "ba_cfg1" in which:
    p1 = synth.cfg1()
    cmp_ba("cfg1", p1)
    cmp_ba("cfg1 again", p1)
    # batch of windows through the plan API
    probs = [synth.make_ba(2000 + i, 10, 2000, 7.7, 10, 3, 0.05) for i in range(8)]
    batch = pack_ba_batch(probs)
    plan = U.BAPlan(ctx, batch)
    plan.run(); ctx.sync()
    t = time.time(); plan.run(); ctx.sync(); dt = time.time() - t
    poses, pts, inl, sts = plan.download()
    its = sum(s.iters[0] + s.iters[1] for s in sts)
    print(f"[batch8 cfg1-shaped] {dt*1e3:.3f} ms, {its} LM iters -> {its/dt:.0f} it/s")
    o = po.local_ba(probs[3])
    c0, c1 = batch["cam_off"][3], batch["cam_off"][4]
    print("   window 3 pose maxdiff vs oracle", np.abs(poses[c0:c1] - o[0]).max(), "chi", list(sts[3].chi2_final), list(o[3].chi2_final)[:2])
if "ba_cfg4" in which:
    cmp_ba("cfg4", synth.cfg4())
if "pose" in which:
    b = synth.make_pose_batch(5, B=16, n_obs=300)
    t = time.time(); gp, gi, gn = ctx.pose_only_batch(b); tg = time.time() - t
    op, oi, on = po.pose_only_batch(b)
    print(f"[pose16] gpu {tg*1e3:.2f} ms; pose maxdiff {np.abs(gp-op).max():.3e}; inlier mismatches {(gi!=oi).sum()}; n_inl diff {np.abs(gn-on).max()}")
    b2 = synth.cfg2()
    plan = U.PosePlan(ctx, b2)
    plan.run(); ctx.sync()
    t = time.time(); plan.run(); ctx.sync(); dt = time.time() - t
    gp, gi, gn, it = plan.download()
    t = time.time(); op, oi, on = po.pose_only_batch(b2); to = time.time() - t
    print(f"[cfg2] gpu {dt*1e3:.3f} ms ({it.sum()} LM it -> {it.sum()/dt:.0f} it/s) oracle {to*1e3:.1f} ms; pose maxdiff {np.abs(gp-op).max():.3e}; inl mism {(gi!=oi).sum()}; n_inl diff {np.abs(gn-on).max()}")
if "tv" in which:
    tv = synth.cfg3(n_hyp=512)
    plan = U.TVPlan(ctx, tv)
    plan.run_ransac(); ctx.sync()
    for model in (0, 1):
        gs, gm, gM = plan.download_hyps(model)
        os_, om, oM = po.score_all(tv, model)
        print(f"[tv model {model}] score bit-mismatches {(gs.view(np.uint32)!=os_.view(np.uint32)).sum()} / {gs.size}; mask word mismatches {(gm!=om).sum()}; model bit-mismatches {(gM.view(np.uint32)!=oM.view(np.uint32)).sum()}; max |score diff| {np.abs(gs-os_).max()}")
    r = plan.reconstruct(); o = po.two_view(tv)
    gs, os_ = r["stats"], o["stats"]
    print("   ok", r["ok"], o["ok"], "SH/SF", gs.SH, gs.SF, os_.SH, os_.SF, "best", gs.best_H, gs.best_F, os_.best_H, os_.best_F)
    print("   n_good", list(gs.n_good), list(os_.n_good), "parallax", list(gs.parallax)[:4], list(os_.parallax)[:4], "motion", gs.best_motion, os_.best_motion)
    print("   T21 bit-equal", np.array_equal(r["T21"], o["T21"]), "P3D bit-equal", np.array_equal(r["P3D"].view(np.uint32), o["P3D"].view(np.uint32)),
          "tri equal", np.array_equal(r["triangulated"], o["triangulated"]), "maskF equal", np.array_equal(r["mask_F"], o["mask_F"]), "maskH equal", np.array_equal(r["mask_H"], o["mask_H"]))
    tv = synth.cfg3()
    plan = U.TVPlan(ctx, tv)
    plan.run_ransac(); ctx.sync()
    t = time.time(); plan.run_ransac(); ctx.sync(); dt = time.time() - t
    print(f"[cfg3] 2x8192 hyps in {dt*1e3:.3f} ms -> {2*8192/dt:.0f} hyps/s")
    tvp = synth.make_two_view(5, planar=True); tvp["sets"] = synth.draw_sets(1000, 300, 0)
    r = ctx.two_view(tvp); o = po.two_view(tvp)
    print("   planar ok", r["ok"], o["ok"], "used_H", r["stats"].used_H, o["stats"].used_H, "n_good", list(r["stats"].n_good), list(o["stats"].n_good),
          "T21 equal", np.array_equal(r["T21"], o["T21"]))
print("launches", ctx.launches)
