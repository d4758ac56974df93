"""Small problems through every kernel family added or changed in round 2 (for compute-sanitizer memcheck / racecheck):
cyclic reduction (2 ... 4 levels), tile-mode LIN / BACKSUB, tiled Cholesky of the small windows, warp EPnP + score +
refine, register-resident homography fit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0)
for n_cams, n_pts, span in ((40, 600, 14), (130, 1500, 16)):
    p = synth.make_ba(500 + n_cams, n_cams, n_pts, 0.6 * span, span, 2, 0.02)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p), opts=U.BAOptions(0, 0, 0, 0, 0, 0, 0, 2), it0=2, it1=1)
    plan.run(); st = plan.download()[3]; print("bcr", n_cams, plan.phase_info()["band_solver"], st.chi2_final[1]); plan.close()
for solver in (0, 1, 2):
    r = ctx.local_ba(synth.small_ba(seed=7), it0=3, it1=1, opts=U.BAOptions(0, 0, 0, 0, 0, solver))
    print("small window, dense solver", solver, r[3].chi2_final[1])
f = synth.make_pnp(1008, 200, 0.3)
g = ctx.pnp_ransac(f["obj"], f["img"], f["intr"]); print("pnp", g["iters"], g["n_inliers"])
gb = ctx.pnp_ransac_batch([(f["obj"], f["img"]), (f["obj"][:50], f["img"][:50])], f["intr"]); print("pnp batch", [x["n_inliers"] for x in gb])
tv = synth.make_two_view(1003, n_keys=300); tv["sets"] = synth.draw_sets(300, 64, 0)
print("two-view ok", ctx.two_view(tv)["ok"])
ctx.close()
