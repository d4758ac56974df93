"""Development: small BA windows through every accumulation mode + one FM RANSAC, for compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0)
p = synth.small_ba(seed=7)
for force in (0, 1, 2):
    for solver in (0, 1):
        ctx.local_ba(p, opts=U.BAOptions(0, 0, 0, 0, force, solver))
ctx.fm_ransac(*synth.make_fm(3001, 60, 0.7))
tv = synth.make_two_view(1003, n_keys=120)
tv["sets"] = synth.draw_sets(120, 32, 0)
ctx.two_view(tv)
ctx.pose_only_batch(synth.make_pose_batch(5, B=2, n_obs=100))
t = synth.make_triangulation(13, n_pts=40)
ctx.triangulate_batch(t["obs_off"], t["obs_pose"], t["obs_uv"], t["poses_Rp"], t["intr"])
print("done")
