"""Small problems through the kernel paths added at the end of round 2 (for compute-sanitizer memcheck / racecheck):
per-constraint camera models in the 3-row BA phases (accumulation modes 5 / 6) and in the pose-only kernel, and the
sub-15-match branches of the fundamental-matrix call (median score warp, per-problem thresholds of the mask kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0)
for force in (0, 1):
    p = synth.add_camera_models(synth.add_stereo(synth.small_ba(seed=7, n_pts=150), 107), 207, n_models=3)
    r = ctx.local_ba_multicam(p, 10.0, 75.0, it0=3, it1=1, opts=U.BAOptions(0, 0, 0, 0, force))
    print("multicam BA, force_atomic", force, r[3].chi2_final[1])
b = synth.add_camera_models(synth.make_pose_batch_stereo(36, B=5, n_obs=120), 5, n_models=4)
print("multicam pose-only", ctx.pose_only_batch_multicam(b, 10.0, 75.0)[2])
pairs = [synth.make_fm(3600 + n, n, 0.8, 0.5, 4.0) for n in (7, 8, 11, 14, 40)]
masks, stats = ctx.fm_ransac_batch(pairs)
print("fm small", [int(m.sum()) for m in masks], [s.iters for s in stats])
ctx.close()
