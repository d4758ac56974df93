"""Regenerates tests/golden/golden_r01.npz from the CPU oracle (run from the repo root:
python tests/golden/make_golden.py).  The reference itself cannot be run here (g2o / Eigen / OpenCV
C++ are absent), so these vectors pin the ORACLE (and through it the kernels) against regressions;
the oracle's own pins are the analytic / scipy / OpenCV cross-checks in tests/test_oracle_*.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from urmvo_b200 import synth  # noqa: E402
import pyoracle as po  # noqa: E402

out = {}
# R2: glibc rand() sets
out["sets_1000x16"] = po.draw_sets(1000, 16, 0)
# B1-B6: small local BA (inputs are regenerated from the seed; stored too so that a change of the
# generator is caught as well)
p = synth.small_ba(seed=7)
poses, pts, inl, st = po.local_ba(p)
for k in ("poses", "fixed", "pts", "uv", "obs_cam", "obs_pt", "intr"):
    out["ba_in_" + k] = p[k]
out["ba_poses"], out["ba_pts"], out["ba_inlier"] = poses, pts, inl
out["ba_trace"] = np.array(st.rows())
out["ba_iters"] = np.array(list(st.iters)[:2])
# B7: pose-only
b = synth.make_pose_batch(5, B=4, n_obs=200)
pp, pi, pn = po.pose_only_batch(b)
for k in ("poses", "obs_offset", "uv", "Xw", "intr"):
    out["po_in_" + k] = b[k]
out["po_poses"], out["po_inlier"], out["po_n_inlier"] = pp, pi, pn
# R1-R8: two-view
tv = synth.make_two_view(1003, n_keys=400)
tv["sets"] = synth.draw_sets(400, 64, 0)
r = po.two_view(tv)
for k in ("keys1", "keys2", "matches12", "K", "sets"):
    out["tv_in_" + k] = tv[k]
sF, mF, MF = po.score_all(tv, 0)
sH, mH, MH = po.score_all(tv, 1)
out.update(tv_ok=np.array(r["ok"]), tv_T21=r["T21"], tv_P3D=r["P3D"], tv_tri=r["triangulated"],
           tv_mask_F=r["mask_F"], tv_mask_H=r["mask_H"], tv_scores_F=sF, tv_scores_H=sH,
           tv_masks_F=mF, tv_masks_H=mH, tv_models_F=MF, tv_models_H=MH,
           tv_best=np.array([r["stats"].best_F, r["stats"].best_H, r["stats"].used_H, r["stats"].best_motion]),
           tv_n_good=np.array(list(r["stats"].n_good)))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_r01.npz"), **out)
print("wrote golden_r01.npz", {k: np.asarray(v).shape for k, v in out.items()})
