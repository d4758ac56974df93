import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
import pyoracle as po
ctx = U.Context(0)
probs = [synth.small_ba(seed=40 + i, n_cams=5 + i % 3, n_pts=100 + 10 * i) for i in range(6)]
batch = pack_ba_batch(probs)
plan = U.BAPlan(ctx, batch)
runs = []
for r in range(3):
    plan.run(); runs.append(plan.download())
for w, p in enumerate(probs):
    op, ox, oi, os_ = po.local_ba(p)
    c = slice(batch["cam_off"][w], batch["cam_off"][w + 1])
    print("win", w, "Nc", p["poses"].shape[0], "Np", p["pts"].shape[0], "oracle it", list(os_.iters)[:2], "trials", [r_[3] for r_ in os_.rows()])
    for r in range(3):
        s = runs[r][3][w]
        print("   run", r, "posediff", np.abs(runs[r][0][c] - op).max(), "it", list(s.iters), "tr", list(s.trials), "pcg", list(s.pcg_iters), "chi", list(s.chi2_final), "vs", list(os_.chi2_final)[:2])
