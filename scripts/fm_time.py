"""Development: timing of the per-frame fundamental-matrix RANSAC. Usage: fm_time.py [B] [matches] [inlier_frac]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
inl = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
ctx = U.Context(0)
pairs = synth.make_fm_batch(1006, B, n, inl)
plan = U.FMPlan(ctx, pairs)
plan.run(); plan.finish()
t0 = time.perf_counter()
for _ in range(5): plan.run()
t_run = (time.perf_counter() - t0) / 5
t0 = time.perf_counter()
for _ in range(5): masks, st = plan.finish()
t_fin = (time.perf_counter() - t0) / 5
print(f"B={B} N={n} inliers={inl}: run {t_run*1e3:.3f} ms  finish {t_fin*1e3:.3f} ms  iterations evaluated {plan.hypotheses}, needed {sum(s.iters for s in st)}")
ctx.fm_ransac(*pairs[0])
for k in range(3):
    t0 = time.perf_counter(); ctx.fm_ransac(*pairs[k % B]); print(f"single call: {(time.perf_counter()-t0)*1e6:.0f} us")
t0 = time.perf_counter(); ctx.fm_ransac_batch(pairs); print(f"batch call: {(time.perf_counter()-t0)*1e3:.3f} ms")
