"""Racecheck / memcheck driver for the double-buffered staging of k_lg_lin: a long camera chain with few points per
camera, so that every CTA works through SEVERAL chunks (the staging-buffer parity runs on across chunks) of several
rounds each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0)
p = synth.make_ba(909, 2000, 36000, 6.0, 14, 2, 0.02)
plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p), it0=2, it1=1)
plan.run()
info = plan.phase_info()
st = plan.download()[3]
print("tile", info["tile_mode"], info["band_solver"], "chi2", st.chi2_final[1], "obs", p["uv"].shape[0])
plan.close(); ctx.close()
