"""ncu driver: one tile-mode run of cfg5 (or cfg4) on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import urmvo_b200 as U
from urmvo_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
p = synth.cfg5() if which == "cfg5" else synth.cfg4()
ctx = U.Context(0)
loc = U.shard_points(p, 0, 1)
plan = U.ShardedBAPlan(ctx, loc, covis=U.ba_covisibility(loc), it0=3, it1=0)
plan.run()
print(plan.phase_info())
plan.close(); ctx.close()
