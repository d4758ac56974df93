"""Hypothesis-sharded two-view RANSAC: every rank scores a contiguous slice of the 8-point sets on
its GPU, the ranks exchange one (score, index) pair per model (earliest index wins ties, like the
reference's strict '>' update), the owner of the winning F / H reconstructs.  Run under torchrun.
  python scripts/sharded_ransac.py [n_hyp] [--check]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import torch.distributed as dist
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.dist import shard_range, merge_best_hypothesis

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8192
tv = synth.cfg3(n_hyp=n_hyp)
a, b = shard_range(n_hyp, rank, world)
ctx = U.Context(lr)
plan = U.TVPlan(ctx, tv, sets=tv["sets"][a:b])
plan.run_ransac(); ctx.sync()
best = {}
for model, name in ((0, "F"), (1, "H")):
    s, m, M = plan.download_hyps(model)
    li = int(np.argmax(s)) if s.max() > 0 else -1
    best[name] = merge_best_hypothesis(float(s[li]) if li >= 0 else 0.0, li, a, dist if world > 1 else None)
if rank == 0:
    print(f"[sharded RANSAC x{world}] best F: score {best['F'][0]:.4f} hyp {best['F'][1]} (rank {best['F'][2]}); best H: score {best['H'][0]:.4f} hyp {best['H'][1]} (rank {best['H'][2]})", flush=True)
    if "--check" in sys.argv:
        import pyoracle as po
        o = po.two_view(tv)
        ok = (o["stats"].best_F == best["F"][1] and o["stats"].best_H == best["H"][1] and
              np.float32(o["stats"].SF) == np.float32(best["F"][0]) and np.float32(o["stats"].SH) == np.float32(best["H"][0]))
        print("   oracle best F/H:", o["stats"].best_F, o["stats"].best_H, "match" if ok else "MISMATCH", flush=True)
plan.close(); ctx.close()
if world > 1:
    dist.destroy_process_group()
