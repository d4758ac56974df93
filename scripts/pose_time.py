"""Development: pose-only (FrameOptimization) batch timing. Usage: pose_time.py [frames] [matches]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np, torch
import urmvo_b200 as U
from urmvo_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = U.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
pb = synth.make_pose_batch(1002, B=B, n_obs=n)
plan = U.PosePlan(ctx, pb)
plan.run(); ctx.sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(10): plan.run()
b.record(stream); ctx.sync()
ms = a.elapsed_time(b) / 10
its = int(plan.download()[3].sum())
print(f"B={B} n={n}: {ms:.4f} ms  {its/ms*1e3:.0f} LM it/s  ({its} its)")
