"""Development: phase cycle counters of window 0 in a batch of 148 cfg1-shaped windows (one CTA per window)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import torch, urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
ctx = U.Context(0); stream = torch.cuda.ExternalStream(ctx.stream)
probs = [synth.cfg1(seed=1001 + (i % 37)) for i in range(148)]
plan = U.BAPlan(ctx, pack_ba_batch(probs))
plan.run(); ctx.sync(); ctx.ba_timing()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(5): plan.run()
b.record(stream); ctx.sync()
t = ctx.ba_timing(); st = plan.download()[3][0]
tr = st.trials[0] + st.trials[1]
tot = sum(t)
print(f"148 windows: {a.elapsed_time(b)/5:.3f} ms, window 0 trials {tr}; cycles per trial: " +
      " ".join(f"{n}={v/5/tr:.0f}({100*v/tot:.0f}%)" for n, v in zip(["lin0", "lin", "red", "solve", "cam", "back", "red2", "-"], t)))
