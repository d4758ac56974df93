"""Development: phase cycle counters of ONE cfg1 window on its 16-CTA cluster."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import torch, urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
ctx = U.Context(0); stream = torch.cuda.ExternalStream(ctx.stream)
for cs in (0, 8, 16):
    plan = U.BAPlan(ctx, pack_ba_batch([synth.cfg1()]), opts=U.BAOptions(0, 0, cs, 0, 0, 0))
    plan.run(); ctx.sync(); ctx.ba_timing()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(10): plan.run()
    b.record(stream); ctx.sync()
    t = ctx.ba_timing(); st = plan.download()[3][0]
    tr = st.trials[0] + st.trials[1]
    print(f"cluster {cs}: {a.elapsed_time(b)/10:.3f} ms, trials {tr}; cycles per trial: " +
          " ".join(f"{n}={v/10/tr:.0f}" for n, v in zip(["lin0", "lin", "red", "solve", "cam", "back", "red2", "-"], t)))
    plan.close()
