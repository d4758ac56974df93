"""Development: one large window on the cooperative grid kernel with the phase cycle counters."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np, torch
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
prob = synth.cfg4() if which == "cfg4" else synth.cfg5(n_cams=int(which.split(",")[0]), n_pts=int(which.split(",")[1]))
ctx = U.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
plan = U.BAPlan(ctx, pack_ba_batch([prob]))
plan.run(); ctx.sync(); ctx.ba_timing()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream); plan.run(); b.record(stream); ctx.sync()
ms = a.elapsed_time(b)
st = plan.download()[3][0]
t = ctx.ba_timing(); tt = sum(t) or 1
print(f"[{which}] Nc={prob['poses'].shape[0]} No={prob['uv'].shape[0]}: {ms:.2f} ms, iters {list(st.iters)} trials {list(st.trials)} pcg {list(st.pcg_iters)} -> {(st.iters[0]+st.iters[1])/ms*1e3:.1f} it/s")
print("   phase us/run: " + " ".join(f"{n}={v/1.9e3:.0f}us({v/tt*100:.0f}%)" for n, v in zip(["lin0","lin","red","pcg","cam","back","red2","-"], t)))
