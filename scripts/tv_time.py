import sys
sys.path.insert(0, "ur-mvo_b200/python")
import torch, urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0); stream = torch.cuda.ExternalStream(ctx.stream)
tv = synth.cfg3(); plan = U.TVPlan(ctx, tv); plan.run_ransac(); ctx.sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(stream)
for _ in range(10): plan.run_ransac()
b.record(stream); ctx.sync()
print("cfg3 ransac ms", a.elapsed_time(b)/10, "hyps/s", 16384/(a.elapsed_time(b)/10*1e-3))
