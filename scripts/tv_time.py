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
import time
r = plan.reconstruct()
t0 = time.perf_counter()
for _ in range(5): r = plan.reconstruct()
print("reconstruct (motion kernel + host decisions) ms", (time.perf_counter() - t0) / 5 * 1e3, "ok", r["ok"])
ctx.two_view(tv)
t0 = time.perf_counter()
for _ in range(5): ctx.two_view(tv)
print("urmvo_two_view host-buffer call, 8192 iterations, ms", (time.perf_counter() - t0) / 5 * 1e3)
tv200 = dict(tv); tv200["sets"] = tv["sets"][:200]
ctx.two_view(tv200)
t0 = time.perf_counter()
for _ in range(5): ctx.two_view(tv200)
print("urmvo_two_view host-buffer call, 200 iterations (the reference's default), ms", (time.perf_counter() - t0) / 5 * 1e3)
