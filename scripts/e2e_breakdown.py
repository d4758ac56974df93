import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np, torch
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
ctx = U.Context(0)
distinct = [synth.make_ba(1001 + i, 10, 2000, 7.7, 10, 3, 0.05) for i in range(37)]
batch = pack_ba_batch([distinct[i % 37] for i in range(148)])
for k in ("poses", "fixed", "pts", "uv", "obs_cam", "obs_pt"):
    batch[k] = torch.from_numpy(batch[k]).pin_memory().numpy()
for rep in range(3):
    t0 = time.perf_counter(); plan = U.BAPlan(ctx, batch); t1 = time.perf_counter()
    plan.run(); ctx.sync(); t2 = time.perf_counter()
    out = plan.download(); t3 = time.perf_counter()
    plan.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  run {1e3*(t2-t1):.2f} ms  download {1e3*(t3-t2):.2f} ms  destroy {1e3*(t4-t3):.2f} ms")
out = {"poses": torch.empty(batch["poses"].shape, dtype=torch.float64).pin_memory().numpy(), "pts": torch.empty(batch["pts"].shape, dtype=torch.float64).pin_memory().numpy(), "inlier": torch.empty(batch["uv"].shape[0], dtype=torch.uint8).pin_memory().numpy()}
for rep in range(3):
    t0 = time.perf_counter(); ctx.local_ba_batch(batch, out=out); t1 = time.perf_counter()
    print(f"one-shot {1e3*(t1-t0):.2f} ms")
