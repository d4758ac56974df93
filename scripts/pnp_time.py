"""Development: SolvePnPWithCV call latency (single frame) and batch throughput."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
ctx = U.Context(0)
for n, frac in ((1000, 0.2), (300, 0.2), (1000, 0.5)):
    f = synth.make_pnp(1008, n, frac)
    ctx.pnp_ransac(f["obj"], f["img"], f["intr"])
    t0 = time.perf_counter()
    for _ in range(20):
        r = ctx.pnp_ransac(f["obj"], f["img"], f["intr"])
    print(f"N={n} outliers={frac}: {1e3 * (time.perf_counter() - t0) / 20:.3f} ms per call, iters {r['iters']} inliers {r['n_inliers']}")
frames = [synth.make_pnp(1008 + 31 * b, 1000, 0.2) for b in range(64)]
pr = [(f["obj"], f["img"]) for f in frames]
ctx.pnp_ransac_batch(pr, frames[0]["intr"])
t0 = time.perf_counter(); ctx.pnp_ransac_batch(pr, frames[0]["intr"]); print(f"batch of 64: {1e3 * (time.perf_counter() - t0):.3f} ms")
ctx.close()
