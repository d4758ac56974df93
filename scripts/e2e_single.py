"""Development: latency of ONE LocalmapOptimization-sized call through the host-buffer C ABI
(the reference's call pattern: one window per keyframe) and of one FrameOptimization call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
ctx = U.Context(0)
w = synth.cfg1()
ctx.local_ba(w)
for rep in range(5):
    t0 = time.perf_counter(); ctx.local_ba(w); t1 = time.perf_counter()
    print(f"urmvo_local_ba (1 cfg1 window, host buffers): {1e3*(t1-t0):.3f} ms")
b = pack_ba_batch([w])
for rep in range(3):
    t0 = time.perf_counter(); plan = U.BAPlan(ctx, b); t1 = time.perf_counter()
    plan.run(); ctx.sync(); t2 = time.perf_counter()
    out = plan.download(); t3 = time.perf_counter()
    plan.close(); t4 = time.perf_counter()
    print(f"  create {1e3*(t1-t0):.3f}  run {1e3*(t2-t1):.3f}  download {1e3*(t3-t2):.3f}  destroy {1e3*(t4-t3):.3f} ms")
pb = synth.make_pose_batch(5, B=1, n_obs=1000)
ctx.pose_only_batch(pb)
for rep in range(3):
    t0 = time.perf_counter(); ctx.pose_only_batch(pb); t1 = time.perf_counter()
    print(f"urmvo_pose_only_batch (1 frame x 1000 matches): {1e3*(t1-t0):.3f} ms")
# the same cfg1 window through the device-resident map (values already in HBM; reset outside the timed region)
m = U.DeviceMap(ctx, w["intr"])
Nc, Np = w["poses"].shape[0], w["pts"].shape[0]
kf_ids = np.arange(Nc, dtype=np.int32) * 3 + 10; pt_ids = np.arange(Np, dtype=np.int32) * 7 + 1000
m.set_keyframes(kf_ids, w["poses"]); m.set_points(pt_ids, w["pts"])
order = np.argsort(w["obs_pt"], kind="stable")
m.add_observations(kf_ids[w["obs_cam"][order]], pt_ids[w["obs_pt"][order]], w["uv"][order])
for rep in range(5):
    m.set_keyframes(kf_ids, w["poses"]); m.set_points(pt_ids, w["pts"]); ctx.sync()
    t0 = time.perf_counter(); okf, opt, inl, st = m.local_ba(kf_ids, w["fixed"], pt_ids, max_obs=len(w["uv"])); t1 = time.perf_counter()
    print(f"urmvo_map_local_ba (same window, values resident in HBM): {1e3*(t1-t0):.3f} ms, iters {list(st.iters)}")
