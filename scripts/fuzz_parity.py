"""Randomised parity sweep on a B200 (development / evidence, results in profiles/fuzz_r02f.txt): many seeded problems of random shape through the C ABI against the CPU restatement
(and the real cv2 where it is importable).  Prints one summary line per path; exit code 1 on any violation of the
parity bar of tests/test_gpu_parity.py.  Usage: fuzz_parity.py [n_problems]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
import pyoracle as po

N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(20251018)
ctx = U.Context(0)
bad = 0

# ---- local BA windows: random camera / point counts, outlier rates, start noise, fixed cameras
w = dict(n=0, cost=0.0, pose=0.0, flags=0, iters=0)
for k in range(N):
    nc = int(rng.integers(3, 13)); npt = int(rng.integers(40, 700)); nf = int(rng.integers(2, max(3, nc - 1)))  # >= 2 fixed keyframes: with one, the monocular scale is a free gauge mode
    p = synth.small_ba(seed=10000 + k, n_cams=nc, n_pts=npt, n_fixed=nf, outlier_frac=float(rng.uniform(0, 0.2)),
                       rot_sigma_deg=float(rng.uniform(0.1, 3.0)), trans_sigma=float(rng.uniform(0.005, 0.2)),
                       pt_sigma=float(rng.uniform(0.01, 0.4)))
    g = ctx.local_ba(p); o = po.local_ba(p)
    w["n"] += 1
    w["cost"] = max(w["cost"], abs(g[3].chi2_final[1] - o[3].chi2_final[1]) / max(abs(o[3].chi2_final[1]), 1e-300))
    w["pose"] = max(w["pose"], float(np.abs(g[0] - o[0]).max()))
    w["flags"] += int((g[2] != o[2]).sum())
    # at a converged minimum the last accept / reject decisions are rounding noise: an iteration count only counts as a
    # mismatch when the cost differs as well
    rel = abs(g[3].chi2_final[1] - o[3].chi2_final[1]) / max(abs(o[3].chi2_final[1]), 1e-300)
    w["iters"] += int(list(g[3].iters)[:2] != list(o[3].iters)[:2] and rel > 1e-9)
print("local BA", w)
bad += w["cost"] > 1e-6 or w["pose"] > 1e-5 or w["flags"] > 0 or w["iters"] > 0

# ---- stereo / several camera models (3-row phases), large windows in tile mode with either band solver
m = dict(n=0, cost=0.0, pose=0.0, flags=0)
for k in range(max(10, N // 8)):
    base = synth.add_stereo(synth.small_ba(seed=50000 + k, n_cams=int(rng.integers(4, 11)), n_pts=int(rng.integers(60, 500)),
                                           n_fixed=2, outlier_frac=float(rng.uniform(0, 0.15))), 50500 + k,
                            stereo_frac=float(rng.uniform(0, 1)))
    if k % 2:
        p = synth.add_camera_models(base, 51000 + k, n_models=int(rng.integers(2, 6)))
        g = ctx.local_ba_multicam(p, 10.0, 75.0); o = po.local_ba_multicam(p, 10.0, 75.0)
    else:
        g = ctx.local_ba_stereo(base, 10.0, 75.0); o = po.local_ba_stereo(base, 10.0, 75.0)
    m["n"] += 1
    m["cost"] = max(m["cost"], abs(g[3].chi2_final[1] - o[3].chi2_final[1]) / abs(o[3].chi2_final[1]))
    m["pose"] = max(m["pose"], float(np.abs(g[0] - o[0]).max()))
    m["flags"] += int((g[2] != o[2]).sum())
print("stereo / camera models", m)
bad += m["cost"] > 1e-6 or m["pose"] > 1e-5 or m["flags"] > 0
L = dict(n=0, cost=0.0, pose=0.0, flags=0, tile=0)
for k in range(max(6, N // 40)):
    ncam = int(rng.integers(30, 420)); span = int(rng.integers(5, 17))
    p = synth.make_ba(60000 + k, ncam, int(ncam * rng.integers(20, 45)), 0.6 * span, span, 2, 0.02)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p), opts=U.BAOptions(0, 0, 0, 0, 0, 0, 0, 1 + k % 2))
    plan.run(); L["tile"] += int(plan.phase_info()["tile_mode"]); g = plan.download(); plan.close()
    o = po.local_ba(p)
    L["n"] += 1
    L["cost"] = max(L["cost"], abs(g[3].chi2_final[1] - o[3].chi2_final[1]) / abs(o[3].chi2_final[1]))
    L["pose"] = max(L["pose"], float(np.abs(g[0] - o[0]).max()))
    L["flags"] += int((g[2] != o[2]).sum())
print("large windows (tile mode, band solve / cyclic reduction alternating)", L)
bad += L["cost"] > 1e-6 or L["pose"] > 1e-5 or L["flags"] > 0

# ---- pose-only frames
b = synth.make_pose_batch(777, B=N, n_obs=int(rng.integers(60, 1200)), outlier_frac=0.15, rot_deg=3.0, trans=0.2)
gp, gi, gn = ctx.pose_only_batch(b); op, oi, on = po.pose_only_batch(b)
r = dict(n=N, pose=float(np.abs(gp - op).max()), flags=int((gi != oi).sum()), counts=int((gn != on).sum()))
print("pose-only", r)
bad += r["pose"] > 1e-5 or r["flags"] > 0 or r["counts"] > 0

# ---- per-frame fundamental matrix (all three OpenCV branches), cv2 beside the restatement when importable
try:
    import cv2
except Exception:
    cv2 = None
f = dict(n=0, mask_vs_port=0, mask_vs_cv2=0, cv2_cases=0)
for k in range(N):
    n = int(rng.choice([7, 14, 15, 16, 40, 120, 500, 1500]))
    p0, p1 = synth.make_fm(20000 + k, n, float(rng.uniform(0.3, 0.95)), float(rng.uniform(0.2, 1.2)), float(rng.uniform(1, 8)))
    g = ctx.fm_ransac(p0, p1); o = po.find_fundamental(p0, p1)
    f["n"] += 1
    f["mask_vs_port"] += int(not np.array_equal(g["mask"], o["mask"]))
    if cv2 is not None:
        _, m = cv2.findFundamentalMat(p0, p1, cv2.FM_RANSAC, 3, 0.99)
        f["cv2_cases"] += 1
        f["mask_vs_cv2"] += int(not np.array_equal(g["mask"], m.ravel().astype(np.uint8)))
print("fundamental", f)
bad += f["mask_vs_port"] > 0 or f["mask_vs_cv2"] > 0

# ---- SolvePnPWithCV
q = dict(n=0, mask=0, pose=0.0)
for k in range(N):
    pp = synth.make_pnp(30000 + k, int(rng.integers(8, 1500)), float(rng.uniform(0, 0.5)))
    g = ctx.pnp_ransac(pp["obj"], pp["img"], pp["intr"]); o = po.pnp_ransac(pp["obj"], pp["img"], pp["intr"])
    q["n"] += 1
    q["mask"] += int(not np.array_equal(g["mask"], o["mask"]) or g["iters"] != o["iters"])
    if g["found"] and o["found"]:
        q["pose"] = max(q["pose"], float(max(np.abs(g["R"] - o["R"]).max(), np.abs(g["t"] - o["t"]).max())))
print("pnp", q)
bad += q["mask"] > 0 or q["pose"] > 1e-6

# ---- two-view reconstruct (the reference's 200 iterations), bit-exact
t = dict(n=0, ok=0, masks=0, T=0)
for k in range(max(10, N // 4)):
    nk = int(rng.integers(100, 1000))
    tv = synth.make_two_view(40000 + k, n_keys=nk, inlier_frac=float(rng.uniform(0.4, 0.9)), planar=bool(k % 3 == 0))
    tv["sets"] = synth.draw_sets(nk, 200, k)
    try:
        g = ctx.two_view(tv); o = po.two_view(tv)
    except Exception as e:
        print("two-view setup:", e); break
    t["n"] += 1
    t["ok"] += int(g["ok"] != o["ok"])
    t["masks"] += int(not (np.array_equal(g["mask_F"], o["mask_F"]) and np.array_equal(g["mask_H"], o["mask_H"])))
    t["T"] += int(not np.array_equal(g["T21"], o["T21"]))
print("two-view", t)
bad += t["ok"] > 0 or t["masks"] > 0 or t["T"] > 0
ctx.close()
print("violations:", int(bad))
sys.exit(1 if bad else 0)
