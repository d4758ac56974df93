"""Development check of the tile-mode large-BA path (csrc/ba_large.cu) on one GPU:
parity against the CPU oracle (small sharded problem, cfg4) and against the round-1 atomic/PCG path
(cfg5), plus timings with the phase breakdown.   python scripts/dev/large_check.py [small|cfg4|cfg5|all]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
import pyoracle as po

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = U.Context(0)


def report(tag, g, o):
    gp, gx, gi, gs = g
    op, ox, oi, os_ = o
    rel = abs(gs.chi2_final[1] - os_.chi2_final[1]) / abs(os_.chi2_final[1])
    print(f"[{tag}] rel cost diff {rel:.2e}; pose maxdiff {np.abs(gp-op).max():.2e}; pts maxdiff {np.abs(gx-ox).max():.2e}; "
          f"inlier mismatches {(gi!=oi).sum()}; iters {list(gs.iters)} vs {list(os_.iters)[:2]}; trials {list(gs.trials)} vs {sum(r[3] for r in os_.rows())}; "
          f"chi {list(gs.chi2_final)} vs {list(os_.chi2_final)[:2]}", flush=True)


def timed(plan, reps=3):
    plan.run(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        plan.run()
    ctx.sync()
    return (time.perf_counter() - t0) / reps


if which in ("small", "all"):
    p = synth.make_ba(77, 40, 1500, 8.0, 14, 2, 0.02)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p))
    plan.run()
    print("small phase info", plan.phase_info(), flush=True)
    report("small sharded x1 tile", plan.download(), po.local_ba(p))
    plan.close()
    # noisy start: rejected trials
    p = synth.make_ba(78, 30, 1200, 8.0, 12, 2, 0.1, rot_sigma_deg=6.0, trans_sigma=0.4, pt_sigma=0.6)
    plan = U.ShardedBAPlan(ctx, U.shard_points(p, 0, 1), covis=U.ba_covisibility(p))
    plan.run()
    print("noisy phase info", plan.phase_info(), flush=True)
    report("noisy sharded x1 tile", plan.download(), po.local_ba(p))
    plan.close()

if which in ("cfg4", "all"):
    p = synth.cfg4()
    t0 = time.perf_counter(); o = po.local_ba(p); to = time.perf_counter() - t0
    for mode in (0, 1):
        plan = U.BAPlan(ctx, pack_ba_batch([p]), opts=U.BAOptions(0, 0, 0, 0, 0, 0, mode))
        dt = timed(plan)
        g = plan.download()
        info = plan.phase_info()
        report(f"cfg4 large_mode={mode} {dt*1e3:.2f} ms ({(g[3][0].iters[0]+g[3][0].iters[1])/dt:.0f} it/s; oracle {to*1e3:.0f} ms)", (g[0], g[1], g[2], g[3][0]), o)
        print("   ", info, flush=True)
        plan.close()

if which in ("cfg5", "all"):
    p = synth.cfg5()
    res = {}
    for mode in (0, 1):
        loc = U.shard_points(p, 0, 1)
        plan = U.ShardedBAPlan(ctx, loc, covis=U.ba_covisibility(loc), opts=U.BAOptions(0, 0, 0, 0, 0, 0, mode))
        ctx.lg_timing(True)
        dt = timed(plan, 2)
        lt = ctx.lg_timing(True)
        if mode == 0 and lt[6]:
            print("    band solve cycles/step: panel seg %.0f, trailing seg %.0f (thread 0: update %.0f, fetch+take %.0f; diagonal warp factor %.0f), backsub %.0f; whole kernel %.0f cyc/step, steps %d"
                  % (lt[1] / lt[6], lt[2] / lt[6], lt[0] / lt[6], lt[7] / lt[6], lt[5] / lt[6], lt[3] / lt[6], lt[4] / lt[6], lt[6]), flush=True)
        g = plan.download()
        res[mode] = g
        st = g[3]
        print(f"[cfg5 large_mode={mode}] {dt*1e3:.2f} ms, iters {list(st.iters)} trials {list(st.trials)} pcg {list(st.pcg_iters)} chi {list(st.chi2_final)}", flush=True)
        print("   ", plan.phase_info(), flush=True)
        plan.close()
    a, b = res[0], res[1]
    rel = abs(a[3].chi2_final[1] - b[3].chi2_final[1]) / abs(b[3].chi2_final[1])
    print(f"[cfg5 tile vs round-1 path] rel cost diff {rel:.2e}; pose maxdiff {np.abs(a[0]-b[0]).max():.2e}; pts maxdiff {np.abs(a[1]-b[1]).max():.2e}; inlier mismatches {(a[2]!=b[2]).sum()}")
ctx.close()
