"""Development: tile-mode timing of the large problems (cfg4 window, cfg5 on one rank) with the phase
times of the first trial and the final cost (compare runs with each other / with the oracle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch

which = sys.argv[1:] or ["cfg4", "cfg5"]
check = "--oracle" in which
ctx = U.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
for name in [w for w in which if not w.startswith("--")]:
    prob = synth.cfg4() if name == "cfg4" else synth.cfg5()
    plan = U.BAPlan(ctx, pack_ba_batch([prob]))
    plan.run(); ctx.sync()
    ms = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); plan.run(); b.record(stream); ctx.sync()
        ms.append(a.elapsed_time(b))
    poses, pts, inl, st = plan.download()
    st = st[0]
    info = plan.phase_info()
    nb = max(1, info["host_syncs"])
    ph = {k: round(v / nb, 4) for k, v in info["phase_ms_first_trial_of_each_batch"].items()}
    print(f"[{name}] No={prob['uv'].shape[0]} {min(ms):.3f} ms  iters {list(st.iters)} trials {list(st.trials)} "
          f"chi2 {st.chi2_final[1]!r} tile={info['tile_mode']} bw={info['half_bandwidth_blocks']} phases {ph}", flush=True)
    if check:
        import pyoracle as po
        o = po.local_ba(prob)
        print(f"   oracle: rel cost diff {abs(st.chi2_final[1] - o[3].chi2_final[1]) / abs(o[3].chi2_final[1]):.2e} "
              f"pose maxdiff {np.abs(poses - o[0]).max():.2e} flags differ {(inl != o[2]).sum()}")
    plan.close()
ctx.close()
