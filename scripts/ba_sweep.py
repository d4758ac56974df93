"""Development: sweep cluster size / threads for the cfg1 window batch. Usage: ba_sweep.py [windows]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python"))
import numpy as np, torch
import urmvo_b200 as U
from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(8, 256), (4, 256), (2, 256), (1, 256), (16, 256), (8, 128), (4, 128)]
ctx = U.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
distinct = [synth.make_ba(1001 + i, 10, 2000, 7.7, 10, 3, 0.05) for i in range(min(37, nw))]
batch = pack_ba_batch([distinct[i % len(distinct)] for i in range(nw)])
for cfg in cfgs:
    cs, th = cfg[0], cfg[1]
    force = cfg[2] if len(cfg) > 2 else 0
    plan = U.BAPlan(ctx, batch, opts=U.BAOptions(0, 0, cs, th, force))
    plan.run(); ctx.sync(); ctx.ba_timing()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(3): plan.run()
    b.record(stream); ctx.sync()
    ms = a.elapsed_time(b) / 3
    st = plan.download()[3]
    its = sum(s.iters[0] + s.iters[1] for s in st); tr = sum(s.trials[0] + s.trials[1] for s in st); pcg = sum(s.pcg_iters[0] + s.pcg_iters[1] for s in st)
    print(f"cluster {cs:2d} threads {th:3d} force {force}: {ms:8.3f} ms/step  {its/ms*1e3:10.0f} LM it/s  trials {tr} pcg_iters {pcg}", flush=True)
    t = ctx.ba_timing(); tt = sum(t) or 1
    print("      phase cycles/run (window 0): " + " ".join(f"{n}={v/3/1e3:.0f}k({v/tt*100:.0f}%)" for n, v in zip(["lin0","lin","red","pcg","cam","back","red2","-"], t)), flush=True)
    plan.close()
