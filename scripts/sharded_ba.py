"""Point-sharded BA over NCCL: run under torchrun (one rank per GPU) or alone.
  python scripts/sharded_ba.py [small|cfg5|<n_cams>,<n_pts>] [--check]
Every rank generates the same problem from the seed, keeps its share of the points, and the ranks
all-reduce the reduced camera system each trial.  --check compares with the CPU oracle on rank 0."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ur-mvo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist
import urmvo_b200 as U
from urmvo_b200 import synth

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
which = sys.argv[1] if len(sys.argv) > 1 else "small"
check = "--check" in sys.argv
tol = float([a for a in sys.argv if a.startswith("--tol=")][0][6:]) if any(a.startswith("--tol=") for a in sys.argv) else 0.0
mode = int([a for a in sys.argv if a.startswith("--mode=")][0][7:]) if any(a.startswith("--mode=") for a in sys.argv) else 0
if which == "small":
    prob = synth.make_ba(77, 40, 3000, 9.0, 14, 2, 0.02)
elif which == "cfg5":
    prob = synth.cfg5()
else:
    nc, npt = (int(v) for v in which.split(","))
    prob = synth.cfg5(n_cams=nc, n_pts=npt)
ctx = U.Context(lr)
if world > 1:
    uid = torch.from_numpy(U.nccl_unique_id() if rank == 0 else np.zeros(128, dtype=np.uint8)).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(rank, world, uid.cpu().numpy())
loc = U.shard_points(prob, rank, world)
cov = torch.from_numpy(U.ba_covisibility(loc).astype(np.int32)).cuda()
if world > 1:
    dist.all_reduce(cov, op=dist.ReduceOp.MAX)
cov = cov.cpu().numpy().astype(np.uint8)
plan = U.ShardedBAPlan(ctx, loc, covis=cov, opts=U.BAOptions(tol, 0, 0, 0, 0, 0, mode))
plan.run()  # warm-up
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
plan.run()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
poses, pts, inl, st = plan.download()
its = st.iters[0] + st.iters[1]
if rank == 0:
    print(f"[sharded BA x{world}] Nc={prob['poses'].shape[0]} Np={prob['pts'].shape[0]} No={prob['uv'].shape[0]} local No={loc['uv'].shape[0]} "
          f"S blocks={int(cov.sum())} tile={plan.phase_info()['tile_mode']}: {dt*1e3:.2f} ms, {its} LM it ({st.trials[0]+st.trials[1]} trials, {st.pcg_iters[0]+st.pcg_iters[1]} PCG it) -> {its/dt:.1f} it/s; chi {list(st.chi2_final)}", flush=True)
if rank == 0:
    print("   ", plan.phase_info(), flush=True)
if check:
    import pyoracle as po
    if rank == 0:
        t0 = time.perf_counter(); op, ox, oi, os_ = po.local_ba(prob); to = time.perf_counter() - t0
        p0, p1 = loc["point_range"]; o0, o1 = loc["obs_range"]
        rel = abs(st.chi2_final[1] - os_.chi2_final[1]) / abs(os_.chi2_final[1])
        print(f"   oracle {to*1e3:.0f} ms ({(os_.iters[0]+os_.iters[1])/to:.2f} it/s); rel cost diff {rel:.2e}; pose maxdiff {np.abs(poses-op).max():.2e}; "
              f"pts maxdiff {np.abs(pts-ox[p0:p1]).max():.2e}; inlier mismatches {(inl!=oi[o0:o1]).sum()}; iters {list(st.iters)} vs {list(os_.iters)[:2]}", flush=True)
plan.close(); ctx.close()
if world > 1: dist.destroy_process_group()
