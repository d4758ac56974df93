"""CPU: host-side logic — batch packing, sharding, and the N>1 paths under gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from urmvo_b200 import synth
from urmvo_b200.capi import pack_ba_batch
from urmvo_b200.dist import shard_range, shard_windows, merge_best_hypothesis, max_over_ranks, sum_over_ranks


def test_pack_ba_batch_offsets_and_local_indices():
    probs = [synth.small_ba(seed=s, n_cams=4 + s, n_pts=30 + 5 * s) for s in range(3)]
    b = pack_ba_batch(probs)
    assert b["cam_off"].tolist() == [0, 4, 9, 15]
    assert b["pt_off"][-1] == sum(p["pts"].shape[0] for p in probs)
    for w, p in enumerate(probs):
        s = slice(b["obs_off"][w], b["obs_off"][w + 1])
        assert np.array_equal(b["obs_cam"][s], p["obs_cam"]) and b["obs_cam"][s].max() < p["poses"].shape[0]
        assert np.array_equal(b["uv"][s], p["uv"])
    assert b["poses"].flags["C_CONTIGUOUS"] and b["obs_pt"].dtype == np.int32


@pytest.mark.parametrize("n,world", [(148, 8), (10, 4), (3, 8), (0, 2), (8192, 3)])
def test_shard_range_partitions_exactly(n, world):
    cover = []
    sizes = []
    for r in range(world):
        a, b = shard_range(n, r, world)
        cover += list(range(a, b))
        sizes.append(b - a)
    assert cover == list(range(n))
    assert max(sizes) - min(sizes) <= 1


def test_shard_windows_slices():
    probs = [synth.small_ba(seed=s, n_cams=4, n_pts=20) for s in range(5)]
    b = pack_ba_batch(probs)
    (w0, w1), cs, ps, os_ = shard_windows((b["cam_off"], b["pt_off"], b["obs_off"]), 1, 2)
    assert (w0, w1) == (3, 5)
    assert os_.start == b["obs_off"][3] and os_.stop == b["obs_off"][5]
    assert cs.stop - cs.start == 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, scores, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # hypothesis-sharded arg-max: each rank scores its contiguous slice
        a, b = shard_range(len(scores), rank, world)
        local = np.asarray(scores[a:b], dtype=np.float32)
        li = int(np.argmax(local)) if len(local) and local.max() > 0 else -1
        ls = float(local[li]) if li >= 0 else 0.0
        best = merge_best_hypothesis(ls, li, a, dist)
        # timing / unit aggregation as bench.py does it
        t = max_over_ranks(1.0 + rank, dist)
        u = sum_over_ranks(10.0 * (rank + 1), dist)
        q.put((rank, best, t, u))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scores", [
    [0.0, 3.0, 7.5, 7.5, 1.0, 7.5, 2.0, 0.0],      # tie across ranks: earliest index (2) wins
    [0.0, 0.0, 0.0, 0.0],                          # nothing scores > 0
    [1.0, 0.5, 0.25, 9.0, 9.0],                    # tie inside the last rank
])
def test_gloo_world2_argmax_and_reductions(scores):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, scores, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = np.asarray(scores, dtype=np.float32)
    want_idx = int(np.argmax(s)) if s.max() > 0 else -1
    for rank, best, t, u in res:
        assert best[1] == want_idx
        if want_idx >= 0:
            assert best[0] == float(s[want_idx])
            assert best[2] == (0 if want_idx < shard_range(len(scores), 0, world)[1] else 1)
        assert t == 2.0 and u == 30.0


def test_reference_arm_non_zero_ranks_do_no_work():
    """bench.py --impl reference under torchrun: only rank 0 runs and prints."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_points_covers_every_point_and_observation_once(world):
    from urmvo_b200.capi import shard_points
    prob = synth.make_ba(5, 30, 800, 6.0, 10, 2, 0.02)
    seen_p, seen_o, sizes = [], [], []
    for r in range(world):
        loc = shard_points(prob, r, world)
        p0, p1 = loc["point_range"]; o0, o1 = loc["obs_range"]
        seen_p += list(range(p0, p1)); seen_o += list(range(o0, o1)); sizes.append(o1 - o0)
        assert loc["poses"].shape == prob["poses"].shape            # cameras replicated
        assert loc["obs_pt"].min(initial=0) >= 0 and loc["obs_pt"].max(initial=0) < max(p1 - p0, 1)
        assert np.array_equal(prob["obs_pt"][o0:o1] - p0, loc["obs_pt"])
        assert np.array_equal(prob["uv"][o0:o1], loc["uv"])
    assert seen_p == list(range(prob["pts"].shape[0])) and seen_o == list(range(prob["uv"].shape[0]))
    assert max(sizes) - min(sizes) <= 2 * np.bincount(prob["obs_pt"]).max()   # balanced by observations


def test_covisibility_union_over_shards_equals_global():
    from urmvo_b200.capi import shard_points, ba_covisibility
    prob = synth.make_ba(6, 24, 500, 5.0, 8, 2, 0.0)
    full = ba_covisibility(prob)
    acc = np.zeros_like(full)
    for r in range(3):
        acc |= ba_covisibility(shard_points(prob, r, 3))
    assert np.array_equal(acc, full)
    assert np.array_equal(full, np.triu(full)) and full.diagonal().all()


def test_points_numbered_along_the_trajectory_is_the_same_problem(oracle):
    """synth.sort_points_by_first_camera only renumbers the points (bench.py's weak-scaling sharded BA uses it so that a
    rank's contiguous range of points is local to a stretch of cameras): same minimum, point-major observations."""
    import numpy as np
    from urmvo_b200 import synth
    p = synth.small_ba(seed=3, n_pts=300)
    q = synth.sort_points_by_first_camera(p)
    assert np.all(np.diff(q["obs_pt"]) >= 0) and sorted(map(tuple, np.c_[q["uv"], q["obs_cam"]])) == sorted(map(tuple, np.c_[p["uv"], p["obs_cam"]]))
    first = np.full(q["pts"].shape[0], 10 ** 9)
    np.minimum.at(first, q["obs_pt"], q["obs_cam"])
    assert np.all(np.diff(first) >= 0)
    a, b = oracle.local_ba(p), oracle.local_ba(q)
    assert abs(a[3].chi2_final[1] - b[3].chi2_final[1]) <= 1e-9 * a[3].chi2_final[1]
    assert np.abs(a[0] - b[0]).max() < 1e-8


def test_trajectory_ordered_shards_are_camera_local():
    """bench.py's weak-scaling sharded BA: with points numbered along the trajectory a rank's contiguous range of points
    touches only its own stretch of cameras (plus the co-visibility halo), the union of the shards is the problem."""
    import numpy as np
    import urmvo_b200 as U
    from urmvo_b200 import synth
    p = synth.sort_points_by_first_camera(synth.make_ba(77, 200, 4000, 6.0, 12, 2, 0.01))
    world = 4
    shards = [U.shard_points(p, r, world) for r in range(world)]
    assert sum(s["uv"].shape[0] for s in shards) == p["uv"].shape[0]
    assert sum(s["pts"].shape[0] for s in shards) == p["pts"].shape[0]
    lo = [int(s["obs_cam"].min()) for s in shards]
    hi = [int(s["obs_cam"].max()) for s in shards]
    assert lo == sorted(lo) and hi == sorted(hi)
    for r in range(world):
        assert hi[r] - lo[r] <= 200 // world + 2 * 12 + 12  # its quarter of the trajectory + the halo of the span
    # the generator's own (random) numbering spreads every shard over the whole trajectory
    q = synth.make_ba(77, 200, 4000, 6.0, 12, 2, 0.01)
    s0 = U.shard_points(q, 0, world)
    assert int(s0["obs_cam"].max()) - int(s0["obs_cam"].min()) > 150
