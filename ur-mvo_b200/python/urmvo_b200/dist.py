"""Host-side sharding of the embarrassingly parallel parts of the path over one-process-per-GPU ranks
(SURVEY.md §8e): independent BA windows, pose-only frames and RANSAC hypotheses need no data-path
collective; only the winner selection of a hypothesis-sharded RANSAC exchanges one (score, index)
pair per rank.  torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of `n_items` units for `rank` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_windows(batch_offsets, rank, world):
    """Window range of this rank plus the slices of the concatenated cam / point / observation arrays."""
    cam_off, pt_off, obs_off = batch_offsets
    w0, w1 = shard_range(len(cam_off) - 1, rank, world)
    return (w0, w1), slice(cam_off[w0], cam_off[w1]), slice(pt_off[w0], pt_off[w1]), slice(obs_off[w0], obs_off[w1])


def merge_best_hypothesis(local_score, local_index, index_offset, dist=None):
    """Deterministic arg-max across ranks with the reference's tie rule (earliest hypothesis wins,
    src/epipolar_geometry.cc:153-157): all-gather (score, global index) and pick max score, then min
    index.  Returns (score, global_index, owner_rank); index -1 means no hypothesis scored > 0."""
    import torch
    gi = -1 if local_index < 0 else int(local_index) + int(index_offset)
    mine = torch.tensor([float(local_score), float(gi)], dtype=torch.float64)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_score), gi, 0
    world = dist.get_world_size()
    if dist.get_backend() == "nccl":
        mine = mine.cuda()
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    best = (0.0, -1, 0)
    for r, t in enumerate(out):
        s, i = float(t[0]), int(t[1])
        if i < 0:
            continue
        if s > best[0] or (s == best[0] and best[1] >= 0 and i < best[1]) or (best[1] < 0 and s > 0):
            best = (s, i, r)
    return best


def max_over_ranks(value, dist=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist=None):
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
