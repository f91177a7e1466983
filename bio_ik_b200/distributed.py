"""Query sharding across GPUs: IK queries are independent, so rank r of W solves the contiguous
slice [r*ceil(B/W), ...) with no data-path collective; ONE all-gather of the per-rank result slab
[S][n_vars + 3] (solution | fitness | success | steps) per solve round returns every answer to every rank.
Backend: NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous slice of a B-query batch owned by `rank` (the last ranks may be shorter or empty)."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def pack_slab(res, n_vars, rows):
    """solutions | fitness | success | steps as one float64 slab, zero-padded to `rows`."""
    slab = np.zeros((rows, n_vars + 3))
    k = len(res["fitness"])
    slab[:k, :n_vars] = res["solutions"]
    slab[:k, n_vars] = res["fitness"]
    slab[:k, n_vars + 1] = res["success"]
    slab[:k, n_vars + 2] = res["steps"]
    return slab


def unpack_slab(slab, n_vars, B):
    slab = slab[:B]
    return dict(solutions=np.ascontiguousarray(slab[:, :n_vars]), fitness=np.ascontiguousarray(slab[:, n_vars]),
                success=slab[:, n_vars + 1].astype(np.int32), steps=slab[:, n_vars + 2].astype(np.int32))


def solve_sharded(solve_fn, n_vars, goal_params, seeds, rng_seeds, steps, early_exit=False, device=None):
    """Every rank passes the FULL batch; each solves its slice with `solve_fn(goal_params, seeds, rng_seeds, steps,
    early_exit) -> dict` and all ranks return the full result after one all-gather."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    B = len(rng_seeds)
    per = -(-B // world)
    lo, hi = shard_range(B, rank, world)
    if hi > lo:
        gp = None if goal_params is None else goal_params[lo:hi]
        res = solve_fn(gp, seeds[lo:hi], rng_seeds[lo:hi], steps, early_exit)
    else:
        res = dict(solutions=np.zeros((0, n_vars)), fitness=np.zeros(0), success=np.zeros(0, dtype=np.int32), steps=np.zeros(0, dtype=np.int32))
    slab = torch.from_numpy(pack_slab(res, n_vars, per))
    if device is not None:
        slab = slab.to(device)
    gathered = torch.empty((world * per, n_vars + 3), dtype=torch.float64, device=slab.device)
    dist.all_gather_into_tensor(gathered, slab)
    # ranks own consecutive `per`-row blocks, so the gathered slab is already in query order
    return unpack_slab(gathered.cpu().numpy(), n_vars, B)
