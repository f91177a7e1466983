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


class DeviceShardedSolver:
    """Device-resident form of the same scheme for throughput runs (bench.py): every rank owns a shard of B queries whose
    inputs already live in its HBM; one pass = bioik_solve_batch_device on the shard, bioik_pack_results_device into the
    [B][n_vars + 3] slab and - world > 1 - ONE all_gather_into_tensor of the slabs (NCCL over NVLink), all enqueued on `stream`."""

    def __init__(self, solver, B, device, stream):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.solver, self.B, self.stream = solver, B, stream
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        n_vars = solver.robot_model.n_vars
        self.sol = torch.empty((B, n_vars), dtype=torch.float64, device=device)
        self.fit = torch.empty(B, dtype=torch.float64, device=device)
        self.succ = torch.empty(B, dtype=torch.int32, device=device)
        self.steps = torch.empty(B, dtype=torch.int32, device=device)
        self.slab = torch.empty((B, n_vars + 3), dtype=torch.float64, device=device)
        self.gathered = torch.empty((self.world * B, n_vars + 3), dtype=torch.float64, device=device) if self.world > 1 else self.slab

    def solve(self, d_goal_params, d_seeds, d_rng_seeds, steps, early_exit=False):
        """d_*: torch tensors on this rank's device.  Returns the gathered slab [world * B][n_vars + 3] (not synchronised)."""
        s, st = self.solver, self.stream.cuda_stream
        s.solve_batch_device(self.B, d_goal_params.data_ptr(), d_seeds.data_ptr(), d_rng_seeds.data_ptr(), steps, early_exit, self.sol.data_ptr(), self.fit.data_ptr(), self.succ.data_ptr(), self.steps.data_ptr(),
                             stream=st)
        s.pack_results_device(self.B, self.sol.data_ptr(), self.fit.data_ptr(), self.succ.data_ptr(), self.steps.data_ptr(), self.slab.data_ptr(), stream=st)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered, self.slab)
        return self.gathered

    def solve_host(self, goal_params, seeds, rng_seeds, steps, out, early_exit=False):
        """The host-buffer path of one rank (bioik_solve_batch: H2D + solve + D2H), then the gather of the result slabs: the
        rank's results go back to the device slab, one all-gather, and every rank reads the full slab into `out["gathered"]`
        (pinned).  world == 1: just the host call."""
        res = self.solver.solve_batch(goal_params, seeds, rng_seeds, steps, early_exit=early_exit, out=out)
        if self.world > 1:
            torch = self.torch
            n_vars = self.solver.robot_model.n_vars
            host_slab = out["slab"]
            host_slab[:, :n_vars] = torch.from_numpy(res["solutions"])
            host_slab[:, n_vars] = torch.from_numpy(res["fitness"])
            host_slab[:, n_vars + 1] = torch.from_numpy(res["success"]).to(torch.float64)
            host_slab[:, n_vars + 2] = torch.from_numpy(res["steps"]).to(torch.float64)
            with torch.cuda.stream(self.stream):
                self.slab.copy_(host_slab, non_blocking=True)
                self.dist.all_gather_into_tensor(self.gathered, self.slab)
                out["gathered"].copy_(self.gathered, non_blocking=True)
            self.stream.synchronize()
        return res
