"""The N>1 path on CPU: world_size-2 gloo processes shard a batch, each solves its slice (with the oracle standing in
for the GPU solve — the sharding / gather logic is compute-agnostic) and one all-gather returns the full result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import oracle_lib
    from bio_ik_b200 import distributed, workloads
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = oracle_lib.Oracle()
    w = workloads.make("cfg2", lambda rm, pr, v: o.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=18)

    def solve_fn(gp, seeds, rs, steps, early_exit):
        return o.solve(w.robot, w.problem, cfg, gp, seeds, rs, steps, early_exit=early_exit, nthreads=2)

    res = distributed.solve_sharded(solve_fn, w.robot.n_vars, w.goal_params, w.seeds, w.rng_seeds, 6)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 11), (2, 4), (3, 2)])
def test_sharded_solve_equals_single_process(tmp_path, oracle, world, B):
    from bio_ik_b200 import distributed, workloads
    import oracle_lib
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    ref = oracle.solve(w.robot, w.problem, oracle_lib.make_cfg(population=18), w.goal_params, w.seeds, w.rng_seeds, 6)
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        for k in ("solutions", "fitness", "success", "steps"):
            assert np.array_equal(got[k], ref[k]), (r, k)
    # slices tile the batch
    cover = []
    for r in range(world):
        lo, hi = distributed.shard_range(B, r, world)
        cover += list(range(lo, hi))
    assert cover == list(range(B))
