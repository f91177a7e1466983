"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the strict
oracle): the CPU leg guards the oracle against regressions, the GPU leg checks the CUDA path against
the same files through the C ABI."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

import oracle_lib  # noqa: E402
from bio_ik_b200 import workloads  # noqa: E402

KEYS = ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps")


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_reproduces_golden(oracle, name):
    w, cfg, res = make_golden.run_case(oracle, name)
    g = load(name)
    assert np.array_equal(w.goal_params, g["goal_params"]) and np.array_equal(w.seeds, g["seeds"])
    for k in KEYS:
        assert np.array_equal(res[k], g[k]), k


class _NoFK:
    """stands in for the oracle where only the structure of a custom workload is needed"""

    def fk(self, rm, pr, v):
        return np.zeros((len(v), len(pr.tip_link_indices), 7))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_gpu_matches_golden(name):
    import gpu_util
    cfgname, B, pop, mode, gens, steps = make_golden.CASES[name]
    g = load(name)
    if cfgname in workloads.CONFIGS:
        f, _ = workloads.CONFIGS[cfgname]
        w = f() if cfgname == "cfg1" else f(B)
    else:  # robot and problem only (the inputs come from the file): the oracle is not needed to build them
        w = make_golden.custom_workload(_NoFK(), cfgname, B)
    solver = gpu_util.make_solver(w, pop, mode, gens)
    got = solver.trace(g["goal_params"], g["seeds"], g["rng_seeds"], steps)
    gpu_util.assert_bit_equal(got, g, what=name)
    res = solver.solve_batch(g["goal_params"], g["seeds"], g["rng_seeds"], steps)
    gpu_util.assert_bit_equal(res, g, keys=("solutions", "fitness", "success", "steps"), what=name)
