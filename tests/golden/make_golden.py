"""Generates the committed golden vectors under tests/golden/ from the IEEE-strict CPU oracle
(oracle/liboracle_strict.so) in THIS container.  The reference itself cannot be built here
(SURVEY.md §8(c)), so the vectors pin the oracle restatement, not the upstream binary.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib  # noqa: E402
from bio_ik_b200 import workloads  # noqa: E402

# name -> (config, batch, population, mode, generations, steps)
CASES = {
    "cfg1_pop64_25steps": ("cfg1", 1, 64, "q", 8, 25),
    "cfg2_pop18_12steps": ("cfg2", 6, 18, "q", 8, 12),
    "cfg2_pop128_25steps": ("cfg2", 4, 128, "q", 8, 25),
    "cfg2_bio2_pop18": ("cfg2", 4, 18, 0, 16, 6),
    "cfg2_memetic_l_pop18": ("cfg2", 4, 18, "l", 8, 6),
    "cfg3_pop128_6steps": ("cfg3", 3, 128, "q", 8, 6),
    "cfg4_pop128_5steps": ("cfg4", 3, 128, "q", 8, 5),
    "cfg5_pop128_5steps": ("cfg5", 3, 128, "q", 8, 5),
}


def run_case(oracle, name):
    cfgname, B, pop, mode, gens, steps = CASES[name]
    w = workloads.make(cfgname, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=pop, memetic=mode, generations=gens)
    res = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps)
    return w, cfg, res


if __name__ == "__main__":
    o = oracle_lib.Oracle()
    for name in CASES:
        w, cfg, res = run_case(o, name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), goal_params=w.goal_params, seeds=w.seeds, rng_seeds=w.rng_seeds, targets=w.targets, **res)
        print(name, "success", res["success"].tolist(), "fitness", res["fitness"].max())
