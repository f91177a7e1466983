#!/usr/bin/env python
"""What "results within 1e-5 of the reference" means for a chaotic solver.

Runs the BASELINE cfg2 problem (PR2-like right arm, one PoseGoal, pop 128) on the CPU with
  ref_strict   the reference's own sources (oracle/_ref/libbioik_ref_strict.so: -O2, IEEE, libm sin/cos) - "the reference as shipped,
               minus -ffast-math"
  ref_fast     the same sources with the reference's Release flags (-O3 -ffast-math ..., CMakeLists.txt:85-88)
  contract     the arithmetic contract of the product (oracle default flags = what the GPU computes, bit for bit): det_sincos
and reports, after 1 step() and after 25 steps, on how many queries the solutions agree bit for bit / within 1e-5, and the
success rate and median fitness of each.  Usage: python profiles/tolerance_study.py [n_queries] > profiles/r02_tolerance_study.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from bio_ik_b200 import workloads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    oracle = oracle_lib.Oracle("strict")
    strict, fast = oracle_lib.Reference("strict"), oracle_lib.Reference("fast")
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=n)
    cfg = oracle_lib.make_cfg(population=128)
    robot, gp = strict.effective_robot(w.robot), strict.effective_goal_params(w.robot, w.problem, w.goal_params, n)
    out = {"workload": w.name, "queries": n, "population": 128, "what": __doc__.split("\n")[0], "steps": {}}
    for steps in (1, 25):
        runs = {
            "ref_strict": strict.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps),
            "ref_fast": fast.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps),
            "contract": oracle.solve(robot, w.problem, cfg, gp, w.seeds, w.rng_seeds, steps),
        }
        rec = {"runs": {k: {"success": int(v["success"].sum()), "median_fitness": float(np.median(v["fitness"]))} for k, v in runs.items()}, "pairs": {}}
        for a, b in (("ref_strict", "ref_fast"), ("ref_strict", "contract"), ("ref_fast", "contract")):
            d = np.abs(runs[a]["solutions"] - runs[b]["solutions"]).max(axis=1)
            dg = np.abs(runs[a]["genes"][:, 0, 0] - runs[b]["genes"][:, 0, 0]).max(axis=1)  # best individual of the best species
            rec["pairs"][f"{a} vs {b}"] = {"solutions_identical": int((d == 0).sum()), "solutions_within_1e-5": int((d < 1e-5).sum()), "solutions_within_1e-3": int((d < 1e-3).sum()),
                                           "best_genes_within_1e-5": int((dg < 1e-5).sum()), "median_abs_diff": float(np.median(d))}
        out["steps"][str(steps)] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
