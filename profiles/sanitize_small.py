"""Small solves of every kernel form for compute-sanitizer (memcheck / racecheck): every BASELINE configuration, the lane-group
generation kernel, the group memetic kernel in both modes, floating / planar joints, islands with the query-level early exit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bio_ik_b200 import goals as G, robots, workloads
from bio_ik_b200.problem import Problem
from bio_ik_b200.solver import IKSolver

for name, B in (("cfg2", 37), ("cfg3", 9), ("cfg4", 5), ("cfg5", 7), ("cfg1", 1)):
    f, cid = workloads.CONFIGS[name]
    w = f(B) if name != "cfg1" else f()
    for stale in (False, True):
        s = IKSolver(w.robot, mode="bio2_memetic", population=128, random_seed=1, device=0, reference_stale_tips=stale).initialize(w.problem)
        w.generate(lambda rm, pr, v: s.fk(v), B=(B if name != "cfg1" else 1), cfg_id=cid)
        r = s.solve_batch(w.goal_params, w.seeds, w.rng_seeds, 5, early_exit=True)
        print(name, "stale" if stale else "plain", "ok", float(np.mean(r["success"])))
        s.close()
w = workloads.cfg2(5)
s = IKSolver(w.robot, mode="bio2_memetic", population=18, random_seed=1, device=0).initialize(w.problem)
w.generate(lambda rm, pr, v: s.fk(v), B=5, cfg_id=2)
r = s.solve_islands(w.goal_params, w.seeds, 6, 12, early_exit=2)
print("islands ok", r["success"])
for maker in (robots.floating_base_arm, robots.planar_base_arm):
    rm, groups = maker()
    g = groups["all"]
    pr = Problem().initialize(rm, g, [G.PoseGoal(t) for t in g.tip_links])
    s = IKSolver(rm, mode="bio2_memetic", population=64, random_seed=1, device=0).initialize(pr)
    rng = np.random.default_rng(1)
    seeds = workloads.sample_configurations(rm, pr.active_variables, 6, rng)
    r = s.solve_batch(None, seeds, 1 + np.arange(6), 3)
    print(maker.__name__, "ok")
