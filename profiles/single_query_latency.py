"""Latency of ONE MoveIt-style query answered by many islands (bioik_solve_islands): host buffers in, wrapped solution
out, early exit at the driver's 4-step checks - with the islands as clones of one random stream (the reference's threads) and with
BIOIK_OPT_ISLAND_STREAM_STRIDE = 1.  usage: python profiles/single_query_latency.py [population]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bio_ik_b200 import workloads
from bio_ik_b200.solver import IKSolver

pop = int(sys.argv[1]) if len(sys.argv) > 1 else 18
w = workloads.cfg2(256)
from bio_ik_b200 import _abi
solver = IKSolver(w.robot, mode="bio2_memetic", population=pop, random_seed=1, device=0).initialize(w.problem)
w.generate(lambda rm, pr, v: solver.fk(v), B=256, cfg_id=2)
for stride, islands in [(0, 1), (0, 16), (1, 16), (0, 64), (1, 64), (0, 256), (1, 256)]:
    solver.set_option(_abi.OPT_ISLAND_STREAM_STRIDE, stride)
    for steps in (4, 8, 25):
        lat, ok = [], []
        for q in range(64):
            gp, sd = w.goal_params[q:q + 1], w.seeds[q:q + 1]
            if q < 4:
                solver.solve_islands(gp, sd, islands, steps, early_exit=2)  # warm-up: sizes the state
            t0 = time.perf_counter()
            r = solver.solve_islands(gp, sd, islands, steps, early_exit=2)
            lat.append(time.perf_counter() - t0)
            ok.append(int(r["success"][0]))
        lat = np.array(lat) * 1e3
        print(f"pop {pop} stream stride {stride} islands {islands:5d} step budget {steps:2d}: median {np.median(lat):.2f} ms  p90 {np.percentile(lat, 90):.2f} ms  success {np.mean(ok):.2f}")
