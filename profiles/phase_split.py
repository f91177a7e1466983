"""Phase timing of the fused serial kernel: BIOIK_SERIAL_SPLIT=1 launches MEMETIC / SPECIES / PREPARE separately
(same results, state passes through HBM) and the context reports the CUDA-event time of each.
usage: BIOIK_SERIAL_SPLIT=1 python profiles/phase_split.py cfg2 10000 [population]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bio_ik_b200 import workloads
from bio_ik_b200.solver import IKSolver

name, B = sys.argv[1], int(sys.argv[2])
pop = int(sys.argv[3]) if len(sys.argv) > 3 else 128
f, cid = workloads.CONFIGS[name]
w = f(B) if name != "cfg1" else f()
solver = IKSolver(w.robot, mode="bio2_memetic", population=pop, random_seed=1, device=0).initialize(w.problem)
w.generate(lambda rm, pr, v: solver.fk(v), B=B, cfg_id=cid, seed_noise=(0.1 if name == "cfg4" else None))
solver.kernel_time(reset=True)
for it in range(4):
    if it == 1:
        solver.kernel_time(reset=True)
        t0 = time.perf_counter()
    r = solver.solve_batch(w.goal_params, w.seeds, w.rng_seeds, 25)
ev, ne, se, ns = solver.kernel_time(reset=True)
dt = (time.perf_counter() - t0) / 3
print(f"{name} B={B} pop={pop}: evolve {ev / 3:.3f} ms/pass ({ne // 3} launches), serial {se / 3:.3f} ms/pass ({ns // 3} launches), wall {1e3 * dt:.2f} ms/pass, {B / dt:.0f} solves/s, success {np.mean(r['success']):.3f}")
