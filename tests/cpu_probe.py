"""Probe of the CPU-port scaling on the current host (diagnostic; not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle_lib
from bio_ik_b200 import workloads
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e:
    print("cpu.max n/a", e)
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
o = oracle_lib.Oracle("strict"); f = oracle_lib.Oracle("fast")
w = workloads.make("cfg2", lambda rm, pr, v: o.fk(rm, pr, v), batch=4096)
cfg = oracle_lib.make_cfg(population=128)
f.tables(1)
for nt in (1, 8, 16, 32, 64, 128):
    n = min(4096, 64 * nt)
    t0 = time.perf_counter()
    f.solve(w.robot, w.problem, cfg, w.goal_params[:n], w.seeds[:n], w.rng_seeds[:n], 25, nthreads=nt)
    dt = time.perf_counter() - t0
    print(f"threads {nt:4d}: {n/dt:9.1f} solves/s  ({n/dt/nt:7.1f} per thread)")
