"""ctypes loader for the TEST-ONLY host simulation of the CUDA kernels (tests/hostsim)."""
import ctypes as C
import os
import subprocess

import numpy as np

from bio_ik_b200 import _abi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")


class HostSim:
    def __init__(self):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
        self.lib = lib = C.CDLL(os.path.join(HERE, "libhostsim.so"))
        dp, ip, up = _abi.c_double_p, _abi.c_int32_p, _abi.c_uint32_p
        lib.hostsim_last_error.restype = C.c_char_p
        lib.hostsim_tables.argtypes = [C.c_uint32, C.c_int]
        lib.hostsim_tables.restype = dp
        lib.hostsim_minstd_uniform.argtypes = [C.c_uint32, C.c_int, dp]
        lib.hostsim_minstd_index.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
        lib.hostsim_sincos.argtypes = [C.c_int, dp, dp, dp]
        lib.hostsim_solve.argtypes = [C.POINTER(_abi.BioikRobot), C.POINTER(_abi.BioikProblem), C.POINTER(_abi.BioikSolverCfg), C.c_int, dp, dp, up, C.c_int, C.c_int,
                                      dp, dp, ip, ip, dp, dp, dp, C.c_int]
        lib.hostsim_select_islands.argtypes = [C.POINTER(_abi.BioikRobot), C.POINTER(_abi.BioikProblem), C.c_int, C.c_int, dp, dp, dp, dp, ip, ip, C.c_int, dp, dp, ip, ip, ip]
        lib.hostsim_fk.argtypes = [C.POINTER(_abi.BioikRobot), C.POINTER(_abi.BioikProblem), C.c_int, dp, dp, dp]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"hostsim rc={rc}: " + self.lib.hostsim_last_error().decode())

    def solve(self, robot, problem, cfg, goal_params, seeds, rng_seeds, steps, early_exit=False, fast=False, islands=0, evolve_lanes=16, stale_tips=False, island_stride=0):
        self.lib.hostsim_set_islands(int(islands))
        self.lib.hostsim_set_island_stride(int(island_stride))
        self.lib.hostsim_set_stale_tips(int(stale_tips))
        self.lib.hostsim_set_evolve_lanes(int(evolve_lanes))
        seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, robot.n_vars)
        B, n = seeds.shape[0], len(problem.active_variables)
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(B, problem.n_goals, _abi.GOAL_NPARAM)
        rs = np.ascontiguousarray(rng_seeds, dtype=np.uint32)
        res = dict(solutions=np.zeros((B, robot.n_vars)), fitness=np.zeros(B), success=np.zeros(B, dtype=np.int32), steps=np.zeros(B, dtype=np.int32),
                   genes=np.zeros((B, 2, 2, n)), gradients=np.zeros((B, 2, 2, n)), species_fitness=np.zeros((B, 2)))
        r, p = robot.to_abi(), problem.to_abi()
        self._check(self.lib.hostsim_solve(C.byref(r), C.byref(p), C.byref(cfg), B, _abi.dptr(gp), _abi.dptr(seeds), _abi.uptr(rs), steps, int(early_exit),
                                           _abi.dptr(res["solutions"]), _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["steps"]),
                                           _abi.dptr(res["genes"]), _abi.dptr(res["gradients"]), _abi.dptr(res["species_fitness"]), int(fast)))
        return res

    def fk(self, robot, problem, variables, delta=False):
        v = np.ascontiguousarray(variables, dtype=np.float64).reshape(-1, robot.n_vars)
        B, T, n = v.shape[0], len(problem.tip_link_indices), len(problem.active_variables)
        tips = np.zeros((B, T, 7))
        d = np.zeros((B, T, n, 7)) if delta else None
        r, p = robot.to_abi(), problem.to_abi()
        self._check(self.lib.hostsim_fk(C.byref(r), C.byref(p), B, _abi.dptr(v), _abi.dptr(tips), _abi.dptr(d)))
        return (tips, d) if delta else tips

    def select_islands(self, robot, problem, islands, goal_params, seeds, runs, wrap=True):
        """k_select_islands on the per-run results `runs` (dict of solutions / fitness / success / steps of Q * islands runs)"""
        B = len(runs["fitness"])
        Q = B // islands
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(B, problem.n_goals, _abi.GOAL_NPARAM)
        sd = np.ascontiguousarray(seeds, dtype=np.float64).reshape(B, robot.n_vars)
        sol, fit, succ, stp = (np.ascontiguousarray(runs[k]) for k in ("solutions", "fitness", "success", "steps"))
        res = dict(solutions=np.zeros((Q, robot.n_vars)), fitness=np.zeros(Q), success=np.zeros(Q, dtype=np.int32), island=np.zeros(Q, dtype=np.int32), steps=np.zeros(Q, dtype=np.int32))
        r, p = robot.to_abi(), problem.to_abi()
        self._check(self.lib.hostsim_select_islands(C.byref(r), C.byref(p), Q, islands, _abi.dptr(gp), _abi.dptr(sd), _abi.dptr(sol), _abi.dptr(fit), _abi.iptr(succ), _abi.iptr(stp), int(wrap),
                                                    _abi.dptr(res["solutions"]), _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["island"]), _abi.iptr(res["steps"])))
        return res
