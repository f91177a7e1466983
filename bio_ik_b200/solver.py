"""Host-side solver object over the C ABI, mirroring the reference's solver interface
(IKBase::initialize / step / getSolution, src/ik_base.h:128-210) in its batched form and the
IKFactory mode names (src/ik_evolution_2.cpp:652-654)."""
import ctypes as C

import numpy as np

from . import _abi

# IKFactory::Class registrations of the bio2 family: name -> (memetic, generations per step)
MODES = {"bio2": (0, 16), "bio2_memetic": (ord("q"), 8), "bio2_memetic_l": (ord("l"), 8)}


class BioIKError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bioik error {code}: {msg}")
        self.code = code


class IKSolver:
    """One solver context on one GPU.  `population` is children.size() of the reference (2 parents +
    child_count, 18 in the reference; BASELINE "pop")."""

    def __init__(self, robot_model, mode="bio2_memetic", population=18, generations=None, memetic_iters=8, random_seed=1, device=0, reference_stale_tips=False, island_stream_stride=0):
        if mode not in MODES:
            raise BioIKError(_abi.E_INVALID, f"class not found {mode}")  # IKFactory::create, src/utils.h:432-437
        self.lib = _abi.load_library()
        self.robot_model = robot_model
        self.problem = None
        cfg = _abi.BioikSolverCfg()
        cfg.memetic, default_gens = MODES[mode]
        cfg.population, cfg.generations = population, generations or default_gens
        cfg.memetic_iters, cfg.table_seed, cfg.device = memetic_iters, random_seed, device
        self.cfg = cfg
        self._ctx = C.c_void_p()
        r = robot_model.to_abi()
        rc = self.lib.bioik_create(C.byref(r), C.byref(cfg), C.byref(self._ctx))
        if rc != _abi.OK:
            raise BioIKError(rc, self.lib.bioik_last_error(None).decode())
        if reference_stale_tips:
            self.set_option(_abi.OPT_REFERENCE_STALE_TIPS, 1)
        if island_stream_stride:
            self.set_option(_abi.OPT_ISLAND_STREAM_STRIDE, island_stream_stride)

    def cancel(self):
        """IKBase::canceled: stop the solve in flight (callable from another thread)"""
        self._check(self.lib.bioik_cancel(self._ctx))

    def set_option(self, option, value):
        """bioik_set_option; OPT_REFERENCE_STALE_TIPS = 1 reproduces quirk Q2 of the reference's memetic step (see include/bioik_b200.h)"""
        self._check(self.lib.bioik_set_option(self._ctx, int(option), int(value)))

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.bioik_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _abi.OK:
            raise BioIKError(rc, self.lib.bioik_last_error(self._ctx).decode())

    def initialize(self, problem):
        """IKBase::initialize(problem)"""
        self.problem = problem
        p = problem.to_abi()
        self._check(self.lib.bioik_set_problem(self._ctx, C.byref(p)))
        return self

    def _shape_inputs(self, goal_params, seeds, rng_seeds):
        rm, pr = self.robot_model, self.problem
        seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, rm.n_vars)
        B = seeds.shape[0]
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(B, pr.n_goals, _abi.GOAL_NPARAM)
        rs = np.ascontiguousarray(rng_seeds, dtype=np.uint32).reshape(B)
        return B, gp, seeds, rs

    def solve_batch(self, goal_params, seeds, rng_seeds, steps, early_exit=False, out=None):
        """B independent queries, `steps` step() calls each (host buffers; H2D/D2H inside)."""
        B, gp, seeds, rs = self._shape_inputs(goal_params, seeds, rng_seeds)
        rm = self.robot_model
        res = out or dict(solutions=np.empty((B, rm.n_vars)), fitness=np.empty(B), success=np.empty(B, dtype=np.int32), steps=np.empty(B, dtype=np.int32))
        self._check(self.lib.bioik_solve_batch(self._ctx, B, _abi.dptr(gp), _abi.dptr(seeds), _abi.uptr(rs), steps, int(early_exit), _abi.dptr(res["solutions"]),
                                               _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["steps"])))
        return res

    def solve_islands(self, goal_params, seeds, islands, steps, rng_seeds=None, early_exit=2, wrap=True):
        """Q MoveIt-style queries, each solved by `islands` differently seeded runs in one batch and reduced like
        IKParallel::solve reduces its threads (src/ik_parallel.h:218-258: best successful island by primary +
        secondary fitness, else best primary fitness); wrap=True applies the plugin's angle wrap towards the seed
        (src/kinematics_plugin.cpp:580-611).  rng_seeds: [Q * islands] (default 1 + run index).
        early_exit: 0 = every island runs `steps` steps, 1 = an island stops at its own success, 2 = all islands of a query stop
        after the check at which the first of them succeeded (the reference driver's `finished` flag, src/ik_parallel.h:160-186)."""
        rm, pr = self.robot_model, self.problem
        seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, rm.n_vars)
        Q = seeds.shape[0]
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(Q, pr.n_goals, _abi.GOAL_NPARAM)
        rs = (1 + np.arange(Q * islands)).astype(np.uint32) if rng_seeds is None else np.ascontiguousarray(rng_seeds, dtype=np.uint32).reshape(Q * islands)
        res = dict(solutions=np.empty((Q, rm.n_vars)), fitness=np.empty(Q), success=np.empty(Q, dtype=np.int32), island=np.empty(Q, dtype=np.int32), steps=np.empty(Q, dtype=np.int32))
        self._check(self.lib.bioik_solve_islands(self._ctx, Q, int(islands), _abi.dptr(gp), _abi.dptr(seeds), _abi.uptr(rs), steps, int(early_exit), int(wrap), _abi.dptr(res["solutions"]),
                                                 _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["island"]), _abi.iptr(res["steps"])))
        return res

    # ---- the reference's solver interface in its resumable form (IKBase::initialize / step / getSolution, src/ik_base.h:138-154)
    def begin(self, goal_params, seeds, islands=1, rng_seeds=None, max_steps=0, early_exit=0):
        """initialize(problem) for Q queries x `islands` runs; the state stays on the device until the next begin / solve_*"""
        rm, pr = self.robot_model, self.problem
        seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, rm.n_vars)
        Q = seeds.shape[0]
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(Q, pr.n_goals, _abi.GOAL_NPARAM)
        rs = (1 + np.arange(Q * islands)).astype(np.uint32) if rng_seeds is None else np.ascontiguousarray(rng_seeds, dtype=np.uint32).reshape(Q * islands)
        self._check(self.lib.bioik_begin(self._ctx, Q, int(islands), _abi.dptr(gp), _abi.dptr(seeds), _abi.uptr(rs), int(max_steps), int(early_exit)))
        self._queries = Q
        return self

    def step(self, nsteps=1):
        """step() x nsteps; returns the number of runs that would execute a further step"""
        active = C.c_int32(0)
        self._check(self.lib.bioik_step(self._ctx, int(nsteps), C.byref(active)))
        return active.value

    def get_solution(self, wrap=False):
        """getSolution() per query after the driver's selection among its runs"""
        Q, rm = self._queries, self.robot_model
        res = dict(solutions=np.empty((Q, rm.n_vars)), fitness=np.empty(Q), success=np.empty(Q, dtype=np.int32), island=np.empty(Q, dtype=np.int32), steps=np.empty(Q, dtype=np.int32))
        self._check(self.lib.bioik_get_solution(self._ctx, int(wrap), _abi.dptr(res["solutions"]), _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["island"]), _abi.iptr(res["steps"])))
        return res

    def solve_batch_device(self, B, d_goal_params, d_seeds, d_rng_seeds, steps, early_exit, d_solutions, d_fitness, d_success, d_steps, stream=None):
        """Same with raw device pointers (ints); enqueues on `stream` (or the context stream), no sync."""
        self._check(self.lib.bioik_solve_batch_device(self._ctx, B, d_goal_params, d_seeds, d_rng_seeds, steps, int(early_exit), d_solutions, d_fitness, d_success, d_steps, stream))

    def pack_results_device(self, B, d_solutions, d_fitness, d_success, d_steps, d_slab, stream=None):
        """[B][n_vars + 3] float64 slab solution | fitness | success | steps from the device outputs of solve_batch_device"""
        self._check(self.lib.bioik_pack_results_device(self._ctx, B, d_solutions, d_fitness, d_success, d_steps, d_slab, stream))

    def kernel_name(self):
        return self.lib.bioik_kernel_name(self._ctx).decode()

    def synchronize(self):
        self._check(self.lib.bioik_synchronize(self._ctx))

    def trace(self, goal_params, seeds, rng_seeds, steps):
        """Solver state after `steps` steps (trajectory-level parity)."""
        B, gp, seeds, rs = self._shape_inputs(goal_params, seeds, rng_seeds)
        n, rm = len(self.problem.active_variables), self.robot_model
        res = dict(genes=np.empty((B, 2, 2, n)), gradients=np.empty((B, 2, 2, n)), species_fitness=np.empty((B, 2)), solutions=np.empty((B, rm.n_vars)), fitness=np.empty(B))
        self._check(self.lib.bioik_solve_batch_trace(self._ctx, B, _abi.dptr(gp), _abi.dptr(seeds), _abi.uptr(rs), steps, _abi.dptr(res["genes"]), _abi.dptr(res["gradients"]),
                                                     _abi.dptr(res["species_fitness"]), _abi.dptr(res["solutions"]), _abi.dptr(res["fitness"])))
        return res

    def fk(self, variables):
        v = np.ascontiguousarray(variables, dtype=np.float64).reshape(-1, self.robot_model.n_vars)
        out = np.empty((v.shape[0], len(self.problem.tip_link_indices), 7))
        self._check(self.lib.bioik_fk_batch(self._ctx, v.shape[0], _abi.dptr(v), _abi.dptr(out)))
        return out

    def approx(self, variables):
        v = np.ascontiguousarray(variables, dtype=np.float64).reshape(-1, self.robot_model.n_vars)
        out = np.empty((v.shape[0], len(self.problem.tip_link_indices), len(self.problem.active_variables), 7))
        self._check(self.lib.bioik_approx_batch(self._ctx, v.shape[0], _abi.dptr(v), _abi.dptr(out)))
        return out

    def approx_fitness(self, goal_params, seeds, base_variables, genotypes):
        rm, n = self.robot_model, len(self.problem.active_variables)
        base = np.ascontiguousarray(base_variables, dtype=np.float64).reshape(-1, rm.n_vars)
        B = base.shape[0]
        g = np.ascontiguousarray(genotypes, dtype=np.float64).reshape(B, -1, n)
        M = g.shape[1]
        seeds = np.ascontiguousarray(seeds, dtype=np.float64).reshape(B, rm.n_vars)
        gp = None if goal_params is None else np.ascontiguousarray(goal_params, dtype=np.float64).reshape(B, self.problem.n_goals, _abi.GOAL_NPARAM)
        prim, sec = np.empty((B, M)), np.empty((B, M))
        self._check(self.lib.bioik_approx_fitness_batch(self._ctx, B, M, _abi.dptr(gp), _abi.dptr(seeds), _abi.dptr(base), _abi.dptr(g), _abi.dptr(prim), _abi.dptr(sec)))
        return prim, sec

    def launch_count(self):
        return int(self.lib.bioik_launch_count(self._ctx))

    def kernel_time(self, reset=True, disable=False):
        """(ms in the generation kernel, its launches, ms in the serial kernels, their launches) since the last reset.
        Per-launch CUDA-event timing is switched ON by the first call (it costs host time per launch and excludes
        CUDA-graph replay) and OFF again with disable=True."""
        a, b = C.c_double(), C.c_double()
        na, nb = C.c_int64(), C.c_int64()
        self._check(self.lib.bioik_kernel_time(self._ctx, 2 if disable else int(reset), C.byref(a), C.byref(na), C.byref(b), C.byref(nb)))
        return a.value, na.value, b.value, nb.value


def create_solver(name, robot_model, **kw):
    """IKFactory::create(params.solver_class_name, params) for the bio2 family on the GPU."""
    return IKSolver(robot_model, mode=name, **kw)
