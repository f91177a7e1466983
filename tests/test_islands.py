"""many differently seeded islands per query, reduced like IKParallel::solve reduces its
threads (src/ik_parallel.h:218-258), then the plugin's angle wrap (src/kinematics_plugin.cpp:580-611).
CPU part: the oracle's restatement against an independent numpy statement and against the simulated kernel."""
import math

import numpy as np
import pytest

import hostsim_lib
import oracle_lib
from bio_ik_b200 import goals as G, robots, workloads
from bio_ik_b200.problem import Problem


@pytest.fixture(scope="module")
def sim():
    return hostsim_lib.HostSim()


def numpy_wrap(v, r, lo, hi):
    """kinematics_plugin.cpp:586-610, one variable"""
    if r < v - math.pi or r > v + math.pi:
        v -= r
        v /= 2 * math.pi
        v += 0.5
        v -= math.floor(v)
        v -= 0.5
        v *= 2 * math.pi
        v += r
    if v > hi:
        v -= math.ceil(max(0.0, v - hi) / (2 * math.pi)) * (2 * math.pi)
    if v < lo:
        v += math.ceil(max(0.0, lo - v) / (2 * math.pi)) * (2 * math.pi)
    return min(max(v, lo), hi)


def fake_runs(rng, robot, problem, Q, islands, success_rate):
    B = Q * islands
    runs = dict(solutions=rng.uniform(-12, 12, (B, robot.n_vars)), fitness=rng.uniform(0, 1, B), success=(rng.uniform(0, 1, B) < success_rate).astype(np.int32),
                steps=rng.integers(1, 30, B).astype(np.int32))
    runs["fitness"][rng.integers(0, B, B // 4)] = 0.25  # ties: the first island must win
    return runs


@pytest.mark.parametrize("secondary", [False, True])
def test_selection_and_wrap_against_numpy(oracle, sim, secondary):
    rm, groups = robots.pr2_like()
    g = groups["right_arm"]
    gl = [G.PoseGoal("r_wrist_roll_link")] + ([G.MinimalDisplacementGoal(1.0), G.AvoidJointLimitsGoal(0.5)] if secondary else [])
    pr = Problem().initialize(rm, g, gl)
    rng = np.random.default_rng(3)
    Q, islands = 40, 6
    seeds_q = workloads.sample_configurations(rm, pr.active_variables, Q, rng)
    seeds = np.repeat(seeds_q, islands, axis=0)
    gp = np.repeat(np.repeat(pr.default_goal_params()[None], Q, 0), islands, axis=0)
    for rate in (0.0, 0.3, 1.0):
        runs = fake_runs(rng, rm, pr, Q, islands, rate)
        import ctypes as C
        from bio_ik_b200 import _abi
        res = dict(solutions=np.zeros((Q, rm.n_vars)), fitness=np.zeros(Q), success=np.zeros(Q, dtype=np.int32), island=np.zeros(Q, dtype=np.int32), steps=np.zeros(Q, dtype=np.int32))
        r, p = rm.to_abi(), pr.to_abi()
        oracle._check(oracle.lib.oracle_select_islands(C.byref(r), C.byref(p), Q, islands, _abi.dptr(gp), _abi.dptr(seeds), _abi.dptr(runs["solutions"]), _abi.dptr(runs["fitness"]), _abi.iptr(runs["success"]),
                                                       _abi.iptr(runs["steps"]), 1, _abi.dptr(res["solutions"]), _abi.dptr(res["fitness"]), _abi.iptr(res["success"]), _abi.iptr(res["island"]), _abi.iptr(res["steps"])))
        # independent statement
        act = pr.active_variables
        sec = np.zeros(Q * islands)
        if secondary:
            _, s = oracle.approx_fitness(rm, pr, gp, seeds, seeds, runs["solutions"][:, act][:, None, :])
            sec = s[:, 0]
        lo, hi = rm.arrays["var_min"], rm.arrays["var_max"]
        for q in range(Q):
            sl = slice(q * islands, (q + 1) * islands)
            f, ok = runs["fitness"][sl], runs["success"][sl]
            if ok.any():
                score = np.where(ok != 0, f + sec[sl], np.inf)
                k = int(np.argmin(score))  # argmin returns the first minimum, like the strict '<' scan
                best = score[k]
            else:
                k = int(np.argmin(f))
                best = f[k]
            assert res["island"][q] == k and res["fitness"][q] == best and res["success"][q] == ok[k] and res["steps"][q] == runs["steps"][sl][k]
            want = runs["solutions"][q * islands + k].copy()
            for iv in act:  # every active variable of the PR2-like arm is revolute and the robot has no mimic joints
                want[iv] = numpy_wrap(want[iv], seeds_q[q, iv], lo[iv], hi[iv])
            assert np.array_equal(res["solutions"][q], want)
            for iv in act:
                assert lo[iv] <= res["solutions"][q, iv] <= hi[iv]
        # the simulated kernel gives the same bits
        got = sim.select_islands(rm, pr, islands, gp, seeds, runs, wrap=True)
        for k in res:
            assert np.array_equal(got[k], res[k]), k
        # without wrap the selected run comes back untouched
        got = sim.select_islands(rm, pr, islands, gp, seeds, runs, wrap=False)
        assert all(np.array_equal(got["solutions"][q], runs["solutions"][q * islands + res["island"][q]]) for q in range(Q))


def test_wrap_properties():
    """the wrapped angle is the same rotation (when no clamp was needed) and the closest copy to the seed inside the limits"""
    rng = np.random.default_rng(0)
    for _ in range(2000):
        lo, hi = sorted(rng.uniform(-7, 7, 2))
        if hi - lo < 0.5:
            continue
        r = rng.uniform(lo, hi)
        v = rng.uniform(-40, 40)
        w = numpy_wrap(v, r, lo, hi)
        assert lo <= w <= hi
        if hi - lo >= 2 * math.pi + 1e-9:  # some copy of the angle always fits: no clamping
            assert abs(math.remainder(w - v, 2 * math.pi)) < 1e-9


def test_wrap_is_skipped_for_robots_with_mimic_joints_and_for_prismatic_variables(oracle, sim):
    for maker, gname in ((robots.mimic_gripper_arm, None), (lambda: robots.random_tree(1), "all")):
        rm, groups = maker()
        g = groups[gname] if gname else groups[sorted(groups)[0]]
        pr = Problem().initialize(rm, g, [G.PositionGoal(t) for t in g.tip_links])
        rng = np.random.default_rng(1)
        Q, islands = 8, 3
        seeds = np.repeat(workloads.sample_configurations(rm, pr.active_variables, Q, rng), islands, axis=0)
        runs = fake_runs(rng, rm, pr, Q, islands, 0.5)
        got = sim.select_islands(rm, pr, islands, None, seeds, runs, wrap=True)
        has_mimic = (rm.arrays["joint_mimic"] >= 0).any()
        for q in range(Q):
            src = runs["solutions"][q * islands + got["island"][q]]
            for iv in range(rm.n_vars):
                j = rm.getJointOfVariable(iv)
                revolute = rm.links[j].joint_type == 1
                if has_mimic or not revolute or iv not in pr.active_variables:
                    assert got["solutions"][q, iv] == src[iv]


def test_islands_oracle_end_to_end(oracle):
    """more islands never hurt: the selected fitness is the best over the islands, success if any island succeeded"""
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=6)
    cfg = oracle_lib.make_cfg(population=18)
    res = oracle_lib.oracle_solve_islands(oracle, w.robot, w.problem, cfg, w.goal_params, w.seeds, 4, 8, wrap=True)
    runs = res["runs"]
    # the lock-step island solver without early exit is just the batch solver on repeated inputs
    plain = oracle.solve(w.robot, w.problem, cfg, np.repeat(w.goal_params, 4, 0), np.repeat(w.seeds, 4, 0), 1 + np.arange(24), 8)
    for k in ("solutions", "fitness", "success", "steps"):
        assert np.array_equal(plain[k], runs[k]), k
    for q in range(6):
        sl = slice(4 * q, 4 * q + 4)
        assert res["success"][q] == int(runs["success"][sl].any())
        if not runs["success"][sl].any():
            assert res["fitness"][q] == runs["fitness"][sl].min()


def test_query_level_early_exit(oracle, sim):
    """early_exit = 2: the reference driver's `finished` flag (src/ik_parallel.h:160-186) - once one island has passed the 4-step
    success test, no island of that query starts another burst.  Oracle (lock-step islands) vs the simulated kernels."""
    Q = 2
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=5)
    w.goal_params, w.seeds = w.goal_params[3:5], w.seeds[3:5]
    cfg = oracle_lib.make_cfg(population=18)
    islands, steps = 3, 16
    ref = oracle_lib.oracle_solve_islands(oracle, w.robot, w.problem, cfg, w.goal_params, w.seeds, islands, steps, early_exit=2)
    runs = ref["runs"]
    st = runs["steps"].reshape(Q, islands)
    ok = runs["success"].reshape(Q, islands)
    assert (st.max(axis=1) == st.min(axis=1)).all()           # lock step: all islands of a query stop together
    assert (st[:, 0] < steps).any() and (st % 4 == 0).all()    # ...and before the budget is used up, at a 4-step check
    assert (ok.sum(axis=1) > 0)[st[:, 0] < steps].all()      # a query stops early only because an island succeeded
    gp, sd = np.repeat(w.goal_params, islands, 0), np.repeat(w.seeds, islands, 0)
    got = sim.solve(w.robot, w.problem, cfg, gp, sd, 1 + np.arange(Q * islands), steps, early_exit=2, fast=True, islands=islands)
    for k in ("solutions", "fitness", "success", "steps"):
        assert np.array_equal(got[k], runs[k]), k
    sel = sim.select_islands(w.robot, w.problem, islands, gp, sd, got, wrap=True)
    for k in ("solutions", "fitness", "success", "island", "steps"):
        assert np.array_equal(sel[k], ref[k]), k
    # per-island exit lets the other islands go on: never fewer steps than with the query-level flag
    per = oracle_lib.oracle_solve_islands(oracle, w.robot, w.problem, cfg, w.goal_params, w.seeds, islands, steps, early_exit=1)
    assert (per["runs"]["steps"] >= runs["steps"]).all()


@pytest.mark.parametrize("name,pop,islands,stride,fast", [("cfg2", 18, 4, 1, True), ("cfg2", 20, 3, 2, False), ("cfg4", 20, 3, 1, 6)])
def test_island_stream_stride(oracle, sim, name, pop, islands, stride, fast):
    """BIOIK_OPT_ISLAND_STREAM_STRIDE: island i starts i * stride steps into the query-independent random streams.  The kernels
    (simulated) against the oracle's statement of the option; island 0 is the plain run; the islands differ from the first step on."""
    Q, steps = 2, 3
    w = workloads.make(name, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=Q)
    cfg = oracle_lib.make_cfg(population=pop)
    B = Q * islands
    gp, seeds = np.repeat(w.goal_params, islands, 0), np.repeat(w.seeds, islands, 0)
    rs = (1 + np.arange(B)).astype(np.uint32)
    want = oracle_lib.oracle_solve_islands(oracle, w.robot, w.problem, cfg, w.goal_params, w.seeds, islands, steps, rng_seeds=rs, early_exit=0, island_stride=stride)["runs"]
    got = sim.solve(w.robot, w.problem, cfg, gp, seeds, rs, steps, fast=fast, islands=islands, island_stride=stride)
    for k in ("solutions", "fitness", "success", "steps"):
        assert np.array_equal(got[k], want[k]), k
    plain = oracle.solve(w.robot, w.problem, cfg, gp, seeds, rs, steps)
    for q in range(Q):
        assert np.array_equal(got["solutions"][q * islands], plain["solutions"][q * islands])  # island 0 reads the streams from their start
    if not (name == "cfg2" and fast is True):
        return
    clones = sim.solve(w.robot, w.problem, cfg, gp, seeds, rs, 1, fast=fast, islands=islands, island_stride=0)
    ahead = sim.solve(w.robot, w.problem, cfg, gp, seeds, rs, 1, fast=fast, islands=islands, island_stride=stride)
    g0, g1 = clones["genes"].reshape(Q, islands, -1), ahead["genes"].reshape(Q, islands, -1)
    assert all(np.array_equal(g0[q, 0, : g0.shape[2] // 2], g0[q, i, : g0.shape[2] // 2]) for q in range(Q) for i in range(islands))  # clones: species 0 identical after one step
    assert all(not np.array_equal(g1[q, 0], g1[q, i]) for q in range(Q) for i in range(1, islands))
