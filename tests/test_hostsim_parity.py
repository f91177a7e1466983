"""CPU check of the CUDA kernel SOURCE before any GPU time is spent: tests/hostsim compiles
bioik_kernels.cuh + bioik_host.hpp with g++ (CUDA constructs shimmed, the warp kernel on 32 OS
threads) and must agree BIT-FOR-BIT with the oracle.  Test infrastructure only."""
import ctypes as C

import numpy as np
import pytest

import hostsim_lib
import oracle_lib
from bio_ik_b200 import _abi, workloads


@pytest.fixture(scope="module")
def sim():
    return hostsim_lib.HostSim()


def test_host_table_generation_equals_oracle_tables(sim, oracle):
    u, g = oracle.table_arrays(7)
    su = np.ctypeslib.as_array(sim.lib.hostsim_tables(7, 0), shape=(1 << 23,))
    sg = np.ctypeslib.as_array(sim.lib.hostsim_tables(7, 1), shape=(1 << 23,))
    assert np.array_equal(u, su) and np.array_equal(g, sg)


def test_device_minstd_matches_libstdcpp(sim, oracle):
    for seed in (1, 2, 12345, 2147483647, 0, 4000000000):
        a, b = np.zeros(64), np.zeros(64)
        oracle.lib.oracle_minstd_uniform(seed, 64, _abi.dptr(a))
        sim.lib.hostsim_minstd_uniform(seed, 64, _abi.dptr(b))
        assert np.array_equal(a, b)
        for m in (1, 2, 15, 125, 253, 1 << 20):
            x, y = (C.c_uint64 * 64)(), (C.c_uint64 * 64)()
            oracle.lib.oracle_minstd_index(seed, m, 64, x)
            sim.lib.hostsim_minstd_index(seed, m, 64, y)
            assert list(x) == list(y)


def test_device_sincos_is_the_oracle_sincos(sim, oracle):
    x = np.concatenate([np.random.default_rng(0).uniform(-50, 50, 20000), [0.0, 1e5, 1.00001e5, -3e8, 1e300]])
    s, c = oracle.sincos(x)
    s2, c2 = np.empty_like(x), np.empty_like(x)
    sim.lib.hostsim_sincos(len(x), _abi.dptr(x), _abi.dptr(s2), _abi.dptr(c2))
    assert np.array_equal(s, s2) and np.array_equal(c, c2)


CASES = [("cfg2", 3, 18, "q", 8, 4), ("cfg2", 2, 128, "q", 8, 2), ("cfg2", 2, 21, 0, 16, 2), ("cfg2", 2, 33, "l", 8, 3), ("cfg3", 2, 40, "q", 8, 3), ("cfg4", 2, 40, "q", 8, 3),
         ("cfg5", 2, 36, "q", 8, 2)]


CASES_FAST = CASES + [("cfg2", 2, 64, "q", 8, 3), ("cfg2", 2, 200, "q", 8, 2), ("cfg2", 3, 4, "q", 8, 3), ("cfg3", 2, 128, "q", 8, 2), ("cfg4", 2, 70, "l", 8, 2)]


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name,B,pop,mode,gens,steps", CASES_FAST)
def test_simulated_kernels_are_bit_identical_to_the_oracle(sim, oracle, name, B, pop, mode, gens, steps, fast):
    """fast=False: generic generation kernel (k_evolve); fast=True: register-blocked k_evolve_fast + mutation table."""
    w = workloads.make(name, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=pop, memetic=mode, generations=gens)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps)
    b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, fast=fast)
    for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("name,B,pop,mode,steps", [("cfg3", 3, 40, "q", 4), ("cfg5", 2, 36, "q", 3), ("cfg3", 2, 20, "l", 3)])
def test_reference_stale_tip_mode(sim, oracle, name, B, pop, mode, steps):
    """BIOIK_OPT_REFERENCE_STALE_TIPS: the memetic probe scores the tips a variable cannot move on what the reference's
    phenotypes3 buffer held (quirk Q2) - one lane group per query, species in order, the carried frames in the state.
    The oracle's emulation of the same quirk (flags bit 3) is pinned against the reference's own code in test_reference_pin.py."""
    w = workloads.make(name, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=pop, memetic=mode)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, flags=8)
    b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, fast=True, stale_tips=True)
    for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
        assert np.array_equal(a[k], b[k]), k
    plain = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps)
    assert not np.array_equal(plain["genes"], a["genes"])  # the quirk really changes multi-tip runs


@pytest.mark.parametrize("maker", ["floating_base_arm", "planar_base_arm"])
def test_floating_and_planar_base_joints(sim, oracle, maker):
    """FLOATING joint on the chain: joint frame, numeric delta frames (frameTwist), quaternion-gene normalisation; the production
    sequence falls back to the generic generation kernel (no fast instantiation with quaternion genes)"""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = getattr(robots, maker)()
    g = groups["all"]
    pr = Problem().initialize(rm, g, [G.PoseGoal(t) for t in g.tip_links])
    rng = np.random.default_rng(1)
    B = 3
    targets = workloads.sample_configurations(rm, pr.active_variables, B, rng)
    seeds = workloads.sample_configurations(rm, pr.active_variables, B, rng)
    tips = oracle.fk(rm, pr, targets)
    t2, dd = sim.fk(rm, pr, targets, delta=True)
    d, mask = oracle.approx(rm, pr, targets)
    d = d.copy()
    d[mask == 0, 6] = 0.0
    assert np.array_equal(t2, tips) and np.array_equal(dd, d)
    gp = np.repeat(pr.default_goal_params()[None], B, 0)
    for gi, rec in enumerate(pr.goal_list):
        gp[:, gi, 0:7] = tips[:, rec["tip"], 0:7]
    cfg = oracle_lib.make_cfg(population=20)
    a = oracle.solve(rm, pr, cfg, gp, seeds, 1 + np.arange(B), 3)
    for fast in (False, True):
        b = sim.solve(rm, pr, cfg, gp, seeds, 1 + np.arange(B), 3, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
            assert np.array_equal(a[k], b[k]), (k, fast)


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_generation_kernel_lane_groups(sim, oracle, lanes):
    """k_evolve_fast with 8 / 16 / 32 lanes per task (4 / 2 / 1 tasks per warp) on the single-pose problem: an odd task count
    leaves lane groups of the last warp without work, early exit retires groups of a warp at different steps."""
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=5)
    for pop, steps, early in ((128, 3, False), (200, 2, False)) + (((128, 8, True),) if lanes == 16 else ()):
        cfg = oracle_lib.make_cfg(population=pop)
        a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, early_exit=early)
        b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, early_exit=early, fast=True, evolve_lanes=lanes)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps") if not early else ("solutions", "fitness", "success", "steps"):
            assert np.array_equal(a[k], b[k]), (k, pop)


@pytest.mark.parametrize("name,B,pop,mode,steps,variant", [("cfg2", 3, 18, "q", 4, 6), ("cfg2", 3, 40, "l", 3, 6), ("cfg2", 3, 18, "q", 3, 9), ("cfg3", 3, 40, "q", 3, 6), ("cfg3", 1, 20, "q", 2, 7),
                                                            ("cfg4", 2, 40, "q", 2, 6), ("cfg4", 1, 20, "l", 2, 8), ("cfg5", 3, 36, "q", 2, 6), ("cfg5", 1, 20, "q", 2, 7),
                                                            ("cfg4", 1, 128, "q", 2, 6), ("cfg4", 1, 210, "q", 1, 6)])
def test_group_memetic_kernel(sim, oracle, name, B, pop, mode, steps, variant):
    """k_memetic_group: the memetic line search on W lanes per task (6 = the width the library picks, 7/8/9 = W forced
    to 8/16/32, so every problem also runs with more variables than lanes and with idle lanes); odd task counts leave
    groups of a warp without work."""
    w = workloads.make(name, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=pop, memetic=mode)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps)
    b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, fast=variant)
    for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
        assert np.array_equal(a[k], b[k]), k


def test_simulated_fk_and_delta_frames(sim, oracle):
    w = workloads.make("cfg3", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=6)
    tips, delta = sim.fk(w.robot, w.problem, w.targets, delta=True)
    assert np.array_equal(tips, oracle.fk(w.robot, w.problem, w.targets))
    d, mask = oracle.approx(w.robot, w.problem, w.targets)
    d = d.copy()
    d[mask == 0, 6] = 0.0  # device convention: unmasked delta frames are all-zero
    assert np.array_equal(delta, d)


def test_early_exit_contract(sim, oracle):
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=3)
    cfg = oracle_lib.make_cfg(population=18)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 14, early_exit=True)
    b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 14, early_exit=True, fast=True)
    for k in ("solutions", "fitness", "success", "steps"):
        assert np.array_equal(a[k], b[k]), k
    assert set(b["steps"].tolist()) <= {4, 8, 12, 14}


def test_mixed_goals_fast_kernel(sim, oracle):
    """link goals of several kinds + primary and secondary joint-space goals + a JointVariableGoal, 2 tips."""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = robots.pr2_like()
    g = groups["all"]
    r, l = "r_wrist_roll_link", "l_wrist_roll_link"
    gl = [G.PositionGoal(r, (0.6, -0.3, 0.9)), G.LookAtGoal(l, (1, 0, 0), (2, 0.5, 1), 0.3), G.JointVariableGoal("torso_lift_joint", 0.2, 2.0), G.CenterJointsGoal(0.5, secondary=False),
          G.LineGoal(l, (0.5, 0.2, 1), (1, 1, 0), 0.7), G.MinimalDisplacementGoal(1.5), G.AvoidJointLimitsGoal(0.8), G.DirectionGoal(r, (1, 0, 0), (0, 0, 1), 0.4),
          G.ConeGoal(l, (1, 0, 0), (0, 0.6, 0.8), 0.4, weight=0.3)]
    pr = Problem().initialize(rm, g, gl)
    rng = np.random.default_rng(5)
    seeds = workloads.sample_configurations(rm, pr.active_variables, 3, rng)
    rs = np.array([3, 4, 5], dtype=np.uint32)
    cfg = oracle_lib.make_cfg(population=45)
    a = oracle.solve(rm, pr, cfg, None, seeds, rs, 3)
    for fast in (False, True):
        b = sim.solve(rm, pr, cfg, None, seeds, rs, 3, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness"):
            assert np.array_equal(a[k], b[k]), (k, fast)


@pytest.mark.parametrize("pop", [20, 128, 200])
def test_preselection_with_tied_secondary_fitness(sim, oracle, pop):
    """AvoidJointLimitsGoal as the only secondary goal: most children score exactly 0.0, so the pre-selection order (:366-378)
    is decided by the child slot - the generation kernel's fast rank pass must notice the ties and take its exact pass."""
    from bio_ik_b200 import goals as G
    w = workloads.cfg2(3)
    pr = w.problem.__class__().initialize(w.robot, w.group, [G.PoseGoal("r_wrist_roll_link"), G.AvoidJointLimitsGoal(1.0)])
    w.problem = pr
    w.generate(lambda rm, p, v: oracle.fk(rm, p, v), B=3, cfg_id=2)
    cfg = oracle_lib.make_cfg(population=pop)
    a = oracle.solve(w.robot, pr, cfg, w.goal_params, w.seeds, w.rng_seeds, 3)
    for fast in (False, True):
        b = sim.solve(w.robot, pr, cfg, w.goal_params, w.seeds, w.rng_seeds, 3, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness"):
            assert np.array_equal(a[k], b[k]), (k, fast)


def test_mixed_goals_tip_major_kernel(sim, oracle):
    """Five tips (the tip-major generation kernel): link goals of several kinds with a tip read again later in the goal list, a
    secondary link goal (a setSecondary link goal sees the identity frames), primary and secondary joint-space goals."""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = robots.shadow_like_hand()
    g = groups["hand"]
    t = g.tip_links
    late = G.LineGoal(t[3], (0.0, 0.0, 0.35), (0, 1, 0), 0.4)
    late.secondary_ = True
    gl = [G.PositionGoal(t[0], (0.03, 0.0, 0.36)), G.OrientationGoal(t[1], (0, 0, 0, 1), 0.5), G.MaxDistanceGoal(t[0], (0.0, 0.0, 0.3), 0.05, 0.8), G.JointVariableGoal("WRJ1", 0.1, 2.0),
          G.PoseGoal(t[2], (0.0, 0.02, 0.37), (0, 0, 0, 1), 0.7), G.CenterJointsGoal(0.5, secondary=False), late, G.MinimalDisplacementGoal(1.5), G.PlaneGoal(t[4], (0, 0, 0.33), (0, 0, 1), 0.6),
          G.ConeGoal(t[1], (0, 0, 1), (0, 0.6, 0.8), 0.4, weight=0.3)]
    pr = Problem().initialize(rm, g, gl)
    rng = np.random.default_rng(9)
    seeds = workloads.sample_configurations(rm, pr.active_variables, 3, rng)
    rs = np.array([7, 8, 9], dtype=np.uint32)
    cfg = oracle_lib.make_cfg(population=45)
    a = oracle.solve(rm, pr, cfg, None, seeds, rs, 3)
    for fast in (False, True, 6):
        b = sim.solve(rm, pr, cfg, None, seeds, rs, 3, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness"):
            assert np.array_equal(a[k], b[k]), (k, fast)


@pytest.mark.parametrize("variant", [2, 3, 4, 5])
def test_serial_kernel_placement_variants(sim, oracle, variant):
    """The fused serial kernel keeps delta frames / link frames in shared memory, in the HBM state or in local
    memory depending on what fits (make_serial_plan): every placement gives the same bits."""
    for name, B, pop, steps in (("cfg3", 2, 36, 3), ("cfg4", 2, 36, 2)):
        w = workloads.make(name, lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
        cfg = oracle_lib.make_cfg(population=pop)
        a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps)
        b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, fast=variant)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
            assert np.array_equal(a[k], b[k]), (name, k)


def test_mimic_joints(sim, oracle):
    """updateMimic + the mimic branches of the Jacobian (chained mimic joints, two tips below them)."""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = robots.mimic_gripper_arm()
    pr = Problem().initialize(rm, groups["all"], [G.PositionGoal("pad_a"), G.PoseGoal("pad_b")])
    assert [rm.variable_names[i] for i in pr.active_variables] == ["j1", "j2", "j3", "j4", "finger_a_joint"]  # mimic joints are not genes
    rng = np.random.default_rng(1)
    tg = workloads.sample_configurations(rm, pr.active_variables, 3, rng)
    seeds = workloads.sample_configurations(rm, pr.active_variables, 3, rng)
    tips = oracle.fk(rm, pr, tg)
    gp = np.repeat(pr.default_goal_params()[None], 3, 0)
    gp[:, 0, 0:3], gp[:, 1, 0:7] = tips[:, 0, 0:3], tips[:, 1, :]
    cfg = oracle_lib.make_cfg(population=40)
    rs = np.arange(3, dtype=np.uint32) + 1
    a = oracle.solve(rm, pr, cfg, gp, seeds, rs, 5)
    for fast in (0, 1):
        b = sim.solve(rm, pr, cfg, gp, seeds, rs, 5, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness"):
            assert np.array_equal(a[k], b[k]), (k, fast)


def test_virtual_joints_that_mimic(sim, oracle):
    """a PLANAR joint mimicking a prismatic joint, a FLOATING joint mimicking a revolute one (first variable only, forward_kinematics.h:
    230-246) and the numeric Jacobian branch reached through the mimic dependency (ivar2, :698-699)"""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = robots.mimic_virtual_joint_arm()
    pr = Problem().initialize(rm, groups["all"], [G.PoseGoal("ee"), G.PositionGoal("probe")])
    assert [rm.variable_names[i] for i in pr.active_variables] == ["j1", "j2", "j3"]
    rng = np.random.default_rng(2)
    base = robots.mimic_virtual_joint_base(rm)
    tg = workloads.sample_configurations(rm, pr.active_variables, 3, rng, base=base)
    seeds = workloads.sample_configurations(rm, pr.active_variables, 3, rng, base=base)
    tips = oracle.fk(rm, pr, tg)
    gp = np.repeat(pr.default_goal_params()[None], 3, 0)
    gp[:, 0, 0:7], gp[:, 1, 0:3] = tips[:, 0, :], tips[:, 1, 0:3]
    tips_s, delta_s = sim.fk(rm, pr, seeds, delta=True)
    assert np.array_equal(tips_s, oracle.fk(rm, pr, seeds))
    d, mask = oracle.approx(rm, pr, seeds)
    d = d.copy()
    d[mask == 0, 6] = 0.0  # device convention: unmasked delta frames are all-zero
    assert np.array_equal(delta_s, d) and np.all(mask[:, 0, :] == 1)  # every gene moves the end effector, two of them through a mimicking virtual joint
    cfg = oracle_lib.make_cfg(population=40)
    rs = np.arange(3, dtype=np.uint32) + 1
    a = oracle.solve(rm, pr, cfg, gp, seeds, rs, 5)
    for fast in (0, 1):
        b = sim.solve(rm, pr, cfg, gp, seeds, rs, 5, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness"):
            assert np.array_equal(a[k], b[k]), (k, fast)


@pytest.mark.parametrize("B,pop,mode,steps,early", [(5, 128, "q", 5, False), (19, 64, "q", 3, False), (33, 128, "l", 2, False), (18, 128, "q", 9, True), (3, 200, 0, 3, False)])
def test_persistent_kernel_is_bit_identical_to_the_oracle(sim, oracle, B, pop, mode, steps, early):
    """bioik_persist.cuh: the whole solve as work items (evolve(query, step), serial(group of 16 queries, step)) taken from two
    queues by resident warps - here by ONE simulated warp, in two launches with a cut in the middle (the resume path of
    bioik_step).  Partial groups (B not a multiple of 16), several groups, early exit and the non-memetic mode included."""
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=B)
    cfg = oracle_lib.make_cfg(population=pop, memetic=mode, generations=8 if mode else 16)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, early_exit=early)
    b = sim.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, steps, early_exit=early, fast=10)
    for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps") if not early else ("solutions", "fitness", "success", "steps"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("first", [False, True])
def test_balance_goal_on_the_device_code(sim, oracle, first):
    """BalanceGoal (every link with mass a tip link: 12 tips here) through the kernel source: generic generation kernel and the
    fused serial kernel, goal first or in the middle of the goal list."""
    from bio_ik_b200 import goals as G, robots
    from bio_ik_b200.problem import Problem
    rm, groups = robots.balancing_tree()
    g = groups["all"]
    bal = G.BalanceGoal((0.05, -0.02, 0.3), 0.8, axis=(0.1, 0.2, 0.97))
    gl = ([bal] if first else []) + [G.PoseGoal(g.tip_links[0])] + ([] if first else [bal]) + [G.PositionGoal(g.tip_links[1], weight=0.5)]
    pr = Problem().initialize(rm, g, gl)
    rng = np.random.default_rng(2)
    B = 3
    seeds = workloads.sample_configurations(rm, pr.active_variables, B, rng)
    gp = np.repeat(pr.default_goal_params()[None], B, 0)
    cfg = oracle_lib.make_cfg(population=20)
    a = oracle.solve(rm, pr, cfg, gp, seeds, 1 + np.arange(B), 4)
    for fast in (False, True):
        b = sim.solve(rm, pr, cfg, gp, seeds, 1 + np.arange(B), 4, fast=fast)
        for k in ("genes", "gradients", "species_fitness", "solutions", "fitness", "success", "steps"):
            assert np.array_equal(a[k], b[k]), (k, fast)
