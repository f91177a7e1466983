"""Validation of the CPU oracle BY CONSTRUCTION: the reference's own
tests pin only the frame algebra, so everything else is checked against
independent implementations and analytic properties.  CPU only."""
import math

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

import oracle_lib
from bio_ik_b200 import _abi, goals as G, robots, workloads
from bio_ik_b200.problem import Problem


def rand_frames(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([rng.uniform(-1, 1, (n, 3)), q], axis=1)


def call3(fn, *frames):
    out = np.zeros(7)
    fn(*[_abi.dptr(np.ascontiguousarray(f)) for f in frames], _abi.dptr(out))
    return out


def frames_close(a, b, tol):
    # quaternions equal up to sign
    return np.allclose(a[:3], b[:3], atol=tol) and (np.allclose(a[3:], b[3:], atol=tol) or np.allclose(a[3:], -b[3:], atol=tol))


def test_change_algebra(oracle):
    """Port of TEST(BioIK, change), test/utest.cpp:63-81: change(fc, fb, concat(fb, fa)) == concat(fc, fa)
    for 10 000 random frames (reference tolerance 1e-3; the IEEE-strict oracle holds 1e-12)."""
    rng = np.random.default_rng(0)
    fa, fb, fc = rand_frames(rng, 10000), rand_frames(rng, 10000), rand_frames(rng, 10000)
    for i in range(10000):
        fx = call3(oracle.lib.oracle_concat, fb[i], fa[i])
        fy = call3(oracle.lib.oracle_change, fc[i], fb[i], fx)
        fz = call3(oracle.lib.oracle_concat, fc[i], fa[i])
        assert frames_close(fy, fz, 1e-12)


def test_concat_invert_vs_scipy(oracle):
    rng = np.random.default_rng(1)
    a, b = rand_frames(rng, 200), rand_frames(rng, 200)
    for i in range(200):
        c = call3(oracle.lib.oracle_concat, a[i], b[i])
        ra, rb = R.from_quat(a[i, 3:]), R.from_quat(b[i, 3:])
        assert np.allclose(c[:3], a[i, :3] + ra.apply(b[i, :3]), atol=1e-13)
        qc = (ra * rb).as_quat()
        assert np.allclose(c[3:], qc, atol=1e-13) or np.allclose(c[3:], -qc, atol=1e-13)
        inv = call3(oracle.lib.oracle_invert, a[i])
        ident = call3(oracle.lib.oracle_concat, a[i], inv)
        assert frames_close(ident, np.array([0, 0, 0, 0, 0, 0, 1.0]), 1e-13)


def test_rng_known_answers(oracle):
    """SURVEY.md §8(c) known-answer values (g++ 13.3 libstdc++)."""
    import ctypes as C
    x = (C.c_uint64 * 3)()
    oracle.lib.oracle_xorshift(3, x)
    assert list(x) == [8748534153485358512, 3040900993826735515, 3453997556048239312]
    assert x[0] & ((1 << 23) - 1) == 6165936 and x[1] >= (1 << 23) and x[2] % 16 == 0
    u = np.zeros(2)
    oracle.lib.oracle_minstd_uniform(1, 2, _abi.dptr(u))
    assert u[0] == 0.085032448717433665 and u[1] == 0.89161127730485767
    n = np.zeros(3)
    oracle.lib.oracle_minstd_normal(1, 3, _abi.dptr(n))
    # measured with g++ 13.3 (SURVEY.md lists the same three values in reverse order)
    assert list(n) == [-0.40747178673787671, -1.2397359036589184, 0.39977066170883935]
    u0 = np.zeros(2)
    oracle.lib.oracle_minstd_uniform(0, 2, _abi.dptr(u0))  # minstd_rand(0) == seed 1
    assert np.array_equal(u, u0)


def test_tables_are_the_random_ctor_streams(oracle):
    """Random::Random fills uniform first, then gauss, from ONE engine (src/ik_base.h:118-125)."""
    u, g = oracle.table_arrays(1)
    assert u[0] == 0.085032448717433665 and u[1] == 0.89161127730485767
    assert np.all((u >= 0) & (u < 1))
    assert abs(u.mean() - 0.5) < 1e-3 and abs(g.mean()) < 2e-3 and abs(g.std() - 1) < 2e-3
    # the gauss table continues the engine after 2 * 8Mi draws, so it differs from a fresh normal stream
    n = np.zeros(3)
    oracle.lib.oracle_minstd_normal(1, 3, _abi.dptr(n))
    assert g[0] != n[0]


def test_det_sincos_accuracy(oracle):
    """det_sincos (the arithmetic-contract sin/cos) stays within 2 ulp of libm on the working range
    and within 1e-9 absolute after the fmod fallback."""
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-7, 7, 200000), rng.uniform(-1e5, 1e5, 50000), np.linspace(-math.pi, math.pi, 10001), [0.0, -0.0, 1e-300, math.pi / 4, -math.pi / 4]])
    s, c = oracle.sincos(x)
    es = np.abs(s - np.sin(x)) / np.spacing(np.maximum(np.abs(np.sin(x)), 1e-300))
    ec = np.abs(c - np.cos(x)) / np.spacing(np.maximum(np.abs(np.cos(x)), 1e-300))
    small = np.abs(x) <= 7
    assert es[small].max() <= 2.0 and ec[small].max() <= 2.0
    assert np.abs(s - np.sin(x)).max() < 1e-11 and np.abs(c - np.cos(x)).max() < 1e-11
    assert np.allclose(s * s + c * c, 1.0, atol=1e-15)
    big = np.array([1e6, -3.3e7, 1e12, 1e300])
    s, c = oracle.sincos(big)
    assert np.all(np.isfinite(s)) and np.abs(s[:2] - np.sin(big[:2])).max() < 1e-8
    s, c = oracle.sincos(np.array([np.nan, np.inf]))
    assert np.all(np.isnan(s)) and np.all(np.isnan(c))


def test_det_acos_accuracy(oracle):
    """det_acos (ConeGoal's arithmetic-contract acos, fdlibm algorithm) is within 1 ulp of libm."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-1, 1, 200000), np.linspace(-1, 1, 20001), [0.0, 0.5, -0.5, 1.0, -1.0, 1e-20, 0.4999999999, 0.5000000001]])
    a = oracle.acos(x)
    ref = np.arccos(x)
    err = np.abs(a - ref) / np.spacing(np.maximum(ref, 1e-300))
    assert err.max() <= 1.0
    assert oracle.acos(np.array([1.0]))[0] == 0.0 and np.isnan(oracle.acos(np.array([1.5]))[0])


def numpy_fk(robot, variables):
    """Independent exact FK (scipy rotations), returns link frames [L][7]."""
    a = robot.arrays
    variables = np.array(variables, dtype=float)
    for l, link in enumerate(robot.links):  # mimic joints follow their source joint
        if link.mimic:
            variables[a["joint_first_var"][l]] = variables[a["joint_first_var"][robot.joint_index[link.mimic]]] * link.mimic_factor + link.mimic_offset
    frames = []
    for l, link in enumerate(robot.links):
        o = a["link_origin"][l]
        ro, po = R.from_quat(o[3:]), o[:3]
        jt = link.joint_type
        fv = a["joint_first_var"][l]
        if jt == _abi.JOINT_REVOLUTE:
            rj, pj = R.from_rotvec(np.array(link.axis) * variables[fv]), np.zeros(3)
        elif jt == _abi.JOINT_PRISMATIC:
            rj, pj = R.identity(), np.array(link.axis) * variables[fv]
        else:
            rj, pj = R.identity(), np.zeros(3)
        rl, pl = ro * rj, po + ro.apply(pj)
        p = a["link_parent"][l]
        if p >= 0:
            rp, pp = frames[p]
            rl, pl = rp * rl, pp + rp.apply(pl)
        frames.append((rl, pl))
    return frames


@pytest.mark.parametrize("maker", [robots.pr2_like, robots.snake, robots.shadow_like_hand, lambda: robots.random_tree(3), robots.mimic_gripper_arm])
def test_exact_fk_vs_independent_numpy(oracle, maker):
    rm, groups = maker()
    g = list(groups.values())[-1]
    pr = Problem().initialize(rm, g, [G.PositionGoal(t) for t in g.tip_links])
    rng = np.random.default_rng(5)
    v = workloads.sample_configurations(rm, range(rm.n_vars), 20, rng)
    tips, links = oracle.fk(rm, pr, v, links=True)
    tips_libm = oracle.fk(rm, pr, v, libm=True)
    assert np.abs(tips - tips_libm).max() < 1e-14  # det_sincos vs libm: a few ulp
    for b in range(20):
        ref = numpy_fk(rm, v[b])
        for t, li in enumerate(pr.tip_link_indices):
            rl, pl = ref[li]
            assert np.allclose(tips[b, t, :3], pl, atol=1e-12)
            q = rl.as_quat()
            assert np.allclose(tips[b, t, 3:], q, atol=1e-12) or np.allclose(tips[b, t, 3:], -q, atol=1e-12)


def body_twist(fc, f0, f1, h):
    """twist between two tip frames / h, expressed in the (central) tip frame fc"""
    rc, r0, r1 = R.from_quat(fc[3:]), R.from_quat(f0[3:]), R.from_quat(f1[3:])
    v = rc.inv().apply(f1[:3] - f0[:3]) / h
    w = rc.inv().apply((r1 * r0.inv()).as_rotvec()) / h
    return np.concatenate([v, w])


@pytest.mark.parametrize("maker", [robots.pr2_like, lambda: robots.random_tree(4), robots.mimic_gripper_arm])
def test_jacobian_vs_central_differences(oracle, maker):
    """computeJacobian (src/forward_kinematics.h:600-730) = tip-local twist per unit variable change."""
    rm, groups = maker()
    g = list(groups.values())[-1]
    pr = Problem().initialize(rm, g, [G.PoseGoal(t) for t in g.tip_links])
    rng = np.random.default_rng(7)
    v = workloads.sample_configurations(rm, range(rm.n_vars), 4, rng)
    _, _, jac = oracle.approx(rm, pr, v, jacobian=True)
    h = 1e-6
    for b in range(4):
        for i, ivar in enumerate(pr.active_variables):
            vp, vm = v[b].copy(), v[b].copy()
            vp[ivar] += h
            vm[ivar] -= h
            fp, fm, fc = oracle.fk(rm, pr, vp)[0], oracle.fk(rm, pr, vm)[0], oracle.fk(rm, pr, v[b])[0]
            for t in range(len(pr.tip_link_indices)):
                num = body_twist(fc[t], fm[t], fp[t], 2 * h)
                assert np.allclose(jac[b, 6 * t:6 * t + 6, i], num, atol=1e-7), (b, i, t)


def test_approximator_error_is_second_order(oracle):
    """Procedure of the reference's self-check solver (src/ik_test.cpp:92-129): the linear tip-frame
    extrapolation differs from exact FK by O(d^2)."""
    rm, groups = robots.pr2_like()
    g = groups["all"]
    pr = Problem().initialize(rm, g, [G.PoseGoal(t) for t in g.tip_links])
    rng = np.random.default_rng(9)
    base = workloads.sample_configurations(rm, pr.active_variables, 8, rng)
    delta, mask = oracle.approx(rm, pr, base)
    tip0 = oracle.fk(rm, pr, base)
    errs = []
    for d in (1e-2, 1e-3):
        pert = base.copy()
        dv = rng.uniform(-d, d, (8, len(pr.active_variables)))
        for i, ivar in enumerate(pr.active_variables):
            pert[:, ivar] += dv[:, i]
        exact = oracle.fk(rm, pr, pert)
        approx = tip0 + np.einsum("bi,btik->btk", dv, delta)
        errs.append(np.abs(exact - approx).max())
    assert errs[0] < 5e-4 and errs[1] < 5e-6 and errs[1] < errs[0] / 50
    # structural zeros: left-arm joints do not move the right tip
    li = [i for i, ivar in enumerate(pr.active_variables) if rm.variable_names[ivar].startswith("l_")]
    assert np.all(mask[:, 0, li] == 0) and np.all(delta[:, 0, li, :] == 0)


def test_approx_fitness_matches_formula(oracle):
    rm, groups = robots.pr2_like()
    g = groups["right_arm"]
    goal = G.PoseGoal("r_wrist_roll_link", (0.6, -0.2, 0.9), (0, 0, 0.3, 1.0))
    cone = G.ConeGoal("r_wrist_roll_link", (1, 0, 0), (0, 0.6, 0.8), 0.3, weight=0.5, position=(0.5, 0, 1), position_weight=0.7)
    pr = Problem().initialize(rm, g, [goal, G.MinimalDisplacementGoal(2.0), cone])
    rng = np.random.default_rng(11)
    base = workloads.sample_configurations(rm, pr.active_variables, 3, rng)
    n = len(pr.active_variables)
    genes = base[:, pr.active_variables][:, None, :] + rng.normal(0, 0.05, (3, 5, n))
    prim, sec = oracle.approx_fitness(rm, pr, None, base, base, genes)
    delta, _ = oracle.approx(rm, pr, base)
    tip0 = oracle.fk(rm, pr, base)
    p = np.array(goal.params())
    vmax = np.array([rm.arrays["var_max_velocity"][i] for i in pr.active_variables])
    vw = (1 / vmax) / (1 / vmax).sum()
    for b in range(3):
        for m in range(5):
            d = genes[b, m] - base[b, pr.active_variables]
            f = tip0[b, 0] + d @ delta[b, 0]
            e = ((f[:3] - p[:3]) ** 2).sum() + min(((p[3:7] - f[3:]) ** 2).sum(), ((p[3:7] + f[3:]) ** 2).sum()) * 0.25
            # ConeGoal: quat_mul_vec(q, axis) by its defining formula r = v + 2 (q_w t + q_xyz x t), t = q_xyz x v (q is NOT normalised here)
            qv, qw, ax = f[3:6], f[6], np.array([1.0, 0, 0])
            t = np.cross(qv, ax)
            v = ax + 2 * (qw * t + np.cross(qv, t))
            ang = math.acos(max(-1.0, min(1.0, float(np.dot(v, [0, 0.6, 0.8]) / math.sqrt((v @ v) * 1.0)))))
            ec = max(0.0, ang - 0.3) ** 2 + 0.49 * ((np.array([0.5, 0, 1]) - f[:3]) ** 2).sum()
            assert math.isclose(prim[b, m], e + 0.25 * ec, rel_tol=1e-11)
            assert math.isclose(sec[b, m], 4.0 * ((d * vw) ** 2).sum(), rel_tol=1e-12)


def test_fk_ik_fk_round_trip(oracle):
    """Integration test of the reference family (README.md:404-447): random reachable poses are solved."""
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=48)
    cfg = oracle_lib.make_cfg(population=18)
    res = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 25)
    assert res["success"].mean() >= 0.8
    ok = res["success"] == 1
    tips = oracle.fk(w.robot, w.problem, res["solutions"])
    assert np.abs(tips[ok, 0, :3] - w.goal_params[ok, 0, :3]).max() < 2e-5
    assert np.all(res["fitness"][ok] < 1e-9)
    # the solution respects the clip limits
    a = w.robot.arrays
    for ivar in w.problem.active_variables:
        if a["var_bounded"][ivar]:
            assert np.all(res["solutions"][:, ivar] >= a["var_min"][ivar]) and np.all(res["solutions"][:, ivar] <= a["var_max"][ivar])
    # early exit returns the same answers for solved queries that were solved at a checkpoint
    res2 = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 25, early_exit=True)
    assert np.all(res2["steps"] <= 25) and np.all((res2["steps"] % 4 == 0) | (res2["steps"] == 25))
    assert res2["success"].sum() >= res["success"].sum() - 2


@pytest.mark.parametrize("memetic,gens", [("l", 8), (0, 16)])
def test_other_bio2_modes_run(oracle, memetic, gens):
    """bio2_memetic_l and bio2 (src/ik_evolution_2.cpp:652-654)."""
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=16)
    cfg = oracle_lib.make_cfg(population=18, memetic=memetic, generations=gens)
    res = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 25)
    assert np.all(np.isfinite(res["fitness"]))
    assert np.median(res["fitness"]) < 1e-2


def test_determinism_and_seed_sensitivity(oracle):
    w = workloads.make("cfg2", lambda rm, pr, v: oracle.fk(rm, pr, v), batch=8)
    cfg = oracle_lib.make_cfg(population=18)
    a = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 6)
    b = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 6, nthreads=1)
    assert np.array_equal(a["genes"], b["genes"]) and np.array_equal(a["fitness"], b["fitness"])
    # libm sincos deviates from det_sincos by ulps only: trajectories may diverge chaotically, fitness scale must not
    c = oracle.solve(w.robot, w.problem, cfg, w.goal_params, w.seeds, w.rng_seeds, 6, flags=1)
    assert np.all(np.isfinite(c["fitness"]))
