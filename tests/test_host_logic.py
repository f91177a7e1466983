"""Host-side logic: Problem construction (src/problem.cpp:72-228 mirror), goal flattening, workloads."""
import numpy as np
import pytest

from bio_ik_b200 import _abi, goals as G, robots, workloads
from bio_ik_b200.problem import Problem
from bio_ik_b200.solver import MODES


def test_active_variables_follow_the_active_subtree():
    rm, groups = robots.pr2_like()
    pr = Problem().initialize(rm, groups["all"], [G.PoseGoal("r_wrist_roll_link")])
    names = [rm.variable_names[i] for i in pr.active_variables]
    # torso + right arm only: left-arm joints are in the group but not on the chain to the tip
    assert names == ["torso_lift_joint"] + [f"r_{j}_joint" for j in robots._ARM_JOINTS]
    pr2 = Problem().initialize(rm, groups["right_arm"], [G.PoseGoal("r_wrist_roll_link")])
    assert [rm.variable_names[i] for i in pr2.active_variables] == [f"r_{j}_joint" for j in robots._ARM_JOINTS]
    assert pr2.tip_link_indices == [rm.link_index["r_wrist_roll_link"]]


def test_goal_variables_come_first_and_fixed_joints_drop_out():
    rm, groups = robots.pr2_like()
    gl = [G.PoseGoal("r_wrist_roll_link"), G.JointVariableGoal("r_elbow_flex_joint", -1.0)]
    pr = Problem().initialize(rm, groups["right_arm"], gl)
    assert rm.variable_names[pr.active_variables[0]] == "r_elbow_flex_joint"
    assert len(pr.active_variables) == 7
    pr = Problem().initialize(rm, groups["right_arm"], gl, fixed_joints=["r_forearm_roll_joint"])
    assert "r_forearm_roll_joint" not in [rm.variable_names[i] for i in pr.active_variables]
    with pytest.raises(RuntimeError):
        Problem().initialize(rm, groups["right_arm"], [G.PoseGoal("no_such_link")])


def test_tips_are_deduplicated_and_goals_keep_order():
    rm, groups = robots.pr2_like()
    gl = [G.PoseGoal("r_wrist_roll_link"), G.PoseGoal("l_wrist_roll_link"), G.OrientationGoal("r_wrist_roll_link", (0, 0, 1, 1), 0.5), G.MinimalDisplacementGoal(2.0)]
    pr = Problem().initialize(rm, groups["all"], gl)
    assert len(pr.tip_link_indices) == 2 and len(pr.active_variables) == 15
    p = pr.to_abi()
    assert [p.goals[i].type for i in range(4)] == [_abi.GOAL_POSE, _abi.GOAL_POSE, _abi.GOAL_ORIENTATION, _abi.GOAL_MINIMAL_DISPLACEMENT]
    assert [p.goals[i].tip for i in range(4)] == [0, 1, 0, 0]
    assert [p.goals[i].secondary for i in range(4)] == [0, 0, 0, 1]
    q = np.array(list(p.goals[2].p)[3:7])
    assert np.isclose(np.linalg.norm(q), 1.0) and p.goals[2].weight == 0.5
    assert p.goals[0].p[7] == 0.5  # PoseGoal rotation_scale default (goal_types.h:133)


def test_host_only_goals_are_refused():
    for cls in (G.JointFunctionGoal, G.LinkFunctionGoal, G.TouchGoal):
        with pytest.raises(G.UnsupportedGoal):
            cls()


def test_mode_table_matches_the_factory_registrations():
    assert MODES == {"bio2": (0, 16), "bio2_memetic": (ord("q"), 8), "bio2_memetic_l": (ord("l"), 8)}


def test_workload_shapes():
    fake_fk = lambda rm, pr, v: np.tile(np.array([0, 0, 0, 0, 0, 0, 1.0]), (v.shape[0], len(pr.tip_link_indices), 1))
    for name, n, T, G_ in (("cfg2", 7, 1, 1), ("cfg3", 15, 2, 3), ("cfg4", 30, 1, 3), ("cfg5", 24, 5, 5)):
        w = workloads.make(name, fake_fk, batch=5)
        assert len(w.problem.active_variables) == n and len(w.problem.tip_link_indices) == T and w.problem.n_goals == G_
        assert w.goal_params.shape == (5, G_, _abi.GOAL_NPARAM) and w.seeds.shape == (5, w.robot.n_vars)
        assert w.rng_seeds.tolist() == [1, 2, 3, 4, 5]
    w = workloads.make("cfg1", fake_fk)
    assert w.seeds.shape[0] == 1 and w.population == 64
