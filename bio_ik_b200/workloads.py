"""Synthetic inputs for the BASELINE.json configurations.

Targets are generated FK-reachable the way the reference's own self-test does
(README.md:410-418): sample a random valid configuration, run exact FK, use the tip
poses as goals; the seed is a second independent sample.  The FK used to make the
targets is passed in as a callable (the GPU FK of the product in bench.py, the
oracle FK in CPU tests) — this module computes nothing itself.
"""
import numpy as np

from . import _abi, goals as G, robots
from .problem import Problem

BASE_SEED = 20240924


def sample_configurations(robot, active_vars, B, rng, base=None):
    """[B][n_vars]: active variables ~ U(lo, hi) (continuous joints U(-pi, pi)), others = base (default 0)."""
    v = np.zeros((B, robot.n_vars)) if base is None else np.repeat(np.asarray(base, dtype=np.float64)[None, :], B, 0)
    for ivar in active_vars:
        lo, hi = robot.sampling_bounds(ivar)
        v[:, ivar] = rng.uniform(lo, hi, B)
    return v


class Workload:
    def __init__(self, name, robot, group, problem, population, steps, batch):
        self.name, self.robot, self.group, self.problem = name, robot, group, problem
        self.population, self.steps, self.batch = population, steps, batch
        self.goal_params = None  # [B][n_goals][NPARAM]
        self.seeds = None  # [B][n_vars]
        self.rng_seeds = None  # [B] uint32
        self.targets = None  # [B][n_vars] configurations that generated the goals

    def generate(self, fk, B=None, cfg_id=0, seed_noise=None):
        """fk(robot, problem, variables[B][n_vars]) -> tip frames [B][T][7]"""
        B = B or self.batch
        rng = np.random.default_rng(BASE_SEED + cfg_id)
        pr, rm = self.problem, self.robot
        self.targets = sample_configurations(rm, pr.active_variables, B, rng)
        if seed_noise is None:
            self.seeds = sample_configurations(rm, pr.active_variables, B, rng)
        else:  # cfg4: seed = target + N(0, noise^2), clipped to the limits
            self.seeds = self.targets.copy()
            for ivar in pr.active_variables:
                lo, hi = rm.sampling_bounds(ivar)
                self.seeds[:, ivar] = np.clip(self.targets[:, ivar] + rng.normal(0, seed_noise, B), lo, hi)
        tips = fk(rm, pr, self.targets)  # [B][T][7]
        gp = np.repeat(pr.default_goal_params()[None, :, :], B, 0)
        for gi, rec in enumerate(pr.goal_list):
            t = rec["goal"].type
            f = tips[:, rec["tip"], :]
            if t == _abi.GOAL_POSITION:
                gp[:, gi, 0:3] = f[:, 0:3]
            elif t == _abi.GOAL_ORIENTATION:
                gp[:, gi, 3:7] = f[:, 3:7]
            elif t == _abi.GOAL_POSE:
                gp[:, gi, 0:7] = f[:, 0:7]
        self.goal_params = np.ascontiguousarray(gp)
        self.rng_seeds = (1 + np.arange(B)).astype(np.uint32)
        return self


def _problem(rm, group, goal_list):
    return Problem().initialize(rm, group, goal_list)


def cfg1():
    """PR2-like right_arm 7-DOF, single PoseGoal, pop=64 (plumbing case)."""
    rm, groups = robots.pr2_like()
    g = groups["right_arm"]
    return Workload("cfg1_pr2_right_arm_pose_pop64", rm, g, _problem(rm, g, [G.PoseGoal("r_wrist_roll_link")]), 64, 25, 1)


def cfg2(batch=10000):
    """PR2-like right_arm 7-DOF, batch of random PoseGoals, pop=128 x 200 gens (25 steps)."""
    rm, groups = robots.pr2_like()
    g = groups["right_arm"]
    return Workload("cfg2_pr2_right_arm_pose_pop128", rm, g, _problem(rm, g, [G.PoseGoal("r_wrist_roll_link")]), 128, 25, batch)


def cfg3(batch=4096):
    """PR2-like 'all' group (torso + both arms, 15 DOF), Pose(r) + Pose(l) + Orientation(r)."""
    rm, groups = robots.pr2_like()
    g = groups["all"]
    gl = [G.PoseGoal("r_wrist_roll_link"), G.PoseGoal("l_wrist_roll_link"), G.OrientationGoal("r_wrist_roll_link", weight=0.5)]
    return Workload("cfg3_pr2_all_multitip", rm, g, _problem(rm, g, gl), 128, 25, batch)


def cfg4(batch=2048):
    """30-DOF snake, Pose + MinimalDisplacement + AvoidJointLimits (both secondary)."""
    rm, groups = robots.snake(30)
    g = groups["all"]
    gl = [G.PoseGoal("tip"), G.MinimalDisplacementGoal(1.0), G.AvoidJointLimitsGoal(1.0)]
    return Workload("cfg4_snake30_weighted", rm, g, _problem(rm, g, gl), 128, 25, batch)


def cfg5(batch=65536):
    """Shadow-like hand 24-DOF, 5 fingertip PositionGoals."""
    rm, groups = robots.shadow_like_hand()
    g = groups["hand"]
    gl = [G.PositionGoal(t) for t in g.tip_links]
    return Workload("cfg5_shadow_hand_5tips", rm, g, _problem(rm, g, gl), 128, 25, batch)


CONFIGS = {"cfg1": (cfg1, 1), "cfg2": (cfg2, 2), "cfg3": (cfg3, 3), "cfg4": (cfg4, 4), "cfg5": (cfg5, 5)}


def make(name, fk, batch=None):
    f, cid = CONFIGS[name]
    w = f() if batch is None or name == "cfg1" else f(batch)
    return w.generate(fk, B=batch if name != "cfg1" else 1, cfg_id=cid, seed_noise=(0.1 if name == "cfg4" else None))
