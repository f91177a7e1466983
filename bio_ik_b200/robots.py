"""Hand-authored kinematic tables for the BASELINE.json configurations.

No URDF/SRDF ships with the reference and none exists offline, so
these are SYNTHETIC look-alikes ("PR2-like", "Shadow-like"): joint order, axes,
link offsets and limits follow public recollection of pr2_description etc. and
must not be mistaken for the vendor models.
"""
import math

import numpy as np

from ._abi import JOINT_FIXED, JOINT_PRISMATIC, JOINT_REVOLUTE
from .model import JointModelGroup, Link, RobotModel

PI = math.pi


def _arm(side, y):
    s = side
    return [
        Link(f"{s}_shoulder_pan_link", "torso_lift_link", JOINT_REVOLUTE, xyz=(0.0, y, 0.0), axis=(0, 0, 1),
             lower=(-2.1354 if s == "r" else -0.5646), upper=(0.5646 if s == "r" else 2.1354), velocity=2.088, joint_name=f"{s}_shoulder_pan_joint"),
        Link(f"{s}_shoulder_lift_link", f"{s}_shoulder_pan_link", JOINT_REVOLUTE, xyz=(0.1, 0, 0), axis=(0, 1, 0), lower=-0.3536, upper=1.2963, velocity=2.082,
             joint_name=f"{s}_shoulder_lift_joint"),
        Link(f"{s}_upper_arm_roll_link", f"{s}_shoulder_lift_link", JOINT_REVOLUTE, axis=(1, 0, 0), lower=(-3.75 if s == "r" else -0.65), upper=(0.65 if s == "r" else 3.75),
             velocity=3.27, joint_name=f"{s}_upper_arm_roll_joint"),
        Link(f"{s}_upper_arm_link", f"{s}_upper_arm_roll_link", JOINT_FIXED, joint_name=f"{s}_upper_arm_joint"),
        Link(f"{s}_elbow_flex_link", f"{s}_upper_arm_link", JOINT_REVOLUTE, xyz=(0.4, 0, 0), axis=(0, 1, 0), lower=-2.1213, upper=-0.15, velocity=3.3,
             joint_name=f"{s}_elbow_flex_joint"),
        Link(f"{s}_forearm_roll_link", f"{s}_elbow_flex_link", JOINT_REVOLUTE, axis=(1, 0, 0), lower=-PI, upper=PI, bounded=False, velocity=3.6,
             joint_name=f"{s}_forearm_roll_joint"),
        Link(f"{s}_forearm_link", f"{s}_forearm_roll_link", JOINT_FIXED, joint_name=f"{s}_forearm_joint"),
        Link(f"{s}_wrist_flex_link", f"{s}_forearm_link", JOINT_REVOLUTE, xyz=(0.321, 0, 0), axis=(0, 1, 0), lower=-2.0, upper=-0.1, velocity=3.078,
             joint_name=f"{s}_wrist_flex_joint"),
        Link(f"{s}_wrist_roll_link", f"{s}_wrist_flex_link", JOINT_REVOLUTE, axis=(1, 0, 0), lower=-PI, upper=PI, bounded=False, velocity=3.6,
             joint_name=f"{s}_wrist_roll_joint"),
    ]


_ARM_JOINTS = ["shoulder_pan", "shoulder_lift", "upper_arm_roll", "elbow_flex", "forearm_roll", "wrist_flex", "wrist_roll"]


def pr2_like():
    """PR2-like torso + both arms (SURVEY.md §8(d) values).  Groups:
    right_arm (7 DOF, tip r_wrist_roll_link), left_arm, all (torso + both arms, 15 DOF)."""
    links = [
        Link("base_link", None, JOINT_FIXED, joint_name="world_joint"),
        Link("torso_lift_link", "base_link", JOINT_PRISMATIC, xyz=(-0.05, 0.0, 0.739675), axis=(0, 0, 1), lower=0.0, upper=0.33, velocity=0.013,
             joint_name="torso_lift_joint"),
    ] + _arm("r", -0.188) + _arm("l", 0.188)
    rm = RobotModel("pr2_like", links)
    groups = {
        "right_arm": JointModelGroup(rm, "right_arm", [f"r_{j}_joint" for j in _ARM_JOINTS], ["r_wrist_roll_link"]),
        "left_arm": JointModelGroup(rm, "left_arm", [f"l_{j}_joint" for j in _ARM_JOINTS], ["l_wrist_roll_link"]),
        "all": JointModelGroup(rm, "all", ["torso_lift_joint"] + [f"r_{j}_joint" for j in _ARM_JOINTS] + [f"l_{j}_joint" for j in _ARM_JOINTS],
                               ["r_wrist_roll_link", "l_wrist_roll_link"]),
    }
    return rm, groups


def snake(n=30, link_length=0.1, limit=2.0, velocity=1.0):
    """n revolute joints with alternating y/z axes, `link_length` m links (cfg4)."""
    links = [Link("base_link", None, JOINT_FIXED, joint_name="world_joint")]
    prev = "base_link"
    for i in range(n):
        name = f"seg{i}"
        links.append(Link(name, prev, JOINT_REVOLUTE, xyz=(link_length if i else 0.0, 0, 0), axis=((0, 1, 0) if i % 2 == 0 else (0, 0, 1)), lower=-limit, upper=limit,
                          velocity=velocity, joint_name=f"j{i}"))
        prev = name
    links.append(Link("tip", prev, JOINT_FIXED, xyz=(link_length, 0, 0), joint_name="tip_joint"))
    rm = RobotModel(f"snake{n}", links)
    groups = {"all": JointModelGroup(rm, "all", [f"j{i}" for i in range(n)], ["tip"])}
    return rm, groups


def shadow_like_hand():
    """Shadow-like hand: 2 wrist joints + fingers with (4,4,4,5,5) joints = 24 DOF, 5 fingertips (cfg5)."""
    links = [
        Link("forearm", None, JOINT_FIXED, joint_name="world_joint"),
        Link("wrist", "forearm", JOINT_REVOLUTE, xyz=(0, 0, 0.213), axis=(0, 1, 0), lower=-0.489, upper=0.140, velocity=2.0, joint_name="WRJ2"),
        Link("palm", "wrist", JOINT_REVOLUTE, xyz=(0, 0, 0.034), axis=(1, 0, 0), lower=-0.698, upper=0.489, velocity=2.0, joint_name="WRJ1"),
    ]
    tips, joints = [], ["WRJ2", "WRJ1"]

    def unit(a):
        a = np.asarray(a, dtype=float)
        return tuple(a / np.linalg.norm(a))

    def finger(prefix, base_xyz, specs, base_rpy=(0, 0, 0)):
        prev = "palm"
        n = len(specs)
        for k, (xyz, axis, lo, hi) in enumerate(specs):
            name, jn = f"{prefix}{k}", f"{prefix.upper()}J{n - k}"
            links.append(Link(name, prev, JOINT_REVOLUTE, xyz=(base_xyz if k == 0 else xyz), rpy=(base_rpy if k == 0 else (0, 0, 0)), axis=unit(axis), lower=lo, upper=hi,
                              velocity=2.0, joint_name=jn))
            joints.append(jn)
            prev = name
        links.append(Link(f"{prefix}tip", prev, JOINT_FIXED, xyz=(0, 0, 0.026), joint_name=f"{prefix}tip_joint"))
        tips.append(f"{prefix}tip")

    f4 = [((0, 0, 0), (0, -1, 0), -0.349, 0.349), ((0, 0, 0), (1, 0, 0), 0.0, 1.571), ((0, 0, 0.045), (1, 0, 0), 0.0, 1.571), ((0, 0, 0.025), (1, 0, 0), 0.0, 1.571)]
    lf = [((0, 0, 0), (0.571, 0, 0.821), 0.0, 0.785), ((0, 0, 0.066), (0, -1, 0), -0.349, 0.349)] + f4[1:]
    th = [((0, 0, 0), (0, 0, -1), -1.047, 1.047), ((0, 0, 0), (1, 0, 0), 0.0, 1.222), ((0, 0, 0.038), (1, 0, 0), -0.209, 0.209), ((0, 0, 0), (0, -1, 0), -0.698, 0.698),
          ((0, 0, 0.032), (0, -1, 0), 0.0, 1.571)]
    finger("ff", (0.033, 0, 0.095), f4)
    finger("mf", (0.011, 0, 0.099), f4)
    finger("rf", (-0.011, 0, 0.095), f4)
    finger("lf", (-0.033, 0, 0.0207), lf)
    finger("th", (0.034, -0.0085, 0.029), th, base_rpy=(0, 0.785, 0))
    rm = RobotModel("shadow_like_hand", links)
    groups = {"hand": JointModelGroup(rm, "hand", joints, tips)}
    return rm, groups


def random_tree(seed=0, n_joints=9, branch_at=4, prismatic_every=4):
    """Randomised two-tip tree with rotated joint origins and mixed revolute /
    prismatic / fixed joints: a test robot that exercises every quaternion path."""
    rng = np.random.default_rng(seed)
    links = [Link("root", None, JOINT_FIXED, xyz=rng.uniform(-0.1, 0.1, 3), rpy=rng.uniform(-1, 1, 3), joint_name="root_joint")]
    names = ["root"]
    joints = []

    def add(parent, name):
        k = len(links)
        if k % 5 == 3:
            jt = JOINT_FIXED
        elif k % prismatic_every == 2:
            jt = JOINT_PRISMATIC
        else:
            jt = JOINT_REVOLUTE
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        if jt == JOINT_PRISMATIC:
            lo, hi = -0.2, 0.3
        else:
            lo, hi = sorted(rng.uniform(-2.5, 2.5, 2))
            if hi - lo < 0.5:
                hi = lo + 0.5
        continuous = jt == JOINT_REVOLUTE and k % 7 == 6
        links.append(Link(name, parent, jt, xyz=rng.uniform(-0.3, 0.3, 3), rpy=rng.uniform(-PI, PI, 3), axis=ax, lower=(-PI if continuous else lo), upper=(PI if continuous else hi),
                          bounded=not continuous, velocity=float(rng.uniform(0.5, 4.0)), joint_name=name + "_joint"))
        if jt != JOINT_FIXED:
            joints.append(name + "_joint")
        return name

    prev = "root"
    trunk = []
    for i in range(n_joints):
        prev = add(prev, f"a{i}")
        trunk.append(prev)
    prev = trunk[branch_at]
    for i in range(3):
        prev = add(prev, f"b{i}")
    rm = RobotModel(f"random_tree{seed}", links)
    groups = {"all": JointModelGroup(rm, "all", joints, [trunk[-1], prev])}
    return rm, groups


def mimic_gripper_arm():
    """4-DOF arm with a two-finger gripper whose second finger MIMICS the first (factor -1.5, offset 0.01) and a
    second mimic chained on it: exercises updateMimic, the mimic branches of the Jacobian (joint_dependencies, scale
    products) and a goal on a link below a mimic joint (src/forward_kinematics.h:230-246,581-587,623-630)."""
    links = [
        Link("base", None, JOINT_FIXED, joint_name="world_joint"),
        Link("l1", "base", JOINT_REVOLUTE, xyz=(0, 0, 0.3), axis=(0, 0, 1), lower=-2.5, upper=2.5, velocity=2.0, joint_name="j1"),
        Link("l2", "l1", JOINT_REVOLUTE, xyz=(0.1, 0, 0.2), rpy=(0.3, 0, 0), axis=(0, 1, 0), lower=-1.8, upper=1.8, velocity=2.0, joint_name="j2"),
        Link("l3", "l2", JOINT_REVOLUTE, xyz=(0.35, 0, 0), axis=(0, 1, 0), lower=-2.2, upper=2.2, velocity=3.0, joint_name="j3"),
        Link("l4", "l3", JOINT_PRISMATIC, xyz=(0.3, 0, 0), rpy=(0, 0.2, 0.1), axis=(1, 0, 0), lower=0.0, upper=0.15, velocity=0.5, joint_name="j4"),
        Link("finger_a", "l4", JOINT_REVOLUTE, xyz=(0.05, 0.03, 0), axis=(0, 0, 1), lower=-0.1, upper=0.8, velocity=1.0, joint_name="finger_a_joint"),
        Link("finger_b", "l4", JOINT_REVOLUTE, xyz=(0.05, -0.03, 0), axis=(0, 0, 1), lower=-1.3, upper=0.2, velocity=1.0, joint_name="finger_b_joint", mimic="finger_a_joint",
             mimic_factor=-1.5, mimic_offset=0.01),
        Link("finger_b_tip", "finger_b", JOINT_REVOLUTE, xyz=(0.04, 0, 0), axis=(0, 0, 1), lower=-1.0, upper=1.0, velocity=1.0, joint_name="finger_b_tip_joint", mimic="finger_b_joint",
             mimic_factor=0.5, mimic_offset=0.0),
        Link("pad_a", "finger_a", JOINT_FIXED, xyz=(0.04, 0, 0), joint_name="pad_a_joint"),
        Link("pad_b", "finger_b_tip", JOINT_FIXED, xyz=(0.03, 0, 0), joint_name="pad_b_joint"),
    ]
    rm = RobotModel("mimic_gripper_arm", links)
    groups = {"all": JointModelGroup(rm, "all", ["j1", "j2", "j3", "j4", "finger_a_joint", "finger_b_joint", "finger_b_tip_joint"], ["pad_a", "pad_b"])}
    return rm, groups


def floating_base_arm():
    """A 3-DOF arm on a FLOATING base joint (7 variables: translation + quaternion): the reference's floating branch of
    getJointFrame (src/forward_kinematics.h:120-127), the numeric-differentiation branch of the Jacobian (:695-726) and the
    quaternion-gene normalisation of reproduce() (src/ik_evolution_2.cpp:118-126,320-324).  The translation is given finite
    bounds: with MoveIt's default infinite ones the reference's wipeout draws random(getMin, getMax) = inf * 0."""
    from ._abi import JOINT_FLOATING
    links = [
        Link("world", None, JOINT_FIXED, joint_name="world_joint"),
        Link("base", "world", JOINT_FLOATING, xyz=(0.1, 0.0, 0.2), rpy=(0.0, 0.1, 0.0), velocity=1.0, joint_name="virtual_joint",
             var_lower=[-0.6, -0.6, -0.3, -1.0, -1.0, -1.0, -1.0], var_upper=[0.6, 0.6, 0.5, 1.0, 1.0, 1.0, 1.0], var_bounded=[1, 1, 1, 1, 1, 1, 1]),
        Link("l1", "base", JOINT_REVOLUTE, xyz=(0, 0, 0.25), axis=(0, 0, 1), lower=-2.5, upper=2.5, velocity=2.0, joint_name="j1"),
        Link("l2", "l1", JOINT_REVOLUTE, xyz=(0.05, 0, 0.2), rpy=(0.2, 0, 0), axis=(0, 1, 0), lower=-1.8, upper=1.8, velocity=2.0, joint_name="j2"),
        Link("l3", "l2", JOINT_REVOLUTE, xyz=(0.3, 0, 0), axis=(0, 1, 0), lower=-2.2, upper=2.2, velocity=3.0, joint_name="j3"),
        Link("ee", "l3", JOINT_FIXED, xyz=(0.25, 0, 0), rpy=(0, 0.3, 0), joint_name="ee_joint"),
        Link("camera", "base", JOINT_FIXED, xyz=(0.1, 0, 0.4), joint_name="camera_joint"),
    ]
    rm = RobotModel("floating_base_arm", links)
    groups = {"all": JointModelGroup(rm, "all", ["virtual_joint", "j1", "j2", "j3"], ["ee", "camera"]), "whole_arm": JointModelGroup(rm, "whole_arm", ["virtual_joint", "j1", "j2", "j3"], ["ee"])}
    return rm, groups


def planar_base_arm():
    """The same 3-DOF arm on a PLANAR base joint (x, y, theta): the default branch of the reference's getJointFrame
    (src/forward_kinematics.h:128-135, MoveIt's computeTransform + Frame(Isometry3d)) and the numeric Jacobian (:695-726).
    x and y are given finite bounds for the same reason as in floating_base_arm."""
    from ._abi import JOINT_PLANAR
    links = [
        Link("world", None, JOINT_FIXED, joint_name="world_joint"),
        Link("base", "world", JOINT_PLANAR, xyz=(0.0, 0.1, 0.05), rpy=(0.0, 0.0, 0.2), velocity=1.0, joint_name="virtual_joint",
             var_lower=[-1.0, -1.0, -3.0], var_upper=[1.0, 1.0, 3.0], var_bounded=[1, 1, 1]),
        Link("l1", "base", JOINT_REVOLUTE, xyz=(0, 0, 0.25), axis=(0, 0, 1), lower=-2.5, upper=2.5, velocity=2.0, joint_name="j1"),
        Link("l2", "l1", JOINT_REVOLUTE, xyz=(0.05, 0, 0.2), rpy=(0.2, 0, 0), axis=(0, 1, 0), lower=-1.8, upper=1.8, velocity=2.0, joint_name="j2"),
        Link("l3", "l2", JOINT_REVOLUTE, xyz=(0.3, 0, 0), axis=(0, 1, 0), lower=-2.2, upper=2.2, velocity=3.0, joint_name="j3"),
        Link("ee", "l3", JOINT_FIXED, xyz=(0.25, 0, 0), rpy=(0, 0.3, 0), joint_name="ee_joint"),
        Link("camera", "base", JOINT_FIXED, xyz=(0.1, 0, 0.4), joint_name="camera_joint"),
    ]
    rm = RobotModel("planar_base_arm", links)
    groups = {"all": JointModelGroup(rm, "all", ["virtual_joint", "j1", "j2", "j3"], ["ee", "camera"]), "whole_arm": JointModelGroup(rm, "whole_arm", ["virtual_joint", "j1", "j2", "j3"], ["ee"])}
    return rm, groups


def mimic_virtual_joint_arm():
    """An arm whose PLANAR joint mimics a prismatic joint and whose FLOATING joint mimics a revolute joint - and a revolute joint that
    mimics the planar one's source again below them.  The reference resolves mimic for any joint type: updateMimic copies the FIRST
    variable (src/forward_kinematics.h:230-246: the x translation of both virtual joints follows its source, their other variables
    keep the seed values), and the Jacobian reaches such a joint through the numeric branch with the variable at the same position
    inside the mimicking joint moved (ivar2, :698-699), scaled by the mimic factor (:624-630)."""
    from ._abi import JOINT_FLOATING, JOINT_PLANAR
    links = [
        Link("base", None, JOINT_FIXED, joint_name="world_joint"),
        Link("l1", "base", JOINT_REVOLUTE, xyz=(0, 0, 0.3), axis=(0, 0, 1), lower=-2.5, upper=2.5, velocity=2.0, joint_name="j1"),
        Link("l2", "l1", JOINT_PRISMATIC, xyz=(0.1, 0, 0.2), rpy=(0.3, 0, 0), axis=(1, 0, 0), lower=-0.2, upper=0.3, velocity=0.5, joint_name="j2"),
        Link("sled", "l2", JOINT_PLANAR, xyz=(0.05, 0.0, 0.1), rpy=(0.0, 0.1, 0.2), velocity=1.0, joint_name="sled_joint", mimic="j2", mimic_factor=0.7, mimic_offset=0.02,
             var_lower=[-1.0, -1.0, -3.0], var_upper=[1.0, 1.0, 3.0], var_bounded=[1, 1, 1]),
        Link("l3", "sled", JOINT_REVOLUTE, xyz=(0.3, 0, 0), axis=(0, 1, 0), lower=-2.2, upper=2.2, velocity=3.0, joint_name="j3"),
        Link("pod", "l3", JOINT_FLOATING, xyz=(0.1, 0.0, 0.05), rpy=(0.1, 0.0, 0.0), velocity=1.0, joint_name="pod_joint", mimic="j3", mimic_factor=-0.25, mimic_offset=0.01,
             var_lower=[-0.6, -0.6, -0.3, -1.0, -1.0, -1.0, -1.0], var_upper=[0.6, 0.6, 0.5, 1.0, 1.0, 1.0, 1.0], var_bounded=[1, 1, 1, 1, 1, 1, 1]),
        Link("l4", "pod", JOINT_REVOLUTE, xyz=(0.2, 0, 0), axis=(0, 0, 1), lower=-1.5, upper=1.5, velocity=1.0, joint_name="j4", mimic="j2", mimic_factor=2.0, mimic_offset=-0.1),
        Link("ee", "l4", JOINT_FIXED, xyz=(0.15, 0, 0), rpy=(0, 0.3, 0), joint_name="ee_joint"),
        Link("probe", "sled", JOINT_FIXED, xyz=(0.0, 0.1, 0.2), joint_name="probe_joint"),
    ]
    rm = RobotModel("mimic_virtual_joint_arm", links)
    groups = {"all": JointModelGroup(rm, "all", ["j1", "j2", "sled_joint", "j3", "pod_joint", "j4"], ["ee", "probe"])}
    return rm, groups


def mimic_virtual_joint_base(rm):
    """a configuration of mimic_virtual_joint_arm whose non-mimicked virtual-joint variables (y, theta; y, z, quaternion) are set"""
    base = np.zeros(rm.n_vars)
    for name, v in (("sled_joint/y", 0.04), ("sled_joint/theta", 0.3), ("pod_joint/trans_y", -0.03), ("pod_joint/trans_z", 0.05),
                    ("pod_joint/rot_x", 0.1), ("pod_joint/rot_y", -0.2), ("pod_joint/rot_z", 0.05), ("pod_joint/rot_w", 0.97)):
        base[rm.variable_names.index(name)] = v
    return base


def balancing_tree(seed=5):
    """random_tree with URDF-style inertials on most links (mass, centre of mass in the link frame): the robot of the
    BalanceGoal tests.  12 of its 13 links carry mass, so the problem has 12+ tip links (every link with mass becomes one)."""
    rm0, _ = random_tree(seed, n_joints=9, branch_at=4)
    rng = np.random.default_rng(100 + seed)
    links = []
    for i, l in enumerate(rm0.links):
        mass = 0.0 if i == 3 else float(rng.uniform(0.2, 3.0))
        links.append(Link(l.name, l.parent, l.joint_type, xyz=l.xyz, quat=l.quat, axis=l.axis, lower=l.lower, upper=l.upper, bounded=l.bounded, velocity=l.velocity,
                          joint_name=l.joint_name, mass=mass, com=rng.uniform(-0.1, 0.1, 3)))
    rm = RobotModel(f"balancing_tree{seed}", links)
    joints = [l.joint_name for l in links if l.joint_type != JOINT_FIXED]
    groups = {"all": JointModelGroup(rm, "all", joints, [links[9].name, links[-1].name])}
    return rm, groups
