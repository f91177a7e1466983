"""Flattened stand-in for moveit::core::RobotModel / JointModelGroup.

The reference reads the kinematic tree from MoveIt; MoveIt
is not available here, so robot descriptions are authored as plain link tables
and flattened into the BioikRobot POD of include/bioik_b200.h.
"""
import math

import numpy as np

from . import _abi


def quat_from_rpy(r, p, y):
    """URDF fixed-axis roll/pitch/yaw -> quaternion (x, y, z, w)."""
    cr, sr = math.cos(r * 0.5), math.sin(r * 0.5)
    cp, sp = math.cos(p * 0.5), math.sin(p * 0.5)
    cy, sy = math.cos(y * 0.5), math.sin(y * 0.5)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy)


class Link:
    """One link and its parent joint (MoveIt: LinkModel + getParentJointModel())."""

    def __init__(self, name, parent=None, joint_type=_abi.JOINT_FIXED, xyz=(0, 0, 0), rpy=(0, 0, 0), quat=None,
                 axis=(0, 0, 1), lower=0.0, upper=0.0, bounded=True, velocity=1.0, joint_name=None, mimic=None,
                 mimic_factor=1.0, mimic_offset=0.0, var_lower=None, var_upper=None, var_bounded=None, mass=0.0, com=(0, 0, 0)):
        self.name = name
        self.parent = parent
        self.joint_type = joint_type
        self.xyz = tuple(float(v) for v in xyz)
        self.quat = tuple(float(v) for v in (quat if quat is not None else quat_from_rpy(*rpy)))
        self.axis = tuple(float(v) for v in axis)
        self.lower, self.upper, self.bounded, self.velocity = float(lower), float(upper), bool(bounded), float(velocity)
        self.joint_name = joint_name or (name + "_joint")
        self.mimic = mimic  # joint name of the mimicked joint
        self.mimic_factor, self.mimic_offset = float(mimic_factor), float(mimic_offset)
        # per-variable bounds of multi-variable joints (floating: 7, planar: 3); None = the MoveIt defaults
        self.var_lower, self.var_upper, self.var_bounded = var_lower, var_upper, var_bounded
        # URDF inertial (mass, centre of mass in the link frame): read only by BalanceGoal (src/goal_types.cpp:236-250)
        self.mass, self.com = float(mass), tuple(float(v) for v in com)


class RobotModel:
    def __init__(self, name, links):
        self.name = name
        self.links = list(links)
        self.link_index = {l.name: i for i, l in enumerate(self.links)}
        self.joint_index = {l.joint_name: i for i, l in enumerate(self.links)}  # joint -> child link index
        self.variable_names = []
        self.first_var = []
        for l in self.links:
            cnt = _abi.JOINT_VARS[l.joint_type]
            self.first_var.append(len(self.variable_names) if cnt else -1)
            if cnt == 1:
                self.variable_names.append(l.joint_name)
            elif cnt == 7:
                self.variable_names += [f"{l.joint_name}/{s}" for s in ("trans_x", "trans_y", "trans_z", "rot_x", "rot_y", "rot_z", "rot_w")]
            elif cnt == 3:
                self.variable_names += [f"{l.joint_name}/{s}" for s in ("x", "y", "theta")]
        self.n_vars = len(self.variable_names)
        self.variable_index = {n: i for i, n in enumerate(self.variable_names)}
        self._build_arrays()

    def _build_arrays(self):
        L = len(self.links)
        a = self.arrays = {}
        a["link_parent"] = np.array([self.link_index[l.parent] if l.parent is not None else -1 for l in self.links], dtype=np.int32)
        for i, p in enumerate(a["link_parent"]):
            if p >= i:
                raise ValueError("links must be ordered parents first")
        a["joint_type"] = np.array([l.joint_type for l in self.links], dtype=np.int32)
        a["joint_first_var"] = np.array(self.first_var, dtype=np.int32)
        a["link_origin"] = np.ascontiguousarray(np.array([list(l.xyz) + list(l.quat) for l in self.links], dtype=np.float64))
        a["joint_axis"] = np.ascontiguousarray(np.array([l.axis for l in self.links], dtype=np.float64))
        a["joint_mimic"] = np.array([self.joint_index[l.mimic] if l.mimic else -1 for l in self.links], dtype=np.int32)
        a["joint_mimic_factor"] = np.array([l.mimic_factor for l in self.links], dtype=np.float64)
        a["joint_mimic_offset"] = np.array([l.mimic_offset for l in self.links], dtype=np.float64)
        vmin, vmax, vb, vv = [], [], [], []
        for l in self.links:
            cnt = _abi.JOINT_VARS[l.joint_type]
            if cnt == 1:
                vmin.append(l.lower), vmax.append(l.upper), vb.append(int(l.bounded)), vv.append(l.velocity)
            elif cnt == 7:  # MoveIt FloatingJointModel defaults
                vmin += list(l.var_lower) if l.var_lower is not None else [-1e308] * 3 + [-1.0] * 4
                vmax += list(l.var_upper) if l.var_upper is not None else [1e308] * 3 + [1.0] * 4
                vb += list(l.var_bounded) if l.var_bounded is not None else [0] * 3 + [1] * 4
                vv += [l.velocity] * 7
            elif cnt == 3:
                vmin += list(l.var_lower) if l.var_lower is not None else [-1e308, -1e308, -math.pi]
                vmax += list(l.var_upper) if l.var_upper is not None else [1e308, 1e308, math.pi]
                vb += list(l.var_bounded) if l.var_bounded is not None else [0, 0, 0]
                vv += [l.velocity] * 3
        a["var_min"] = np.array(vmin, dtype=np.float64)
        a["var_max"] = np.array(vmax, dtype=np.float64)
        a["var_bounded"] = np.array(vb, dtype=np.int32)
        a["var_max_velocity"] = np.array(vv, dtype=np.float64)
        a["link_mass"] = np.array([l.mass for l in self.links], dtype=np.float64)
        a["link_com"] = np.ascontiguousarray(np.array([l.com for l in self.links], dtype=np.float64))
        assert L == len(a["link_parent"])

    def to_abi(self):
        a = self.arrays
        r = _abi.BioikRobot()
        r.n_links, r.n_vars = len(self.links), self.n_vars
        r.link_parent, r.joint_type, r.joint_first_var = _abi.iptr(a["link_parent"]), _abi.iptr(a["joint_type"]), _abi.iptr(a["joint_first_var"])
        r.link_origin, r.joint_axis = _abi.dptr(a["link_origin"]), _abi.dptr(a["joint_axis"])
        r.joint_mimic = _abi.iptr(a["joint_mimic"])
        r.joint_mimic_factor, r.joint_mimic_offset = _abi.dptr(a["joint_mimic_factor"]), _abi.dptr(a["joint_mimic_offset"])
        r.var_min, r.var_max = _abi.dptr(a["var_min"]), _abi.dptr(a["var_max"])
        r.var_bounded, r.var_max_velocity = _abi.iptr(a["var_bounded"]), _abi.dptr(a["var_max_velocity"])
        r.link_mass, r.link_com = _abi.dptr(a["link_mass"]), _abi.dptr(a["link_com"])
        r._keepalive = self
        return r

    def getVariableCount(self):
        return self.n_vars

    def getLinkModel(self, name):
        return self.links[self.link_index[name]]

    def getJointOfVariable(self, ivar):
        """child link index of the joint owning variable ivar"""
        for i, l in enumerate(self.links):
            cnt = _abi.JOINT_VARS[l.joint_type]
            if cnt and self.first_var[i] <= ivar < self.first_var[i] + cnt:
                return i
        raise IndexError(ivar)

    def sampling_bounds(self, ivar):
        """(lo, hi) used to draw random configurations: the variable bounds,
        U(-pi, pi) for unbounded / continuous variables."""
        a = self.arrays
        lo, hi = a["var_min"][ivar], a["var_max"][ivar]
        if not a["var_bounded"][ivar] or not np.isfinite(hi - lo) or hi - lo > 1e6:
            return -math.pi, math.pi
        return lo, hi


class JointModelGroup:
    """Named subset of joints + end-effector tips (MoveIt JointModelGroup as used
    by src/problem.cpp:117-124,201-204 and src/kinematics_plugin.cpp:232-237)."""

    def __init__(self, robot_model, name, joint_names, tip_links):
        self.robot_model = robot_model
        self.name = name
        self.joint_names = list(joint_names)
        self.tip_links = list(tip_links)

    def getActiveJointModels(self):
        rm = self.robot_model
        out = []
        for jn in self.joint_names:
            l = rm.links[rm.joint_index[jn]]
            if l.joint_type != _abi.JOINT_FIXED and not l.mimic:
                out.append(jn)
        return out

    def getVariableNames(self):
        rm = self.robot_model
        out = []
        for jn in self.joint_names:
            i = rm.joint_index[jn]
            cnt = _abi.JOINT_VARS[rm.links[i].joint_type]
            out += rm.variable_names[rm.first_var[i]:rm.first_var[i] + cnt] if cnt else []
        return out
