"""Host-side Problem construction, mirroring src/problem.cpp:72-228: collects tip
links and active variables from the goals and the joint group, then flattens
everything into the BioikProblem POD that crosses the C ABI."""
import ctypes as C
import sys

import numpy as np

from . import _abi

DBL_MAX = sys.float_info.max


class Problem:
    def __init__(self):
        self.robot_model = None
        self.joint_model_group = None
        self.goal_list = []  # goals in the order given (primary and secondary interleaved)
        self.tip_link_indices = []
        self.active_variables = []
        self.dpos, self.drot, self.dtwist = DBL_MAX, DBL_MAX, 1e-5

    # src/problem.cpp:57-65
    def _addTipLink(self, link_index):
        if self._link_tip_indices[link_index] < 0:
            self._link_tip_indices[link_index] = len(self.tip_link_indices)
            self.tip_link_indices.append(link_index)
        return self._link_tip_indices[link_index]

    # src/problem.cpp:103-126
    def _addActiveVariable(self, name, fixed_joints):
        rm = self.robot_model
        ivar = rm.variable_index.get(name)
        if ivar is None:
            raise RuntimeError(f"joint variable not found {name}")
        joint_name = rm.links[rm.getJointOfVariable(ivar)].joint_name
        if joint_name in fixed_joints:
            return -1 - ivar
        for i, v in enumerate(self.active_variables):
            if v == ivar:
                return i
        if name in self.joint_model_group.getVariableNames():
            self.active_variables.append(ivar)
            return len(self.active_variables) - 1
        raise RuntimeError(f"joint variable not found {name}")

    def initialize(self, robot_model, joint_model_group, goals, fixed_joints=(), dpos=DBL_MAX, drot=DBL_MAX, dtwist=1e-5):
        self.robot_model, self.joint_model_group = robot_model, joint_model_group
        self.dpos, self.drot, self.dtwist = dpos, drot, dtwist
        rm = robot_model
        self._link_tip_indices = [-1] * len(rm.links)
        self.tip_link_indices = []
        self.active_variables = []
        self.goal_list = []
        for goal in goals:
            rec = {"goal": goal, "tip": 0, "var": 0}
            links = []
            for link_name in (goal.describe_links(rm) if hasattr(goal, "describe_links") else goal.link_names()):
                if link_name not in rm.link_index:
                    raise RuntimeError(f"link not found {link_name}")
                links.append(self._addTipLink(rm.link_index[link_name]))
            for variable_name in goal.variable_names():
                self._addActiveVariable(variable_name, fixed_joints)
                rec["var"] = rm.variable_index[variable_name]
            if links:
                rec["tip"] = links[0]
            self.goal_list.append(rec)
        # src/problem.cpp:191-204: active variables from the active subtree
        joint_usage = [0] * len(rm.links)
        for tip in self.tip_link_indices:
            l = tip
            while l >= 0:
                joint_usage[l] = 1
                l = int(rm.arrays["link_parent"][l])
        for jn in fixed_joints:
            joint_usage[rm.joint_index[jn]] = 0
        for jn in joint_model_group.getActiveJointModels():
            ji = rm.joint_index[jn]
            if joint_usage[ji] and not rm.links[ji].mimic:
                cnt = _abi.JOINT_VARS[rm.links[ji].joint_type]
                for n in rm.variable_names[rm.first_var[ji]:rm.first_var[ji] + cnt]:
                    self._addActiveVariable(n, fixed_joints)
        self._flatten()
        return self

    @property
    def n_goals(self):
        return len(self.goal_list)

    def _flatten(self):
        self._tips = np.array(self.tip_link_indices, dtype=np.int32)
        self._active = np.array(self.active_variables, dtype=np.int32)
        G = len(self.goal_list)
        self._goals = (_abi.BioikGoal * max(G, 1))()
        for i, rec in enumerate(self.goal_list):
            g = rec["goal"]
            bg = self._goals[i]
            bg.type, bg.tip, bg.secondary, bg.var = g.type, rec["tip"], int(g.isSecondary()), rec["var"]
            bg.weight = g.getWeight()
            for k, v in enumerate(g.params()):
                bg.p[k] = v

    def default_goal_params(self):
        """[n_goals][GOAL_NPARAM] parameter block of the goals as constructed."""
        out = np.zeros((self.n_goals, _abi.GOAL_NPARAM))
        for i, rec in enumerate(self.goal_list):
            out[i] = rec["goal"].params()
        return out

    def to_abi(self):
        p = _abi.BioikProblem()
        p.n_tips, p.tip_links = len(self._tips), _abi.iptr(self._tips)
        p.n_active, p.active_vars = len(self._active), _abi.iptr(self._active)
        p.n_goals, p.goals = len(self.goal_list), C.cast(self._goals, C.POINTER(_abi.BioikGoal))
        p.dpos, p.drot, p.dtwist = self.dpos, self.drot, self.dtwist
        p._keepalive = self
        return p
