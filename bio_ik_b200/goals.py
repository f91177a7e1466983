"""Goal classes with the reference's names and constructor arguments
(include/bio_ik/goal.h:97-119, include/bio_ik/goal_types.h).  Each goal only
*describes* itself into the flattened BioikGoal record; evaluation happens in the
CUDA kernels (and, independently, in the test oracle).
"""
import math

from . import _abi


class UnsupportedGoal(Exception):
    """Goal needs a host callback or FCL: BIOIK_E_UNSUPPORTED_GOAL.  Keep using the
    stock CPU solver for such queries; this package has no CPU fallback."""


def _v3(v):
    v = tuple(float(x) for x in v)
    assert len(v) == 3
    return v


def _normalized3(v):
    v = _v3(v)
    s = 1.0 / math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])  # tf2: v * (1.0 / length())
    return (v[0] * s, v[1] * s, v[2] * s)


def _normalized4(q):
    q = tuple(float(x) for x in q)
    assert len(q) == 4
    s = 1.0 / math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
    return (q[0] * s, q[1] * s, q[2] * s, q[3] * s)


class Goal:
    """include/bio_ik/goal.h:97-119"""
    type = 0

    def __init__(self):
        self.weight_ = 1.0
        self.secondary_ = False

    def isSecondary(self):
        return self.secondary_

    def getWeight(self):
        return self.weight_

    def setWeight(self, w):
        self.weight_ = float(w)

    # what GoalContext collects in describe(): link names and variable names
    def link_names(self):
        return []

    def variable_names(self):
        return []

    def params(self):
        return [0.0] * _abi.GOAL_NPARAM


class LinkGoalBase(Goal):
    """goal_types.h:56-78"""

    def __init__(self, link_name="", weight=1.0):
        super().__init__()
        self.weight_ = float(weight)
        self.link_name_ = link_name

    def setLinkName(self, n):
        self.link_name_ = n

    def getLinkName(self):
        return self.link_name_

    def link_names(self):
        return [self.link_name_]


def _pad(p):
    return list(p) + [0.0] * (_abi.GOAL_NPARAM - len(p))


class PositionGoal(LinkGoalBase):
    type = _abi.GOAL_POSITION

    def __init__(self, link_name="", position=(0, 0, 0), weight=1.0):
        super().__init__(link_name, weight)
        self.position_ = _v3(position)

    def setPosition(self, p):
        self.position_ = _v3(p)

    def params(self):
        return _pad(self.position_)


class OrientationGoal(LinkGoalBase):
    type = _abi.GOAL_ORIENTATION

    def __init__(self, link_name="", orientation=(0, 0, 0, 1), weight=1.0):
        super().__init__(link_name, weight)
        self.orientation_ = _normalized4(orientation)

    def setOrientation(self, q):
        self.orientation_ = _normalized4(q)

    def params(self):
        return _pad([0.0, 0.0, 0.0] + list(self.orientation_))


class PoseGoal(LinkGoalBase):
    type = _abi.GOAL_POSE

    def __init__(self, link_name="", position=(0, 0, 0), orientation=(0, 0, 0, 1), weight=1.0):
        super().__init__(link_name, weight)
        self.position_ = _v3(position)
        self.orientation_ = _normalized4(orientation)
        self.rotation_scale_ = 0.5

    def setPosition(self, p):
        self.position_ = _v3(p)

    def setOrientation(self, q):
        self.orientation_ = _normalized4(q)

    def setRotationScale(self, s):
        self.rotation_scale_ = float(s)

    def getRotationScale(self):
        return self.rotation_scale_

    def params(self):
        return _pad(list(self.position_) + list(self.orientation_) + [self.rotation_scale_])


class LookAtGoal(LinkGoalBase):
    type = _abi.GOAL_LOOK_AT

    def __init__(self, link_name="", axis=(1, 0, 0), target=(0, 0, 0), weight=1.0):
        super().__init__(link_name, weight)
        self.axis_, self.target_ = _v3(axis), _v3(target)  # ctor does not normalise (goal_types.h:194-199)

    def setAxis(self, a):
        self.axis_ = _normalized3(a)

    def setTarget(self, t):
        self.target_ = _v3(t)

    def params(self):
        return _pad(list(self.axis_) + list(self.target_))


class MaxDistanceGoal(LinkGoalBase):
    type = _abi.GOAL_MAX_DISTANCE

    def __init__(self, link_name="", target=(0, 0, 0), distance=1.0, weight=1.0):
        super().__init__(link_name, weight)
        self.target, self.distance = _v3(target), float(distance)

    def setTarget(self, t):
        self.target = _v3(t)

    def setDistance(self, d):
        self.distance = float(d)

    def params(self):
        return _pad(list(self.target) + [self.distance])


class MinDistanceGoal(MaxDistanceGoal):
    type = _abi.GOAL_MIN_DISTANCE


class LineGoal(LinkGoalBase):
    type = _abi.GOAL_LINE

    def __init__(self, link_name="", position=(0, 0, 0), direction=(0, 0, 1), weight=1.0):
        super().__init__(link_name, weight)
        self.position, self.direction = _v3(position), _normalized3(direction)

    def setPosition(self, p):
        self.position = _v3(p)

    def setDirection(self, d):
        self.direction = _normalized3(d)

    def params(self):
        return _pad(list(self.position) + list(self.direction))


class PlaneGoal(LinkGoalBase):
    type = _abi.GOAL_PLANE

    def __init__(self, link_name="", position=(0, 0, 0), normal=(0, 0, 1), weight=1.0):
        super().__init__(link_name, weight)
        self.position, self.normal = _v3(position), _normalized3(normal)

    def setPosition(self, p):
        self.position = _v3(p)

    def setNormal(self, n):
        self.normal = _normalized3(n)

    def params(self):
        return _pad(list(self.position) + list(self.normal))


class AvoidJointLimitsGoal(Goal):
    type = _abi.GOAL_AVOID_JOINT_LIMITS

    def __init__(self, weight=1.0, secondary=True):
        super().__init__()
        self.weight_, self.secondary_ = float(weight), bool(secondary)


class CenterJointsGoal(Goal):
    type = _abi.GOAL_CENTER_JOINTS

    def __init__(self, weight=1.0, secondary=True):
        super().__init__()
        self.weight_, self.secondary_ = float(weight), bool(secondary)


class RegularizationGoal(Goal):
    type = _abi.GOAL_REGULARIZATION

    def __init__(self, weight=1.0):
        super().__init__()
        self.weight_ = float(weight)


class MinimalDisplacementGoal(Goal):
    type = _abi.GOAL_MINIMAL_DISPLACEMENT

    def __init__(self, weight=1.0, secondary=True):
        super().__init__()
        self.weight_, self.secondary_ = float(weight), bool(secondary)


class JointVariableGoal(Goal):
    type = _abi.GOAL_JOINT_VARIABLE

    def __init__(self, variable_name="", variable_position=0.0, weight=1.0, secondary=False):
        super().__init__()
        self.variable_name, self.variable_position = variable_name, float(variable_position)
        self.weight_, self.secondary_ = float(weight), bool(secondary)

    def setVariablePosition(self, p):
        self.variable_position = float(p)

    def variable_names(self):
        return [self.variable_name]

    def params(self):
        return _pad([self.variable_position])


class SideGoal(LinkGoalBase):
    type = _abi.GOAL_SIDE

    def __init__(self, link_name="", axis=(0, 0, 1), direction=(0, 0, 1), weight=1.0):
        super().__init__(link_name, weight)
        self.axis, self.direction = _v3(axis), _v3(direction)  # ctor does not normalise (goal_types.h:596-601)

    def setAxis(self, a):
        self.axis = _normalized3(a)

    def setDirection(self, d):
        self.direction = _normalized3(d)

    def params(self):
        return _pad(list(self.axis) + list(self.direction))


class DirectionGoal(SideGoal):
    type = _abi.GOAL_DIRECTION


class ConeGoal(LinkGoalBase):
    """goal_types.h:646-712.  The three reference constructors: (link, axis, direction, angle, weight),
    (link, position, axis, direction, angle, weight) -> position_weight 1, and
    (link, position, position_weight, axis, direction, angle, weight)."""
    type = _abi.GOAL_CONE

    def __init__(self, link_name="", axis=(0, 0, 1), direction=(0, 0, 1), angle=0.0, weight=1.0, position=None, position_weight=None):
        super().__init__(link_name, weight)
        self.axis, self.direction, self.angle = _v3(axis), _v3(direction), float(angle)  # constructors do not normalise
        self.position = _v3(position) if position is not None else (0.0, 0.0, 0.0)
        self.position_weight = float(position_weight) if position_weight is not None else (1.0 if position is not None else 0.0)

    def setAxis(self, a):
        self.axis = _normalized3(a)

    def setDirection(self, d):
        self.direction = _normalized3(d)

    def setAngle(self, a):
        self.angle = float(a)

    def setPosition(self, p):
        self.position = _v3(p)

    def setPositionWeight(self, w):
        self.position_weight = float(w)

    def params(self):
        return _pad(list(self.position) + [self.position_weight] + list(self.axis) + list(self.direction) + [self.angle])


class _HostOnlyGoal(Goal):
    def __init__(self, *a, **k):
        raise UnsupportedGoal(f"{type(self).__name__} needs a host callback / FCL and cannot run on the device "
                              "(BIOIK_E_UNSUPPORTED_GOAL); use the reference CPU solver for this query")


class JointFunctionGoal(_HostOnlyGoal):
    """goal_types.h:501-538 (std::function callback)"""


class LinkFunctionGoal(_HostOnlyGoal):
    """goal_types.h:570-583 (std::function callback)"""


class TouchGoal(_HostOnlyGoal):
    """src/goal_types.cpp:46-228 (FCL collision geometry)"""


class BalanceGoal(Goal):
    """goal_types.h:540-568, src/goal_types.cpp:231-272: the centre of mass - the mass-weighted mean of every link's inertial origin,
    taken over the links whose URDF inertial has a positive mass - is pulled onto the line through `target` along `axis`.
    describe() makes each of those links a goal link (in link order), so they all become tip links of the problem."""
    type = _abi.GOAL_BALANCE

    def __init__(self, target=(0, 0, 0), weight=1.0, axis=(0, 0, 1)):
        super().__init__()
        self.weight_ = float(weight)
        self.target_, self.axis_ = _v3(target), _v3(axis)  # neither the constructor nor the setters normalise

    def setTarget(self, t):
        self.target_ = _v3(t)

    def setAxis(self, a):
        self.axis_ = _v3(a)

    def getTarget(self):
        return self.target_

    def getAxis(self):
        return self.axis_

    def describe_links(self, robot_model):
        return [l.name for l in robot_model.links if l.mass > 0]

    def params(self):
        return _pad(list(self.target_) + list(self.axis_))
